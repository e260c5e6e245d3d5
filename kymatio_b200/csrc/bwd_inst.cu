// bwd_inst.cu - instances of the fused first-order backward kernels (bwd2d.cuh) for the static field sizes of the
// BASELINE.json configurations.
#include "bwd2d.cuh"
#include "common.cuh"

namespace sb {

#define SB_TILE_SIZES(X) X(136) X(68) X(128) X(64) X(32) X(40) X(20) X(36) X(18)
#define SB_STREAM_SIZES(X) X(272) X(256) X(240)

template <typename T> TileAdjKernel<T> tile_adj_lookup(int n0, int n1) {
    if (n0 == n1) {
#define SB_CASE(N) if (n0 == N) return k2d_tile_adj<T, N, N>;
        SB_TILE_SIZES(SB_CASE)
#undef SB_CASE
    }
    return k2d_tile_adj_g<T>;           // runtime-size instance (natural-order R, pairs with k2d_tile_bwd<T, 0, 0, 0>)
}
template <typename T> BwdColKernel<T> bwd_col_lookup(int n) {
#define SB_CASE(N) if (n == N) return k2d_bwd_col<T, N>;
    SB_STREAM_SIZES(SB_CASE)
#undef SB_CASE
    return nullptr;
}
template <typename T> BwdRowKernel<T> bwd_row_lookup(int n) {
#define SB_CASE(N) if (n == N) return k2d_bwd_row<T, N>;
    SB_STREAM_SIZES(SB_CASE)
#undef SB_CASE
    return nullptr;
}
template <typename T> void bwd_kernels_enable_smem() {
#define SB_EN(N) enable_big_smem(k2d_tile_adj<T, N, N>);
    SB_TILE_SIZES(SB_EN)
#undef SB_EN
    enable_big_smem(k2d_tile_adj_g<T>);
#define SB_EN(N) enable_big_smem(k2d_bwd_col<T, N>); enable_big_smem(k2d_bwd_row<T, N>);
    SB_STREAM_SIZES(SB_EN)
#undef SB_EN
}

#define SB_INST(T)                                                   \
    template TileAdjKernel<T> tile_adj_lookup<T>(int, int);          \
    template BwdColKernel<T> bwd_col_lookup<T>(int);                 \
    template BwdRowKernel<T> bwd_row_lookup<T>(int);                 \
    template void bwd_kernels_enable_smem<T>();
SB_INST(float)
SB_INST(double)

}  // namespace sb
