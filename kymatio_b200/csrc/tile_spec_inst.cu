// tile_spec_inst.cu - static instances of the fused tile kernel for paths WITH children (SPEC = true: the modulus field is
// kept as (|u|, 0), the forward transform and the natural-order spectrum store are compiled in).  Separate translation
// unit so that it builds in parallel with tile_inst.cu.
#include "tile2d.cuh"
#include "common.cuh"

namespace sb {

#define SB_TILE_SIZES(X) X(136) X(68) X(128) X(64) X(32) X(40) X(20) X(36) X(18)

template <typename T> TileKernel<T> tile_spec_kernel_lookup(int n0, int n1, int k) {
    if (n0 == n1) {
#define SB_CASE(N) if (n0 == N) { if (k == 2) return k2d_tile<T, N, N, 2, true>; if (k == 4) return k2d_tile<T, N, N, 4, true>; return nullptr; }
        SB_TILE_SIZES(SB_CASE)
#undef SB_CASE
    }
    return nullptr;
}

template <typename T> void tile_spec_kernels_enable_smem() {
#define SB_EN(N) enable_big_smem(k2d_tile<T, N, N, 2, true>); enable_big_smem(k2d_tile<T, N, N, 4, true>);
    SB_TILE_SIZES(SB_EN)
#undef SB_EN
}

template TileKernel<float> tile_spec_kernel_lookup<float>(int, int, int);
template TileKernel<double> tile_spec_kernel_lookup<double>(int, int, int);
template void tile_spec_kernels_enable_smem<float>();
template void tile_spec_kernels_enable_smem<double>();

}  // namespace sb
