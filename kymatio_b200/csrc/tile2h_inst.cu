// tile2h_inst.cu - instances of the two-half leaf tile kernel (tile2h.cuh): float, the static field sizes of the
// BASELINE.json configurations.
#include "tile2h.cuh"
#include "common.cuh"

namespace sb {

#define SB_TILE2H_SIZES(X) X(136) X(128) X(68) X(64)

template <typename T> TileKernel<T> tile2h_kernel_lookup(int n0, int n1, int k) {
    if constexpr (std::is_same<T, float>::value) {
        if (n0 == n1) {
#define SB_CASE(N) if (n0 == N) { if (k == 2) return k2d_tile2h<T, N, 2>; if (k == 4) return k2d_tile2h<T, N, 4>; return nullptr; }
            SB_TILE2H_SIZES(SB_CASE)
#undef SB_CASE
        }
    }
    return nullptr;
}

template <typename T> void tile2h_kernels_enable_smem() {
    if constexpr (std::is_same<T, float>::value) {
#define SB_EN(N) enable_big_smem(k2d_tile2h<T, N, 2>); enable_big_smem(k2d_tile2h<T, N, 4>);
        SB_TILE2H_SIZES(SB_EN)
#undef SB_EN
    }
}

template TileKernel<float> tile2h_kernel_lookup<float>(int, int, int);
template TileKernel<double> tile2h_kernel_lookup<double>(int, int, int);
template void tile2h_kernels_enable_smem<float>();
template void tile2h_kernels_enable_smem<double>();

}  // namespace sb
