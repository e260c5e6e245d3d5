// tile_inst.cu - explicit instances of the fused tile kernel for the field sizes of the
// BASELINE.json configurations, plus the generic runtime-size fallback.
//   272-padded (256^2, J=3): 136, 68        256-padded (224^2, J=4): 128, 64, 32        40-padded (32^2, J=2): 40, 20        36-padded (28^2, J=2): 36, 18
#include "tile2d.cuh"
#include "common.cuh"

namespace sb {

#define SB_TILE_SIZES(X) X(136) X(68) X(128) X(64) X(32) X(40) X(20) X(36) X(18)

template <typename T> TileKernel<T> tile_kernel_lookup(int n0, int n1, int k, bool* is_static) {
    if (is_static) *is_static = true;
    if (n0 == n1) {
#define SB_CASE(N)                                           \
        if (n0 == N) {                                       \
            if (k == 2) return k2d_tile<T, N, N, 2>;         \
            if (k == 4) return k2d_tile<T, N, N, 4>;         \
            return k2d_tile<T, N, N, 0, true>;               \
        }
        SB_TILE_SIZES(SB_CASE)
#undef SB_CASE
    }
    if (is_static) *is_static = false;
    return k2d_tile<T, 0, 0, 0, true>;
}

template <typename T> void tile_kernels_enable_smem() {
#define SB_EN(N) enable_big_smem(k2d_tile<T, N, N, 2>); enable_big_smem(k2d_tile<T, N, N, 4>); \
                 enable_big_smem(k2d_tile<T, N, N, 0, true>);
    SB_TILE_SIZES(SB_EN)
#undef SB_EN
    enable_big_smem(k2d_tile<T, 0, 0, 0, true>);
    tile_bwd_kernels_enable_smem<T>();      // tile_bwd_inst.cu
}

// profiling build: copy out (and optionally reset) the per-phase cycle counters; the production build reports 0 slots
int phase_prof_read(unsigned long long* out, int max_n, bool reset) {
#ifdef SB_PHASE_PROF
    const int n = std::min(max_n, kPhaseKinds * kPhaseSlots);
    SB_CUDA(cudaDeviceSynchronize());
    SB_CUDA(cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * n));
    {   // the counters are per translation unit: add the backward tiles' copy (kinds 24..31, tile_bwd_inst.cu)
        unsigned long long tmp[kPhaseKinds * kPhaseSlots];
        phase_prof_read_bwd(tmp, reset);
        for (int i = 0; i < n; ++i) out[i] += tmp[i];
    }
    if (reset) {
        static const unsigned long long zeros[kPhaseKinds * kPhaseSlots] = {};
        SB_CUDA(cudaMemcpyToSymbol(g_phase_cycles, zeros, sizeof zeros));
    }
    return n;
#else
    (void)out; (void)max_n; (void)reset;
    return 0;
#endif
}

template TileKernel<float> tile_kernel_lookup<float>(int, int, int, bool*);
template TileKernel<double> tile_kernel_lookup<double>(int, int, int, bool*);
template void tile_kernels_enable_smem<float>();
template void tile_kernels_enable_smem<double>();

}  // namespace sb
