// tile_bwd_inst.cu - explicit instances of the fused BACKWARD tile kernel (tile2d.cuh: k2d_tile_bwd) for the field sizes
// of tile_inst.cu, plus the runtime-size fallback.  Separate translation unit so that it builds in parallel.
#include "tile2d.cuh"
#include "common.cuh"

namespace sb {

#define SB_TILE_SIZES(X) X(136) X(68) X(128) X(64) X(32) X(40) X(20) X(36) X(18)

// backward instances: alias counts 2 and 4 static, the generic alias-count version (K = 0) for the rest
template <typename T> TileKernel<T> tile_bwd_kernel_lookup(int n0, int n1, int k, bool* is_static) {
    if (is_static) *is_static = true;
    if (n0 == n1) {
#define SB_CASE(N) if (n0 == N) { if (k == 2) return k2d_tile_bwd<T, N, N, 2>; if (k == 4) return k2d_tile_bwd<T, N, N, 4>; \
                                 return k2d_tile_bwd<T, N, N, 0>; }
        SB_TILE_SIZES(SB_CASE)
#undef SB_CASE
    }
    if (is_static) *is_static = false;
    return k2d_tile_bwd<T, 0, 0, 0>;
}

template <typename T> void tile_bwd_kernels_enable_smem() {
#define SB_ENB(N) enable_big_smem(k2d_tile_bwd<T, N, N, 0>); enable_big_smem(k2d_tile_bwd<T, N, N, 2>); \
                  enable_big_smem(k2d_tile_bwd<T, N, N, 4>);
    SB_TILE_SIZES(SB_ENB)
#undef SB_ENB
    enable_big_smem(k2d_tile_bwd<T, 0, 0, 0>);
}

// profiling build: this translation unit's copy of the per-phase counters (see tile_inst.cu: phase_prof_read)
void phase_prof_read_bwd(unsigned long long* out, bool reset) {
#ifdef SB_PHASE_PROF
    SB_CUDA(cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * kPhaseKinds * kPhaseSlots));
    if (reset) {
        static const unsigned long long zeros[kPhaseKinds * kPhaseSlots] = {};
        SB_CUDA(cudaMemcpyToSymbol(g_phase_cycles, zeros, sizeof zeros));
    }
#else
    (void)out; (void)reset;
#endif
}

template TileKernel<float> tile_bwd_kernel_lookup<float>(int, int, int, bool*);
template TileKernel<double> tile_bwd_kernel_lookup<double>(int, int, int, bool*);
template void tile_bwd_kernels_enable_smem<float>();
template void tile_bwd_kernels_enable_smem<double>();

}  // namespace sb
