// slab.cuh - block-level batched 1-D transforms on a slab of lines held in shared memory.
#pragma once
#include <cuda_runtime.h>
#include "fft_core.cuh"

namespace sb {

// upper bound on threads per CTA for every slab kernel (caps registers at 65536/576 = 113)
constexpr int kMaxThreads = 576;

__device__ __forceinline__ int flat_tid() { return threadIdx.y * blockDim.x + threadIdx.x; }
__device__ __forceinline__ int flat_nt() { return blockDim.x * blockDim.y; }

// Transform `lines` lines of length P.n in shared memory, in place.
//   element e of line l lives at s[l*lstride + e*estride]
//   INV=false: DIF forward (natural -> scrambled); INV=true: DIT inverse, unnormalised
// Work items (butterfly, line) are flattened over the whole CTA with the line index
// fastest, so consecutive lanes touch consecutive lines (which every caller lays out
// conflict-free) and any lines x butterflies shape keeps all threads busy.
// One __syncthreads per pass; ends synchronised.
template <bool INV, typename T>
__device__ __noinline__ void slab_fft(cx<T>* s, int lines, int lstride, int estride, const Plan1& P,
                                      const cx<T>* tw) {
    const int tid = flat_tid(), nt = flat_nt();
    for (int pp = 0; pp < P.npass; ++pp) {
        const int p = INV ? P.npass - 1 - pp : pp;
        const int r = P.radix[p], m = P.blen[p];
        const int q = m / r, nbf = P.n / r, tws = P.n / m;
        const int items = nbf * lines;
        for (int it = tid; it < items; it += nt) {
            const int bf = it / lines, line = it - bf * lines;
            const int blk = bf / q, i = bf - blk * q;
            butterfly_dispatch<INV, T>(r, s + line * lstride, estride, blk * m + i, q, i * tws, tw, q > 1, P.n);
        }
        __syncthreads();
    }
}

// Compile-time variant: length N, line stride LS and element stride ES are constants, so the
// pass loop is fully unrolled and every butterfly access is base + immediate offset.
//   DIT = false: decimation in frequency, natural -> scrambled;  DIT = true: scrambled -> natural
//   SIGN = -1 forward, +1 inverse (unnormalised)
// MODULUS = 1 (2) replaces every output of the LAST pass by (|v|, 0) ((|v|, |v|)) while it is still in registers.
// PFA = true runs the prime-factor variant (fft_core.cuh: no twiddles between the two passes; positions pfa_in/pfa_out).
template <int N, bool DIT, int SIGN, int LS, int ES, typename T, int MODULUS = 0, bool PFA = false>
__device__ __forceinline__ void slab_fft_s(cx<T>* s, const int lines, const cx<T>* tw) {
    constexpr int NP = ct_plan1(N).npass;
    static_assert(!PFA || ct_pfa_ok(N), "prime-factor variant needs N = 2^a * odd prime with one pass each");
    const int tid = flat_tid(), nt = flat_nt();
    static_for<0, NP>([&](auto pp_) {
        constexpr int pp = decltype(pp_)::value;
        constexpr int p = DIT ? NP - 1 - pp : pp;
        constexpr int r = ct_plan1(N).radix[p], m = ct_plan1(N).blen[p];
        constexpr int q = m / r, nbf = N / r, tws = N / m;
        const int items = nbf * lines;
        for (int it = tid; it < items; it += nt) {
            const int bf = it / lines, line = it - bf * lines;
            const int blk = bf / q, i = bf - blk * q;
            butterfly_s<r, DIT, SIGN, q, ES, (pp == NP - 1 ? MODULUS : 0), T, (q > 1 && !PFA)>(s + line * LS + (blk * m + i) * ES, i * tws, tw);
        }
        __syncthreads();
    });
}

template <typename T>
__device__ __forceinline__ void copy_tw(cx<T>* dst, const cx<T>* __restrict__ src, int n) {
    for (int i = flat_tid(); i < n; i += flat_nt()) dst[i] = src[i];
}

// periodic reflection index (numpy 'reflect' / torch ReflectionPad2d semantics extended
// to pad == size, kymatio/scattering2d/backend/torch_backend.py:49-54,80-83)
__device__ __forceinline__ int reflect_idx(int t, int n) {
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    t %= period;
    if (t < 0) t += period;
    return t < n ? t : period - t;
}

}  // namespace sb
