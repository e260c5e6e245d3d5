// tile2d.cuh - the fused one-CTA-per-path tile kernel.
//
// One CTA = one scattering path (image b, parent spectrum, band-pass filter):
//   Z      = periodise_k(parent * filt) * scale                       (n0 x n1, Fourier)
//   u      = ifft2(Z)                                                 (shared memory)
//   U      = |u|
//   S      = unpad( (U conv g)[::kl, ::kl] ),  g = a (x) b           -> out[b][ch]
//   spec   = fft2(U)   (only when spec_out != nullptr)                -> spec_out[g]
// i.e. cdgmm -> subsample_fourier -> ifft -> modulus -> rfft -> cdgmm(phi) -> subsample_fourier ->
// irfft -> unpad of kymatio/scattering2d/core/scattering2d.py:33-47 (and :59-75) in one pass.
// The separable spatial low-pass equals the reference's Fourier-domain low-pass whenever
// phi_hat = a_hat (x) b_hat / phi_hat[0][0] (checked when the filters are bound); it is applied as
// two small dense products with the decimation matrices G1[x][xo], G0[y][yo] built at bind time.
//
// Template parameters N0, N1 (field size) and K (periodisation factor) select a fully
// specialised instance (static FFT plans, immediate-offset shared-memory accesses, unrolled alias
// loops); N0 = 0 is the generic runtime-size fallback.
#pragma once
#include "kernels2d.cuh"

namespace sb {

// Optional per-phase cycle accounting (profiling build only, -DSB_PHASE_PROF -> libscat_b200_prof.so): thread 0 of
// every CTA adds the cycles between phase boundaries to g_phase_cycles[kernel id][phase]; read back through
// scat_phase_prof_read (tile_inst.cu).  The production build compiles none of it.
#ifdef SB_PHASE_PROF
constexpr int kPhaseKinds = 32, kPhaseSlots = 8;   // 0..23 forward tiles, 24..31 backward tiles
static __device__ unsigned long long g_phase_cycles[kPhaseKinds * kPhaseSlots];
#define SB_PHASE_INIT(kid_expr) const int sb_kid = (kid_expr); long long sb_t_last = clock64();
#define SB_PHASE(i)                                                                            \
    do {                                                                                       \
        __syncthreads();                                                                       \
        if (flat_tid() == 0) {                                                                 \
            const long long sb_t = clock64();                                                  \
            atomicAdd(&g_phase_cycles[sb_kid * kPhaseSlots + (i)], (unsigned long long)(sb_t - sb_t_last)); \
            sb_t_last = sb_t;                                                                  \
        }                                                                                      \
    } while (0)
#else
#define SB_PHASE_INIT(kid_expr)
#define SB_PHASE(i)
#endif

template <typename T> struct TileArgs {
    const cx<T>* parent; const T* const* filt; const int2* supp;
    cx<T>* spec_out; T* out;
    int P0, P1, k, n0, n1, W, NF;
    T scale;
    Plan1 plan0, plan1; const cx<T>* tw0; const cx<T>* tw1; const int* pos0; const int* pos1;
    // prime-factor forward instances (fft_core.cuh): input positions of the natural Fourier indices; pos0/pos1 then hold
    // the matching prime-factor output positions
    const int* pin0; const int* pin1;
    const T* G0; const T* G1;              // [n0][o0p], [n1][o1p] dense low-pass + decimation + unpad matrices
    int y0lo, y0cnt, x1lo, x1cnt;          // input window (start offset rel. to kl*(4*group+1), length) per 4-output group
    int kl, o0, o1, o0p, o1p;              // o?p = o? rounded up to a multiple of 4
    int PP, NFch, ch0, chs, K;
    int G;                                 // number of paths; CTAs are persistent and stride over them
    int use_mma;                           // 1: dense low-pass products on the tensor cores (3xTF32 mma.sync)
    int tt;                                // 1: static forward, CUDA-core low-pass from the [cnt][4] tap tables TT0/TT1
                                           //    (shift invariance, see tile2h.cuh) instead of the dense G0/G1 matrices
    int prefetch;                          // 1: bulk L2 prefetch of the next path's parent spectrum (kernels2d.cuh)
    int stagger_ns;                        // > 0: CTA i starts (i % 4) * stagger_ns late, so that the L2-bound load phases
                                           // of the persistent CTAs do not all coincide
    // backward kernel only: gradient w.r.t. the output planes (same layout / channel mapping as `out`) and the
    // parent-gradient accumulator [NPAR][P0][P1] (atomically added to; zeroed by the caller)
    OutPeers<T> peers;                     // additional destinations of the output planes (peer GPUs), see kernels2d.cuh
    const T* gout; cx<T>* gparent;
    // backward of a path WITH children: R = Re(F^H gU1) per path in storage order (k2d_tile_adj, bwd2d.cuh), added to gA
    const T* radd;
    // two-half leaf kernel only (tile2h.cuh): twiddles / scramble table of the half-length column transform and the
    // [cnt][4] tap tables of the shift-invariant low-pass
    const cx<T>* twh; const int* posh; const T* TT0; const T* TT1;
};

template <typename T> struct TileSmem {
    cx<T>* tile; cx<T>* tw0; cx<T>* tw1; unsigned* supp; T* w1; T* G0; T* G1; int* pos0; int* pos1; T* gs; int2* xr;
    int* pin0; int* pin1; T* pb;
};
template <typename T> __host__ __device__ inline size_t tile_smem_layout(const TileArgs<T>& a, TileSmem<T>* L) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 15) / 16 * 16; return o; };
    const size_t o_tile = take(sizeof(cx<T>) * (size_t)a.n0 * a.W);
    const size_t o_tw0 = take(sizeof(cx<T>) * a.n0), o_tw1 = take(sizeof(cx<T>) * a.n1);
    const size_t o_supp = take(sizeof(unsigned) * 2 * a.P0);  // packed (start | len << 16), double buffered: the next path's rows are staged early
    // pitches: +4 (CUDA-core path, 16-byte rows) or +8 (mma path, conflict-free fragment loads); K padded to 8.
    // tt mode: two w1 halves (the x window is split over two work items) and the tap tables in place of G0/G1
    const size_t o_w1 = take(a.tt ? sizeof(T) * 2 * (size_t)a.n0 * (a.o1p + 4) : sizeof(T) * (size_t)((a.n0 + 7) & ~7) * (a.o1p + 8));
    const size_t o_g0 = take(a.tt ? sizeof(T) * 4 * (size_t)a.y0cnt : sizeof(T) * (size_t)((a.n0 + 7) & ~7) * (a.o0p + 8));
    // backward only (the forward kernel must not pay for them: 68 x 68 tiles fit three per SM without)
    const bool bwd = a.gparent != nullptr;
    // the backward kernel keeps G1 transposed, [o1p][n1 rounded up to 32] (tile_bwd_body)
    const size_t g1_fwd = a.tt ? sizeof(T) * 4 * (size_t)a.x1cnt : sizeof(T) * (size_t)((a.n1 + 7) & ~7) * (a.o1p + 8);
    const size_t g1_bwd = bwd ? sizeof(T) * (size_t)a.o1p * ((a.n1 + 31) & ~31) : 0;
    const size_t o_g1 = take(g1_fwd > g1_bwd ? g1_fwd : g1_bwd);
    const size_t o_p0 = take(sizeof(int) * a.n0), o_p1 = take(sizeof(int) * a.n1);
    const size_t o_gs = take(bwd ? sizeof(T) * (size_t)a.o0p * a.o1p : 0);   // staged output-plane gradient
    const size_t o_xr = take(bwd ? sizeof(int2) * (size_t)a.n1 : 0);         // nonzero output range of each G1 row
    const size_t o_i0 = take(a.pin0 ? sizeof(int) * a.n0 : 0), o_i1 = take(a.pin0 ? sizeof(int) * a.n1 : 0);
    const size_t o_pb = take(a.tt ? sizeof(T) * (size_t)a.o0p * a.o1p : 0);    // 4b partial sums of the second half window
    if (L) {
#ifdef __CUDA_ARCH__
        unsigned char* base = dyn_smem<unsigned char>();
        L->tile = reinterpret_cast<cx<T>*>(base + o_tile);
        L->tw0 = reinterpret_cast<cx<T>*>(base + o_tw0); L->tw1 = reinterpret_cast<cx<T>*>(base + o_tw1);
        L->supp = reinterpret_cast<unsigned*>(base + o_supp);
        L->w1 = reinterpret_cast<T*>(base + o_w1);
        L->G0 = reinterpret_cast<T*>(base + o_g0); L->G1 = reinterpret_cast<T*>(base + o_g1);
        L->pos0 = reinterpret_cast<int*>(base + o_p0); L->pos1 = reinterpret_cast<int*>(base + o_p1);
        L->gs = reinterpret_cast<T*>(base + o_gs);
        L->xr = reinterpret_cast<int2*>(base + o_xr);
        L->pin0 = reinterpret_cast<int*>(base + o_i0); L->pin1 = reinterpret_cast<int*>(base + o_i1);
        L->pb = reinterpret_cast<T*>(base + o_pb);
#endif
    }
    return off;
}

// per-row filter supports are staged packed: start | len << 16 (both < 65536 for any field that fits a tile)
__device__ __forceinline__ void stage_supp(unsigned* dst, const int2* __restrict__ src, int n) {
    for (int i = flat_tid(); i < n; i += flat_nt()) { const int2 v = src[i]; dst[i] = (unsigned)v.x | ((unsigned)v.y << 16); }
}
__device__ __forceinline__ int2 unpack_supp(unsigned v) { return make_int2((int)(v & 0xffffu), (int)(v >> 16)); }

__device__ __forceinline__ float fast_abs(float x, float y) {
    const float m2 = x * x + y * y;
    return m2 > 0.f ? m2 * rsqrtf(m2) : 0.f;      // <= 2 ulp; the reference computes sqrt(x^2+y^2)
}
__device__ __forceinline__ double fast_abs(double x, double y) { return sqrt(x * x + y * y); }

// threads per CTA / CTAs per SM the instances are compiled for: a field that fills more than half of
// the SM's shared memory runs alone with up to 640 threads (<= 102 registers); smaller fields share
// the SM three at a time with up to 320 threads each (<= 68 registers).
// forward static instances of the 272-padded sizes run prime-factor transforms (fft_core.cuh)
__host__ __device__ constexpr bool tile_pfa(int n0, int n1) { return n0 > 0 && n0 == n1 && ct_pfa_ok(n0 > 0 ? n0 : 2); }
__host__ __device__ constexpr bool tile_is_big(int n0, int n1) { return n0 == 0 || (long long)n0 * n1 > 10000; }
__host__ __device__ constexpr int tile_max_threads(int n0, int n1) { return tile_is_big(n0, n1) ? 640 : 320; }
__host__ __device__ constexpr int tile_min_blocks(int n0, int n1) { return tile_is_big(n0, n1) ? 1 : 3; }

// Static instances (alias count KT and field size known at compile time, parent = KT*N0 x KT*N1): product + periodise of
// the 4 adjacent columns e..e+3 of output row r.  One base address per operand, every alias at an immediate offset; the
// loads of an alias outside the filter's support are predicated off (zero-filled).
template <typename T, int N0, int N1, int KT, bool PFA>
__device__ __forceinline__ void tile_load_item_s(cx<T>* s, const unsigned* supp, const cx<T>* __restrict__ pb,
                                                 const T* __restrict__ fb, int r, int e, T scale, int lane,
                                                 const int* pin0, const int* pin1) {
    constexpr int P1 = N1 * KT, W = N1 | 1;
    T ax[4], ay[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ax[i] = T(0); ay[i] = T(0); }
    const cx<T>* __restrict__ p0 = pb + (size_t)r * P1 + e;
    const T* __restrict__ f0 = fb + (size_t)r * P1 + e;
#pragma unroll
    for (int c = 0; c < KT; ++c) {
        const int2 sp = unpack_supp(supp[r + c * N0]);
        if (KT > 2 && sp.y == 0) continue;                // whole filter row negligible (common for coarse psi)
        cx2<T> v0[KT], v1[KT];
        re4<T> f[KT];
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            int rel = e + d * N1 - sp.x;
            if (rel < 0) rel += P1;
            const bool in = (rel < sp.y) | ((rel > P1 - 4) & (sp.y > 0));
            const size_t off = (size_t)c * N0 * P1 + (size_t)d * N1;
            v0[d] = ld_pred<cx2<T>>(p0 + off, in);
            v1[d] = ld_pred<cx2<T>>(p0 + off + 2, in);
            f[d] = ld_pred<re4<T>>(f0 + off, in);
        }
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            ax[0] += v0[d].a.x * f[d].a; ay[0] += v0[d].a.y * f[d].a;
            ax[1] += v0[d].b.x * f[d].b; ay[1] += v0[d].b.y * f[d].b;
            ax[2] += v1[d].a.x * f[d].c; ay[2] += v1[d].a.y * f[d].c;
            ax[3] += v1[d].b.x * f[d].d; ay[3] += v1[d].b.y * f[d].d;
        }
    }
    if constexpr (PFA) {
        // prime-factor input order: Fourier bin (r, e + i) goes to position (pin0[r], pin1[e + i]); the four targets of a
        // thread are scattered, which also spreads the banks - no store rotation needed
        cx<T>* row = s + pin0[r] * W;
        const int4 pc = *reinterpret_cast<const int4*>(pin1 + e);
        row[pc.x] = mk<T>(ax[0] * scale, ay[0] * scale); row[pc.y] = mk<T>(ax[1] * scale, ay[1] * scale);
        row[pc.z] = mk<T>(ax[2] * scale, ay[2] * scale); row[pc.w] = mk<T>(ax[3] * scale, ay[3] * scale);
        return;
    }
    cx<T>* dst = s + r * W + e;
    // rotate which of its 4 columns a lane writes in each of the 4 store instructions by (lane>>2)&3:
    // lanes l, l+4, l+8, l+12 (same bank group, columns 16 apart) then hit distinct banks
    const int rot = (lane >> 2) & 3;
    const bool p0r = rot & 1, p1r = rot & 2;
    T bx[4], by[4], cxr[4], cyr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { bx[i] = p0r ? ax[(i + 1) & 3] : ax[i]; by[i] = p0r ? ay[(i + 1) & 3] : ay[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { cxr[i] = p1r ? bx[(i + 2) & 3] : bx[i]; cyr[i] = p1r ? by[(i + 2) & 3] : by[i]; }
#pragma unroll
    for (int t = 0; t < 4; ++t) dst[(t + rot) & 3] = mk<T>(cxr[t] * scale, cyr[t] * scale);
}

// product + periodise for VEC adjacent columns starting at column e of output row r.
//   NATURAL = true : result stored at s[r*W + e + i] (static instances: the inverse runs as DIF)
//   NATURAL = false: result scattered to s[pos0[r]*W + pos1[e+i]] (generic instance: DIT inverse)
template <typename T, int VEC, int KT, bool NATURAL>
__device__ __forceinline__ void tile_load_item(cx<T>* s, const TileSmem<T>& m, const unsigned* supp,
                                               const cx<T>* __restrict__ pb, const T* __restrict__ fb, int r, int e,
                                               int k, int n0, int n1, int W, int P1, T scale, int lane) {
    T ax[VEC], ay[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { ax[i] = T(0); ay[i] = T(0); }
    if constexpr (KT > 0 && VEC == 4) {
        // compile-time alias count: predicated loads, no branches inside a filter row
#pragma unroll
        for (int c = 0; c < KT; ++c) {
            const int R = r + c * n0;
            const int2 sp = unpack_supp(supp[R]);
            if (KT > 2 && sp.y == 0) continue;            // whole filter row negligible (common for coarse psi)
            const size_t rowoff = (size_t)R * P1;
            cx2<T> v0[KT], v1[KT];
            re4<T> f[KT];
#pragma unroll
            for (int d = 0; d < KT; ++d) {
                const int C = e + d * n1;
                int rel = C - sp.x;
                if (rel < 0) rel += P1;
                const bool in = (rel < sp.y) | ((rel > P1 - 4) & (sp.y > 0));
                v0[d] = ld_pred<cx2<T>>(pb + rowoff + C, in);
                v1[d] = ld_pred<cx2<T>>(pb + rowoff + C + 2, in);
                f[d] = ld_pred<re4<T>>(fb + rowoff + C, in);
            }
#pragma unroll
            for (int d = 0; d < KT; ++d) {
                ax[0] += v0[d].a.x * f[d].a; ay[0] += v0[d].a.y * f[d].a;
                ax[1] += v0[d].b.x * f[d].b; ay[1] += v0[d].b.y * f[d].b;
                ax[2] += v1[d].a.x * f[d].c; ay[2] += v1[d].a.y * f[d].c;
                ax[3] += v1[d].b.x * f[d].d; ay[3] += v1[d].b.y * f[d].d;
            }
        }
    } else {
        for (int c = 0; c < k; ++c) {
            const int R = r + c * n0;
            const int2 sp = unpack_supp(supp[R]);
            if (sp.y == 0) continue;
            const size_t rowoff = (size_t)R * P1;
            for (int d = 0; d < k; ++d) {
                const int C = e + d * n1;
                int rel = C - sp.x;
                if (rel < 0) rel += P1;
                if ((rel < sp.y) | (rel > P1 - VEC)) {
                    if constexpr (VEC == 4) {
                        const cx2<T> v0 = *reinterpret_cast<const cx2<T>*>(pb + rowoff + C);
                        const cx2<T> v1 = *reinterpret_cast<const cx2<T>*>(pb + rowoff + C + 2);
                        const re4<T> f = *reinterpret_cast<const re4<T>*>(fb + rowoff + C);
                        ax[0] += v0.a.x * f.a; ay[0] += v0.a.y * f.a;
                        ax[1] += v0.b.x * f.b; ay[1] += v0.b.y * f.b;
                        ax[2] += v1.a.x * f.c; ay[2] += v1.a.y * f.c;
                        ax[3] += v1.b.x * f.d; ay[3] += v1.b.y * f.d;
                    } else {
                        const cx2<T> v = *reinterpret_cast<const cx2<T>*>(pb + rowoff + C);
                        const re2<T> f = *reinterpret_cast<const re2<T>*>(fb + rowoff + C);
                        ax[0] += v.a.x * f.a; ay[0] += v.a.y * f.a;
                        ax[1] += v.b.x * f.b; ay[1] += v.b.y * f.b;
                    }
                }
            }
        }
    }
    if constexpr (NATURAL) {
        cx<T>* dst = s + r * W + e;
        if constexpr (VEC == 4) {
            // rotate which of its 4 columns a lane writes in each of the 4 store instructions by (lane>>2)&3:
            // lanes l, l+4, l+8, l+12 (same bank group, columns 16 apart) then hit distinct banks
            const int rot = (lane >> 2) & 3;
            const bool p0 = rot & 1, p1 = rot & 2;
            T bx[4], by[4], cxr[4], cyr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { bx[i] = p0 ? ax[(i + 1) & 3] : ax[i]; by[i] = p0 ? ay[(i + 1) & 3] : ay[i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i) { cxr[i] = p1 ? bx[(i + 2) & 3] : bx[i]; cyr[i] = p1 ? by[(i + 2) & 3] : by[i]; }
#pragma unroll
            for (int t = 0; t < 4; ++t) dst[(t + rot) & 3] = mk<T>(cxr[t] * scale, cyr[t] * scale);
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) dst[i] = mk<T>(ax[i] * scale, ay[i] * scale);
        }
    } else {
        const int prow = m.pos0[r] * W;
#pragma unroll
        for (int i = 0; i < VEC; ++i) s[prow + m.pos1[e + i]] = mk<T>(ax[i] * scale, ay[i] * scale);
    }
}

// ---------------------------------------------------------------------------------------------------
// Tensor-core low-pass (float only).  Both products of the separable low-pass,
//     W1[q][xo] = sum_p U[q][p] G1s[p][xo]            (n0 x n1) x (n1 x o1)
//     S[yo][xo] = sum_q G0s[q][yo] W1[q][xo]          (o0 x n0) x (n0 x o1)
// are dense when the spatial taps of phi cover the whole circle (the finest resolution).  They run on
// mma.sync.m16n8k8 TF32 with the 3xTF32 split a = a_hi + a_lo, b = b_hi + b_lo,
//     a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi        (fp32 accumulate)
// which keeps fp32-level accuracy (dropped term ~2^-22).  G0s/G1s are G0/G1 with rows permuted into the
// storage (scrambled) order of the field, so no position lookups are needed.
// ---------------------------------------------------------------------------------------------------
// hi = x truncated to the 10 explicit mantissa bits of TF32 (one LOP3), lo = x - hi exactly (one FADD); the tensor core
// reads only the TF32 bits of lo, i.e. truncates it again: |x - (hi + tf32(lo))| < 2^-20 |x|.  (cvt.rna.tf32.f32 is a
// multi-instruction emulation on sm_100a - it was a quarter of the instructions of the 68 x 68 tile kernel.)
__device__ __forceinline__ void tf32_split(float x, unsigned& hi, unsigned& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// C[M x N] += A[M x K] * B[K x N] for one 16 x (8*NT) warp tile.  A(row, col) = Aptr[row*ars + col*acs] (rows
// clamped to M-1), B(k, n) = Bptr[k*brs + n] (the caller pads B with zero rows up to a multiple of 8).
template <int NT>
__device__ __forceinline__ void warp_gemm_3xtf32(float (&c)[NT][4], const float* Aptr, int ars, int acs, int M, int K,
                                                 const float* Bptr, int brs, int m0, int n0, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int r0 = min(m0 + g, M - 1), r1 = min(m0 + g + 8, M - 1);
    for (int k0 = 0; k0 < K; k0 += 8) {
        const int ka = min(k0 + t, K - 1), kb = min(k0 + t + 4, K - 1);     // clamped: B is zero beyond K
        unsigned ah[4], al[4];
        tf32_split(Aptr[r0 * ars + ka * acs], ah[0], al[0]);
        tf32_split(Aptr[r1 * ars + ka * acs], ah[1], al[1]);
        tf32_split(Aptr[r0 * ars + kb * acs], ah[2], al[2]);
        tf32_split(Aptr[r1 * ars + kb * acs], ah[3], al[3]);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            unsigned bh[2], bl[2];
            tf32_split(Bptr[(k0 + t) * brs + n0 + 8 * j + g], bh[0], bl[0]);
            tf32_split(Bptr[(k0 + t + 4) * brs + n0 + 8 * j + g], bh[1], bl[1]);
            mma_tf32(c[j], al, bh);
            mma_tf32(c[j], ah, bl);
            mma_tf32(c[j], ah, bh);
        }
    }
}

// SPEC (static instances): the path has children - the modulus is stored as (|u|, 0) and the forward transform + spectrum
// store are compiled in; leaf instances (SPEC = false) store (|u|, |u|) (packed FFMA2 operand of the low-pass) and contain
// no forward transform.  The generic instance decides at run time (a.spec_out).
template <typename T, int N0, int N1, int KT, bool SPEC>
__device__ __forceinline__ void tile_body(const TileArgs<T>& a) {
    constexpr bool ST = N0 > 0;
    // prime-factor transforms (no twiddles between the 2^a and the 17 pass) for the 272-padded sizes; the host passes the
    // matching position tables (plan2d.cuh: tile())
    constexpr bool PFA = ST && KT > 0 && tile_pfa(N0, N1);
    const int n0 = ST ? N0 : a.n0, n1 = ST ? N1 : a.n1;
    const int W = ST ? (N1 | 1) : a.W;
    const int k = KT > 0 ? KT : a.k;
    const int wp = a.o1p + 4;            // w1 pitch: 16-byte stores from consecutive rows hit distinct banks
    TileSmem<T> m;
    tile_smem_layout(a, &m);
    cx<T>* s = m.tile;
    const int tid = flat_tid(), nt = flat_nt();
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;

    // constants shared by every path this (persistent) CTA processes
    stage(m.tw0, a.tw0, n0); stage(m.tw1, a.tw1, n1);
    stage(m.pos0, a.pos0, n0); stage(m.pos1, a.pos1, n1);
    if constexpr (PFA) { stage(m.pin0, a.pin0, n0); stage(m.pin1, a.pin1, n1); }
    const bool mma = ST && std::is_same<T, float>::value && a.use_mma;
    const int gp0 = a.o0p + 8, gp1 = a.o1p + 8;      // mma-path pitches of G0s / G1s (and of w1)
    if (mma) {
        // rows in storage order, zero rows up to a multiple of 8 (positions are read after the barrier below)
        __syncthreads();
        const int k0p = (n0 + 7) & ~7, k1p = (n1 + 7) & ~7;
        for (int i = tid; i < k0p * gp0; i += nt) m.G0[i] = T(0);
        for (int i = tid; i < k1p * gp1; i += nt) m.G1[i] = T(0);
        for (int i = tid; i < (k0p - n0) * gp1; i += nt) m.w1[n0 * gp1 + i] = T(0);   // K padding of the 4b B operand
        __syncthreads();
        for (int i = tid; i < n0 * a.o0p; i += nt) { const int y = i / a.o0p, o = i - y * a.o0p; m.G0[m.pos0[y] * gp0 + o] = a.G0[i]; }
        for (int i = tid; i < n1 * a.o1p; i += nt) { const int x = i / a.o1p, o = i - x * a.o1p; m.G1[m.pos1[x] * gp1 + o] = a.G1[i]; }
    } else if (ST && a.tt) {
        stage(reinterpret_cast<re4<T>*>(m.G0), reinterpret_cast<const re4<T>*>(a.TT0), a.y0cnt);
        stage(reinterpret_cast<re4<T>*>(m.G1), reinterpret_cast<const re4<T>*>(a.TT1), a.x1cnt);
    } else {
        stage(reinterpret_cast<re4<T>*>(m.G0), reinterpret_cast<const re4<T>*>(a.G0), n0 * a.o0p / 4);
        stage(reinterpret_cast<re4<T>*>(m.G1), reinterpret_cast<const re4<T>*>(a.G1), n1 * a.o1p / 4);
    }

    SB_PHASE_INIT((N0 == 136 ? 0 : N0 == 68 ? 1 : 2) * 8 + (KT == 4 ? 4 : 0) + (a.spec_out ? 2 : 0) + (a.PP != a.NFch ? 1 : 0))
    // peer destinations (multi-GPU gather fused into this kernel): the low-pass tail stores the plane locally; after the
    // barrier that ends the path, warp 0 pushes it to the peers with 16-byte (multicast) stores (kernels2d.cuh: push_plane)
    OutPeers<T> local_only; local_only.n = 0;
    if ((int)blockIdx.x < a.G) stage_supp(m.supp, a.supp + (size_t)(blockIdx.x % a.NF) * a.P0, a.P0);
    if (a.stagger_ns > 0) {
        for (int i = (int)(blockIdx.x & 3); i > 0; --i) __nanosleep((unsigned)a.stagger_ns);
    }
    int sbuf = 0;
    for (int g = blockIdx.x; g < a.G; g += gridDim.x, sbuf ^= 1) {
        const int fi = g % a.NF, pg = g / a.NF;
        const int b = g / a.PP, path = g - b * a.PP;
        const int ch = a.ch0 + (path / a.NFch) * a.chs + (path % a.NFch);
        const unsigned* supp = m.supp + sbuf * a.P0;
        const int gn = g + gridDim.x;
        // the next path of this CTA: pull its slice of the parent spectrum into L2 now (the NF CTAs that share a
        // parent cover it together), and stage its support rows into the other buffer
        if (gn < a.G) {
            if (a.prefetch) prefetch_rows_slice(a.parent + (size_t)(gn / a.NF) * a.P0 * a.P1, a.P0, a.P1, gn % a.NF, a.NF, tid, nt);
            stage_supp(m.supp + (sbuf ^ 1) * a.P0, a.supp + (size_t)(gn % a.NF) * a.P0, a.P0);
        }
        __syncthreads();
        SB_PHASE(0);

        // 1. product + periodise: 4 (or 2) adjacent columns per thread, 128-bit loads, aliases outside the
        //    filter support skipped
        {
            const cx<T>* __restrict__ pb = a.parent + (size_t)pg * a.P0 * a.P1;
            const T* __restrict__ fb = a.filt[fi];
            const int P1 = a.P1;
            if constexpr (ST && KT > 0 && (N1 & 3) == 0) {
                constexpr int per_row = N1 >> 2, items = N0 * per_row;
                for (int it = tid; it < items; it += nt) {
                    const int r0 = it / per_row, e0 = 4 * (it - r0 * per_row);
                    tile_load_item_s<T, N0, N1, KT, PFA>(s, supp, pb, fb, r0, e0, a.scale, lane, m.pin0, m.pin1);
                }
            } else if ((n1 & 3) == 0 && (P1 & 3) == 0) {
                const int per_row = n1 >> 2, items = n0 * per_row;
                for (int it = tid; it < items; it += nt) {
                    const int r0 = it / per_row, e0 = 4 * (it - r0 * per_row);
                    tile_load_item<T, 4, KT, ST>(s, m, supp, pb, fb, r0, e0, k, n0, n1, W, P1, a.scale, lane);
                }
            } else {
                const int per_row = n1 >> 1, items = n0 * per_row;
                for (int it = tid; it < items; it += nt) {
                    const int r0 = it / per_row, e0 = 2 * (it - r0 * per_row);
                    tile_load_item<T, 2, KT, ST>(s, m, supp, pb, fb, r0, e0, k, n0, n1, W, P1, a.scale, lane);
                }
            }
        }
        __syncthreads();
        SB_PHASE(1);
        // 2+3. inverse 2-D FFT and modulus.
        //   static : DIF (natural Fourier in -> scrambled spatial out), modulus in the registers of the last pass;
        //            the spatial field stays scrambled: U[y][x] lives at s[pos0[y]*W + pos1[x]]
        //   generic: DIT (scattered input -> natural spatial), separate modulus sweep
        if constexpr (ST) {
            slab_fft_s<N1, false, +1, (N1 | 1), 1, T, false, PFA>(s, N0, m.tw1);
            SB_PHASE(2);
            slab_fft_s<N0, false, +1, 1, (N1 | 1), T, (SPEC ? 1 : 2), PFA>(s, N1, m.tw0);
            SB_PHASE(3);
        } else {
            slab_fft<true, T>(s, n0, W, 1, a.plan1, m.tw1);
            slab_fft<true, T>(s, n1, 1, W, a.plan0, m.tw0);
            for (int y = warp; y < n0; y += nwarps)
                for (int x = lane; x < n1; x += 32) s[y * W + x] = mk<T>(cabs_fast<T>(s[y * W + x]), T(0));
            __syncthreads();
        }
        if constexpr (ST && std::is_same<T, float>::value) {
          if (mma) {
            const int nwarp = nt >> 5, wid = tid >> 5;
            // 4a on tensor cores: W1 = U * G1s; warp task = (16-row tile, 16-column half)
            {
                const int mt = (n0 + 15) >> 4, nh = (a.o1p + 15) >> 4;
                for (int task = wid; task < mt * nh; task += nwarp) {
                    const int m0 = (task / nh) * 16, c0 = (task % nh) * 16;
                    float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                    warp_gemm_3xtf32<2>(c, reinterpret_cast<const float*>(s), 2 * W, 2, n0, n1,
                                        reinterpret_cast<const float*>(m.G1), gp1, m0, c0, lane);
                    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int col = c0 + 8 * j + 2 * tq;
                        if (m0 + gq < n0) { m.w1[(m0 + gq) * gp1 + col] = c[j][0]; m.w1[(m0 + gq) * gp1 + col + 1] = c[j][1]; }
                        if (m0 + gq + 8 < n0) { m.w1[(m0 + gq + 8) * gp1 + col] = c[j][2]; m.w1[(m0 + gq + 8) * gp1 + col + 1] = c[j][3]; }
                    }
                }
                // zero the K-padding rows of w1 read (clamped) by 4b - B operand must be finite; rows >= n0 unused
            }
            __syncthreads();
            // 4b on tensor cores: S = G0s^T * W1; warp task = (16 output rows, 8 output columns)
            {
                const OutRef<T> ob = out_ref(a.out, local_only, ((size_t)b * a.K + ch) * a.o0 * a.o1);
                const int mt = (a.o0p + 15) >> 4, ntl = (a.o1p + 7) >> 3;
                for (int task = wid; task < mt * ntl; task += nwarp) {
                    const int m0 = (task / ntl) * 16, c0 = (task % ntl) * 8;
                    float c[1][4] = {{0.f, 0.f, 0.f, 0.f}};
                    // A(row = yo, col = q) = G0s[q][yo]; rows beyond o0p are clamped (their outputs are discarded)
                    warp_gemm_3xtf32<1>(c, reinterpret_cast<const float*>(m.G0), 1, gp0, a.o0p, n0,
                                        reinterpret_cast<const float*>(m.w1), gp1, m0, c0, lane);
                    const int gq = lane >> 2, tq = lane & 3;
                    const int col = c0 + 2 * tq;
                    if (m0 + gq < a.o0) {
                        if (col < a.o1) ob[(m0 + gq) * a.o1 + col] = c[0][0];
                        if (col + 1 < a.o1) ob[(m0 + gq) * a.o1 + col + 1] = c[0][1];
                    }
                    if (m0 + gq + 8 < a.o0) {
                        if (col < a.o1) ob[(m0 + gq + 8) * a.o1 + col] = c[0][2];
                        if (col + 1 < a.o1) ob[(m0 + gq + 8) * a.o1 + col + 1] = c[0][3];
                    }
                }
            }
          }
        }
        bool lowpass_done = mma;
        if constexpr (ST) {
          if (!mma && a.tt) {
            lowpass_done = true;
            // 4a (tap tables). w1[h][row][xo] = sum over half h of the x window of U[row][x] * g1[kl (xo+1) - x]:
            //   4 storage rows x 4 outputs per work item, the window split in two so that every thread has an item;
            //   leaf instances fetch (|u|, |u|) with one 64-bit load and issue two FFMA2 per row and tap
            {
                constexpr int rgroups = (N0 + 3) >> 2;
                const int xgroups = a.o1p >> 2, half_items = rgroups * xgroups;
                const int h0 = (a.x1cnt + 1) >> 1;
                for (int it = tid; it < 2 * half_items; it += nt) {
                    const int hf = it >= half_items ? 1 : 0;
                    const int rem = it - hf * half_items;
                    const int xg = rem / rgroups, rg = rem - xg * rgroups;
                    int yy[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) yy[j] = min(rg + j * rgroups, N0 - 1) * W;
                    const int st0 = hf ? h0 : 0, st1 = hf ? a.x1cnt : h0;
                    int x = (a.kl * (4 * xg + 1) + a.x1lo + st0) % N1;
                    if (x < 0) x += N1;
                    cx<T> acc01[4], acc23[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { acc01[j] = mk<T>(T(0), T(0)); acc23[j] = mk<T>(T(0), T(0)); }
#pragma unroll 4
                    for (int st = st0; st < st1; ++st) {
                        const re4<T> gq = *reinterpret_cast<const re4<T>*>(m.G1 + 4 * st);
                        const cx<T> g01 = mk<T>(gq.a, gq.b), g23 = mk<T>(gq.c, gq.d);
                        const int xs = m.pos1[x];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            cx<T> uu = s[yy[j] + xs];
                            if constexpr (SPEC) uu.y = uu.x;
                            acc01[j] = fma_cc(acc01[j], uu, g01);
                            acc23[j] = fma_cc(acc23[j], uu, g23);
                        }
                        x = (x + 1 == N1) ? 0 : x + 1;
                    }
                    T* w1h = m.w1 + hf * N0 * wp;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int y = rg + j * rgroups;
                        if (y < N0) {
                            re4<T> o; o.a = acc01[j].x; o.b = acc01[j].y; o.c = acc23[j].x; o.d = acc23[j].y;
                            *reinterpret_cast<re4<T>*>(w1h + y * wp + 4 * xg) = o;
                        }
                    }
                }
            }
            __syncthreads();
            SB_PHASE(4);
            // 4b (tap tables). S[yo][xo] = sum_y g0[kl (yo+1) - y] * (w1[0] + w1[1])[row(y)][xo]; the y window is split in two
            //   work items as well: the second half leaves its partial sums in shared memory, the first adds them and stores
            {
                const OutRef<T> ob = out_ref(a.out, local_only, ((size_t)b * a.K + ch) * a.o0 * a.o1);
                const int ygroups = a.o0p >> 2, nitem = ygroups * a.o1p;
                const T* w1b = m.w1 + N0 * wp;
                const int g0 = (a.y0cnt + 1) >> 1;
                T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
                const bool two = 2 * nitem <= nt;                 // enough threads for both halves at once
                const int hf = (two && tid >= nitem) ? 1 : 0;
                const int it = tid - hf * nitem;
                const bool active = it < nitem;
                const int yg = active ? it / a.o1p : 0, xo = active ? it - yg * a.o1p : 0;
                if (active) {
                    const int st0 = two ? (hf ? g0 : 0) : 0, st1 = two ? (hf ? a.y0cnt : g0) : a.y0cnt;
                    int y = (a.kl * (4 * yg + 1) + a.y0lo + st0) % N0;
                    if (y < 0) y += N0;
#pragma unroll 4
                    for (int st = st0; st < st1; ++st) {
                        const re4<T> gq = *reinterpret_cast<const re4<T>*>(m.G0 + 4 * st);
                        const int ys = m.pos0[y] * wp + xo;
                        const T w = m.w1[ys] + w1b[ys];
                        acc0 += w * gq.a; acc1 += w * gq.b; acc2 += w * gq.c; acc3 += w * gq.d;
                        y = (y + 1 == N0) ? 0 : y + 1;
                    }
                    if (hf) {
                        re4<T> o; o.a = acc0; o.b = acc1; o.c = acc2; o.d = acc3;
                        *reinterpret_cast<re4<T>*>(m.pb + 4 * it) = o;
                    }
                }
                if (two) __syncthreads();
                if (active && !hf) {
                    if (two) {
                        const re4<T> o = *reinterpret_cast<const re4<T>*>(m.pb + 4 * it);
                        acc0 += o.a; acc1 += o.b; acc2 += o.c; acc3 += o.d;
                    }
                    if (xo < a.o1) {
                        const int yo = 4 * yg;
                        if (yo + 0 < a.o0) ob[(yo + 0) * a.o1 + xo] = acc0;
                        if (yo + 1 < a.o0) ob[(yo + 1) * a.o1 + xo] = acc1;
                        if (yo + 2 < a.o0) ob[(yo + 2) * a.o1 + xo] = acc2;
                        if (yo + 3 < a.o0) ob[(yo + 3) * a.o1 + xo] = acc3;
                    }
                }
                // (requires nitem <= nt: checked by the host, which otherwise leaves tt off)
            }
          }
        }
        if (!lowpass_done) {
        // 4a. horizontal low-pass + decimation + unpad: w1[row][xo] = sum_x U[row][x] * G1[x][xo]
        //     register tile: 4 (storage) rows x 4 outputs per thread, x restricted to the group's input window
        {
            const int rgroups = (n0 + 3) >> 2, xgroups = a.o1p >> 2;
            for (int it = tid; it < rgroups * xgroups; it += nt) {
                const int xg = it / rgroups, rg = it - xg * rgroups;
                int yy[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) yy[j] = min(rg + j * rgroups, n0 - 1) * W;
                T acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
                int x = (a.kl * (4 * xg + 1) + a.x1lo) % n1;
                if (x < 0) x += n1;
                for (int st = 0; st < a.x1cnt; ++st) {
                    const re4<T> gq = *reinterpret_cast<const re4<T>*>(m.G1 + x * a.o1p + 4 * xg);
                    const int xs = ST ? m.pos1[x] : x;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const T u = s[yy[j] + xs].x;
                        acc[j][0] += u * gq.a; acc[j][1] += u * gq.b; acc[j][2] += u * gq.c; acc[j][3] += u * gq.d;
                    }
                    x = (x + 1 == n1) ? 0 : x + 1;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int y = rg + j * rgroups;
                    if (y < n0) {
                        re4<T> o; o.a = acc[j][0]; o.b = acc[j][1]; o.c = acc[j][2]; o.d = acc[j][3];
                        *reinterpret_cast<re4<T>*>(m.w1 + y * wp + 4 * xg) = o;
                    }
                }
            }
        }
        __syncthreads();
        SB_PHASE(4);
        // 4b. vertical low-pass + decimation + unpad, straight to the output plane:
        //     S[yo][xo] = sum_y G0[y][yo] * w1[row(y)][xo]; 4 output rows per thread, lanes along xo
        {
            const OutRef<T> ob = out_ref(a.out, local_only, ((size_t)b * a.K + ch) * a.o0 * a.o1);
            const int ygroups = a.o0p >> 2;
            for (int it = tid; it < ygroups * a.o1p; it += nt) {
                const int yg = it / a.o1p, xo = it - yg * a.o1p;
                T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
                int y = (a.kl * (4 * yg + 1) + a.y0lo) % n0;
                if (y < 0) y += n0;
                for (int st = 0; st < a.y0cnt; ++st) {
                    const re4<T> gq = *reinterpret_cast<const re4<T>*>(m.G0 + y * a.o0p + 4 * yg);
                    const int ys = ST ? m.pos0[y] : y;
                    const T w = m.w1[ys * wp + xo];
                    acc0 += w * gq.a; acc1 += w * gq.b; acc2 += w * gq.c; acc3 += w * gq.d;
                    y = (y + 1 == n0) ? 0 : y + 1;
                }
                if (xo < a.o1) {
                    const int yo = 4 * yg;
                    if (yo + 0 < a.o0) ob[(yo + 0) * a.o1 + xo] = acc0;
                    if (yo + 1 < a.o0) ob[(yo + 1) * a.o1 + xo] = acc1;
                    if (yo + 2 < a.o0) ob[(yo + 2) * a.o1 + xo] = acc2;
                    if (yo + 3 < a.o0) ob[(yo + 3) * a.o1 + xo] = acc3;
                }
            }
        }
        }   // dense CUDA-core low-pass
        SB_PHASE(5);
        // 5. forward 2-D FFT of U for the children of this path, natural-order store
        //    (static: DIT, scrambled spatial in -> natural Fourier out; generic: DIF + gather)
        if ((!ST || SPEC) && a.spec_out) {
            cx<T>* ob = a.spec_out + (size_t)g * n0 * n1;
            if constexpr (ST) {
                slab_fft_s<N0, true, -1, 1, (N1 | 1), T, false, PFA>(s, N1, m.tw0);
                slab_fft_s<N1, true, -1, (N1 | 1), 1, T, false, PFA>(s, N0, m.tw1);
                constexpr int half = N1 / 2;
                for (int it = tid; it < N0 * half; it += nt) {
                    const int r = it / half, e = 2 * (it - r * half);
                    cx2<T> v;
                    if constexpr (PFA) {      // natural bin (r, e) sits at the prime-factor input position
                        const cx<T>* row = s + m.pin0[r] * W;
                        v.a = row[m.pin1[e]]; v.b = row[m.pin1[e + 1]];
                    } else {
                        v.a = s[r * W + e]; v.b = s[r * W + e + 1];
                    }
                    *reinterpret_cast<cx2<T>*>(ob + (size_t)r * N1 + e) = v;
                }
            } else {
                slab_fft<false, T>(s, n0, W, 1, a.plan1, m.tw1);
                slab_fft<false, T>(s, n1, 1, W, a.plan0, m.tw0);
                for (int r = warp; r < n0; r += nwarps) {
                    const int prow = m.pos0[r] * W;
                    for (int e = lane; e < n1; e += 32) ob[(size_t)r * n1 + e] = s[prow + m.pos1[e]];
                }
            }
        }
        __syncthreads();   // the next path rewrites the tile, the support rows and w1
        if (a.peers.n != 0 && warp == 0) push_plane(a.out, a.peers, ((size_t)b * a.K + ch) * a.o0 * a.o1, a.o0 * a.o1, lane);
        SB_PHASE(6);
    }
}

template <typename T, int N0, int N1, int KT, bool SPEC = false>
__global__ void __launch_bounds__(tile_max_threads(N0, N1), tile_min_blocks(N0, N1)) k2d_tile(TileArgs<T> a) {
    tile_body<T, N0, N1, KT, SPEC>(a);
}

// ---------------------------------------------------------------------------------------------------
// Backward of one leaf path, same tile, same persistent structure:
//   u      = ifft2(periodise_k(parent * filt) * scale)            recomputed, no modulus
//   gA     = G0 . gS . G1^T                                        adjoint of low-pass + decimation + unpad
//   gu     = gA * u / |u|   (0 where |u| = 0)                      ModulusStable.backward, backend/torch_backend.py:85-96
//   gV     = F(gu)                                                 adjoint of the unnormalised inverse transform
//   gparent[r + c*n0][e + d*n1] += scale * filt[..] * gV[r][e]     adjoint of periodisation and filter multiply
// The last step uses atomic adds: the children of one parent run in different CTAs.
// ---------------------------------------------------------------------------------------------------
template <typename T, int N0, int N1, int KT>
__device__ __forceinline__ void tile_bwd_body(const TileArgs<T>& a) {
    constexpr bool ST = N0 > 0;
    const int n0 = ST ? N0 : a.n0, n1 = ST ? N1 : a.n1;
    const int W = ST ? (N1 | 1) : a.W;
    const int k = KT > 0 ? KT : a.k;
    const int wp = a.o1p + 4;
    TileSmem<T> m;
    tile_smem_layout(a, &m);
    cx<T>* s = m.tile;
    const int tid = flat_tid(), nt = flat_nt();
    const int lane = tid & 31;

    stage(m.tw0, a.tw0, n0); stage(m.tw1, a.tw1, n1);
    stage(m.pos0, a.pos0, n0); stage(m.pos1, a.pos1, n1);
    stage(reinterpret_cast<re4<T>*>(m.G0), reinterpret_cast<const re4<T>*>(a.G0), n0 * a.o0p / 4);
    __syncthreads();
    // G1 is held TRANSPOSED and in the storage order of the tile's columns, G1s[xo][xs] = G1[x][xo] with xs = pos1[x]:
    // step 3b walks the tile in storage order, so its three shared-memory streams (tile, G1s, xr) are all unit-stride
    // across a warp (the natural-order [x][o1p] layout put a whole warp on one or two banks).
    // The low-pass is banded: column x touches only outputs [first, last] (a few of them, all for the full-circle
    // level); found once per persistent CTA
    const int g1p = (n1 + 31) & ~31;           // pitch = 0 mod 32 banks: lanes with different xo never collide
    int* s_maxband = reinterpret_cast<int*>(m.gs);      // scratch word (gs is staged per path, later); no static shared
    if (tid == 0) *s_maxband = 0;                       // memory: the kernels opt in to the full dynamic carve-out
    __syncthreads();
    for (int x = tid; x < n1; x += nt) {
        const int xs = ST ? m.pos1[x] : x;
        const T* __restrict__ gr = a.G1 + (size_t)x * a.o1p;
        int first = a.o1, last = -1;
        for (int xo = 0; xo < a.o1p; ++xo) {
            const T v = xo < a.o1 ? gr[xo] : T(0);
            m.G1[xo * g1p + xs] = v;
            if (v != T(0)) { if (first == a.o1) first = xo; last = xo; }
        }
        m.xr[xs] = make_int2(first, last);
        if (last >= first) atomicMax(s_maxband, last - first + 1);
    }
    __syncthreads();
    const int maxband = *s_maxband;
    __syncthreads();
    // step 3b work split: a thread owns ONE column xs and a group of rows, so the column's taps stay in registers.
    // RG row groups: minimise (rounds over the threads) x (rows per group)
    int RG = 1;
    {
        int best = 1 << 30;
        for (int rg = 1; rg <= 16 && rg <= n0; ++rg) {
            const int cost = ((n1 * rg + nt - 1) / nt) * ((n0 + rg - 1) / rg);
            if (cost < best) { best = cost; RG = rg; }
        }
    }
    const int RP = (n0 + RG - 1) / RG;

    SB_PHASE_INIT(24 + (n0 >= 128 ? 0 : n0 >= 64 ? 1 : n0 >= 32 ? 2 : 3) * 2 + (k > 2 ? 1 : 0))
    for (int g = blockIdx.x; g < a.G; g += gridDim.x) {
        const int fi = g % a.NF, pg = g / a.NF;
        const int b = g / a.PP, path = g - b * a.PP;
        const int ch = a.ch0 + (path / a.NFch) * a.chs + (path % a.NFch);
        stage_supp(m.supp, a.supp + (size_t)fi * a.P0, a.P0);
        {
            const T* gb = a.gout + ((size_t)b * a.K + ch) * a.o0 * a.o1;
            for (int i = tid; i < a.o0p * a.o1p; i += nt) {
                const int yo = i / a.o1p, xo = i - yo * a.o1p;
                m.gs[i] = (yo < a.o0 && xo < a.o1) ? gb[yo * a.o1 + xo] : T(0);
            }
        }
        __syncthreads();
        SB_PHASE(0);
        const cx<T>* __restrict__ pb = a.parent + (size_t)pg * a.P0 * a.P1;
        const T* __restrict__ fb = a.filt[fi];
        const int P1 = a.P1;
        // 1. recompute the product + periodise
        if constexpr (ST && KT > 0 && (N1 & 3) == 0) {
            constexpr int per_row = N1 >> 2, items = N0 * per_row;
            for (int it = tid; it < items; it += nt) {
                const int r0 = it / per_row, e0 = 4 * (it - r0 * per_row);
                tile_load_item_s<T, N0, N1, KT, false>(s, m.supp, pb, fb, r0, e0, a.scale, lane, nullptr, nullptr);
            }
        } else if ((n1 & 3) == 0 && (P1 & 3) == 0) {
            const int per_row = n1 >> 2, items = n0 * per_row;
            for (int it = tid; it < items; it += nt) {
                const int r0 = it / per_row, e0 = 4 * (it - r0 * per_row);
                tile_load_item<T, 4, KT, ST>(s, m, m.supp, pb, fb, r0, e0, k, n0, n1, W, P1, a.scale, lane);
            }
        } else {
            const int per_row = n1 >> 1, items = n0 * per_row;
            for (int it = tid; it < items; it += nt) {
                const int r0 = it / per_row, e0 = 2 * (it - r0 * per_row);
                tile_load_item<T, 2, KT, ST>(s, m, m.supp, pb, fb, r0, e0, k, n0, n1, W, P1, a.scale, lane);
            }
        }
        SB_PHASE(1);
        // 3a. T[row][xo] = sum_yo G0[y][yo] * gS[yo][xo]   (row = storage row of y) - independent of the tile
        for (int it = tid; it < n0 * a.o1p; it += nt) {
            const int y = it / a.o1p, xo = it - y * a.o1p;
            T acc = T(0);
            for (int yo = 0; yo < a.o0; ++yo) acc += m.G0[y * a.o0p + yo] * m.gs[yo * a.o1p + xo];
            m.w1[(ST ? m.pos0[y] : y) * wp + xo] = acc;
        }
        __syncthreads();
        SB_PHASE(2);
        // 2. inverse 2-D FFT (no modulus): static -> scrambled spatial order, generic -> natural
        if constexpr (ST) {
            slab_fft_s<N1, false, +1, (N1 | 1), 1, T>(s, N0, m.tw1);
            slab_fft_s<N0, false, +1, 1, (N1 | 1), T>(s, N1, m.tw0);
        } else {
            slab_fft<true, T>(s, n0, W, 1, a.plan1, m.tw1);
            slab_fft<true, T>(s, n1, 1, W, a.plan0, m.tw0);
        }
        SB_PHASE(3);
        // 3b. gu = (sum_xo T[row][xo] G1[x][xo]) * u / |u|
        auto modulus_bwd = [&](int q, int xs, T gA) {
            if (a.radd) gA += a.radd[(size_t)g * n0 * n1 + q * n1 + xs];
            const int idx = q * W + xs;
            const cx<T> v = s[idx];
            const T m2 = v.x * v.x + v.y * v.y;
            T sc;
            if constexpr (std::is_same<T, float>::value) sc = m2 > 0.f ? gA * rsqrtf(m2) : 0.f;
            else sc = m2 > T(0) ? gA / sqrt(m2) : T(0);
            s[idx] = mk<T>(v.x * sc, v.y * sc);
        };
        // banded low-pass (every level but the full-circle one): MT taps of the thread's column in registers, static inner
        // loop; the tap window is clamped into [0, o1p - MT] (G1 is zero outside the band, so the extra taps are zeros)
        auto cols = [&](auto mt_) {
            constexpr int MT = decltype(mt_)::value;
            for (int it = tid; it < n1 * RG; it += nt) {
                const int rg = it / n1, xs = it - rg * n1;
                const int lo = min(max(m.xr[xs].x, 0), a.o1p - MT);
                T tap[MT];
#pragma unroll
                for (int j = 0; j < MT; ++j) tap[j] = m.G1[(lo + j) * g1p + xs];
                const int q1 = min(n0, (rg + 1) * RP);
#pragma unroll 2
                for (int q = rg * RP; q < q1; ++q) {
                    const T* __restrict__ tr = m.w1 + q * wp + lo;
                    T gA = T(0);
#pragma unroll
                    for (int j = 0; j < MT; ++j) gA += tr[j] * tap[j];
                    modulus_bwd(q, xs, gA);
                }
            }
        };
        if (maxband <= 8 && a.o1p >= 8) cols(std::integral_constant<int, 8>{});
        else if (maxband <= 16 && a.o1p >= 16) cols(std::integral_constant<int, 16>{});
        else {
            for (int it = tid; it < n0 * n1; it += nt) {
                const int q = it / n1, xs = it - q * n1;
                const T* __restrict__ tr = m.w1 + q * wp;
                const T* __restrict__ gc = m.G1 + xs;
                T gA = T(0);
                const int2 rng = m.xr[xs];
                for (int xo = rng.x; xo <= rng.y; ++xo) gA += tr[xo] * gc[xo * g1p];
                modulus_bwd(q, xs, gA);
            }
        }
        __syncthreads();
        SB_PHASE(4);
        // 4. forward 2-D FFT: static DIT(-) -> natural Fourier order; generic DIF -> scrambled (read through pos)
        if constexpr (ST) {
            slab_fft_s<N0, true, -1, 1, (N1 | 1), T>(s, N1, m.tw0);
            slab_fft_s<N1, true, -1, (N1 | 1), 1, T>(s, N0, m.tw1);
        } else {
            slab_fft<false, T>(s, n0, W, 1, a.plan1, m.tw1);
            slab_fft<false, T>(s, n1, 1, W, a.plan0, m.tw0);
        }
        SB_PHASE(5);
        // 5. adjoint of periodise + filter multiply, accumulated into the parent gradient.  float, even sizes: TWO adjacent
        //    bins per thread and one 16-byte vector reduction (red.global.add.v4.f32, sm_90+) instead of four scalar
        //    atomics - the scatter of the children of one parent is bound by the number of L2 reduction operations
        {
            cx<T>* gp = a.gparent + (size_t)pg * a.P0 * a.P1;
            bool done = false;
            if constexpr (std::is_same<T, float>::value) {
                if ((n1 & 1) == 0 && (P1 & 1) == 0) {
                    done = true;
                    const int half = n1 >> 1;
                    for (int it = tid; it < n0 * half; it += nt) {
                        const int r = it / half, e = 2 * (it - r * half);
                        const cx<T> g0 = scal(ST ? s[r * W + e] : s[m.pos0[r] * W + m.pos1[e]], a.scale);
                        const cx<T> g1 = scal(ST ? s[r * W + e + 1] : s[m.pos0[r] * W + m.pos1[e + 1]], a.scale);
                        for (int c = 0; c < k; ++c) {
                            const int R = r + c * n0;
                            const int2 sp = unpack_supp(m.supp[R]);
                            if (sp.y == 0) continue;
                            for (int d = 0; d < k; ++d) {
                                const int C = e + d * n1;
                                int rel = C - sp.x;
                                if (rel < 0) rel += P1;
                                // either bin inside the support (the filter is negligible, not zero, just outside it)
                                if ((rel < sp.y) | (rel == P1 - 1)) {
                                    const float2 f = *reinterpret_cast<const float2*>(fb + (size_t)R * P1 + C);
                                    float4 v; v.x = g0.x * f.x; v.y = g0.y * f.x; v.z = g1.x * f.y; v.w = g1.y * f.y;
                                    atomicAdd(reinterpret_cast<float4*>(gp + (size_t)R * P1 + C), v);
                                }
                            }
                        }
                    }
                }
            }
            if (!done) {
                for (int it = tid; it < n0 * n1; it += nt) {
                    const int r = it / n1, e = it - r * n1;
                    const cx<T> gv = scal(ST ? s[r * W + e] : s[m.pos0[r] * W + m.pos1[e]], a.scale);
                    for (int c = 0; c < k; ++c) {
                        const int R = r + c * n0;
                        const int2 sp = unpack_supp(m.supp[R]);
                        if (sp.y == 0) continue;
                        for (int d = 0; d < k; ++d) {
                            const int C = e + d * n1;
                            int rel = C - sp.x;
                            if (rel < 0) rel += P1;
                            if (rel < sp.y) {
                                const T f = fb[(size_t)R * P1 + C];
                                cx<T>* dst = gp + (size_t)R * P1 + C;
                                atomicAdd(&dst->x, gv.x * f);
                                atomicAdd(&dst->y, gv.y * f);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        SB_PHASE(6);
    }
}

template <typename T, int N0, int N1, int KT>
__global__ void __launch_bounds__(tile_max_threads(N0, N1), tile_min_blocks(N0, N1)) k2d_tile_bwd(TileArgs<T> a) {
    tile_bwd_body<T, N0, N1, KT>(a);
}

// specialised instances are compiled in tile_inst_*.cu; returns the kernel for (n0, n1, k) or the
// generic one
template <typename T> using TileKernel = void (*)(TileArgs<T>);
template <typename T> TileKernel<T> tile_kernel_lookup(int n0, int n1, int k, bool* is_static);
template <typename T> TileKernel<T> tile_spec_kernel_lookup(int n0, int n1, int k);     // static instances with children
template <typename T> void tile_spec_kernels_enable_smem();
template <typename T> TileKernel<T> tile_bwd_kernel_lookup(int n0, int n1, int k, bool* is_static);
template <typename T> void tile_kernels_enable_smem();
template <typename T> void tile_bwd_kernels_enable_smem();
void phase_prof_read_bwd(unsigned long long* out, bool reset);
int phase_prof_read(unsigned long long* out, int max_n, bool reset);

}  // namespace sb
