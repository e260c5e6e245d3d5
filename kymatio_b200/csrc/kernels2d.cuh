// kernels2d.cuh - fused 2-D scattering kernels.
//
// Reference semantics being fused (paths relative to the kymatio tree):
//   pad            kymatio/scattering2d/backend/torch_backend.py:36-86
//   rfft/ifft      kymatio/scattering2d/backend/torch_backend.py:134-155
//   cdgmm          kymatio/backend/torch_backend.py:148-219  (real filter branch :205-206)
//   subsample_f.   kymatio/scattering2d/backend/torch_backend.py:93-129
//   modulus        kymatio/backend/torch_backend.py:138-141
//   irfft + unpad  kymatio/scattering2d/backend/torch_backend.py:144-176
//   cascade        kymatio/scattering2d/core/scattering2d.py:14-86
//
// Conventions
//   * global memory: spatial AND Fourier data in natural order (filters are read in place
//     from the frontend's buffers);
//   * shared memory: spatial data natural, Fourier data in the plan's scrambled order
//     (fft_core.cuh).  Inverse transforms are DIT (scatter through pos[] while staging in),
//     forward transforms are DIF (gather through pos[] while staging out).
//   * filter sparsity: every filter row R carries a circular column interval
//     (start, len) outside which the filter is negligible (|f| <= 1e-7 max|f|); the
//     product/periodise prologues skip loads outside it.
//
// Two execution shapes:
//   tile kernel       one CTA holds a whole n0 x n1 field: product+periodise, 2-D inverse FFT,
//                     modulus, separable spatial low-pass + decimation + unpad, and (for
//                     parents of second-order paths) the forward 2-D FFT - one pass over HBM/L2;
//   streaming passes  row-slab and column-slab kernels for fields that do not fit one CTA
//                     (the full-resolution first-order band) and the generic Fourier
//                     low-pass (any, also non-separable, phi).
#pragma once
#include "slab.cuh"

namespace sb {

template <typename T> __device__ __forceinline__ T* dyn_smem() {
    extern __shared__ __align__(16) unsigned char sb_dyn_smem[];
    return reinterpret_cast<T*>(sb_dyn_smem);
}

// Coefficient planes may be written to SEVERAL buffers at once: the caller's own output and the same location of the
// output tensors of the peer GPUs (symmetric memory mapped over NVLink, parallel.py: PeerGatherScattering).  The final
// low-pass stores are a few KB per path, so replicating them costs nothing and the all-gather of the batch-sharded
// result needs no extra pass - the transfer rides on the kernels that produce the coefficients.
constexpr int kMaxPeers = 7;
// n >= 0: p[0..n) are the peers' buffers (plain remote stores, one per peer, beside the local store);
// n == -1: p[0] is the MULTICAST address of the block (NVLS): one multimem.st per value is replicated by the NVSwitch
//          into every rank's buffer, the caller's own included - 1/8 of the NVLink egress of the unicast form at 8 GPUs.
template <typename T> struct OutPeers { T* p[kMaxPeers]; int n; };
__device__ __forceinline__ void multimem_st(float* mc, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}
__device__ __forceinline__ void multimem_st(double* mc, double v) {
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc), "d"(v) : "memory");
}
template <typename T> struct OutRef {
    T* out; const OutPeers<T>* pe; size_t off;
    struct Slot {
        const OutRef& r; size_t i;
        __device__ __forceinline__ void operator=(T v) const {
            if (r.pe->n < 0) { multimem_st(r.pe->p[0] + r.off + i, v); return; }
            r.out[r.off + i] = v;
            for (int k = 0; k < r.pe->n; ++k) r.pe->p[k][r.off + i] = v;
        }
    };
    __device__ __forceinline__ Slot operator[](size_t i) const { return Slot{*this, i}; }
};
template <typename T> __device__ __forceinline__ OutRef<T> out_ref(T* out, const OutPeers<T>& pe, size_t off) {
    OutRef<T> r; r.out = out; r.pe = &pe; r.off = off; return r;
}

// Push a finished coefficient plane (n values at out + off, already stored locally and made visible by a CTA barrier) to
// the peers with 16-byte stores issued by ONE warp: multicast -> one multimem.st.v4 per 16 bytes, unicast -> one st.v4 per
// peer.  Decoupled from the low-pass tail that produced the plane: the other warps are already loading the next path.
__device__ __forceinline__ void multimem_st4(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
template <typename T>
__device__ __forceinline__ void push_plane(const T* out, const OutPeers<T>& pe, size_t off, int n, int lane) {
    if constexpr (sizeof(T) == 4) {
        if ((n & 3) == 0 && ((off & 3) == 0)) {
            const float4* src = reinterpret_cast<const float4*>(out + off);
            for (int i = lane; i < (n >> 2); i += 32) {
                const float4 v = __ldcg(src + i);
                if (pe.n < 0) multimem_st4(reinterpret_cast<float*>(pe.p[0] + off) + 4 * i, v);
                else for (int k = 0; k < pe.n; ++k) reinterpret_cast<float4*>(pe.p[k] + off)[i] = v;
            }
            return;
        }
    }
    for (int i = lane; i < n; i += 32) {
        const T v = __ldcg(out + off + i);
        if (pe.n < 0) multimem_st(pe.p[0] + off + i, v);
        else for (int k = 0; k < pe.n; ++k) pe.p[k][off + i] = v;
    }
}

template <typename U> __device__ __forceinline__ void stage(U* dst, const U* __restrict__ src, int n) {
    for (int i = flat_tid(); i < n; i += flat_nt()) dst[i] = src[i];
}

template <typename T> struct alignas(2 * sizeof(cx<T>)) cx2 { cx<T> a, b; };
template <typename T> struct alignas(2 * sizeof(T)) re2 { T a, b; };
template <typename T> struct alignas(4 * sizeof(T)) re4 { T a, b, c, d; };

// 16-byte predicated read-only global load (zero when the predicate is false): keeps the alias loads of
// the product/periodise prologue branch-free so that all of them are in flight together
__device__ __forceinline__ uint4 ldg16_pred(const void* ptr, bool pred) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w)
        : "l"(ptr), "r"((int)pred));
    return r;
}
template <typename V> __device__ __forceinline__ V ld_pred(const void* ptr, bool pred) {
    static_assert(sizeof(V) % 16 == 0, "ld_pred needs 16-byte multiples");
    union { V v; uint4 q[sizeof(V) / 16]; } u;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(V) / 16); ++i) u.q[i] = ldg16_pred(static_cast<const char*>(ptr) + 16 * i, pred);
    return u.v;
}

// Bulk L2 prefetch (cp.async.bulk.prefetch.L2, the TMA unit's prefetch form; SASS UBLKPF): pulls `bytes` (multiple of 16,
// 16-byte aligned) from HBM into L2 without occupying registers or shared memory.  The persistent tile kernels issue
// it one path ahead, so the product/periodise loads of the next path hit L2 instead of stalling on DRAM latency.
__device__ __forceinline__ void prefetch_l2_bulk(const void* ptr, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}
// rows [slice * ceil(P0/nslices), ...) of a P0 x P1 field at `base`: the slice is contiguous, so ONE thread issues ONE
// bulk prefetch for all of it (a per-row, per-lane version costs a serialising "waterfall" loop: UBLKPF takes uniform
// registers)
template <typename E>
__device__ __forceinline__ void prefetch_rows_slice(const E* base, int P0, int P1, int slice, int nslices, int tid, int nt) {
    (void)nt;
    if (tid != 0) return;
    const int per = (P0 + nslices - 1) / nslices;
    const int r0 = slice * per, r1 = min(P0, r0 + per);
    if (r1 <= r0) return;
    const size_t b0 = (size_t)r0 * P1 * sizeof(E), b1 = (size_t)r1 * P1 * sizeof(E);
    const size_t a0 = (b0 + 15) & ~(size_t)15, a1 = b1 & ~(size_t)15;      // 16-byte granules inside the slice
    if (a1 > a0) prefetch_l2_bulk(reinterpret_cast<const char*>(base) + a0, (unsigned)(a1 - a0));
}

// sum over the k x k aliases of (parent * filter) for output bin (r, e), skipping aliases outside
// the filter's per-row support interval.  supp may live in shared or global memory.
template <typename T>
__device__ __forceinline__ cx<T> prod_fold(const cx<T>* __restrict__ pb, const T* __restrict__ fb, const int2* supp,
                                           int r, int e, int k, int n0, int n1, int P1) {
    T ax = T(0), ay = T(0);
    for (int c = 0; c < k; ++c) {
        const int R = r + c * n0;
        const int2 sp = supp[R];
        if (sp.y == 0) continue;
        const size_t rowoff = (size_t)R * P1;
        for (int d = 0; d < k; ++d) {
            const int C = e + d * n1;
            int rel = C - sp.x;
            if (rel < 0) rel += P1;
            if (rel < sp.y) {
                const cx<T> v = pb[rowoff + C];
                const T f = fb[rowoff + C];
                ax += v.x * f; ay += v.y * f;
            }
        }
    }
    return mk<T>(ax, ay);
}

__device__ __forceinline__ int wrap(int t, int n) {   // t in (-n, 2n)
    if (t < 0) t += n; else if (t >= n) t -= n;
    return t;
}

// ===================================================================================
// Streaming slab kernels.  Template parameter NS > 0 selects a compile-time line length
// (static FFT plan, 16 lines per CTA, pitch 17): every index computation folds to constants.
// NS = 0 is the generic runtime-size version.
// ===================================================================================
constexpr int kSLines = 16;           // lines per CTA in the static variants
constexpr int kSLP = kSLines | 1;     // odd shared-memory pitch
// launch bounds of the slab kernels: static instances run 16 lines x <= 18 butterflies (<= 288 threads)
// and are compiled for three CTAs per SM (<= 75 registers); generic ones may use up to kMaxThreads
#define SB_SLAB_BOUNDS(NS) __launch_bounds__((NS) > 0 ? 288 : kMaxThreads, (NS) > 0 ? 4 : 1)

template <typename T> struct alignas(2 * sizeof(cx<T>)) cxpair { cx<T> a, b; };
template <typename T> struct alignas(2 * sizeof(T)) repair { T a, b; };

__device__ __forceinline__ float abs2f(float x, float y) {
    const float m2 = x * x + y * y;
    return m2 > 0.f ? m2 * rsqrtf(m2) : 0.f;
}
__device__ __forceinline__ double abs2f(double x, double y) { return sqrt(x * x + y * y); }

// ------------------------------------------------------------------ pad + row FFT
template <typename T> struct PadRowArgs {
    const T* x; cx<T>* out;
    int M, N, top, left, P0, P1;
    int lines, LP;
    Plan1 plan; const cx<T>* tw; const int* pos;
};
// grid (B, ceil(P0/lines)); reflect-pad rows on the fly, forward DIF along rows, natural-order store.
template <typename T, int NS> __global__ void SB_SLAB_BOUNDS(NS) k2d_pad_rowfft(PadRowArgs<T> a) {
    const int P1 = NS ? NS : a.P1, LP = NS ? kSLP : a.LP, lines = NS ? kSLines : a.lines;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)P1 * LP;
    int* pos = reinterpret_cast<int*>(tw + P1);
    const int b = blockIdx.x, r0 = blockIdx.y * lines;
    const int nl = min(lines, a.P0 - r0);
    stage(tw, a.tw, P1);
    stage(pos, a.pos, P1);
    const T* xb = a.x + (size_t)b * a.M * a.N;
    for (int idx = flat_tid(); idx < nl * P1; idx += flat_nt()) {
        const int l = idx / P1, e = idx - l * P1;
        const int sr = reflect_idx(r0 + l - a.top, a.M), sc = reflect_idx(e - a.left, a.N);
        s[e * LP + l] = mk<T>(xb[(size_t)sr * a.N + sc], T(0));
    }
    __syncthreads();
    if constexpr (NS > 0) slab_fft_s<NS, false, -1, 1, kSLP, T>(s, nl, tw);
    else slab_fft<false, T>(s, nl, 1, LP, a.plan, tw);
    cx<T>* ob = a.out + ((size_t)b * a.P0 + r0) * P1;
    for (int idx = flat_tid(); idx < nl * P1; idx += flat_nt()) {
        const int l = idx / P1, e = idx - l * P1;
        ob[(size_t)l * P1 + e] = s[pos[e] * LP + l];
    }
}

// ------------------------------------------------------------------ column pass
enum { COL_FWD = 0, COL_INV = 1, COL_INV_MOD_FWD = 2 };
template <typename T> struct ColArgs {
    const cx<T>* in; cx<T>* out;
    int n0, n1;
    int lines, LP;
    Plan1 plan; const cx<T>* tw; const int* pos;
};
// grid (G, ceil(n1/lines)); slab = `lines` adjacent columns, all n0 rows.
//   COL_FWD          spatial rows in  -> forward DIF along columns -> Fourier rows out
//   COL_INV          Fourier rows in  -> inverse DIT along columns -> spatial rows out
//   COL_INV_MOD_FWD  Fourier rows in  -> inverse DIT, modulus, forward DIF -> Fourier rows out
template <typename T, int MODE, int NS> __global__ void SB_SLAB_BOUNDS(NS) k2d_colpass(ColArgs<T> a) {
    const int n0 = NS ? NS : a.n0, LP = NS ? kSLP : a.LP, lines = NS ? kSLines : a.lines;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)n0 * LP;
    int* pos = reinterpret_cast<int*>(tw + n0);
    const int g = blockIdx.x, c0 = blockIdx.y * lines;
    const int nl = min(lines, a.n1 - c0);
    stage(tw, a.tw, n0);
    stage(pos, a.pos, n0);
    if (MODE != COL_FWD) __syncthreads();
    const cx<T>* ib = a.in + (size_t)g * n0 * a.n1 + c0;
    cx<T>* ob = a.out + (size_t)g * n0 * a.n1 + c0;
    const int tid = flat_tid(), nt = flat_nt();
    // static COL_INV_MOD_FWD runs DIF(+) -> modulus -> DIT(-): natural order on both global sides, no scatter
    constexpr bool PLAIN = (NS > 0 && MODE == COL_INV_MOD_FWD);
    if (NS > 0 && nl == kSLines && (a.n1 & 1) == 0) {
        // full slab: 8 lanes x 16 bytes per row segment
        for (int idx = tid; idx < n0 * (kSLines / 2); idx += nt) {
            const int e = idx / (kSLines / 2), l = 2 * (idx - e * (kSLines / 2));
            const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(ib + (size_t)e * a.n1 + l);
            const int se = (MODE == COL_FWD || PLAIN) ? e : pos[e];
            s[se * LP + l] = v.a; s[se * LP + l + 1] = v.b;
        }
    } else {
        for (int idx = tid; idx < n0 * lines; idx += nt) {
            const int e = idx / lines, l = idx - e * lines;
            if (l < nl) s[((MODE == COL_FWD || PLAIN) ? e : pos[e]) * LP + l] = ib[(size_t)e * a.n1 + l];
        }
    }
    __syncthreads();
    if (MODE == COL_FWD) {
        if constexpr (NS > 0) slab_fft_s<NS, false, -1, 1, kSLP, T>(s, nl, tw);
        else slab_fft<false, T>(s, nl, 1, LP, a.plan, tw);
    } else if constexpr (PLAIN) {
        slab_fft_s<NS, false, +1, 1, kSLP, T, true>(s, nl, tw);
        slab_fft_s<NS, true, -1, 1, kSLP, T>(s, nl, tw);
    } else {
        if constexpr (NS > 0) slab_fft_s<NS, true, +1, 1, kSLP, T>(s, nl, tw);
        else slab_fft<true, T>(s, nl, 1, LP, a.plan, tw);
        if (MODE == COL_INV_MOD_FWD) {
            for (int idx = tid; idx < n0 * lines; idx += nt) {
                const int e = idx / lines, l = idx - e * lines;
                const cx<T> v = s[e * LP + l];
                s[e * LP + l] = mk<T>(abs2f(v.x, v.y), T(0));
            }
            __syncthreads();
            if constexpr (NS > 0) slab_fft_s<NS, false, -1, 1, kSLP, T>(s, nl, tw);
            else slab_fft<false, T>(s, nl, 1, LP, a.plan, tw);
        }
    }
    if (NS > 0 && nl == kSLines && (a.n1 & 1) == 0) {
        for (int idx = tid; idx < n0 * (kSLines / 2); idx += nt) {
            const int e = idx / (kSLines / 2), l = 2 * (idx - e * (kSLines / 2));
            const int se = (MODE == COL_INV || PLAIN) ? e : pos[e];
            cxpair<T> v; v.a = s[se * LP + l]; v.b = s[se * LP + l + 1];
            *reinterpret_cast<cxpair<T>*>(ob + (size_t)e * a.n1 + l) = v;
        }
    } else {
        for (int idx = tid; idx < n0 * lines; idx += nt) {
            const int e = idx / lines, l = idx - e * lines;
            if (l < nl) ob[(size_t)e * a.n1 + l] = s[((MODE == COL_INV || PLAIN) ? e : pos[e]) * LP + l];
        }
    }
}

// ------------------------------------------------------------------ row pass, product + periodise prologue
template <typename T> struct RowProdArgs {
    const cx<T>* parent;      // [Bp][P0][P1] natural-order spectra
    const T* const* filt;     // [NF] pointers to real (P0, P1) filters, natural order; nullptr = unit filter (k = 1 only)
    const int2* supp;         // [NF][P0] per-row circular support (start, len)
    cx<T>* out;               // [Bp*NF][n0][n1]: Fourier rows (natural), spatial columns
    int P0, P1, k, n0, n1, NF;
    T scale;
    int lines, LP;
    Plan1 plan; const cx<T>* tw; const int* pos;   // length n1
};
// grid (G = Bp*NF, ceil(n0/lines)).  rows of out = inverse DIT along the row of
//   V[r][e] = scale * sum_{c,d<k} parent[r+c*n0][e+d*n1] * filt[r+c*n0][e+d*n1]
template <typename T, int NS> __global__ void SB_SLAB_BOUNDS(NS) k2d_rowpass_prod(RowProdArgs<T> a) {
    const int n1 = NS ? NS : a.n1, LP = NS ? kSLP : a.LP, lines = NS ? kSLines : a.lines;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)n1 * LP;
    int* pos = reinterpret_cast<int*>(tw + n1);
    const int g = blockIdx.x, r0 = blockIdx.y * lines;
    const int nl = min(lines, a.n0 - r0);
    const int fi = g % a.NF, pg = g / a.NF;
    stage(tw, a.tw, n1);
    stage(pos, a.pos, n1);
    __syncthreads();
    const cx<T>* __restrict__ pb = a.parent + (size_t)pg * a.P0 * a.P1;
    const T* __restrict__ fb = a.filt ? a.filt[fi] : nullptr;
    const int2* sp = a.supp + (size_t)fi * a.P0;
    const int tid = flat_tid(), nt = flat_nt();
    // static instances keep natural order in shared memory (the inverse runs as DIF and leaves the row
    // scrambled, which the following column/row passes of the chain expect); generic ones scatter for DIT
    if ((n1 & 1) == 0) {
        const int half = n1 >> 1, P1 = a.P1, k = a.k;
        for (int idx = tid; idx < nl * half; idx += nt) {
            const int l = idx / half, e = 2 * (idx - l * half);
            T ax0 = T(0), ay0 = T(0), ax1 = T(0), ay1 = T(0);
            if (k == 1) {
                // no aliases: plain vector loads (values outside the support interval are the true, tiny ones)
                const size_t off = (size_t)(r0 + l) * P1 + e;
                const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(pb + off);
                repair<T> f; f.a = T(1); f.b = T(1);
                if (fb) f = *reinterpret_cast<const repair<T>*>(fb + off);
                ax0 = v.a.x * f.a; ay0 = v.a.y * f.a; ax1 = v.b.x * f.b; ay1 = v.b.y * f.b;
            } else {
                for (int c = 0; c < k; ++c) {
                    const int R = r0 + l + c * a.n0;
                    const int2 iv = sp[R];
                    if (iv.y == 0) continue;
                    const size_t rowoff = (size_t)R * P1;
                    for (int d = 0; d < k; ++d) {
                        const int C = e + d * n1;
                        int rel = C - iv.x;
                        if (rel < 0) rel += P1;
                        if ((rel < iv.y) | (rel == P1 - 1)) {
                            const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(pb + rowoff + C);
                            const repair<T> f = *reinterpret_cast<const repair<T>*>(fb + rowoff + C);
                            ax0 += v.a.x * f.a; ay0 += v.a.y * f.a;
                            ax1 += v.b.x * f.b; ay1 += v.b.y * f.b;
                        }
                    }
                }
            }
            s[(NS ? e : pos[e]) * LP + l] = mk<T>(ax0 * a.scale, ay0 * a.scale);
            s[(NS ? e + 1 : pos[e + 1]) * LP + l] = mk<T>(ax1 * a.scale, ay1 * a.scale);
        }
    } else {
        for (int idx = tid; idx < nl * n1; idx += nt) {
            const int l = idx / n1, e = idx - l * n1;
            const cx<T> v = prod_fold<T>(pb, fb, sp, r0 + l, e, a.k, a.n0, n1, a.P1);
            s[(NS ? e : pos[e]) * LP + l] = scal(v, a.scale);
        }
    }
    __syncthreads();
    if constexpr (NS > 0) slab_fft_s<NS, false, +1, 1, kSLP, T>(s, nl, tw);
    else slab_fft<true, T>(s, nl, 1, LP, a.plan, tw);
    cx<T>* ob = a.out + ((size_t)g * a.n0 + r0) * n1;
    if ((n1 & 1) == 0) {
        const int half = n1 >> 1;
        for (int idx = tid; idx < nl * half; idx += nt) {
            const int l = idx / half, e = 2 * (idx - l * half);
            cxpair<T> v; v.a = s[e * LP + l]; v.b = s[(e + 1) * LP + l];
            *reinterpret_cast<cxpair<T>*>(ob + (size_t)l * n1 + e) = v;
        }
    } else {
        for (int idx = tid; idx < nl * n1; idx += nt) {
            const int l = idx / n1, e = idx - l * n1;
            ob[(size_t)l * n1 + e] = s[e * LP + l];
        }
    }
}

// ------------------------------------------------------------------ plain row pass
template <typename T> struct RowArgs {
    const cx<T>* in; cx<T>* out;
    int n0, n1;
    int lines, LP;
    Plan1 plan; const cx<T>* tw; const int* pos;
    // optional (forward static instances): also emit the low-pass product folded along the row,
    //   low_out[g][u][v'] = sum_d out[g][u][v' + d*low_m1] * low_filt[u][v' + d*low_m1],  v' < low_m1
    // so that the Fourier low-pass of this band needs n0 x low_m1 values per path instead of n0 x n1
    const T* low_filt; const int2* low_supp; cx<T>* low_out; int low_m1;
};
// INV=false: spatial row in -> DIF -> Fourier row out; INV=true: Fourier row in -> DIT -> spatial row out
template <typename T, bool INV, int NS> __global__ void SB_SLAB_BOUNDS(NS) k2d_rowpass(RowArgs<T> a) {
    const int n1 = NS ? NS : a.n1, LP = NS ? kSLP : a.LP, lines = NS ? kSLines : a.lines;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)n1 * LP;
    int* pos = reinterpret_cast<int*>(tw + n1);
    const int g = blockIdx.x, r0 = blockIdx.y * lines;
    const int nl = min(lines, a.n0 - r0);
    stage(tw, a.tw, n1);
    stage(pos, a.pos, n1);
    __syncthreads();
    const cx<T>* ib = a.in + ((size_t)g * a.n0 + r0) * n1;
    cx<T>* ob = a.out + ((size_t)g * a.n0 + r0) * n1;
    const int tid = flat_tid(), nt = flat_nt();
    // static forward instance: the row arrives scrambled (left so by the static chain's DIF inverse), DIT(-)
    // returns natural-order Fourier data - no permutation on either side
    constexpr bool PLAIN = (NS > 0 && !INV);
    if ((n1 & 1) == 0) {
        const int half = n1 >> 1;
        for (int idx = tid; idx < nl * half; idx += nt) {
            const int l = idx / half, e = 2 * (idx - l * half);
            const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(ib + (size_t)l * n1 + e);
            s[((INV && !PLAIN) ? pos[e] : e) * LP + l] = v.a;
            s[((INV && !PLAIN) ? pos[e + 1] : e + 1) * LP + l] = v.b;
        }
    } else {
        for (int idx = tid; idx < nl * n1; idx += nt) {
            const int l = idx / n1, e = idx - l * n1;
            s[(INV ? pos[e] : e) * LP + l] = ib[(size_t)l * n1 + e];
        }
    }
    __syncthreads();
    if constexpr (PLAIN) slab_fft_s<NS, true, -1, 1, kSLP, T>(s, nl, tw);
    else if constexpr (NS > 0) slab_fft_s<NS, true, +1, 1, kSLP, T>(s, nl, tw);
    else slab_fft<INV, T>(s, nl, 1, LP, a.plan, tw);
    if (PLAIN && a.low_out) {
        // natural-order Fourier rows are in shared memory: fold (row * phi) onto low_m1 columns
        const int m1 = a.low_m1, kf = n1 / m1;
        cx<T>* lo = a.low_out + ((size_t)g * a.n0 + r0) * m1;
        for (int idx = tid; idx < nl * m1; idx += nt) {
            const int l = idx / m1, e = idx - l * m1;
            const int u = r0 + l;
            const int2 sp = a.low_supp[u];
            T ax = T(0), ay = T(0);
            if (sp.y > 0) {
                const T* __restrict__ fr = a.low_filt + (size_t)u * n1;
                for (int d = 0; d < kf; ++d) {
                    const int C = e + d * m1;
                    int rel = C - sp.x;
                    if (rel < 0) rel += n1;
                    if (rel < sp.y) {
                        const cx<T> v = s[C * LP + l];
                        const T f = fr[C];
                        ax += v.x * f; ay += v.y * f;
                    }
                }
            }
            lo[(size_t)l * m1 + e] = mk<T>(ax, ay);
        }
    }
    if ((n1 & 1) == 0) {
        const int half = n1 >> 1;
        for (int idx = tid; idx < nl * half; idx += nt) {
            const int l = idx / half, e = 2 * (idx - l * half);
            cxpair<T> v;
            v.a = s[((INV || PLAIN) ? e : pos[e]) * LP + l];
            v.b = s[((INV || PLAIN) ? e + 1 : pos[e + 1]) * LP + l];
            *reinterpret_cast<cxpair<T>*>(ob + (size_t)l * n1 + e) = v;
        }
    } else {
        for (int idx = tid; idx < nl * n1; idx += nt) {
            const int l = idx / n1, e = idx - l * n1;
            ob[(size_t)l * n1 + e] = s[(INV ? e : pos[e]) * LP + l];
        }
    }
}

// ------------------------------------------------------------------ real-input (Hermitian) variants of the chain
// The field between the inverse and the forward transform is a modulus, i.e. REAL.  Static chains use it:
//   k2d_colpass_imrf : column inverse DIF(+), modulus, then ONE complex forward DIT(-) per PAIR of columns
//                      (z = A[:,x0] + i A[:,x1]); the two spectra are untangled and only rows v <= n0/2 are
//                      stored (the rest is the conjugate mirror);
//   k2d_rowpass_fwdh : forward DIT(-) along the rows v <= n0/2 only; every row is also written as its mirror
//                      U[(n0-v)%n0][(n1-w)%n1] = conj(U[v][w]), and (optionally) both feed the row-folded
//                      low-pass product (see RowArgs).
template <typename T, int NS> __global__ void SB_SLAB_BOUNDS(NS) k2d_colpass_imrf(ColArgs<T> a) {
    static_assert(NS > 0 && NS % 2 == 0, "static even sizes only");
    constexpr int n0 = NS, LP = kSLP, H = NS / 2;
    // prime-factor transforms where the length allows (272 = 16 x 17): the rows are scattered to their prime-factor input
    // positions while staging (free), both transforms run without twiddles between their two passes, and the natural
    // output row v of the forward (DIT) transform is found at pin[v]   (fft_core.cuh)
    // (measured at 272: 0.613 ms with the prime-factor variant against 0.589 ms without - the scattered staging stores and
    //  the pin[] lookups of the untangling step cost more than the 255 twiddle multiplications per column save)
    constexpr bool PFA = false && ct_pfa_ok(NS);
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)n0 * LP;
    int* pin = reinterpret_cast<int*>(tw + n0);
    const int g = blockIdx.x, c0 = blockIdx.y * kSLines;    // launch requires n1 % kSLines == 0
    stage(tw, a.tw, n0);
    if constexpr (PFA) { stage(pin, a.pos, n0); __syncthreads(); }      // a.pos = prime-factor input table for PFA sizes
    const cx<T>* ib = a.in + (size_t)g * n0 * a.n1 + c0;
    cx<T>* ob = a.out + (size_t)g * n0 * a.n1 + c0;
    const int tid = flat_tid(), nt = flat_nt();
    for (int idx = tid; idx < n0 * (kSLines / 2); idx += nt) {
        const int e = idx / (kSLines / 2), l = 2 * (idx - e * (kSLines / 2));
        const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(ib + (size_t)e * a.n1 + l);
        const int se = PFA ? pin[e] : e;
        s[se * LP + l] = v.a; s[se * LP + l + 1] = v.b;
    }
    __syncthreads();
    slab_fft_s<NS, false, +1, 1, kSLP, T, 1, PFA>(s, kSLines, tw);         // inverse + modulus -> (|u|, 0), rows scrambled
    // pack column pairs: (|u|_{2m}, |u|_{2m+1}) -> one complex column at lane 2m
    for (int idx = tid; idx < n0 * (kSLines / 2); idx += nt) {
        const int e = idx / (kSLines / 2), l = 2 * (idx - e * (kSLines / 2));
        s[e * LP + l].y = s[e * LP + l + 1].x;
    }
    __syncthreads();
    slab_fft_s<NS, true, -1, 2, kSLP, T, 0, PFA>(s, kSLines / 2, tw);      // forward on the 8 packed columns
    // untangle: A[v] = (Z[v] + conj Z[n-v]) / 2,  B[v] = (Z[v] - conj Z[n-v]) / (2i);  rows v = 0..n0/2
    for (int idx = tid; idx < (H + 1) * (kSLines / 2); idx += nt) {
        const int v = idx / (kSLines / 2), l = 2 * (idx - v * (kSLines / 2));
        const int vm = v == 0 ? 0 : n0 - v;
        const cx<T> z = s[(PFA ? pin[v] : v) * LP + l], zm = s[(PFA ? pin[vm] : vm) * LP + l];
        cxpair<T> o;
        o.a = mk<T>(T(0.5) * (z.x + zm.x), T(0.5) * (z.y - zm.y));
        o.b = mk<T>(T(0.5) * (z.y + zm.y), T(0.5) * (zm.x - z.x));
        *reinterpret_cast<cxpair<T>*>(ob + (size_t)v * a.n1 + l) = o;
    }
}

template <typename T, int NS> __global__ void SB_SLAB_BOUNDS(NS) k2d_rowpass_fwdh(RowArgs<T> a) {
    static_assert(NS > 0 && NS % 2 == 0, "static even sizes only");
    constexpr int n1 = NS, LP = kSLP;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)n1 * LP;
    const int g = blockIdx.x, r0 = blockIdx.y * kSLines;
    const int H = a.n0 / 2;
    const int nl = min(kSLines, H + 1 - r0);                 // rows v = r0 .. r0+nl-1 <= n0/2
    stage(tw, a.tw, n1);
    const cx<T>* ib = a.in + ((size_t)g * a.n0 + r0) * n1;
    cx<T>* ob = a.out + (size_t)g * a.n0 * n1;
    const int tid = flat_tid(), nt = flat_nt();
    constexpr int half = n1 / 2;
    for (int idx = tid; idx < nl * half; idx += nt) {
        const int l = idx / half, e = 2 * (idx - l * half);
        const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(ib + (size_t)l * n1 + e);
        s[e * LP + l] = v.a; s[(e + 1) * LP + l] = v.b;
    }
    __syncthreads();
    slab_fft_s<NS, true, -1, 1, kSLP, T>(s, nl, tw);
    // row v and its conjugate mirror row (n0 - v): U[n0-v][w] = conj(U[v][(n1-w)%n1])
    for (int idx = tid; idx < nl * half; idx += nt) {
        const int l = idx / half, e = 2 * (idx - l * half);
        const int v = r0 + l;
        cxpair<T> o; o.a = s[e * LP + l]; o.b = s[(e + 1) * LP + l];
        *reinterpret_cast<cxpair<T>*>(ob + (size_t)v * n1 + e) = o;
        if (v > 0 && v < H) {
            const cx<T> m0 = s[(e == 0 ? 0 : n1 - e) * LP + l], m1 = s[(n1 - e - 1) * LP + l];
            cxpair<T> om; om.a = mk<T>(m0.x, -m0.y); om.b = mk<T>(m1.x, -m1.y);
            *reinterpret_cast<cxpair<T>*>(ob + (size_t)(a.n0 - v) * n1 + e) = om;
        }
    }
    if (a.low_out) {
        const int m1 = a.low_m1, kf = n1 / m1;
        for (int idx = tid; idx < 2 * nl * m1; idx += nt) {
            const int mir = idx / (nl * m1), rem = idx - mir * nl * m1;
            const int l = rem / m1, e = rem - l * m1;
            const int v = r0 + l;
            if (mir && !(v > 0 && v < H)) continue;
            const int u = mir ? a.n0 - v : v;
            const int2 sp = a.low_supp[u];
            T ax = T(0), ay = T(0);
            if (sp.y > 0) {
                const T* __restrict__ fr = a.low_filt + (size_t)u * n1;
                for (int d = 0; d < kf; ++d) {
                    const int C = e + d * m1;
                    int rel = C - sp.x;
                    if (rel < 0) rel += n1;
                    if (rel < sp.y) {
                        const cx<T> t = s[(mir ? (C == 0 ? 0 : n1 - C) : C) * LP + l];
                        const T f = fr[C];
                        ax += t.x * f; ay += (mir ? -t.y : t.y) * f;
                    }
                }
            }
            a.low_out[((size_t)g * a.n0 + u) * m1 + e] = mk<T>(ax, ay);
        }
    }
}

// ------------------------------------------------------------------ Fourier low-pass tile (generic phi)
template <typename T> struct LowArgs {
    const cx<T>* in;      // [G][P0][P1] natural-order spectra
    const T* filt;        // [P0][P1] real low-pass at this resolution
    const int2* supp;     // [P0]
    T* out;               // [B][K][m0-2][m1-2]
    int P0, P1, k, m0, m1, W;
    int PP, NF, ch0, chs, K;
    T scale;
    Plan1 plan0, plan1; const cx<T>* tw0; const cx<T>* tw1; const int* pos0; const int* pos1;
    int row_folded;       // 1: `in` is [G][P0][m1], already multiplied by the filter and folded along rows
    OutPeers<T> peers;    // additional destinations of the channel planes (peer GPUs)
};
// grid (G).  One CTA: periodise (in*filt) to m0 x m1, inverse 2-D DIT in shared memory,
// keep the real part, crop one sample per side (unpad) and write the channel plane.
template <typename T> __global__ void __launch_bounds__(kMaxThreads) k2d_lowpass(LowArgs<T> a) {
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw0 = s + (size_t)a.m0 * a.W;
    cx<T>* tw1 = tw0 + a.m0;
    int* pos0 = reinterpret_cast<int*>(tw1 + a.m1);
    int* pos1 = pos0 + a.m0;
    const int g = blockIdx.x;
    const int b = g / a.PP, path = g - b * a.PP;
    const int ch = a.ch0 + (path / a.NF) * a.chs + (path % a.NF);
    stage(tw0, a.tw0, a.m0); stage(tw1, a.tw1, a.m1);
    stage(pos0, a.pos0, a.m0); stage(pos1, a.pos1, a.m1);
    __syncthreads();
    const cx<T>* pb = a.in + (size_t)g * a.P0 * a.P1;
    for (int idx = flat_tid(); idx < a.m0 * a.m1; idx += flat_nt()) {
        const int r = idx / a.m1, e = idx - r * a.m1;
        cx<T> v;
        if (a.row_folded) {
            T ax = T(0), ay = T(0);
            const cx<T>* rb = a.in + (size_t)g * a.P0 * a.m1;
            for (int c = 0; c < a.k; ++c) { const cx<T> t = rb[(size_t)(r + c * a.m0) * a.m1 + e]; ax += t.x; ay += t.y; }
            v = mk<T>(ax, ay);
        } else {
            v = prod_fold<T>(pb, a.filt, a.supp, r, e, a.k, a.m0, a.m1, a.P1);
        }
        s[pos0[r] * a.W + pos1[e]] = scal(v, a.scale);
    }
    __syncthreads();
    slab_fft<true, T>(s, a.m0, a.W, 1, a.plan1, tw1);   // along rows (length m1)
    slab_fft<true, T>(s, a.m1, 1, a.W, a.plan0, tw0);   // along columns (length m0)
    const int o0 = a.m0 - 2, o1 = a.m1 - 2;
    const OutRef<T> ob = out_ref(a.out, a.peers, ((size_t)b * a.K + ch) * o0 * o1);
    for (int idx = flat_tid(); idx < o0 * o1; idx += flat_nt()) {
        const int y = idx / o1, x = idx - y * o1;
        ob[idx] = s[(y + 1) * a.W + (x + 1)].x;
    }
}

// ------------------------------------------------------------------ crop + real (streaming low-pass tail)
template <typename T> struct CropArgs {
    const cx<T>* in;   // [G][m0][m1] natural-order spatial field
    T* out;            // [B][K][m0-2][m1-2]
    int m0, m1, PP, NF, ch0, chs, K;
    OutPeers<T> peers;
};
template <typename T> __global__ void k2d_crop_real(CropArgs<T> a) {
    const int g = blockIdx.x;
    const int b = g / a.PP, path = g - b * a.PP;
    const int ch = a.ch0 + (path / a.NF) * a.chs + (path % a.NF);
    const int o0 = a.m0 - 2, o1 = a.m1 - 2;
    const int idx = blockIdx.y * blockDim.x + threadIdx.x;
    if (idx >= o0 * o1) return;
    const int y = idx / o1, x = idx - y * o1;
    out_ref(a.out, a.peers, ((size_t)b * a.K + ch) * o0 * o1)[idx] = a.in[((size_t)g * a.m0 + y + 1) * a.m1 + x + 1].x;
}

}  // namespace sb

namespace sb {
// Kernel table for one line length (compiled in stream_inst.cu); `is_static` tells whether the
// entries are compile-time specialised (they then require 16 lines per CTA, pitch 17).
template <typename T> struct StreamKernels {
    void (*pad_rowfft)(PadRowArgs<T>);
    void (*col_fwd)(ColArgs<T>);
    void (*col_inv)(ColArgs<T>);
    void (*col_imf)(ColArgs<T>);
    void (*row_prod)(RowProdArgs<T>);
    void (*row_fwd)(RowArgs<T>);
    void (*row_inv)(RowArgs<T>);
    void (*col_imrf)(ColArgs<T>);     // real-input column pass (static even sizes only, else null)
    void (*row_fwdh)(RowArgs<T>);     // Hermitian forward row pass (static even sizes only, else null)
    bool is_static;
};
template <typename T> StreamKernels<T> stream_kernels_lookup(int n, bool allow_static);
template <typename T> void stream_kernels_enable_smem();
}  // namespace sb
