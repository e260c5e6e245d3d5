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

template <typename U> __device__ __forceinline__ void stage(U* dst, const U* __restrict__ src, int n) {
    for (int i = flat_tid(); i < n; i += flat_nt()) dst[i] = src[i];
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

// ------------------------------------------------------------------ pad + row FFT
template <typename T> struct PadRowArgs {
    const T* x; cx<T>* out;
    int M, N, top, left, P0, P1;
    int lines, LP;
    Plan1 plan; const cx<T>* tw; const int* pos;
};
// grid (B, ceil(P0/lines)); reflect-pad rows on the fly, forward DIF along rows, natural-order store.
template <typename T> __global__ void __launch_bounds__(kMaxThreads) k2d_pad_rowfft(PadRowArgs<T> a) {
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)a.P1 * a.LP;
    int* pos = reinterpret_cast<int*>(tw + a.P1);
    const int b = blockIdx.x, r0 = blockIdx.y * a.lines;
    const int nl = min(a.lines, a.P0 - r0);
    stage(tw, a.tw, a.P1);
    stage(pos, a.pos, a.P1);
    const T* xb = a.x + (size_t)b * a.M * a.N;
    for (int idx = flat_tid(); idx < nl * a.P1; idx += flat_nt()) {
        const int l = idx / a.P1, e = idx - l * a.P1;
        const int sr = reflect_idx(r0 + l - a.top, a.M), sc = reflect_idx(e - a.left, a.N);
        s[e * a.LP + l] = mk<T>(xb[(size_t)sr * a.N + sc], T(0));
    }
    __syncthreads();
    slab_fft<false, T>(s, nl, 1, a.LP, a.plan, tw);
    cx<T>* ob = a.out + ((size_t)b * a.P0 + r0) * a.P1;
    for (int idx = flat_tid(); idx < nl * a.P1; idx += flat_nt()) {
        const int l = idx / a.P1, e = idx - l * a.P1;
        ob[(size_t)l * a.P1 + e] = s[pos[e] * a.LP + l];
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
template <typename T, int MODE> __global__ void __launch_bounds__(kMaxThreads) k2d_colpass(ColArgs<T> a) {
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)a.n0 * a.LP;
    int* pos = reinterpret_cast<int*>(tw + a.n0);
    const int g = blockIdx.x, c0 = blockIdx.y * a.lines;
    const int nl = min(a.lines, a.n1 - c0);
    stage(tw, a.tw, a.n0);
    stage(pos, a.pos, a.n0);
    if (MODE != COL_FWD) __syncthreads();
    const cx<T>* ib = a.in + (size_t)g * a.n0 * a.n1 + c0;
    for (int e = threadIdx.y; e < a.n0; e += blockDim.y) {
        const int se = (MODE == COL_FWD) ? e : pos[e];
        for (int l = threadIdx.x; l < nl; l += blockDim.x)
            s[se * a.LP + l] = ib[(size_t)e * a.n1 + l];
    }
    __syncthreads();
    if (MODE == COL_FWD) {
        slab_fft<false, T>(s, nl, 1, a.LP, a.plan, tw);
    } else {
        slab_fft<true, T>(s, nl, 1, a.LP, a.plan, tw);
        if (MODE == COL_INV_MOD_FWD) {
            for (int e = threadIdx.y; e < a.n0; e += blockDim.y)
                for (int l = threadIdx.x; l < nl; l += blockDim.x) {
                    const cx<T> v = s[e * a.LP + l];
                    s[e * a.LP + l] = mk<T>(sqrt(v.x * v.x + v.y * v.y), T(0));
                }
            __syncthreads();
            slab_fft<false, T>(s, nl, 1, a.LP, a.plan, tw);
        }
    }
    cx<T>* ob = a.out + (size_t)g * a.n0 * a.n1 + c0;
    for (int e = threadIdx.y; e < a.n0; e += blockDim.y) {
        const int se = (MODE == COL_INV) ? e : pos[e];
        for (int l = threadIdx.x; l < nl; l += blockDim.x)
            ob[(size_t)e * a.n1 + l] = s[se * a.LP + l];
    }
}

// ------------------------------------------------------------------ row pass, product + periodise prologue
template <typename T> struct RowProdArgs {
    const cx<T>* parent;      // [Bp][P0][P1] natural-order spectra
    const T* const* filt;     // [NF] pointers to real (P0, P1) filters, natural order
    const int2* supp;         // [NF][P0] per-row circular support (start, len)
    cx<T>* out;               // [Bp*NF][n0][n1]: Fourier rows (natural), spatial columns
    int P0, P1, k, n0, n1, NF;
    T scale;
    int lines, LP;
    Plan1 plan; const cx<T>* tw; const int* pos;   // length n1
};
// grid (G = Bp*NF, ceil(n0/lines)).  rows of out = inverse DIT along the row of
//   V[r][e] = scale * sum_{c,d<k} parent[r+c*n0][e+d*n1] * filt[r+c*n0][e+d*n1]
template <typename T> __global__ void __launch_bounds__(kMaxThreads) k2d_rowpass_prod(RowProdArgs<T> a) {
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)a.n1 * a.LP;
    int* pos = reinterpret_cast<int*>(tw + a.n1);
    const int g = blockIdx.x, r0 = blockIdx.y * a.lines;
    const int nl = min(a.lines, a.n0 - r0);
    const int fi = g % a.NF, pg = g / a.NF;
    stage(tw, a.tw, a.n1);
    stage(pos, a.pos, a.n1);
    __syncthreads();
    const cx<T>* pb = a.parent + (size_t)pg * a.P0 * a.P1;
    const T* fb = a.filt[fi];
    const int2* sp = a.supp + (size_t)fi * a.P0;
    for (int idx = flat_tid(); idx < nl * a.n1; idx += flat_nt()) {
        const int l = idx / a.n1, e = idx - l * a.n1;
        const cx<T> v = prod_fold<T>(pb, fb, sp, r0 + l, e, a.k, a.n0, a.n1, a.P1);
        s[pos[e] * a.LP + l] = scal(v, a.scale);
    }
    __syncthreads();
    slab_fft<true, T>(s, nl, 1, a.LP, a.plan, tw);
    cx<T>* ob = a.out + ((size_t)g * a.n0 + r0) * a.n1;
    for (int idx = flat_tid(); idx < nl * a.n1; idx += flat_nt()) {
        const int l = idx / a.n1, e = idx - l * a.n1;
        ob[(size_t)l * a.n1 + e] = s[e * a.LP + l];
    }
}

// ------------------------------------------------------------------ plain row pass
template <typename T> struct RowArgs {
    const cx<T>* in; cx<T>* out;
    int n0, n1;
    int lines, LP;
    Plan1 plan; const cx<T>* tw; const int* pos;
};
// INV=false: spatial row in -> DIF -> Fourier row out; INV=true: Fourier row in -> DIT -> spatial row out
template <typename T, bool INV> __global__ void __launch_bounds__(kMaxThreads) k2d_rowpass(RowArgs<T> a) {
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)a.n1 * a.LP;
    int* pos = reinterpret_cast<int*>(tw + a.n1);
    const int g = blockIdx.x, r0 = blockIdx.y * a.lines;
    const int nl = min(a.lines, a.n0 - r0);
    stage(tw, a.tw, a.n1);
    stage(pos, a.pos, a.n1);
    if (INV) __syncthreads();
    const cx<T>* ib = a.in + ((size_t)g * a.n0 + r0) * a.n1;
    for (int idx = flat_tid(); idx < nl * a.n1; idx += flat_nt()) {
        const int l = idx / a.n1, e = idx - l * a.n1;
        s[(INV ? pos[e] : e) * a.LP + l] = ib[(size_t)l * a.n1 + e];
    }
    __syncthreads();
    slab_fft<INV, T>(s, nl, 1, a.LP, a.plan, tw);
    cx<T>* ob = a.out + ((size_t)g * a.n0 + r0) * a.n1;
    for (int idx = flat_tid(); idx < nl * a.n1; idx += flat_nt()) {
        const int l = idx / a.n1, e = idx - l * a.n1;
        ob[(size_t)l * a.n1 + e] = s[(INV ? e : pos[e]) * a.LP + l];
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
        const cx<T> v = prod_fold<T>(pb, a.filt, a.supp, r, e, a.k, a.m0, a.m1, a.P1);
        s[pos0[r] * a.W + pos1[e]] = scal(v, a.scale);
    }
    __syncthreads();
    slab_fft<true, T>(s, a.m0, a.W, 1, a.plan1, tw1);   // along rows (length m1)
    slab_fft<true, T>(s, a.m1, 1, a.W, a.plan0, tw0);   // along columns (length m0)
    const int o0 = a.m0 - 2, o1 = a.m1 - 2;
    T* ob = a.out + ((size_t)b * a.K + ch) * o0 * o1;
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
};
template <typename T> __global__ void k2d_crop_real(CropArgs<T> a) {
    const int g = blockIdx.x;
    const int b = g / a.PP, path = g - b * a.PP;
    const int ch = a.ch0 + (path / a.NF) * a.chs + (path % a.NF);
    const int o0 = a.m0 - 2, o1 = a.m1 - 2;
    const int idx = blockIdx.y * blockDim.x + threadIdx.x;
    if (idx >= o0 * o1) return;
    const int y = idx / o1, x = idx - y * o1;
    a.out[((size_t)b * a.K + ch) * o0 * o1 + idx] = a.in[((size_t)g * a.m0 + y + 1) * a.m1 + x + 1].x;
}

}  // namespace sb
