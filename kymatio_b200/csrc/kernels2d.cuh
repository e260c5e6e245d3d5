// kernels2d.cuh - fused 2-D scattering kernels (streaming row / column passes + low-pass tile).
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
// Layout: spatial data natural order, Fourier data canonical scrambled order per axis
// (fft_core.cuh).  A "slab" is `lines` adjacent lines staged in shared memory as
// s[e*LP + l] (element-major, LP odd) so that both the coalesced global side and the
// butterfly side are bank-conflict free.
#pragma once
#include "slab.cuh"

namespace sb {

template <typename T> __device__ __forceinline__ T* dyn_smem() {
    extern __shared__ __align__(16) unsigned char sb_dyn_smem[];
    return reinterpret_cast<T*>(sb_dyn_smem);
}

// ------------------------------------------------------------------ pad + row FFT
template <typename T> struct PadRowArgs {
    const T* x; cx<T>* out;
    int M, N, top, left, P0, P1;
    int lines, LP;
    Plan1 plan; const cx<T>* tw;
};
// grid (B, ceil(P0/lines)); reflect-pad rows on the fly, forward DIF along rows.
template <typename T> __global__ void __launch_bounds__(kMaxThreads) k2d_pad_rowfft(PadRowArgs<T> a) {
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)a.P1 * a.LP;
    const int b = blockIdx.x, r0 = blockIdx.y * a.lines;
    const int nl = min(a.lines, a.P0 - r0);
    copy_tw(tw, a.tw, a.P1);
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
        ob[(size_t)l * a.P1 + e] = s[e * a.LP + l];
    }
}

// ------------------------------------------------------------------ column pass
enum { COL_FWD = 0, COL_INV = 1, COL_INV_MOD_FWD = 2 };
template <typename T> struct ColArgs {
    const cx<T>* in; cx<T>* out;
    int n0, n1;
    int lines, LP;
    Plan1 plan; const cx<T>* tw;
};
// grid (G, ceil(n1/lines)); slab = `lines` adjacent columns, all n0 rows.
//   COL_FWD          forward DIF along columns
//   COL_INV          inverse DIT along columns
//   COL_INV_MOD_FWD  inverse DIT, complex modulus, forward DIF (imag = 0)
template <typename T, int MODE> __global__ void __launch_bounds__(kMaxThreads) k2d_colpass(ColArgs<T> a) {
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)a.n0 * a.LP;
    const int g = blockIdx.x, c0 = blockIdx.y * a.lines;
    const int nl = min(a.lines, a.n1 - c0);
    copy_tw(tw, a.tw, a.n0);
    const cx<T>* ib = a.in + (size_t)g * a.n0 * a.n1 + c0;
    for (int e = threadIdx.y; e < a.n0; e += blockDim.y)
        for (int l = threadIdx.x; l < nl; l += blockDim.x)
            s[e * a.LP + l] = ib[(size_t)e * a.n1 + l];
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
    for (int e = threadIdx.y; e < a.n0; e += blockDim.y)
        for (int l = threadIdx.x; l < nl; l += blockDim.x)
            ob[(size_t)e * a.n1 + l] = s[e * a.LP + l];
}

// ------------------------------------------------------------------ row pass, product + periodise prologue
template <typename T> struct RowProdArgs {
    const cx<T>* parent;   // [B*NP][P0][P1] scrambled spectra
    const T* filt;         // [NF][P0][P1] real scrambled filters
    cx<T>* out;            // [B*NP*NF][n0][n1]
    int P0, P1, k, n0, n1, NP, NF;
    T scale;
    int lines, LP;
    Plan1 plan; const cx<T>* tw;   // length n1
};
// grid (G = B*NP*NF, ceil(n0/lines)).  out rows = inverse DIT along rows of
//   V[r][e] = scale * sum_{c,d<k} parent[r*k+c][e*k+d] * filt[r*k+c][e*k+d]
// (the k x k aliases of the Fourier periodisation are adjacent in scrambled order).
template <typename T> __global__ void __launch_bounds__(kMaxThreads) k2d_rowpass_prod(RowProdArgs<T> a) {
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)a.n1 * a.LP;
    const int g = blockIdx.x, r0 = blockIdx.y * a.lines;
    const int nl = min(a.lines, a.n0 - r0);
    const int fi = g % a.NF, pg = g / a.NF;
    copy_tw(tw, a.tw, a.n1);
    const cx<T>* pb = a.parent + (size_t)pg * a.P0 * a.P1;
    const T* fb = a.filt + (size_t)fi * a.P0 * a.P1;
    const int k = a.k;
    for (int idx = flat_tid(); idx < nl * a.n1; idx += flat_nt()) {
        const int l = idx / a.n1, e = idx - l * a.n1;
        T ax = T(0), ay = T(0);
        for (int c = 0; c < k; ++c) {
            const size_t off = (size_t)((r0 + l) * k + c) * a.P1 + (size_t)e * k;
            for (int d = 0; d < k; ++d) {
                const cx<T> v = pb[off + d];
                const T f = fb[off + d];
                ax += v.x * f; ay += v.y * f;
            }
        }
        s[e * a.LP + l] = mk<T>(ax * a.scale, ay * a.scale);
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
    Plan1 plan; const cx<T>* tw;
};
template <typename T, bool INV> __global__ void __launch_bounds__(kMaxThreads) k2d_rowpass(RowArgs<T> a) {
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)a.n1 * a.LP;
    const int g = blockIdx.x, r0 = blockIdx.y * a.lines;
    const int nl = min(a.lines, a.n0 - r0);
    copy_tw(tw, a.tw, a.n1);
    const cx<T>* ib = a.in + ((size_t)g * a.n0 + r0) * a.n1;
    for (int idx = flat_tid(); idx < nl * a.n1; idx += flat_nt()) {
        const int l = idx / a.n1, e = idx - l * a.n1;
        s[e * a.LP + l] = ib[(size_t)l * a.n1 + e];
    }
    __syncthreads();
    slab_fft<INV, T>(s, nl, 1, a.LP, a.plan, tw);
    cx<T>* ob = a.out + ((size_t)g * a.n0 + r0) * a.n1;
    for (int idx = flat_tid(); idx < nl * a.n1; idx += flat_nt()) {
        const int l = idx / a.n1, e = idx - l * a.n1;
        ob[(size_t)l * a.n1 + e] = s[e * a.LP + l];
    }
}

// ------------------------------------------------------------------ low-pass tile
template <typename T> struct LowArgs {
    const cx<T>* in;      // [G][P0][P1] scrambled spectra
    const T* filt;        // [P0][P1] real scrambled low-pass at this resolution
    T* out;               // [B][K][m0-2][m1-2]
    int P0, P1, k, m0, m1, W;
    int PP, NF, ch0, chs, K;
    T scale;
    Plan1 plan0, plan1; const cx<T>* tw0; const cx<T>* tw1;
};
// grid (G).  One CTA: periodise (in*filt) to m0 x m1, inverse 2-D DIT in shared memory,
// keep the real part, crop one sample per side (unpad) and write the channel plane.
template <typename T> __global__ void __launch_bounds__(kMaxThreads) k2d_lowpass(LowArgs<T> a) {
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw0 = s + (size_t)a.m0 * a.W;
    cx<T>* tw1 = tw0 + a.m0;
    const int g = blockIdx.x;
    const int b = g / a.PP, path = g - b * a.PP;
    const int ch = a.ch0 + (path / a.NF) * a.chs + (path % a.NF);
    copy_tw(tw0, a.tw0, a.m0);
    copy_tw(tw1, a.tw1, a.m1);
    const cx<T>* pb = a.in + (size_t)g * a.P0 * a.P1;
    const int k = a.k;
    for (int idx = flat_tid(); idx < a.m0 * a.m1; idx += flat_nt()) {
        const int r = idx / a.m1, e = idx - r * a.m1;
        T ax = T(0), ay = T(0);
        for (int c = 0; c < k; ++c) {
            const size_t off = (size_t)(r * k + c) * a.P1 + (size_t)e * k;
            for (int d = 0; d < k; ++d) {
                const cx<T> v = pb[off + d];
                const T f = a.filt[off + d];
                ax += v.x * f; ay += v.y * f;
            }
        }
        s[r * a.W + e] = mk<T>(ax * a.scale, ay * a.scale);
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

// ------------------------------------------------------------------ filter scramble
// dst[pos0[r]][pos1[c]] = src[r][c]  (natural -> canonical scrambled, per axis)
template <typename T>
__global__ void k2d_scramble(const T* __restrict__ src, T* __restrict__ dst, const int* __restrict__ pos0,
                             const int* __restrict__ pos1, int n0, int n1) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c < n1) dst[(size_t)pos0[r] * n1 + pos1[c]] = src[(size_t)r * n1 + c];
}

}  // namespace sb
