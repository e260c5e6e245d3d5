// kernels3d.cuh - fused 3-D solid-harmonic scattering kernels (float32, power-of-two volumes).
//
// Reference semantics being fused (paths relative to the kymatio tree):
//   cascade            kymatio/scattering3d/core/scattering3d.py:24-73
//   cdgmm3d            kymatio/backend/torch_backend.py:207-219          (complex x complex)
//   ifft / rfft        kymatio/scattering3d/backend/torch_backend.py:81-95 (ifftn / fftn over the last 3 axes)
//   modulus_rotation   kymatio/scattering3d/backend/torch_backend.py:102-124 (sqrt(prev^2 + |x|^2) over m)
//   compute_integrals  kymatio/scattering3d/backend/torch_backend.py:127-151
//
// A volume (M, N, O) is stored with O fastest.  A 3-D transform is TWO passes over HBM:
//   * along M: slabs of 16 adjacent (n, o) columns, all M rows               (k3d_col_prod, k1d_col_fwd)
//   * the (N, O) plane of one m: a 2-D transform held entirely in shared memory (k3d_plane)
// Inverse transforms run decimation-in-frequency (natural in, scrambled out) and forward ones decimation-in-time
// (scrambled in, natural out), so the spatial field between them stays in scrambled positions along all three
// axes - harmless, because only order-agnostic operations (|.|^2 accumulation over m, sqrt, voxel sums) touch it.
//
//   k3d_col_prod   Y[b, mi] = ifft_M( U_hat[b] * Psi[mi] ) / (MNO)            one launch per band (l, j), all m
//   k3d_plane      U = sqrt(sum_mi |ifft_NO(Y[b, mi])|^2) per plane, held in registers across the m loop;
//                  integrals sum U^q accumulated (float64 atomics); parents: (U, 0) -> fft_NO -> spectrum plane
//   k1d_col_fwd    fft_M of the parent spectrum planes (kernels1d.cuh, NA = M, NB = N*O)
#pragma once
#include "kernels1d.cuh"

namespace sb {

constexpr int k3Threads = 512;

// ------------------------------------------------------------------ M-axis inverse with the filter product
template <typename T> struct ColProd3 {
    const cx<T>* U;        // [B][M][NO] natural-order spectrum
    const cx<T>* filt;     // [nm][M][NO] complex filters of this band
    cx<T>* Y;              // [B*nm][M (scrambled)][NO]
    int B, nm, NO;
    T scale;               // 1 / (M N O)
    const cx<T>* twM;
};
template <typename T, int M> __global__ void __launch_bounds__(k1Threads, 3) k3d_col_prod(ColProd3<T> a) {
    constexpr int LP = k1LP;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)M * LP;
    const int ncol = a.NO / k1L;
    const int cg = blockIdx.x % ncol, bm = blockIdx.x / ncol;      // column group fastest, then b, then mi
    const int b = bm % a.B, mi = bm / a.B;
    const int c0 = cg * k1L;
    const int tid = flat_tid(), nt = flat_nt();
    stage(tw, a.twM, M);
    const cx<T>* __restrict__ ub = a.U + (size_t)b * M * a.NO + c0;
    const cx<T>* __restrict__ fb = a.filt + (size_t)mi * M * a.NO + c0;
    for (int idx = tid; idx < M * (k1L / 2); idx += nt) {
        const int r = idx / (k1L / 2), l = 2 * (idx - r * (k1L / 2));
        const cxpair<T> u = *reinterpret_cast<const cxpair<T>*>(ub + (size_t)r * a.NO + l);
        const cxpair<T> f = *reinterpret_cast<const cxpair<T>*>(fb + (size_t)r * a.NO + l);
        s[r * LP + l] = scal(cmul(u.a, f.a), a.scale);
        s[r * LP + l + 1] = scal(cmul(u.b, f.b), a.scale);
    }
    __syncthreads();
    slab_fft_s<M, false, +1, 1, k1LP, T>(s, k1L, tw);
    cx<T>* yb = a.Y + ((size_t)(b * a.nm + mi) * M) * a.NO + c0;
    for (int idx = tid; idx < M * (k1L / 2); idx += nt) {
        const int p = idx / (k1L / 2), l = 2 * (idx - p * (k1L / 2));
        cxpair<T> o; o.a = s[p * LP + l]; o.b = s[p * LP + l + 1];
        *reinterpret_cast<cxpair<T>*>(yb + (size_t)p * a.NO + l) = o;
    }
}

// ------------------------------------------------------------------ plane pass: 2-D inverse, rotation modulus, integrals, 2-D forward
template <typename T> struct Plane3 {
    const cx<T>* Y;        // [B*nm][M][N][O]
    cx<T>* spec;           // parents: [B][M][N][O] plane spectra (M still scrambled); leaves: nullptr
    double* integ;         // integ[b*istride + ioff + p] += sum U^{q_p}
    const float* powers; int P;
    long long istride; int ioff;
    int nm, M;
    const cx<T>* twN; const cx<T>* twO;
};
template <typename T, int N, int O> __global__ void __launch_bounds__(k3Threads, 1) k3d_plane(Plane3<T> a) {
    constexpr int W = O + 1;                                       // odd pitch: both passes are conflict-free
    constexpr int EPT = (N * O + k3Threads - 1) / k3Threads;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twN = s + (size_t)N * W;
    cx<T>* twO = twN + N;
    __shared__ double red[8][k3Threads / 32];
    const int b = blockIdx.x / a.M, p = blockIdx.x - b * a.M;
    const int tid = flat_tid();
    stage(twN, a.twN, N);
    stage(twO, a.twO, O);
    T acc[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) acc[k] = T(0);
    for (int mi = 0; mi < a.nm; ++mi) {
        const cx<T>* __restrict__ yb = a.Y + (((size_t)(b * a.nm + mi) * a.M) + p) * N * O;
        for (int idx = tid; idx < N * O / 2; idx += k3Threads) {
            const int e = 2 * idx, n = e / O, o = e - n * O;
            const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(yb + e);
            s[n * W + o] = v.a; s[n * W + o + 1] = v.b;
        }
        __syncthreads();
        slab_fft_s<O, false, +1, W, 1, T>(s, N, twO);             // rows (along o): N lines
        slab_fft_s<N, false, +1, 1, W, T>(s, O, twN);             // columns (along n): O lines
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * k3Threads;
            if (idx < N * O) { const int n = idx / O, o = idx - n * O; const cx<T> v = s[n * W + o]; acc[k] += v.x * v.x + v.y * v.y; }
        }
        __syncthreads();
    }
    // U = sqrt(sum_m |.|^2)  (== the reference's nested sqrt(prev^2 + |x|^2)); voxel sums of U^q
    T part[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) part[q] = T(0);
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int idx = tid + k * k3Threads;
        if (idx < N * O) {
            const T m2 = acc[k];
            const T u = sqrt(m2);
            acc[k] = u;
            for (int q = 0; q < a.P; ++q) {
                const float pw = a.powers[q];
                part[q] += pw == 1.f ? u : pw == 2.f ? m2 : pw == 0.5f ? sqrt(u) : (u > T(0) ? pow(u, T(pw)) : (pw == 0.f ? T(1) : T(0)));
            }
        }
    }
    const int lane = tid & 31, warp = tid >> 5;
    for (int q = 0; q < a.P; ++q) {
        double v = (double)part[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[q][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
        for (int q = 0; q < a.P; ++q) {
            double v = lane < k3Threads / 32 ? red[q][lane] : 0.0;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(a.integ + (long long)b * a.istride + a.ioff + q, v);
        }
    }
    if (a.spec) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * k3Threads;
            if (idx < N * O) { const int n = idx / O, o = idx - n * O; s[n * W + o] = mk<T>(acc[k], T(0)); }
        }
        __syncthreads();
        slab_fft_s<N, true, -1, 1, W, T>(s, O, twN);              // columns first (transpose of the inverse order)
        slab_fft_s<O, true, -1, W, 1, T>(s, N, twO);
        cx<T>* ob = a.spec + ((size_t)b * a.M + p) * N * O;
        for (int idx = tid; idx < N * O / 2; idx += k3Threads) {
            const int e = 2 * idx, n = e / O, o = e - n * O;
            cxpair<T> v; v.a = s[n * W + o]; v.b = s[n * W + o + 1];
            *reinterpret_cast<cxpair<T>*>(ob + e) = v;
        }
    }
}

template <typename T> void (*kern3d_col_prod(int M))(ColProd3<T>);
template <typename T> void (*kern3d_plane(int N, int O))(Plane3<T>);
void kern3d_enable_smem();

}  // namespace sb
