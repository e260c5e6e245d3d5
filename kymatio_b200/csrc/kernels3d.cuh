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
//   * along M: slabs of 16 (n, o) columns, all M rows                          (k3d_col_prod, k3d_col_fwd)
//   * the (N, O/2) HALF plane of one m: a 2-D transform held entirely in shared memory (k3d_plane)
// The radix-2 stage that couples the two halves of the O axis is done in the M-axis kernels (their slab holds the
// columns o..o+7 and o+O/2..o+O/2+7), so a plane-pass CTA needs only 64 KB for a 128 x 128 plane and three of them
// share an SM, overlapping one CTA's loads with another's butterflies.
// Inverse transforms run decimation-in-frequency (natural in, scrambled out) and forward ones decimation-in-time
// (scrambled in, natural out), so the spatial field between them stays in scrambled positions along all three
// axes - harmless, because only order-agnostic operations (|.|^2 accumulation over m, sqrt, voxel sums) touch it.
//
//   k3d_col_prod   Y[b, mi] = DIF2_O( ifft_M( U_hat[b] * Psi[mi] ) ) / (MNO)    one launch per band (l, j): U_hat[b] is
//                  read once and kept in registers across the loop over the band's 2l+1 filters
//   k3d_plane      U = sqrt(sum_mi |ifft_{N,O/2}(Y[b, mi] half plane)|^2), held in registers across the m loop;
//                  integrals sum U^q accumulated (float64 atomics); parents: (U, 0) -> fft_{N,O/2} -> half spectrum plane
//   k3d_col_fwd    parents: DIT2_O then fft_M of the half spectrum planes -> natural-order U1_hat
#pragma once
#include "kernels1d.cuh"

namespace sb {

// threads per CTA / CTAs per SM the half-plane kernel is compiled for (a 128 x 64 half plane takes 65 KB: three per SM)
constexpr int plane_threads(int n_elems) { return n_elems >= 4096 ? 512 : n_elems >= 1024 ? 256 : 64; }
constexpr int plane_ctas(int n_elems) { return n_elems >= 8192 ? 2 : n_elems >= 4096 ? 2 : 4; }

// ------------------------------------------------------------------ M-axis inverse with the filter product
template <typename T> struct ColProd3 {
    const cx<T>* U;        // [B][M][N][O] natural-order spectrum
    const cx<T>* filt;     // [nm][M][N][O] complex filters of this band
    cx<T>* Y;              // [B*nm][M (scrambled)][N][O]: columns o < O/2 = even output samples, o >= O/2 = odd ones
    int B, nm, NO, O;
    T scale;               // 1 / (M N O)
    const cx<T>* twM; const cx<T>* twO;
};
// slab column l < 8 <-> o0 + l, column 8 + l <-> o0 + O/2 + l (same n); grid = (NO/16 column groups) x B, b fastest so
// that the CTAs sharing a filter slab run together
template <typename T, int M> __global__ void __launch_bounds__(k1Threads, 4) k3d_col_prod(ColProd3<T> a) {
    constexpr int LP = k1LP, CPT = (M * 4 + k1Threads - 1) / k1Threads;     // cells (row, column pair) per thread
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)M * LP;
    const int b = blockIdx.x % a.B, cg = blockIdx.x / a.B;
    const int OH = a.O >> 1;
    const int row = (cg * 8) / OH, o0 = cg * 8 - row * OH;               // (n, first o) of this column group
    const int cA = row * a.O + o0;                                        // offset of segment A inside an m-row
    const int tid = flat_tid(), nt = flat_nt();
    stage(tw, a.twM, M);
    const cx<T>* __restrict__ ub = a.U + (size_t)b * M * a.NO + cA;
    cxpair<T> uA[CPT], uB[CPT]; cx<T> w0[CPT], w1[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int idx = tid + c * k1Threads;
        if (idx < M * 4) {
            const int r = idx >> 2, l = 2 * (idx & 3);
            uA[c] = *reinterpret_cast<const cxpair<T>*>(ub + (size_t)r * a.NO + l);
            uB[c] = *reinterpret_cast<const cxpair<T>*>(ub + (size_t)r * a.NO + OH + l);
            w0[c] = ldg_cx(a.twO + o0 + l); w1[c] = ldg_cx(a.twO + o0 + l + 1);
        }
    }
    for (int mi = 0; mi < a.nm; ++mi) {
        const cx<T>* __restrict__ fb = a.filt + (size_t)mi * M * a.NO + cA;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int idx = tid + c * k1Threads;
            if (idx < M * 4) {
                const int r = idx >> 2, l = 2 * (idx & 3);
                const cxpair<T> fA = *reinterpret_cast<const cxpair<T>*>(fb + (size_t)r * a.NO + l);
                const cxpair<T> fB = *reinterpret_cast<const cxpair<T>*>(fb + (size_t)r * a.NO + OH + l);
                const cx<T> pa0 = cmul(uA[c].a, fA.a), pa1 = cmul(uA[c].b, fA.b);
                const cx<T> pb0 = cmul(uB[c].a, fB.a), pb1 = cmul(uB[c].b, fB.b);
                // radix-2 DIF stage along o (inverse sign): top = a + b, bottom = (a - b) * exp(+2 pi i o / O)
                s[r * LP + l] = scal(pa0 + pb0, a.scale);
                s[r * LP + l + 1] = scal(pa1 + pb1, a.scale);
                s[r * LP + 8 + l] = scal(cmulc(pa0 - pb0, w0[c]), a.scale);
                s[r * LP + 8 + l + 1] = scal(cmulc(pa1 - pb1, w1[c]), a.scale);
            }
        }
        __syncthreads();
        slab_fft_s<M, false, +1, 1, k1LP, T>(s, k1L, tw);
        cx<T>* yb = a.Y + ((size_t)(b * a.nm + mi) * M) * a.NO + cA;
        for (int idx = tid; idx < M * (k1L / 2); idx += nt) {
            const int p = idx / (k1L / 2), l = 2 * (idx - p * (k1L / 2));
            cxpair<T> o; o.a = s[p * LP + l]; o.b = s[p * LP + l + 1];
            *reinterpret_cast<cxpair<T>*>(yb + (size_t)p * a.NO + (l < 8 ? l : OH + l - 8)) = o;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ M-axis forward of the parents' half-plane spectra
template <typename T> struct ColFwd3 {
    const cx<T>* Z; cx<T>* out;    // [B][M (scrambled)][N][O] half-plane spectra -> natural-order spectrum (may alias)
    int B, NO, O;
    const cx<T>* twM; const cx<T>* twO;
    const int* posM;               // non-null: the M rows of Z are in NATURAL order (k3d_plane_real): row m is staged at posM[m]
};
template <typename T, int M> __global__ void __launch_bounds__(k1Threads, 4) k3d_col_fwd(ColFwd3<T> a) {
    constexpr int LP = k1LP;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)M * LP;
    const int ncol = a.NO / k1L;
    const int cg = blockIdx.x % ncol, b = blockIdx.x / ncol;
    const int OH = a.O >> 1;
    const int row = (cg * 8) / OH, o0 = cg * 8 - row * OH;
    const int cA = row * a.O + o0;
    const int tid = flat_tid(), nt = flat_nt();
    stage(tw, a.twM, M);
    const cx<T>* zb = a.Z + (size_t)b * M * a.NO + cA;
    cx<T>* ob = a.out + (size_t)b * M * a.NO + cA;
    for (int idx = tid; idx < M * 4; idx += nt) {
        const int p = idx >> 2, l = 2 * (idx & 3);
        const cxpair<T> e = *reinterpret_cast<const cxpair<T>*>(zb + (size_t)p * a.NO + l);
        const cxpair<T> d = *reinterpret_cast<const cxpair<T>*>(zb + (size_t)p * a.NO + OH + l);
        // radix-2 DIT stage along o (forward sign): X[f] = E[f] + w^f D[f], X[f + O/2] = E[f] - w^f D[f]
        const cx<T> t0 = cmul(d.a, ldg_cx(a.twO + o0 + l)), t1 = cmul(d.b, ldg_cx(a.twO + o0 + l + 1));
        const int sp = a.posM ? __ldg(a.posM + p) : p;
        s[sp * LP + l] = e.a + t0; s[sp * LP + l + 1] = e.b + t1;
        s[sp * LP + 8 + l] = e.a - t0; s[sp * LP + 8 + l + 1] = e.b - t1;
    }
    __syncthreads();
    slab_fft_s<M, true, -1, 1, k1LP, T>(s, k1L, tw);
    for (int idx = tid; idx < M * (k1L / 2); idx += nt) {
        const int f = idx / (k1L / 2), l = 2 * (idx - f * (k1L / 2));
        cxpair<T> o; o.a = s[f * LP + l]; o.b = s[f * LP + l + 1];
        *reinterpret_cast<cxpair<T>*>(ob + (size_t)f * a.NO + (l < 8 ? l : OH + l - 8)) = o;
    }
}

// Inverse DIF transform of LINES lines (same layout conventions as slab_fft_s) whose LAST pass does not store its
// outputs: |v|^2 of output j of the thread's n-th butterfly is added to acc[n*R + j].  The work-item -> thread mapping is
// the fixed one of slab_fft_s (item = tid + n*NT), so a thread owns the same spatial positions for every field it
// processes; acc_store_sqrt() later writes (sqrt(acc), 0) back to exactly those positions.  Everything is compile-time
// so that acc stays in registers.
template <int N> struct LastPass {
    static constexpr int NP = ct_plan1(N).npass;
    static constexpr int R = ct_plan1(N).radix[NP - 1];
    static_assert(ct_plan1(N).blen[NP - 1] == R && ct_is_pow2(R), "last pass must be an untwiddled power-of-two butterfly");
};
template <int N, int LINES, int NT> constexpr int acc_iters() { return ((N / LastPass<N>::R) * LINES + NT - 1) / NT; }
template <int N, int LINES, int NT> constexpr int acc_size() { return acc_iters<N, LINES, NT>() * LastPass<N>::R; }

template <int N, int SIGN, int LS, int ES, int LINES, int NT, typename T>
__device__ __forceinline__ void slab_fft_acc(cx<T>* s, const cx<T>* tw, T (&acc)[acc_size<N, LINES, NT>()]) {
    constexpr int NP = LastPass<N>::NP, R = LastPass<N>::R;
    const int tid = flat_tid();
    static_for<0, NP - 1>([&](auto p_) {
        constexpr int p = decltype(p_)::value;
        constexpr int r = ct_plan1(N).radix[p], m = ct_plan1(N).blen[p];
        constexpr int q = m / r, nbf = N / r, tws = N / m;
        constexpr int items = nbf * LINES;
        for (int it = tid; it < items; it += NT) {
            const int bf = it / LINES, line = it - bf * LINES;
            const int blk = bf / q, i = bf - blk * q;
            butterfly_s<r, false, SIGN, q, ES, false, T>(s + line * LS + (blk * m + i) * ES, i * tws, tw);
        }
        __syncthreads();
    });
    constexpr int items = (N / R) * LINES;
    static_for<0, acc_iters<N, LINES, NT>()>([&](auto n_) {
        constexpr int n = decltype(n_)::value;
        const int it = tid + n * NT;
        if (it < items) {
            const int bf = it / LINES, line = it - bf * LINES;
            const cx<T>* p0 = s + line * LS + bf * R * ES;
            cx<T> v[R];
            static_for<0, R>([&](auto k_) { constexpr int k = decltype(k_)::value; v[k] = p0[k * ES]; });
            dif_pow2<R, SIGN, T>(v);
            static_for<0, R>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                acc[n * R + k] += v[k].x * v[k].x + v[k].y * v[k].y;
            });
        }
    });
    __syncthreads();
}
// write (sqrt(acc), 0) to the positions slab_fft_acc accumulated from
template <int N, int LS, int ES, int LINES, int NT, typename T>
__device__ __forceinline__ void acc_store_sqrt(cx<T>* s, const T (&acc)[acc_size<N, LINES, NT>()]) {
    constexpr int R = LastPass<N>::R;
    constexpr int items = (N / R) * LINES;
    const int tid = flat_tid();
    static_for<0, acc_iters<N, LINES, NT>()>([&](auto n_) {
        constexpr int n = decltype(n_)::value;
        const int it = tid + n * NT;
        if (it < items) {
            const int bf = it / LINES, line = it - bf * LINES;
            cx<T>* p0 = s + line * LS + bf * R * ES;
            static_for<0, R>([&](auto k_) { constexpr int k = decltype(k_)::value; p0[k * ES] = mk<T>(sqrt(acc[n * R + k]), T(0)); });
        }
    });
}

// ------------------------------------------------------------------ plane pass: 2-D inverse, rotation modulus, integrals, 2-D forward
template <typename T> struct Plane3 {
    const cx<T>* Y;        // [B*nm][M][N][O]
    cx<T>* spec;           // parents: [B][M][N][O] half-plane spectra (M still scrambled); leaves: nullptr
    double* integ;         // integ[b*istride + ioff + p] += sum U^{q_p}
    const float* powers; int P;
    long long istride; int ioff;
    int nm, M;
    const cx<T>* twN; const cx<T>* twH;   // lengths N and O/2
};
// grid = B * M * 2: one CTA per HALF plane (columns [h*OH, (h+1)*OH) of every row of plane (b, p))
template <typename T, int N, int OH> __global__ void __launch_bounds__(plane_threads(N * OH), plane_ctas(N * OH)) k3d_plane(Plane3<T> a) {
    constexpr int W = OH + 1, O = 2 * OH;                          // odd pitch: both passes are conflict-free
    constexpr int TH = plane_threads(N * OH);
    constexpr int EPT = acc_size<N, OH, TH>();                     // accumulators per thread (outputs of its last-pass butterflies)
    constexpr int ACC_R = LastPass<N>::R, ACC_ITEMS = (N / ACC_R) * OH;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twN = s + (size_t)N * W;
    cx<T>* twH = twN + N;
    double* red = reinterpret_cast<double*>(twH + OH);             // [8][32]
    const int h = blockIdx.x & 1, bp = blockIdx.x >> 1;
    const int b = bp / a.M, p = bp - b * a.M;
    const int tid = flat_tid();
    stage(twN, a.twN, N);
    stage(twH, a.twH, OH);
    T acc[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) acc[k] = T(0);
    for (int mi = 0; mi < a.nm; ++mi) {
        const cx<T>* __restrict__ yb = a.Y + (((size_t)(b * a.nm + mi) * a.M) + p) * N * O + h * OH;
        constexpr int LU = (N * OH / 2) % (4 * TH) == 0 ? 4 : 1;   // loads in flight per thread
        for (int base = 0; base < N * OH / 2; base += LU * TH) {
            cxpair<T> v[LU];
#pragma unroll
            for (int c = 0; c < LU; ++c) {
                const int idx = base + c * TH + tid;
                if (idx < N * OH / 2) { const int e = 2 * idx, n = e / OH, o = e - n * OH; v[c] = *reinterpret_cast<const cxpair<T>*>(yb + (size_t)n * O + o); }
            }
#pragma unroll
            for (int c = 0; c < LU; ++c) {
                const int idx = base + c * TH + tid;
                if (idx < N * OH / 2) { const int e = 2 * idx, n = e / OH, o = e - n * OH; s[n * W + o] = v[c].a; s[n * W + o + 1] = v[c].b; }
            }
        }
        __syncthreads();
        slab_fft_s<OH, false, +1, W, 1, T>(s, N, twH);            // rows (along o): N lines
        slab_fft_acc<N, +1, 1, W, OH, TH, T>(s, twN, acc);        // columns (along n): |.|^2 of the last pass stays in registers
    }
    // U = sqrt(sum_m |.|^2)  (== the reference's nested sqrt(prev^2 + |x|^2)); voxel sums of U^q
    for (int q = 0; q < a.P; ++q) {
        const float pw = a.powers[q];
        T part = T(0);
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            if (tid + (k / ACC_R) * TH < ACC_ITEMS) {
                const T m2 = acc[k], u = sqrt(m2);
                part += pw == 1.f ? u : pw == 2.f ? m2 : pw == 0.5f ? sqrt(u) : (u > T(0) ? pow(u, T(pw)) : (pw == 0.f ? T(1) : T(0)));
            }
        }
        double v = (double)part;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[q * 32 + (tid >> 5)] = v;
    }
    __syncthreads();
    if (tid < 32) {
        for (int q = 0; q < a.P; ++q) {
            double v = tid < (TH + 31) / 32 ? red[q * 32 + tid] : 0.0;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (tid == 0) atomicAdd(a.integ + (long long)b * a.istride + a.ioff + q, v);
        }
    }
    if (a.spec) {
        acc_store_sqrt<N, 1, W, OH, TH, T>(s, acc);
        __syncthreads();
        slab_fft_s<N, true, -1, 1, W, T>(s, OH, twN);             // columns first (transpose of the inverse order)
        slab_fft_s<OH, true, -1, W, 1, T>(s, N, twH);
        cx<T>* ob = a.spec + ((size_t)b * a.M + p) * N * O + h * OH;
        for (int idx = tid; idx < N * OH / 2; idx += TH) {
            const int e = 2 * idx, n = e / OH, o = e - n * OH;
            cxpair<T> v; v.a = s[n * W + o]; v.b = s[n * W + o + 1];
            *reinterpret_cast<cxpair<T>*>(ob + (size_t)n * O + o) = v;
        }
    }
}

// ------------------------------------------------------------------ rfft of the input volume, plane half
// U0_hat = rfft(x) (core/scattering3d.py:24): half h of plane (b, m) holds the REAL samples x[b][m][n][2 o' + h], staged at
// the scrambled positions of (n, o') so that the forward DIT passes leave natural-order half spectra; k3d_col_fwd
// (posM != null) applies the radix-2 DIT stage along O and the transform along M.
template <typename T> struct PlaneReal3 {
    const T* x; cx<T>* spec;       // x: [B][M][N][O] real; spec: [B][M][N][O] half-plane spectra
    int M;
    const cx<T>* twN; const cx<T>* twH; const int* posN; const int* posH;
};
template <typename T, int N, int OH> __global__ void __launch_bounds__(plane_threads(N * OH), plane_ctas(N * OH)) k3d_plane_real(PlaneReal3<T> a) {
    constexpr int W = OH + 1, O = 2 * OH;
    constexpr int TH = plane_threads(N * OH);
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twN = s + (size_t)N * W;
    cx<T>* twH = twN + N;
    const int h = blockIdx.x & 1, bp = blockIdx.x >> 1;
    const int tid = flat_tid();
    stage(twN, a.twN, N);
    stage(twH, a.twH, OH);
    const T* __restrict__ xb = a.x + (size_t)bp * N * O + h;
    for (int idx = tid; idx < N * OH; idx += TH) {
        const int n = idx / OH, o = idx - n * OH;
        s[__ldg(a.posN + n) * W + __ldg(a.posH + o)] = mk<T>(xb[(size_t)n * O + 2 * o], T(0));
    }
    __syncthreads();
    slab_fft_s<N, true, -1, 1, W, T>(s, OH, twN);
    slab_fft_s<OH, true, -1, W, 1, T>(s, N, twH);
    cx<T>* ob = a.spec + (size_t)bp * N * O + h * OH;
    for (int idx = tid; idx < N * OH / 2; idx += TH) {
        const int e = 2 * idx, n = e / OH, o = e - n * OH;
        cxpair<T> v; v.a = s[n * W + o]; v.b = s[n * W + o + 1];
        *reinterpret_cast<cxpair<T>*>(ob + (size_t)n * O + o) = v;
    }
}

template <typename T> void (*kern3d_col_prod(int M))(ColProd3<T>);
template <typename T> void (*kern3d_col_fwd(int M))(ColFwd3<T>);
template <typename T> void (*kern3d_plane(int N, int O))(Plane3<T>);
template <typename T> void (*kern3d_plane_real(int N, int O))(PlaneReal3<T>);
void kern3d_enable_smem();

}  // namespace sb
