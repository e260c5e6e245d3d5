// kernels1d.cuh - fused 1-D scattering kernels (float32).
//
// Reference semantics being fused (paths relative to the kymatio tree):
//   cascade            kymatio/scattering1d/core/scattering1d.py:40-107
//   cdgmm              kymatio/backend/torch_backend.py:148-219   (real filter branch :205-206)
//   subsample_fourier  kymatio/scattering1d/backend/torch_backend.py:19-48
//   ifft / rfft / irfft kymatio/scattering1d/backend/torch_backend.py:109-141
//   modulus            kymatio/backend/torch_backend.py:138-141
//   unpad              kymatio/scattering1d/backend/torch_backend.py:85-106
//
// One scattering path  U -> |ifft(periodise_k(U_hat * psi))| -> fft  of length N = NA * NB (powers of two,
// N up to 2^18: a line does not fit one CTA) runs as a four-step transform whose two halves share the SAME
// shared-memory slab around the modulus, so a path is three passes over HBM and needs no permutation pass:
//
//   f = NB*f1 + f2 (Fourier index, the natural contiguous layout is the matrix [NA][NB]),  t = t1 + NA*t2 (time)
//
//   k1d_col_prod  (P1)  16 adjacent columns f2, all f1: load parent*filter with the k aliases folded (only rows
//                       that meet the filter's support are touched), inverse DIF over f1 -> t1 (scrambled row
//                       order p), twiddle exp(+2 pi i f2 t1 / N), store Y[p][f2];
//   k1d_row_mod   (P2)  16 rows p, all f2: inverse DIF over f2 -> t2 (scrambled), MODULUS fused into the last
//                       butterfly pass, forward DIT over t2 -> f2' (natural), twiddle exp(-2 pi i t1 f2' / N);
//                         parents: store Z[p][f2'] in place;
//                         leaves : only the Fc lowest bins of the spectrum are needed by the low-pass, so the
//                                  column DFT is pruned to  part[cta][f] = sum_rows w_N^{t1 f} R[p][f mod NB];
//   k1d_col_fwd   (P3)  16 adjacent columns f2', all p: forward DIT over t1 -> f1' (natural), store the
//                       natural-order spectrum U1_hat[NB*f1' + f2'];
//   k1d_finish          low-pass tail  cdgmm(phi) -> subsample_fourier -> irfft -> unpad  on the Fc lowest bins
//                       (Hermitian symmetry of the spectrum of a real field supplies the negative ones):
//                       fold onto M bins, inverse DIF of length M in shared memory (16 paths per CTA),
//                       write samples [i0, i0+W) of the real part into the channel row of the output.
#pragma once
#include "kernels2d.cuh"

namespace sb {

constexpr int k1L = 16;            // lines per CTA
constexpr int k1LP = k1L | 1;      // odd shared-memory pitch
constexpr int k1Threads = 256;

// exp(-2 pi i j / N) = hi[j >> lb] * lo[j & (2^lb - 1)]
template <typename T> struct TwN { const cx<T>* hi; const cx<T>* lo; int lb, nhi; };
template <typename T> __device__ __forceinline__ cx<T> twn(const cx<T>* hi, const cx<T>* lo, int lb, int j) {
    return cmul(hi[j >> lb], lo[j & ((1 << lb) - 1)]);
}

// ------------------------------------------------------------------ P1: product + periodise + column inverse
template <typename T> struct ColProd1 {
    const cx<T>* parent; long long ps_b, ps_i;   // parent spectrum of path g = b*NI + i at parent + b*ps_b + i*ps_i
    const T* const* filt;                        // [NI] real filters on the parent grid (length Npar)
    const int2* supp;                            // [NI] circular support (start, len) of each filter
    cx<T>* Y;                                    // [G][NA][NB]
    int NI, Npar, k, NB;
    T scale;                                     // 1 / (N k)
    const cx<T>* twA; const int* invA;           // length-NA twiddles; invA[p] = t1 held at scrambled row p
    TwN<T> w;
};
template <typename T, int NA> __global__ void __launch_bounds__(k1Threads, 2) k1d_col_prod(ColProd1<T> a) {
    constexpr int LP = k1LP;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twA = s + (size_t)NA * LP;
    cx<T>* hi = twA + NA;
    cx<T>* lo = hi + a.w.nhi;
    int* invA = reinterpret_cast<int*>(lo + (1 << a.w.lb));
    const int g = blockIdx.x, c0 = blockIdx.y * k1L;
    const int i = g % a.NI, b = g / a.NI;
    const int tid = flat_tid(), nt = flat_nt();
    stage(twA, a.twA, NA);
    stage(hi, a.w.hi, a.w.nhi);
    stage(lo, a.w.lo, 1 << a.w.lb);
    stage(invA, a.invA, NA);
    const cx<T>* __restrict__ pb = a.parent + (long long)b * a.ps_b + (long long)i * a.ps_i;
    const T* __restrict__ fb = a.filt[i];
    const int2 sp = a.supp[i];
    const int NB = a.NB, KNA = a.k * NA, Npar = a.Npar;
    const int Rb = sp.x / NB;                                         // first parent row meeting the support
    const int nr = min(KNA, (sp.x - Rb * NB + sp.y + NB - 1) / NB);   // number of such rows (circular)
    for (int idx = tid; idx < NA * (k1L / 2); idx += nt) {
        const int f1 = idx / (k1L / 2), l = 2 * (idx - f1 * (k1L / 2));
        T ax0 = T(0), ay0 = T(0), ax1 = T(0), ay1 = T(0);
        for (int d = (f1 - Rb) & (NA - 1); d < nr; d += NA) {         // aliases f1 + a*NA inside the support rows
            int R = Rb + d;
            if (R >= KNA) R -= KNA;
            const int off = R * NB + c0 + l;
            int rel = off - sp.x;
            if (rel < 0) rel += Npar;
            if ((rel < sp.y) | (rel == Npar - 1)) {
                const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(pb + off);
                const repair<T> f = *reinterpret_cast<const repair<T>*>(fb + off);
                ax0 += v.a.x * f.a; ay0 += v.a.y * f.a;
                ax1 += v.b.x * f.b; ay1 += v.b.y * f.b;
            }
        }
        s[f1 * LP + l] = mk<T>(ax0 * a.scale, ay0 * a.scale);
        s[f1 * LP + l + 1] = mk<T>(ax1 * a.scale, ay1 * a.scale);
    }
    __syncthreads();
    slab_fft_s<NA, false, +1, 1, k1LP, T>(s, k1L, twA);               // inverse DIF over f1: row p holds t1 = invA[p]
    cx<T>* yb = a.Y + (size_t)g * NA * NB + c0;
    for (int idx = tid; idx < NA * (k1L / 2); idx += nt) {
        const int p = idx / (k1L / 2), l = 2 * (idx - p * (k1L / 2));
        const int t1 = invA[p], f2 = c0 + l;
        cxpair<T> o;
        o.a = cmulc(s[p * LP + l], twn(hi, lo, a.w.lb, f2 * t1));
        o.b = cmulc(s[p * LP + l + 1], twn(hi, lo, a.w.lb, (f2 + 1) * t1));
        *reinterpret_cast<cxpair<T>*>(yb + (size_t)p * NB + l) = o;
    }
}

// ------------------------------------------------------------------ P2: row inverse, modulus, row forward
template <typename T> struct RowMod1 {
    cx<T>* Y;                                    // [G][NA][NB] in; parents: Z out (in place)
    int NA, N;
    const cx<T>* twB; const int* invA;
    TwN<T> w;
    cx<T>* part; int Fc;                         // leaves: part[(g*nparts + cta)][Fc]
};
template <typename T, int NB, bool LEAF> __global__ void __launch_bounds__(k1Threads, 2) k1d_row_mod(RowMod1<T> a) {
    constexpr int LP = k1LP;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twB = s + (size_t)NB * LP;
    cx<T>* hi = twB + NB;
    cx<T>* lo = hi + a.w.nhi;
    int* t1s = reinterpret_cast<int*>(lo + (1 << a.w.lb));
    const int g = blockIdx.x, p0 = blockIdx.y * k1L;
    const int nl = min(k1L, a.NA - p0);
    const int tid = flat_tid(), nt = flat_nt();
    stage(twB, a.twB, NB);
    stage(hi, a.w.hi, a.w.nhi);
    stage(lo, a.w.lo, 1 << a.w.lb);
    if (tid < k1L) t1s[tid] = tid < nl ? a.invA[p0 + tid] : 0;
    cx<T>* yb = a.Y + ((size_t)g * a.NA + p0) * NB;
    constexpr int half = NB / 2;
    for (int idx = tid; idx < nl * half; idx += nt) {
        const int l = idx / half, e = 2 * (idx - l * half);
        const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(yb + (size_t)l * NB + e);
        s[e * LP + l] = v.a; s[(e + 1) * LP + l] = v.b;
    }
    __syncthreads();
    slab_fft_s<NB, false, +1, 1, k1LP, T, true>(s, nl, twB);          // inverse DIF + modulus: (|u|, 0), scrambled t2
    slab_fft_s<NB, true, -1, 1, k1LP, T>(s, nl, twB);                 // forward DIT: natural f2'
    if constexpr (!LEAF) {
        for (int idx = tid; idx < nl * half; idx += nt) {
            const int l = idx / half, e = 2 * (idx - l * half);
            const int t1 = t1s[l];
            cxpair<T> o;
            o.a = cmul(s[e * LP + l], twn(hi, lo, a.w.lb, t1 * e));
            o.b = cmul(s[(e + 1) * LP + l], twn(hi, lo, a.w.lb, t1 * (e + 1)));
            *reinterpret_cast<cxpair<T>*>(yb + (size_t)l * NB + e) = o;
        }
    } else {
        cx<T>* pb = a.part + ((size_t)g * gridDim.y + blockIdx.y) * a.Fc;
        const int maskN = a.N - 1;
        for (int f = tid; f < a.Fc; f += nt) {
            const int e = f & (NB - 1);
            T ax = T(0), ay = T(0);
            for (int l = 0; l < nl; ++l) {
                const cx<T> wv = twn(hi, lo, a.w.lb, (t1s[l] * f) & maskN);
                const cx<T> v = s[e * LP + l];
                ax += v.x * wv.x - v.y * wv.y; ay += v.x * wv.y + v.y * wv.x;
            }
            pb[f] = mk<T>(ax, ay);
        }
    }
}

// ------------------------------------------------------------------ P3: column forward, natural-order spectrum out
template <typename T> struct ColFwd1 {
    const cx<T>* Z; cx<T>* out;                  // [G][NA][NB] both; out = natural-order spectrum
    int NB;
    const cx<T>* twA;
};
template <typename T, int NA> __global__ void __launch_bounds__(k1Threads, 2) k1d_col_fwd(ColFwd1<T> a) {
    constexpr int LP = k1LP;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twA = s + (size_t)NA * LP;
    const int g = blockIdx.x, c0 = blockIdx.y * k1L;
    const int tid = flat_tid(), nt = flat_nt();
    stage(twA, a.twA, NA);
    const cx<T>* zb = a.Z + (size_t)g * NA * a.NB + c0;
    cx<T>* ob = a.out + (size_t)g * NA * a.NB + c0;
    for (int idx = tid; idx < NA * (k1L / 2); idx += nt) {
        const int p = idx / (k1L / 2), l = 2 * (idx - p * (k1L / 2));
        const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(zb + (size_t)p * a.NB + l);
        s[p * LP + l] = v.a; s[p * LP + l + 1] = v.b;
    }
    __syncthreads();
    slab_fft_s<NA, true, -1, 1, k1LP, T>(s, k1L, twA);                // forward DIT over t1 (scrambled in, natural out)
    for (int idx = tid; idx < NA * (k1L / 2); idx += nt) {
        const int f1 = idx / (k1L / 2), l = 2 * (idx - f1 * (k1L / 2));
        cxpair<T> o; o.a = s[f1 * LP + l]; o.b = s[f1 * LP + l + 1];
        *reinterpret_cast<cxpair<T>*>(ob + (size_t)f1 * a.NB + l) = o;
    }
}

// ------------------------------------------------------------------ low-pass tail on the lowest Fc bins
template <typename T> struct Finish1 {
    const cx<T>* src; long long ss_g, ss_part; int nparts;   // X[f] = sum_part src[g*ss_g + part*ss_part + f], f < Fc
    const T* phi;                                            // real low-pass on the length-N grid
    int N, Fc;
    T scale;                                                 // 1 / N
    T* out; long long os_b; const int* chan; int NI;         // out[b*os_b + chan[i]*W + n - i0]
    int G, i0, W;
    const cx<T>* twM; const int* posM;
};
template <typename T, int M> __global__ void __launch_bounds__(k1Threads, 2) k1d_finish(Finish1<T> a) {
    constexpr int LP = k1LP;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twM = s + (size_t)M * LP;
    int* posM = reinterpret_cast<int*>(twM + M);
    const int g0 = blockIdx.x * k1L;
    const int nl = min(k1L, a.G - g0);
    const int tid = flat_tid(), nt = flat_nt();
    stage(twM, a.twM, M);
    stage(posM, a.posM, M);
    const int N = a.N, Fc = a.Fc, nyq = N >> 1;
    for (int idx = tid; idx < nl * M; idx += nt) {
        const int l = idx / M, u = idx - l * M;
        const cx<T>* __restrict__ xb = a.src + (long long)(g0 + l) * a.ss_g;
        T ax = T(0), ay = T(0);
        for (int f = u; f < Fc; f += M) {                    // bins f = u + aM on the non-negative side
            T vx = T(0), vy = T(0);
            for (int q = 0; q < a.nparts; ++q) { const cx<T> v = xb[q * a.ss_part + f]; vx += v.x; vy += v.y; }
            const T ph = a.phi[f];
            ax += vx * ph; ay += vy * ph;
        }
        for (int f = (M - u) & (M - 1); f < Fc; f += M) {    // bins N - f == u (mod M): conj(X[f]) * phi[N - f]
            if (f == 0 || f == nyq) continue;
            T vx = T(0), vy = T(0);
            for (int q = 0; q < a.nparts; ++q) { const cx<T> v = xb[q * a.ss_part + f]; vx += v.x; vy += v.y; }
            const T ph = a.phi[N - f];
            ax += vx * ph; ay -= vy * ph;
        }
        s[u * LP + l] = mk<T>(ax * a.scale, ay * a.scale);
    }
    __syncthreads();
    slab_fft_s<M, false, +1, 1, k1LP, T>(s, nl, twM);                 // inverse DIF: sample n at row posM[n]
    for (int idx = tid; idx < nl * a.W; idx += nt) {
        const int l = idx / a.W, n = idx - l * a.W;
        const int g = g0 + l, b = g / a.NI, i = g - b * a.NI;
        a.out[(long long)b * a.os_b + (long long)a.chan[i] * a.W + n] = s[posM[a.i0 + n] * LP + l].x;
    }
}

// kernel tables (instances in scat1d_inst.cu)
template <typename T> struct Kern1d {
    void (*col_prod)(ColProd1<T>);
    void (*col_fwd)(ColFwd1<T>);
};
template <typename T> struct KernRow1d {
    void (*parent)(RowMod1<T>);
    void (*leaf)(RowMod1<T>);
};
template <typename T> Kern1d<T> kern1d_cols(int NA);                 // nullptr entries when NA is not compiled
template <typename T> KernRow1d<T> kern1d_rows(int NB);
template <typename T> void (*kern1d_finish(int M))(Finish1<T>);
void kern1d_enable_smem();

}  // namespace sb
