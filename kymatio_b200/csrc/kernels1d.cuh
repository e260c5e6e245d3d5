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
//                       (Hermitian symmetry of the spectrum of a real field supplies the negative ones): ONE launch per
//                       batch chunk for every path (segment table), fold onto M bins, inverse DIF of length M in shared
//                       memory (8 paths per CTA), real part of samples [i0, i0+W) into the channel row of the output;
//   k1d_tile            transforms up to 8192 points: the WHOLE path (product, four-step inverse, modulus, four-step
//                       forward) in one CTA's shared memory - no Y round trip, no pruned DFT;
//   k1d_row_real        U0_hat = rfft(U_0) from the REAL padded signal (rows staged at scrambled positions), finished by
//                       k1d_col_fwd with scrambled row staging.
// All slab kernels are compiled for four CTAs per SM (<= 64 registers); the two-level twiddle table exp(-2 pi i j / N) =
// hi[j >> lb] * lo[j & mask] is read through the read-only path (L1 resident).
#pragma once
#include "kernels2d.cuh"

namespace sb {

constexpr int k1L = 16;            // lines per CTA
constexpr int k1LP = k1L | 1;      // odd shared-memory pitch
constexpr int k1Threads = 256;

// 8-byte predicated read-only global load (zero when the predicate is false)
template <typename V> __device__ __forceinline__ V ld8_pred(const void* ptr, bool pred) {
    static_assert(sizeof(V) == 8, "ld8_pred loads 8 bytes");
    union { V v; uint2 q; } u;
    u.q = make_uint2(0u, 0u);
    asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p ld.global.nc.v2.u32 {%0, %1}, [%2];\n\t}"
        : "+r"(u.q.x), "+r"(u.q.y)
        : "l"(ptr), "r"((int)pred));
    return u.v;
}

// exp(-2 pi i j / N) = hi[j >> lb] * lo[j & (2^lb - 1)]
template <typename T> struct TwN { const cx<T>* hi; const cx<T>* lo; int lb, nhi; };
template <typename T> __device__ __forceinline__ cx<T> ldg_cx(const cx<T>* p) {
    if constexpr (sizeof(T) == 4) { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); return mk<T>(v.x, v.y); }
    else { const double2 v = __ldg(reinterpret_cast<const double2*>(p)); return mk<T>(v.x, v.y); }
}
// the two small tables are read through the read-only path (L1-resident), not staged per CTA
template <typename T> __device__ __forceinline__ cx<T> twn(const cx<T>* hi, const cx<T>* lo, int lb, int j) {
    return cmul(ldg_cx(hi + (j >> lb)), ldg_cx(lo + (j & ((1 << lb) - 1))));
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
template <typename T, int NA> __global__ void __launch_bounds__(k1Threads, 4) k1d_col_prod(ColProd1<T> a) {
    constexpr int LP = k1LP;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twA = s + (size_t)NA * LP;
    const cx<T>* hi = a.w.hi; const cx<T>* lo = a.w.lo;
    const int* __restrict__ invA = a.invA;
    const int ncol = a.NB / k1L;
    const int g = blockIdx.x / ncol, c0 = (blockIdx.x - g * ncol) * k1L;   // column group fastest: neighbours share DRAM rows
    const int i = g % a.NI, b = g / a.NI;
    const int tid = flat_tid(), nt = flat_nt();
    stage(twA, a.twA, NA);
    const cx<T>* __restrict__ pb = a.parent + (long long)b * a.ps_b + (long long)i * a.ps_i;
    const T* __restrict__ fb = a.filt[i];
    const int2 sp = a.supp[i];
    const int NB = a.NB, KNA = a.k * NA, Npar = a.Npar;
    const int Rb = sp.x / NB;                                         // first parent row meeting the support
    const int nr = min(KNA, (sp.x - Rb * NB + sp.y + NB - 1) / NB);   // number of such rows (circular)
    const int amax = (nr + NA - 1) / NA;                              // aliases f1 + a*NA that can meet the support
    constexpr int CELLS = NA * (k1L / 2), U = 4;
    for (int base = 0; base < CELLS; base += U * k1Threads) {
        T acc[U][4];
#pragma unroll
        for (int c = 0; c < U; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = T(0);
        for (int ai = 0; ai < amax; ++ai) {
            cxpair<T> v[U]; repair<T> f[U];
#pragma unroll
            for (int c = 0; c < U; ++c) {                             // U independent predicated loads in flight
                const int idx = base + c * k1Threads + tid;
                const int f1 = idx / (k1L / 2), l = 2 * (idx - f1 * (k1L / 2));
                const int d = ((f1 - Rb) & (NA - 1)) + ai * NA;
                int R = Rb + d;
                if (R >= KNA) R -= KNA;
                const int off = R * NB + c0 + l;
                int rel = off - sp.x;
                if (rel < 0) rel += Npar;
                const bool in = (idx < CELLS) & (d < nr) & ((rel < sp.y) | (rel == Npar - 1));
                v[c] = ld_pred<cxpair<T>>(pb + off, in);
                f[c] = ld8_pred<repair<T>>(fb + off, in);
            }
#pragma unroll
            for (int c = 0; c < U; ++c) {
                acc[c][0] += v[c].a.x * f[c].a; acc[c][1] += v[c].a.y * f[c].a;
                acc[c][2] += v[c].b.x * f[c].b; acc[c][3] += v[c].b.y * f[c].b;
            }
        }
#pragma unroll
        for (int c = 0; c < U; ++c) {
            const int idx = base + c * k1Threads + tid;
            if (idx < CELLS) {
                const int f1 = idx / (k1L / 2), l = 2 * (idx - f1 * (k1L / 2));
                s[f1 * LP + l] = mk<T>(acc[c][0] * a.scale, acc[c][1] * a.scale);
                s[f1 * LP + l + 1] = mk<T>(acc[c][2] * a.scale, acc[c][3] * a.scale);
            }
        }
    }
    __syncthreads();
    slab_fft_s<NA, false, +1, 1, k1LP, T>(s, k1L, twA);               // inverse DIF over f1: row p holds t1 = invA[p]
    cx<T>* yb = a.Y + (size_t)g * NA * NB + c0;
    for (int idx = tid; idx < NA * (k1L / 2); idx += nt) {
        const int p = idx / (k1L / 2), l = 2 * (idx - p * (k1L / 2));
        const int t1 = __ldg(invA + p), f2 = c0 + l;
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
    // T = 0 (unaveraged output, core/scattering1d.py:75-76,104-105): the modulus field itself is an output.  mod != null:
    // |u| is stored in NATURAL time order t = t1 + NA*t2 at mod[g*N + t]; the CTA then takes the rows of 16 CONSECUTIVE t1
    // (row of t1 = posA[t1]) so that its stores are 64-byte segments; skip_fwd: a leaf - nothing else to do.
    T* mod; const int* posA; const int* posB; int skip_fwd;
};
// Shared-memory layout: line-major, line l at s + l*(NB+1) (odd pitch): both the staging (lanes along a line) and
// the butterflies (lanes across lines) are bank-conflict free.
template <typename T, int NB, bool LEAF> __global__ void __launch_bounds__(k1Threads, NB >= 512 ? 3 : 4) k1d_row_mod(RowMod1<T> a) {
    constexpr int LS = NB + 1;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twB = s + (size_t)k1L * LS;
    int* t1s = reinterpret_cast<int*>(twB + NB);
    int* lk = t1s + k1L;                                              // leaves: row holding the k-th term of the progression
    const cx<T>* hi = a.w.hi; const cx<T>* lo = a.w.lo;
    const int g = blockIdx.x, p0 = blockIdx.y * k1L;
    const int nl = min(k1L, a.NA - p0);
    const int tid = flat_tid(), nt = flat_nt();
    stage(twB, a.twB, NB);
    const bool gather = !LEAF && a.mod != nullptr;                    // rows of consecutive t1 instead of consecutive rows
    if (tid < k1L) t1s[tid] = tid < nl ? (gather ? p0 + tid : a.invA[p0 + tid]) : 0;
    cx<T>* yb = a.Y + ((size_t)g * a.NA + p0) * NB;
    cx<T>* yg = a.Y + (size_t)g * a.NA * NB;
    for (int idx = tid; idx < nl * NB; idx += nt) {
        const int l = idx / NB, e = idx - l * NB;
        s[l * LS + e] = gather ? yg[(size_t)__ldg(a.posA + p0 + l) * NB + e] : yb[idx];
    }
    __syncthreads();
    // the rows of one CTA hold t1 = base + stride*k, k < nl (DIF digit structure of the NA-point column transform; checked
    // on the host when the tables are built): the leaf's pruned column DFT then needs one twiddle and a power ladder per bin
    int base = t1s[0];
    for (int l = 1; l < nl; ++l) base = min(base, t1s[l]);
    const int stride = max(1, a.NA / k1L);
    if (LEAF && tid < nl) lk[(t1s[tid] - base) / stride] = tid;
    slab_fft_s<NB, false, +1, LS, 1, T, true>(s, nl, twB);            // inverse DIF + modulus: (|u|, 0), scrambled t2
    if constexpr (!LEAF) {
        if (gather) {
            T* mb = a.mod + (size_t)g * a.N + p0;
            for (int idx = tid; idx < NB * k1L; idx += nt) {
                const int t2 = idx / k1L, l = idx - t2 * k1L;
                if (l < nl) mb[(size_t)t2 * a.NA + l] = s[l * LS + __ldg(a.posB + t2)].x;
            }
            if (a.skip_fwd) return;
            __syncthreads();
        }
    }
    slab_fft_s<NB, true, -1, LS, 1, T>(s, nl, twB);                   // forward DIT: natural f2'
    if constexpr (!LEAF) {
        for (int idx = tid; idx < nl * NB; idx += nt) {
            const int l = idx / NB, e = idx - l * NB;
            const cx<T> v = cmul(s[l * LS + e], twn(hi, lo, a.w.lb, t1s[l] * e));
            if (gather) yg[(size_t)__ldg(a.posA + p0 + l) * NB + e] = v; else yb[idx] = v;
        }
    } else {
        cx<T>* pb = a.part + ((size_t)g * gridDim.y + blockIdx.y) * a.Fc;
        const int maskN = a.N - 1;
        for (int f = tid; f < a.Fc; f += nt) {
            // X[f] += w_N^{base f} sum_k (w_N^{stride f})^k R[row_k][f mod NB]
            const int e = f & (NB - 1);
            const cx<T> z1 = twn(hi, lo, a.w.lb, (stride * f) & maskN);
            const cx<T> z2 = cmul(z1, z1), z4 = cmul(z2, z2), z8 = cmul(z4, z4);
            T ax = T(0), ay = T(0);
            static_for<0, k1L>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                if (k < nl) {
                    cx<T> pk = mk<T>(T(1), T(0));
                    if constexpr (k & 1) pk = z1;
                    if constexpr (k & 2) pk = (k & 1) ? cmul(pk, z2) : z2;
                    if constexpr (k & 4) pk = (k & 3) ? cmul(pk, z4) : z4;
                    if constexpr (k & 8) pk = (k & 7) ? cmul(pk, z8) : z8;
                    const cx<T> v = s[lk[k] * LS + e];
                    ax += v.x * pk.x - v.y * pk.y; ay += v.x * pk.y + v.y * pk.x;
                }
            });
            pb[f] = cmul(mk<T>(ax, ay), twn(hi, lo, a.w.lb, (base * f) & maskN));
        }
    }
}

// ------------------------------------------------------------------ rfft of the padded input, first half
// U0_hat = rfft(U_0) (core/scattering1d.py:41): rows t1 (natural order) of the real padded signal x[t1 + NA*t2] are staged
// at the scrambled positions of t2, transformed forward (DIT, natural f2' out) and twiddled; k1d_col_fwd (posA != null)
// finishes the transform.  The real -> complex copy of the reference (torch_backend.py:109-113) never touches HBM.
template <typename T> struct RowReal1 {
    const T* x; cx<T>* Z;                        // x: [G][N] real; Z: [G][NA][NB]
    int NA;
    const cx<T>* twB; const int* posB;
    TwN<T> w;
};
template <typename T, int NB> __global__ void __launch_bounds__(k1Threads, NB >= 512 ? 3 : 4) k1d_row_real(RowReal1<T> a) {
    constexpr int LS = NB + 1;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twB = s + (size_t)k1L * LS;
    const cx<T>* hi = a.w.hi; const cx<T>* lo = a.w.lo;
    const int g = blockIdx.x, p0 = blockIdx.y * k1L;
    const int nl = min(k1L, a.NA - p0);
    const int tid = flat_tid(), nt = flat_nt();
    stage(twB, a.twB, NB);
    const T* __restrict__ xb = a.x + (size_t)g * a.NA * NB + p0;
    for (int idx = tid; idx < NB * k1L; idx += nt) {
        const int t2 = idx / k1L, l = idx - t2 * k1L;                 // 16 consecutive samples per t2: 64-byte segments
        if (l < nl) s[l * LS + __ldg(a.posB + t2)] = mk<T>(xb[(size_t)t2 * a.NA + l], T(0));
    }
    __syncthreads();
    slab_fft_s<NB, true, -1, LS, 1, T>(s, nl, twB);                   // forward DIT over t2: natural f2'
    cx<T>* zb = a.Z + ((size_t)g * a.NA + p0) * NB;
    for (int idx = tid; idx < nl * NB; idx += nt) {
        const int l = idx / NB, e = idx - l * NB;
        zb[idx] = cmul(s[l * LS + e], twn(hi, lo, a.w.lb, (p0 + l) * e));
    }
}

// ------------------------------------------------------------------ P3: column forward, natural-order spectrum out
template <typename T> struct ColFwd1 {
    const cx<T>* Z; cx<T>* out;                  // [G][NA][NB] both; out = natural-order spectrum
    int NB;
    const cx<T>* twA;
    const int* posA;                             // non-null: the rows of Z are in NATURAL t1 order (first transform of a real
                                                 // signal, k1d_row_real): row t1 is staged at its scrambled position posA[t1]
};
template <typename T, int NA> __global__ void __launch_bounds__(k1Threads, 3) k1d_col_fwd(ColFwd1<T> a) {
    constexpr int LP = k1LP;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twA = s + (size_t)NA * LP;
    const int ncol = a.NB / k1L;
    const int g = blockIdx.x / ncol, c0 = (blockIdx.x - g * ncol) * k1L;
    const int tid = flat_tid(), nt = flat_nt();
    stage(twA, a.twA, NA);
    const cx<T>* zb = a.Z + (size_t)g * NA * a.NB + c0;
    cx<T>* ob = a.out + (size_t)g * NA * a.NB + c0;
    for (int idx = tid; idx < NA * (k1L / 2); idx += nt) {
        const int p = idx / (k1L / 2), l = 2 * (idx - p * (k1L / 2));
        const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(zb + (size_t)p * a.NB + l);
        const int sp = a.posA ? __ldg(a.posA + p) : p;
        s[sp * LP + l] = v.a; s[sp * LP + l + 1] = v.b;
    }
    __syncthreads();
    slab_fft_s<NA, true, -1, 1, k1LP, T>(s, k1L, twA);                // forward DIT over t1 (scrambled in, natural out)
    for (int idx = tid; idx < NA * (k1L / 2); idx += nt) {
        const int f1 = idx / (k1L / 2), l = 2 * (idx - f1 * (k1L / 2));
        cxpair<T> o; o.a = s[f1 * LP + l]; o.b = s[f1 * LP + l + 1];
        *reinterpret_cast<cxpair<T>*>(ob + (size_t)f1 * a.NB + l) = o;
    }
}

// ------------------------------------------------------------------ whole path in one CTA (N <= 8192)
// Short paths fit one CTA's shared memory as the [NA][NB] matrix (pitch NB+1): product + periodise, four-step inverse
// (columns, twiddle, rows with the modulus fused), four-step forward (rows, twiddle, columns) and the natural-order
// spectrum is read straight out of shared memory - no Y round trip through HBM and no pruned column DFT.
template <typename T> struct Tile1 {
    const cx<T>* parent; long long ps_b, ps_i;
    const T* const* filt; const int2* supp;
    cx<T>* spec;                                 // parents: [G][N] natural-order spectrum; else nullptr
    cx<T>* part; int Fc;                         // leaves: part[g][Fc] (the Fc lowest bins); else nullptr
    int NI, Npar, k;
    T scale;
    const cx<T>* twA; const cx<T>* twB; const int* invA;
    TwN<T> w;
    T* mod; const int* posA; const int* posB;    // T = 0: |u| in natural time order at mod[g*N + t] (see RowMod1)
};
template <typename T, int NA, int NB> __global__ void __launch_bounds__(k1Threads, 4) k1d_tile(Tile1<T> a) {
    constexpr int W = NB + 1, N = NA * NB;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twA = s + (size_t)NA * W;
    cx<T>* twB = twA + NA;
    const cx<T>* hi = a.w.hi; const cx<T>* lo = a.w.lo;
    const int* __restrict__ invA = a.invA;
    const int g = blockIdx.x;
    const int i = g % a.NI, b = g / a.NI;
    const int tid = flat_tid(), nt = flat_nt();
    stage(twA, a.twA, NA);
    stage(twB, a.twB, NB);
    const cx<T>* __restrict__ pb = a.parent + (long long)b * a.ps_b + (long long)i * a.ps_i;
    const T* __restrict__ fb = a.filt[i];
    const int2 sp = a.supp[i];
    const int KNA = a.k * NA, Npar = a.Npar;
    const int Rb = sp.x / NB;
    const int nr = min(KNA, (sp.x - Rb * NB + sp.y + NB - 1) / NB);
    const int amax = (nr + NA - 1) / NA;
    constexpr int CELLS = NA * (NB / 2), U = 4;
    for (int base = 0; base < CELLS; base += U * k1Threads) {
        T acc[U][4];
#pragma unroll
        for (int c = 0; c < U; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = T(0);
        for (int ai = 0; ai < amax; ++ai) {
            cxpair<T> v[U]; repair<T> f[U];
#pragma unroll
            for (int c = 0; c < U; ++c) {
                const int idx = base + c * k1Threads + tid;
                const int f1 = idx / (NB / 2), l = 2 * (idx - f1 * (NB / 2));
                const int d = ((f1 - Rb) & (NA - 1)) + ai * NA;
                int R = Rb + d;
                if (R >= KNA) R -= KNA;
                const int off = R * NB + l;
                int rel = off - sp.x;
                if (rel < 0) rel += Npar;
                const bool in = (idx < CELLS) & (d < nr) & ((rel < sp.y) | (rel == Npar - 1));
                v[c] = ld_pred<cxpair<T>>(pb + off, in);
                f[c] = ld8_pred<repair<T>>(fb + off, in);
            }
#pragma unroll
            for (int c = 0; c < U; ++c) {
                acc[c][0] += v[c].a.x * f[c].a; acc[c][1] += v[c].a.y * f[c].a;
                acc[c][2] += v[c].b.x * f[c].b; acc[c][3] += v[c].b.y * f[c].b;
            }
        }
#pragma unroll
        for (int c = 0; c < U; ++c) {
            const int idx = base + c * k1Threads + tid;
            if (idx < CELLS) {
                const int f1 = idx / (NB / 2), l = 2 * (idx - f1 * (NB / 2));
                s[f1 * W + l] = mk<T>(acc[c][0] * a.scale, acc[c][1] * a.scale);
                s[f1 * W + l + 1] = mk<T>(acc[c][2] * a.scale, acc[c][3] * a.scale);
            }
        }
    }
    __syncthreads();
    slab_fft_s<NA, false, +1, 1, W, T>(s, NB, twA);                    // inverse over f1 (columns): row p holds t1 = invA[p]
    for (int idx = tid; idx < N; idx += nt) {
        const int p = idx / NB, e = idx - p * NB;
        s[p * W + e] = cmulc(s[p * W + e], twn(hi, lo, a.w.lb, e * __ldg(invA + p)));
    }
    __syncthreads();
    slab_fft_s<NB, false, +1, W, 1, T, true>(s, NA, twB);              // inverse over f2 (rows) + modulus
    if (a.mod) {
        T* mb = a.mod + (size_t)g * N;
        for (int t = tid; t < N; t += nt) {
            const int t1 = t & (NA - 1), t2 = t / NA;
            mb[t] = s[__ldg(a.posA + t1) * W + __ldg(a.posB + t2)].x;
        }
        if (!a.spec && !a.part) return;
        __syncthreads();
    }
    slab_fft_s<NB, true, -1, W, 1, T>(s, NA, twB);                     // forward over t2 (rows): natural f2'
    for (int idx = tid; idx < N; idx += nt) {
        const int p = idx / NB, e = idx - p * NB;
        s[p * W + e] = cmul(s[p * W + e], twn(hi, lo, a.w.lb, e * __ldg(invA + p)));
    }
    __syncthreads();
    slab_fft_s<NA, true, -1, 1, W, T>(s, NB, twA);                     // forward over t1 (columns): natural f1'
    if (a.spec) {
        cx<T>* ob = a.spec + (size_t)g * N;
        for (int idx = tid; idx < N; idx += nt) { const int f1 = idx / NB, e = idx - f1 * NB; ob[idx] = s[f1 * W + e]; }
    }
    if (a.part) {
        cx<T>* pb2 = a.part + (size_t)g * a.Fc;
        for (int f = tid; f < a.Fc; f += nt) { const int f1 = f / NB, e = f - f1 * NB; pb2[f] = s[f1 * W + e]; }
    }
}

// ------------------------------------------------------------------ low-pass tail on the lowest Fc bins
// One launch serves every path of a batch chunk: the lines (one per path) are described by segments, one per
// launch group (S0, each first-order group, each (j1, n2) second-order group).
template <typename T> struct FinSeg {
    long long src_off, ss_g, ss_part;      // X[f] = sum_{q<nparts} base[which][src_off + gl*ss_g + q*ss_part + f], f < Fc
    const T* phi;                          // real low-pass on the length-N grid
    const int* chan;                       // [NI] output channel of path i
    int which, nparts, N, Fc, NI, line0;   // gl = line - line0 = b*NI + i
};
template <typename T> struct Finish1 {
    const cx<T>* base[3];                  // U0_hat, U1_hat buffer, leaf partial sums
    const FinSeg<T>* segs; int nseg, total;
    T* out; long long os_b;                // out[b*os_b + chan[i]*W + n - i0]
    int i0, W;
    const cx<T>* twM; const int* posM;
};
constexpr int k1FL = 8, k1FLP = k1FL | 1;  // lines per CTA of the finish kernel
template <typename T, int M> __global__ void __launch_bounds__(k1Threads, 3) k1d_finish(Finish1<T> a) {
    constexpr int LP = k1FLP;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* twM = s + (size_t)M * LP;
    int* posM = reinterpret_cast<int*>(twM + M);
    int* segi = posM + M;
    const int g0 = blockIdx.x * k1FL;
    const int nl = min(k1FL, a.total - g0);
    const int tid = flat_tid(), nt = flat_nt();
    stage(twM, a.twM, M);
    stage(posM, a.posM, M);
    if (tid < nl) {
        int sgi = 0;
        while (sgi + 1 < a.nseg && a.segs[sgi + 1].line0 <= g0 + tid) ++sgi;
        segi[tid] = sgi;
    }
    __syncthreads();
    for (int idx = tid; idx < nl * M; idx += nt) {
        const int l = idx / M, u = idx - l * M;
        const FinSeg<T>& sg = a.segs[segi[l]];
        const int N = sg.N, Fc = sg.Fc, nyq = N >> 1, nparts = sg.nparts;
        const long long ssp = sg.ss_part;
        const cx<T>* __restrict__ xb = a.base[sg.which] + sg.src_off + (long long)(g0 + l - sg.line0) * sg.ss_g;
        const T* __restrict__ phi = sg.phi;
        T ax = T(0), ay = T(0);
        for (int f = u; f < Fc; f += M) {                    // bins f = u + aM on the non-negative side
            T vx = T(0), vy = T(0);
#pragma unroll 4
            for (int q = 0; q < nparts; ++q) { const cx<T> v = xb[q * ssp + f]; vx += v.x; vy += v.y; }
            const T ph = phi[f];
            ax += vx * ph; ay += vy * ph;
        }
        for (int f = (M - u) & (M - 1); f < Fc; f += M) {    // bins N - f == u (mod M): conj(X[f]) * phi[N - f]
            if (f == 0 || f == nyq) continue;
            T vx = T(0), vy = T(0);
#pragma unroll 4
            for (int q = 0; q < nparts; ++q) { const cx<T> v = xb[q * ssp + f]; vx += v.x; vy += v.y; }
            const T ph = phi[N - f];
            ax += vx * ph; ay -= vy * ph;
        }
        const T sc = T(1) / T(N);
        s[u * LP + l] = mk<T>(ax * sc, ay * sc);
    }
    __syncthreads();
    slab_fft_s<M, false, +1, 1, k1FLP, T>(s, nl, twM);                // inverse DIF: sample n at row posM[n]
    for (int idx = tid; idx < nl * a.W; idx += nt) {
        const int l = idx / a.W, n = idx - l * a.W;
        const FinSeg<T>& sg = a.segs[segi[l]];
        const int gl = g0 + l - sg.line0, b = gl / sg.NI, i = gl - b * sg.NI;
        a.out[(long long)b * a.os_b + (long long)sg.chan[i] * a.W + n] = s[posM[a.i0 + n] * LP + l].x;
    }
}

// average='global' tail (kymatio/scattering1d/frontend/base_frontend.py:137-138, backend.average_global): the sum over
// time of a path's modulus field is bin 0 of its spectrum, which the cascade has already produced (complete for parents
// and for U0_hat, as per-CTA partial sums for leaves) - no low-pass, no inverse transform.  One thread per path.
template <typename T> __global__ void k1d_finish_global(Finish1<T> a) {
    const int line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= a.total) return;
    int sgi = 0;
    while (sgi + 1 < a.nseg && a.segs[sgi + 1].line0 <= line) ++sgi;
    const FinSeg<T>& sg = a.segs[sgi];
    const int gl = line - sg.line0, b = gl / sg.NI, i = gl - b * sg.NI;
    const cx<T>* __restrict__ xb = a.base[sg.which] + sg.src_off + (long long)gl * sg.ss_g;
    T acc = T(0);
    for (int q = 0; q < sg.nparts; ++q) acc += xb[q * sg.ss_part].x;
    a.out[(long long)b * a.os_b + sg.chan[i]] = acc;
}

// kernel tables (instances in scat1d_inst.cu)
template <typename T> struct Kern1d {
    void (*col_prod)(ColProd1<T>);
    void (*col_fwd)(ColFwd1<T>);
};
template <typename T> struct KernRow1d {
    void (*parent)(RowMod1<T>);
    void (*leaf)(RowMod1<T>);
    void (*real)(RowReal1<T>);
};
template <typename T> Kern1d<T> kern1d_cols(int NA);                 // nullptr entries when NA is not compiled
template <typename T> KernRow1d<T> kern1d_rows(int NB);
template <typename T> void (*kern1d_finish(int M))(Finish1<T>);
template <typename T> void (*kern1d_tile(int NA, int NB))(Tile1<T>);
constexpr int k1TileMaxN = 8192;
void kern1d_enable_smem();

}  // namespace sb
