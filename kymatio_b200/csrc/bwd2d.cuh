// bwd2d.cuh - fused backward of the FIRST-order block of the 2-D cascade (the reference gets these
// gradients by replaying torch autograd over every primitive, with kymatio/backend/torch_backend.py:64-96 for the modulus).
//
// One first-order path:   V = periodise_k(U0 * psi) * scale;  u = F^H V (unnormalised inverse);  A = |u|;  U1 = F A
// with the gradient gU1 of everything downstream of U1 (the fused second-order block and, at the streaming level, the
// Fourier low-pass) and - at tile levels - the gradient gS1 of the path's own low-passed output:
//     gA  = Re(F^H gU1) [+ G0 gS1 G1^T]        adjoint of the forward transform of a REAL field (+ separable low-pass)
//     gu  = gA u / |u|   (0 where u = 0)       ModulusStable.backward
//     gV  = F gu                               adjoint of the unnormalised inverse transform
//     gU0 += scale * psi * replicate_k(gV)     adjoint of periodisation and filter multiply
//
// Tile levels (field fits one CTA):
//     k2d_tile_adj    R = Re(F^H gU1) of one path, written in the tile's STORAGE order (rows / columns at the scrambled
//                     positions of the decimation-in-frequency inverse), so that
//     k2d_tile_bwd    (tile2d.cuh; recomputes u in the same storage order) just adds R[q][x'] to its low-pass adjoint.
// Streaming level (full resolution, field = 16-line slabs), mirror of the forward chain rowpass_prod -> colpass -> rowpass:
//     (row pass)      Y = rows-inverse(U0 * psi)   and   GX = rows-inverse(gU1)         (k2d_rowpass_prod, unit filter)
//     k2d_bwd_col     per 16-column slab: u = cols-inverse(Y), G = cols-inverse(GX), gu = Re(G) u/|u|, cols-forward(gu)
//     k2d_bwd_row     per (image, 16-row slab): rows-forward of the L paths of the image one after the other,
//                     multiply by psi_theta * scale and ACCUMULATE over theta in registers -> one deterministic
//                     read-modify-write of gU0 (no atomics).
#pragma once
#include "kernels2d.cuh"
#include "tile2d.cuh"

namespace sb {

// ------------------------------------------------------------------ tile level: R = Re(F^H gU1) in storage order
template <typename T> struct TileAdjArgs {
    const cx<T>* gspec;        // [G][n0][n1] natural-order spectrum gradients
    T* R;                      // [G][n0][n1] real, storage order
    const cx<T>* tw0; const cx<T>* tw1;
    int G;
    // runtime-size instance only (k2d_tile_adj_g): field size, transform plans and scramble tables
    int n0, n1;
    Plan1 plan0, plan1;
    const int* pos0; const int* pos1;
};
template <typename T, int N0, int N1>
__global__ void __launch_bounds__(tile_max_threads(N0, N1), tile_min_blocks(N0, N1)) k2d_tile_adj(TileAdjArgs<T> a) {
    constexpr int W = N1 | 1;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw0 = s + (size_t)N0 * W;
    cx<T>* tw1 = tw0 + N0;
    const int tid = flat_tid(), nt = flat_nt();
    stage(tw0, a.tw0, N0); stage(tw1, a.tw1, N1);
    constexpr int half = N1 / 2;
    for (int g = blockIdx.x; g < a.G; g += gridDim.x) {
        const cx<T>* __restrict__ ib = a.gspec + (size_t)g * N0 * N1;
        __syncthreads();
        for (int it = tid; it < N0 * half; it += nt) {
            const int r = it / half, e = 2 * (it - r * half);
            const cx2<T> v = *reinterpret_cast<const cx2<T>*>(ib + (size_t)r * N1 + e);
            s[r * W + e] = v.a; s[r * W + e + 1] = v.b;
        }
        __syncthreads();
        slab_fft_s<N1, false, +1, W, 1, T>(s, N0, tw1);
        slab_fft_s<N0, false, +1, 1, W, T>(s, N1, tw0);
        T* __restrict__ ob = a.R + (size_t)g * N0 * N1;
        for (int it = tid; it < N0 * N1; it += nt) {
            const int q = it / N1, x = it - q * N1;
            ob[it] = s[q * W + x].x;
        }
    }
}

// runtime-size variant for fields without a compiled instance: the generic backward tile (k2d_tile_bwd<T, 0, 0, 0>) keeps its
// field in NATURAL order (decimation-in-time inverse fed through the scramble tables), so R is natural-order too
template <typename T>
__global__ void __launch_bounds__(tile_max_threads(0, 0), tile_min_blocks(0, 0)) k2d_tile_adj_g(TileAdjArgs<T> a) {
    const int n0 = a.n0, n1 = a.n1, W = n1 | 1;
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw0 = s + (size_t)n0 * W;
    cx<T>* tw1 = tw0 + n0;
    int* pos0 = reinterpret_cast<int*>(tw1 + n1);
    int* pos1 = pos0 + n0;
    const int tid = flat_tid(), nt = flat_nt();
    stage(tw0, a.tw0, n0); stage(tw1, a.tw1, n1);
    stage(pos0, a.pos0, n0); stage(pos1, a.pos1, n1);
    for (int g = blockIdx.x; g < a.G; g += gridDim.x) {
        const cx<T>* __restrict__ ib = a.gspec + (size_t)g * n0 * n1;
        __syncthreads();
        for (int it = tid; it < n0 * n1; it += nt) {
            const int r = it / n1, e = it - r * n1;
            s[pos0[r] * W + pos1[e]] = ib[it];
        }
        __syncthreads();
        slab_fft<true, T>(s, n0, W, 1, a.plan1, tw1);
        slab_fft<true, T>(s, n1, 1, W, a.plan0, tw0);
        T* __restrict__ ob = a.R + (size_t)g * n0 * n1;
        for (int it = tid; it < n0 * n1; it += nt) {
            const int q = it / n1, x = it - q * n1;
            ob[it] = s[q * W + x].x;
        }
    }
}

// ------------------------------------------------------------------ streaming level, adjoint of the low-pass (horizontal half)
// The forward S1 = unpad(Re ifft2(periodise(U1 * phi))) of the full-resolution band equals the separable spatial low-pass
// S1 = G0^T A G1 of A = |u| (plan2d.cuh: analyse_lowpass; the fused backward is only offered when that holds), so
//     gA[y][x] = sum_yo G0[y][yo] * Tl[yo][x],     Tl[yo][x] = sum_xo gS1[yo][xo] * G1[x][xo].
// This kernel computes Tl for one path, stored with the columns at the SCRAMBLED positions x' = pos1[x] of the row passes;
// k2d_bwd_col applies the vertical half for its 16 columns.
template <typename T> struct HlowBwdArgs {
    const T* gs1;              // [G][o0][o1]
    const T* G1;               // [n1][o1p] dense
    const int* pos1;           // [n1]
    T* Tl;                     // [G][o0][n1]
    int n1, o0, o1, o1p;
};
template <typename T> __global__ void __launch_bounds__(256) k2d_hlow_bwd(HlowBwdArgs<T> a) {
    T* gs = dyn_smem<T>();                                      // [o0][o1]
    const int g = blockIdx.x;
    const T* __restrict__ gb = a.gs1 + (size_t)g * a.o0 * a.o1;
    for (int i = threadIdx.x; i < a.o0 * a.o1; i += blockDim.x) gs[i] = gb[i];
    __syncthreads();
    T* ob = a.Tl + (size_t)g * a.o0 * a.n1;
    for (int x = threadIdx.x; x < a.n1; x += blockDim.x) {
        const T* __restrict__ gr = a.G1 + (size_t)x * a.o1p;
        int lo = a.o1, hi = -1;
        for (int xo = 0; xo < a.o1; ++xo) if (gr[xo] != T(0)) { if (lo == a.o1) lo = xo; hi = xo; }
        const int xs = a.pos1[x];
        for (int yo = 0; yo < a.o0; ++yo) {
            T acc = T(0);
            for (int xo = lo; xo <= hi; ++xo) acc += gs[yo * a.o1 + xo] * gr[xo];
            ob[(size_t)yo * a.n1 + xs] = acc;
        }
    }
}

// ------------------------------------------------------------------ streaming level, column pass
template <typename T> struct BwdColArgs {
    const cx<T>* Y;            // [G][n0][n1] rows-inverse of the product (rows: natural frequency, columns: scrambled x)
    const cx<T>* GX;           // [G][n0][n1] rows-inverse of gU1, same layout; nullptr: no gradient through U1
    cx<T>* out;                // [G][n0][n1] cols-forward of gu (rows natural frequency, columns scrambled x); may alias Y
    int n1;
    const cx<T>* tw;
    // vertical half of the low-pass adjoint (nullptr: the path's own S1 carries no gradient here)
    const T* Tl;               // [G][o0][n1] from k2d_hlow_bwd
    const T* G0;               // [n0][o0p] dense
    const int* pos0;           // [n0] storage row of natural row y
    int o0, o0p, kl, R;        // R = tap radius (>= n0/2: every output touches every row)
};
constexpr int kBwdThreads = 256;
// grid (G, n1 / 16), kBwdThreads threads; two 16-column slabs in shared memory
template <typename T, int NS> __global__ void __launch_bounds__(kBwdThreads, 3) k2d_bwd_col(BwdColArgs<T> a) {
    constexpr int n0 = NS, LP = kSLP;
    cx<T>* s1 = dyn_smem<cx<T>>();
    cx<T>* s2 = s1 + (size_t)n0 * LP;
    cx<T>* tw = s2 + (size_t)n0 * LP;
    int* ipos = reinterpret_cast<int*>(tw + n0);                 // natural row held at storage row q
    T* tl = reinterpret_cast<T*>(ipos + n0);                     // [o0][16] slice of Tl
    const int g = blockIdx.x, c0 = blockIdx.y * kSLines;
    const int tid = flat_tid(), nt = flat_nt();
    stage(tw, a.tw, n0);
    const size_t off = (size_t)g * n0 * a.n1 + c0;
    const cxpair<T> zero2 = {mk<T>(T(0), T(0)), mk<T>(T(0), T(0))};
    for (int idx = tid; idx < n0 * (kSLines / 2); idx += nt) {
        const int e = idx / (kSLines / 2), l = 2 * (idx - e * (kSLines / 2));
        const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(a.Y + off + (size_t)e * a.n1 + l);
        const cxpair<T> w = a.GX ? *reinterpret_cast<const cxpair<T>*>(a.GX + off + (size_t)e * a.n1 + l) : zero2;
        s1[e * LP + l] = v.a; s1[e * LP + l + 1] = v.b;
        s2[e * LP + l] = w.a; s2[e * LP + l + 1] = w.b;
    }
    if (a.Tl) {
        for (int y = tid; y < n0; y += nt) ipos[a.pos0[y]] = y;
        const T* __restrict__ tb = a.Tl + (size_t)g * a.o0 * a.n1 + c0;
        for (int i = tid; i < a.o0 * kSLines; i += nt) { const int yo = i / kSLines, l = i - yo * kSLines; tl[i] = tb[(size_t)yo * a.n1 + l]; }
    }
    __syncthreads();
    slab_fft_s<NS, false, +1, 1, kSLP, T>(s1, kSLines, tw);      // u   (rows scrambled)
    if (a.GX) slab_fft_s<NS, false, +1, 1, kSLP, T>(s2, kSLines, tw);      // F^H gU1, same positions
    const int mper = n0 / a.kl;                                  // outputs per period before unpadding (= o0 + 2)
    auto modulus_bwd = [&](int e, int l, T gA) {
        const cx<T> v = s1[e * LP + l];
        const T m2 = v.x * v.x + v.y * v.y;
        T sc;
        if constexpr (std::is_same<T, float>::value) sc = m2 > 0.f ? gA * rsqrtf(m2) : 0.f;
        else sc = m2 > T(0) ? gA / sqrt(m2) : T(0);
        s1[e * LP + l] = mk<T>(v.x * sc, v.y * sc);
    };
    constexpr int MT = 8;                                        // taps of a row kept in registers (banded low-pass)
    if (a.Tl && 2 * a.R + 1 < n0 && (2 * a.R) / a.kl + 2 <= MT) {
        // a thread owns one ROW of the slab: the (<= MT) outputs yo its sample y contributes to and their weights are
        // found once, then applied to the 16 columns (G0[y][yo] = a0[(kl (yo+1) - y) mod n0], nonzero within R of a
        // multiple of kl; lanes = different rows at the odd pitch LP: conflict-free)
        for (int e = tid; e < n0; e += nt) {
            const int y = ipos[e];
            const T* __restrict__ g0 = a.G0 + (size_t)y * a.o0p;
            int j0 = y - a.R, j1 = y + a.R;                      // kl*(yo+1) in [j0, j1] (mod n0)
            j0 = (j0 >= 0) ? (j0 + a.kl - 1) / a.kl : -((-j0) / a.kl);
            j1 = (j1 >= 0) ? j1 / a.kl : -((-j1 + a.kl - 1) / a.kl);
            T tap[MT]; int row[MT];
#pragma unroll
            for (int t = 0; t < MT; ++t) {
                const int jj = j0 + t;
                int j = jj % mper; if (j < 0) j += mper;
                const int yo = j - 1;
                const bool ok = jj <= j1 && yo >= 0 && yo < a.o0;
                tap[t] = ok ? g0[yo] : T(0);
                row[t] = ok ? yo * kSLines : 0;
            }
#pragma unroll 4
            for (int l = 0; l < kSLines; ++l) {
                T gA = s2[e * LP + l].x;
#pragma unroll
                for (int t = 0; t < MT; ++t) gA += tap[t] * tl[row[t] + l];
                modulus_bwd(e, l, gA);
            }
        }
    } else {
        for (int idx = tid; idx < n0 * kSLines; idx += nt) {
            const int e = idx / kSLines, l = idx - e * kSLines;
            T gA = s2[e * LP + l].x;
            if (a.Tl) {
                // + sum_yo G0[y][yo] * Tl[yo][x]
                const int y = ipos[e];
                const T* __restrict__ g0 = a.G0 + (size_t)y * a.o0p;
                if (2 * a.R + 1 >= n0) {
                    for (int yo = 0; yo < a.o0; ++yo) gA += g0[yo] * tl[yo * kSLines + l];
                } else {
                    int j0 = y - a.R, j1 = y + a.R;              // kl*(yo+1) in [j0, j1] (mod n0)
                    j0 = (j0 >= 0) ? (j0 + a.kl - 1) / a.kl : -((-j0) / a.kl);
                    j1 = (j1 >= 0) ? j1 / a.kl : -((-j1 + a.kl - 1) / a.kl);
                    for (int jj = j0; jj <= j1; ++jj) {
                        int j = jj % mper; if (j < 0) j += mper;
                        const int yo = j - 1;
                        if (yo >= 0 && yo < a.o0) gA += g0[yo] * tl[yo * kSLines + l];
                    }
                }
            }
            modulus_bwd(e, l, gA);
        }
    }
    __syncthreads();
    slab_fft_s<NS, true, -1, 1, kSLP, T>(s1, kSLines, tw);       // cols-forward: scrambled in -> natural frequency out
    for (int idx = tid; idx < n0 * (kSLines / 2); idx += nt) {
        const int e = idx / (kSLines / 2), l = 2 * (idx - e * (kSLines / 2));
        cxpair<T> v; v.a = s1[e * LP + l]; v.b = s1[e * LP + l + 1];
        *reinterpret_cast<cxpair<T>*>(a.out + off + (size_t)e * a.n1 + l) = v;
    }
}

// ------------------------------------------------------------------ streaming level, row pass + reduction over theta
template <typename T> struct BwdRowArgs {
    const cx<T>* GV;           // [B*NF][n0][n1] output of k2d_bwd_col
    const T* const* filt;      // [NF] real (n0, n1) filters, natural order
    cx<T>* gU0;                // [B][n0][n1], accumulated into
    int n0, NF;
    T scale;
    const cx<T>* tw;
};
// grid (B, ceil(n0 / 16)), kBwdThreads threads
template <typename T, int NS> __global__ void __launch_bounds__(kBwdThreads, 2) k2d_bwd_row(BwdRowArgs<T> a) {
    constexpr int n1 = NS, LP = kSLP, half = NS / 2;
    constexpr int PAIRS = (kSLines * half + kBwdThreads - 1) / kBwdThreads;     // column pairs per thread
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)n1 * LP;
    const int b = blockIdx.x, r0 = blockIdx.y * kSLines;
    const int nl = min(kSLines, a.n0 - r0);
    const int tid = flat_tid(), nt = flat_nt();
    stage(tw, a.tw, n1);
    cx<T> acc[PAIRS][2];
#pragma unroll
    for (int p = 0; p < PAIRS; ++p) { acc[p][0] = mk<T>(T(0), T(0)); acc[p][1] = mk<T>(T(0), T(0)); }
    for (int fi = 0; fi < a.NF; ++fi) {
        const cx<T>* ib = a.GV + (((size_t)b * a.NF + fi) * a.n0 + r0) * n1;
        __syncthreads();                                         // the previous filter's reads of s are done
        for (int idx = tid; idx < nl * half; idx += nt) {
            const int l = idx / half, e = 2 * (idx - l * half);
            const cxpair<T> v = *reinterpret_cast<const cxpair<T>*>(ib + (size_t)l * n1 + e);
            s[e * LP + l] = v.a; s[(e + 1) * LP + l] = v.b;
        }
        __syncthreads();
        slab_fft_s<NS, true, -1, 1, kSLP, T>(s, nl, tw);         // rows-forward: scrambled x in -> natural frequency out
        const T* __restrict__ fb = a.filt[fi] + (size_t)r0 * n1;
#pragma unroll
        for (int p = 0; p < PAIRS; ++p) {
            const int idx = tid + p * nt;
            if (idx < nl * half) {
                const int l = idx / half, e = 2 * (idx - l * half);
                const repair<T> f = *reinterpret_cast<const repair<T>*>(fb + (size_t)l * n1 + e);
                const cx<T> v0 = s[e * LP + l], v1 = s[(e + 1) * LP + l];
                acc[p][0] = fma_rc(acc[p][0], v0, f.a);
                acc[p][1] = fma_rc(acc[p][1], v1, f.b);
            }
        }
    }
    cx<T>* ob = a.gU0 + ((size_t)b * a.n0 + r0) * n1;
#pragma unroll
    for (int p = 0; p < PAIRS; ++p) {
        const int idx = tid + p * nt;
        if (idx < nl * half) {
            const int l = idx / half, e = 2 * (idx - l * half);
            cxpair<T>* dst = reinterpret_cast<cxpair<T>*>(ob + (size_t)l * n1 + e);
            cxpair<T> o = *dst;
            o.a = fma_rc(o.a, acc[p][0], a.scale);
            o.b = fma_rc(o.b, acc[p][1], a.scale);
            *dst = o;
        }
    }
}

// kernel lookup (instances in bwd_inst.cu); null when the size has no static instance
template <typename T> using TileAdjKernel = void (*)(TileAdjArgs<T>);
template <typename T> using BwdColKernel = void (*)(BwdColArgs<T>);
template <typename T> using BwdRowKernel = void (*)(BwdRowArgs<T>);
template <typename T> TileAdjKernel<T> tile_adj_lookup(int n0, int n1);
template <typename T> BwdColKernel<T> bwd_col_lookup(int n);
template <typename T> BwdRowKernel<T> bwd_row_lookup(int n);
template <typename T> size_t bwd_col_smem(int n0, int o0) {
    return ((size_t)2 * n0 * kSLP + n0) * sizeof(cx<T>) + (size_t)n0 * sizeof(int) + (size_t)o0 * kSLines * sizeof(T);
}
template <typename T> void bwd_kernels_enable_smem();

}  // namespace sb
