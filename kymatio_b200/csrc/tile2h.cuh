// tile2h.cuh - the leaf tile kernel in two halves: two CTAs per SM on fields that fill an SM when held whole.
//
// A leaf path (no children: every second-order path and the coarsest first-order scale) only needs
//     S = unpad( (|ifft2(Z)| conv g)[::kl, ::kl] ),      Z = periodise_k(parent * psi) / (k^2 n0 n1)
// (kymatio/scattering2d/core/scattering2d.py:59-75).  The n0 x n1 field |ifft2(Z)| is never needed all at once: its
// even and odd rows y = 2m + h are two independent (n0/2) x n1 inverse transforms of
//     Z_h[r][e] = (Z[r][e] + (-1)^h Z[r + n0/2][e]) * exp(+2 pi i r h / n0),        r < n0/2
// (the first radix-2 decimation-in-frequency stage of the column transform, folded into the product/periodise
// prologue exactly like the alias sums of the periodisation - SURVEY Appendix A identity), and the separable low-pass
// is linear, so the two halves just add:   S = sum_h G0[2m+h][yo] (U_h G1)[m][xo].
// One CTA processes both halves of a path one after the other in a half-size tile and keeps the partial low-pass
// sums in registers.  Shared memory per CTA drops from 220 KB (136 x 136) to ~88 KB, so TWO CTAs share an SM and one
// CTA's product/periodise loads (L2 latency bound) overlap the other's butterflies (issue / shared-memory bound).
// The price is that the parent spectrum is read twice (once per half); the parents are L2 resident.
//
// The dense [n][o] decimation matrices of tile2d.cuh are shift invariant (G[x][o] = a[(kl (o+1) - x) mod n]), so
// this kernel reads the taps from two tiny tables TT[st][4] (taps of the st-th input of a 4-output group).
#pragma once
#include "tile2d.cuh"

namespace sb {

template <typename T> struct Tile2hSmem {
    cx<T>* tile; T* w1; cx<T>* tw1; cx<T>* twh; unsigned* supp; int* pos1; int* posh; T* tt0; T* tt1;
};
template <typename T> __host__ __device__ inline size_t tile2h_smem_layout(const TileArgs<T>& a, Tile2hSmem<T>* L) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 15) / 16 * 16; return o; };
    const int H = a.n0 / 2;
    const size_t o_tile = take(sizeof(cx<T>) * (size_t)H * (a.n1 | 1));
    const size_t o_w1 = take(sizeof(T) * (size_t)H * (a.o1p + 4));
    const size_t o_tw1 = take(sizeof(cx<T>) * a.n1), o_twh = take(sizeof(cx<T>) * H);
    const size_t o_supp = take(sizeof(unsigned) * 2 * a.P0);  // packed, double buffered
    const size_t o_p1 = take(sizeof(int) * a.n1), o_ph = take(sizeof(int) * H);
    const size_t o_t0 = take(sizeof(T) * 4 * (size_t)a.y0cnt), o_t1 = take(sizeof(T) * 4 * (size_t)a.x1cnt);
    if (L) {
#ifdef __CUDA_ARCH__
        unsigned char* base = dyn_smem<unsigned char>();
        L->tile = reinterpret_cast<cx<T>*>(base + o_tile);
        L->w1 = reinterpret_cast<T*>(base + o_w1);
        L->tw1 = reinterpret_cast<cx<T>*>(base + o_tw1); L->twh = reinterpret_cast<cx<T>*>(base + o_twh);
        L->supp = reinterpret_cast<unsigned*>(base + o_supp);
        L->pos1 = reinterpret_cast<int*>(base + o_p1); L->posh = reinterpret_cast<int*>(base + o_ph);
        L->tt0 = reinterpret_cast<T*>(base + o_t0); L->tt1 = reinterpret_cast<T*>(base + o_t1);
#endif
    }
    return off;
}

constexpr int kTile2hMaxThreads = 384;     // two CTAs per SM: <= 85 registers per thread

// product + periodise + first column-DIF stage for 4 adjacent columns e..e+3 of half-row r (< N/2) of half h.
template <typename T, int N, int KT>
__device__ __forceinline__ void tile2h_load_item(cx<T>* s, const unsigned* supp, const cx<T>* __restrict__ pb,
                                                 const T* __restrict__ fb, int r, int e, int P1, T scale, int h,
                                                 int lane, const cx<T>* twN) {
    constexpr int H = N / 2, W = N | 1;
    T sx[4], sy[4];
#pragma unroll
    for (int part = 0; part < 2; ++part) {
        T ax[4], ay[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ax[i] = T(0); ay[i] = T(0); }
#pragma unroll
        for (int c = 0; c < KT; ++c) {
            const int R = r + part * H + c * N;
            const int2 sp = unpack_supp(supp[R]);
            if (KT > 2 && sp.y == 0) continue;
            const size_t rowoff = (size_t)R * P1;
            cx2<T> v0[KT], v1[KT];
            re4<T> f[KT];
#pragma unroll
            for (int d = 0; d < KT; ++d) {
                const int C = e + d * N;
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
        if (part == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { sx[i] = ax[i]; sy[i] = ay[i]; }
        } else if (h) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { sx[i] -= ax[i]; sy[i] -= ay[i]; }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) { sx[i] += ax[i]; sy[i] += ay[i]; }
        }
    }
    // odd half: twiddle exp(+2 pi i r / N) = conj(tw[r]); the normalisation rides along
    T wx = scale, wy = T(0);
    if (h) { const cx<T> w = twN[r]; wx = w.x * scale; wy = -w.y * scale; }
    cx<T>* dst = s + r * W + e;
    // rotate which of its 4 columns a lane writes per store instruction (see tile_load_item)
    const int rot = (lane >> 2) & 3;
    const bool p0 = rot & 1, p1 = rot & 2;
    T bx[4], by[4], cxr[4], cyr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { bx[i] = p0 ? sx[(i + 1) & 3] : sx[i]; by[i] = p0 ? sy[(i + 1) & 3] : sy[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { cxr[i] = p1 ? bx[(i + 2) & 3] : bx[i]; cyr[i] = p1 ? by[(i + 2) & 3] : by[i]; }
#pragma unroll
    for (int t = 0; t < 4; ++t) dst[(t + rot) & 3] = mk<T>(cxr[t] * wx - cyr[t] * wy, cxr[t] * wy + cyr[t] * wx);
}

template <typename T, int N, int KT>
__global__ void __launch_bounds__(kTile2hMaxThreads, 2) k2d_tile2h(TileArgs<T> a) {
    constexpr int H = N / 2, W = N | 1;
    const int wp = a.o1p + 4;
    Tile2hSmem<T> m;
    tile2h_smem_layout(a, &m);
    cx<T>* s = m.tile;
    const int tid = flat_tid(), nt = flat_nt();
    const int lane = tid & 31;

    stage(m.tw1, a.tw1, N); stage(m.twh, a.twh, H);
    stage(m.pos1, a.pos1, N); stage(m.posh, a.posh, H);
    stage(m.tt0, a.TT0, 4 * a.y0cnt); stage(m.tt1, a.TT1, 4 * a.x1cnt);

    SB_PHASE_INIT(16 + (KT == 4 ? 4 : 0) + (a.PP != a.NFch ? 1 : 0))
    const int ygroups = a.o0p >> 2;
    if ((int)blockIdx.x < a.G) stage_supp(m.supp, a.supp + (size_t)(blockIdx.x % a.NF) * a.P0, a.P0);
    int sbuf = 0;
    for (int g = blockIdx.x; g < a.G; g += gridDim.x, sbuf ^= 1) {
        const int fi = g % a.NF, pg = g / a.NF;
        const int b = g / a.PP, path = g - b * a.PP;
        const int ch = a.ch0 + (path / a.NFch) * a.chs + (path % a.NFch);
        const unsigned* supp = m.supp + sbuf * a.P0;
        const int gn = g + gridDim.x;
        if (gn < a.G) {      // next path of this CTA: parent slice -> L2, support rows -> the other buffer
            if (a.prefetch) prefetch_rows_slice(a.parent + (size_t)(gn / a.NF) * a.P0 * a.P1, a.P0, a.P1, gn % a.NF, a.NF, tid, nt);
            stage_supp(m.supp + (sbuf ^ 1) * a.P0, a.supp + (size_t)(gn % a.NF) * a.P0, a.P0);
        }
        __syncthreads();
        SB_PHASE(0);
        const cx<T>* __restrict__ pb = a.parent + (size_t)pg * a.P0 * a.P1;
        const T* __restrict__ fb = a.filt[fi];
        // 4b accumulators of this thread's (4 output rows, 1 output column) item, summed over both halves
        T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            // 1. product + periodise + first column DIF stage -> Z_h (H x N, natural order)
            {
                constexpr int per_row = N >> 2, items = H * per_row;
                for (int it = tid; it < items; it += nt) {
                    const int r0 = it / per_row, e0 = 4 * (it - r0 * per_row);
                    tile2h_load_item<T, N, KT>(s, supp, pb, fb, r0, e0, a.P1, a.scale, h, lane, m.tw1);
                }
            }
            __syncthreads();
            SB_PHASE(1);
            // 2+3. inverse transforms (DIF: natural in -> scrambled out), modulus in the last pass:
            //      U[2m+h][x] lives at s[posh[m]*W + pos1[x]]
            slab_fft_s<N, false, +1, W, 1, T>(s, H, m.tw1);
            SB_PHASE(2);
            slab_fft_s<H, false, +1, 1, W, T, true>(s, N, m.twh);
            SB_PHASE(3);
            // 4a. horizontal low-pass + decimation + unpad: w1[row][xo] = sum_x U[row][x] * g1[kl (xo+1) - x]
            {
                constexpr int rgroups = (H + 3) >> 2;
                const int xgroups = a.o1p >> 2;
                for (int it = tid; it < rgroups * xgroups; it += nt) {
                    const int xg = it / rgroups, rg = it - xg * rgroups;
                    int yy[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) yy[j] = min(rg + j * rgroups, H - 1) * W;
                    T acc[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
                    int x = (a.kl * (4 * xg + 1) + a.x1lo) % N;
                    if (x < 0) x += N;
                    for (int st = 0; st < a.x1cnt; ++st) {
                        const re4<T> gq = *reinterpret_cast<const re4<T>*>(m.tt1 + 4 * st);
                        const int xs = m.pos1[x];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const T u = s[yy[j] + xs].x;
                            acc[j][0] += u * gq.a; acc[j][1] += u * gq.b; acc[j][2] += u * gq.c; acc[j][3] += u * gq.d;
                        }
                        x = (x + 1 == N) ? 0 : x + 1;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int y = rg + j * rgroups;
                        if (y < H) {
                            re4<T> o; o.a = acc[j][0]; o.b = acc[j][1]; o.c = acc[j][2]; o.d = acc[j][3];
                            *reinterpret_cast<re4<T>*>(m.w1 + y * wp + 4 * xg) = o;
                        }
                    }
                }
            }
            __syncthreads();
            SB_PHASE(4);
            // 4b. vertical low-pass over the rows y = 2m + h of this half, accumulated in registers
            if (tid < ygroups * a.o1p) {
                const int yg = tid / a.o1p, xo = tid - yg * a.o1p;
                int y = (a.kl * (4 * yg + 1) + a.y0lo) % N;
                if (y < 0) y += N;
                int st = 0;
                if ((y & 1) != h) { st = 1; y = (y + 1 == N) ? 0 : y + 1; }
                for (; st < a.y0cnt; st += 2) {
                    const re4<T> gq = *reinterpret_cast<const re4<T>*>(m.tt0 + 4 * st);
                    const T w = m.w1[m.posh[y >> 1] * wp + xo];
                    acc0 += w * gq.a; acc1 += w * gq.b; acc2 += w * gq.c; acc3 += w * gq.d;
                    y += 2; if (y >= N) y -= N;
                }
            }
            __syncthreads();     // the next half (or path) rewrites the tile and w1
            SB_PHASE(5);
        }
        if (tid < ygroups * a.o1p) {
            const int yg = tid / a.o1p, xo = tid - yg * a.o1p;
            if (xo < a.o1) {
                const OutRef<T> ob = out_ref(a.out, a.peers, ((size_t)b * a.K + ch) * a.o0 * a.o1);
                const int yo = 4 * yg;
                if (yo + 0 < a.o0) ob[(yo + 0) * a.o1 + xo] = acc0;
                if (yo + 1 < a.o0) ob[(yo + 1) * a.o1 + xo] = acc1;
                if (yo + 2 < a.o0) ob[(yo + 2) * a.o1 + xo] = acc2;
                if (yo + 3 < a.o0) ob[(yo + 3) * a.o1 + xo] = acc3;
            }
        }
    }
}

template <typename T> TileKernel<T> tile2h_kernel_lookup(int n0, int n1, int k);
template <typename T> void tile2h_kernels_enable_smem();

}  // namespace sb
