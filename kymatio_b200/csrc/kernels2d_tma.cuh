// kernels2d_tma.cuh - the streaming row passes of the full-resolution first-order chain, fed by the TMA unit.
//
// Same mathematics as k2d_rowpass_prod / k2d_rowpass_fwdh (kernels2d.cuh; reference primitives cdgmm -> ifft and rfft of
// kymatio/scattering2d/core/scattering2d.py:33-47), different machinery:
//   * persistent CTAs (one wave) loop over 16-row slabs; twiddles are staged once per CTA, not once per slab;
//   * a slab of 16 consecutive rows is contiguous in HBM, so ONE asynchronous bulk copy (cp.async.bulk, SASS UBLKCP)
//     brings it into shared memory and one bulk copy writes the result back - no LDG -> STS staging loops, no index
//     arithmetic, no registers, and the copy of slab i+1 / the write-back of slab i-1 overlap the butterflies of slab i
//     (two buffers, transaction mbarriers `full`, plain mbarriers `done`);
//   * a dedicated producer warp (one elected lane) issues every copy; the 9 compute warps never touch global memory
//     except for the read-only filter values (and the mirrored rows of the Hermitian pass);
//   * shared memory is ROW-major with a dense pitch (what a bulk copy delivers); butterflies are distributed with the
//     butterfly index fastest across lanes, which is bank-conflict free for this layout (consecutive lanes read
//     consecutive elements; the radix-17 pass reads blocks 17 elements = 34 banks apart).
// Instances: line length 272 = 16 x 17 (BASELINE configs[1]: 256 x 256 padded, J = 3), float.
#pragma once
#include "kernels2d.cuh"
#include "tma.cuh"

namespace sb {

constexpr int kTmaRows = 16;                                   // rows per slab
constexpr int kTmaComputeWarps = 9;                            // 288 threads: one butterfly per thread and pass at 272
constexpr int kTmaComputeThreads = kTmaComputeWarps * 32;
constexpr int kTmaThreads = kTmaComputeThreads + 32;           // + the producer warp

template <int NS> constexpr size_t tma_row_smem_bytes() {
    return 2 * (size_t)kTmaRows * NS * sizeof(cx<float>) + (size_t)NS * sizeof(cx<float>) + 4 * sizeof(uint64_t);
}

// passes [PBEGIN, NP) of a DIF (or NP-1-.. of a DIT) transform of `rows` row-major lines of length N; butterfly index
// fastest across the compute threads; named barrier 1 over the compute threads after every pass
template <int N, bool DIT, int SIGN, int PBEGIN, typename T>
__device__ __forceinline__ void slab_fft_rm(cx<T>* s, const int rows, const cx<T>* tw, const int tid) {
    constexpr int NP = ct_plan1(N).npass;
    static_for<PBEGIN, NP>([&](auto pp_) {
        constexpr int pp = decltype(pp_)::value;
        constexpr int p = DIT ? NP - 1 - pp : pp;
        constexpr int r = ct_plan1(N).radix[p], m = ct_plan1(N).blen[p];
        constexpr int q = m / r, nbf = N / r, tws = N / m;
        const int items = nbf * rows;
        for (int it = tid; it < items; it += kTmaComputeThreads) {
            const int row = it / nbf, bf = it - row * nbf;
            const int blk = bf / q, i = bf - blk * q;
            butterfly_s<r, DIT, SIGN, q, 1, false, T>(s + row * N + blk * m + i, i * tws, tw);
        }
        tma::named_sync<1>(kTmaComputeThreads);
    });
}

// ------------------------------------------------------------------------------------------------------------------
// out[g][r][:] = inverse row transform (DIF: natural in, scrambled out) of parent[g / NF][r][:] * filt[g % NF][r][:] * scale
// (no periodisation: parent and output have the same size NS x NS).
// Slab-to-CTA map: a CTA owns ONE (filter, row block) pair for its whole life and walks over the images, so the 16
// filter values of each thread's first-pass butterfly are loaded ONCE into registers - the steady-state loop touches
// global memory only through the two bulk copies.  grid = NF * (NS/16) * m CTAs; CTA (pair, j) takes images j, j+m, ...
// ------------------------------------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(kTmaThreads, 2) k2d_rowprod_tma(RowProdArgs<float> a, int Bp, int m) {
    using T = float;
    constexpr int ROWS = kTmaRows, SPP = NS / ROWS;            // row blocks per field
    static_assert(NS % ROWS == 0, "line count must be a multiple of the slab height");
    constexpr uint32_t SLAB_BYTES = ROWS * NS * sizeof(cx<T>);
    unsigned char* base = dyn_smem<unsigned char>();
    cx<T>* buf[2] = {reinterpret_cast<cx<T>*>(base), reinterpret_cast<cx<T>*>(base + SLAB_BYTES)};
    cx<T>* tw = reinterpret_cast<cx<T>*>(base + 2 * SLAB_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(tw + NS);
    uint64_t* done = full + 2;
    const int tid = threadIdx.x;
    if (tid == 0) {
        tma::mbar_init(&full[0], 1); tma::mbar_init(&full[1], 1);
        tma::mbar_init(&done[0], 1); tma::mbar_init(&done[1], 1);
        tma::fence_mbar_init();
    }
    for (int i = tid; i < NS; i += kTmaThreads) tw[i] = a.tw[i];
    __syncthreads();
    const int npairs = a.NF * SPP;
    const int pair = blockIdx.x % npairs, j = blockIdx.x / npairs;
    const int fi = pair / SPP, r0 = (pair - fi * SPP) * ROWS;
    const int n_my = j < Bp ? (Bp - j + m - 1) / m : 0;       // images j, j + m, ...

    if (tid >= kTmaComputeThreads) {
        // ---------------- producer warp: one lane issues every bulk copy of this CTA
        if (tid != kTmaComputeThreads) return;
        auto issue_load = [&](int i) {
            const int b = i & 1;
            const cx<T>* src = a.parent + ((size_t)(j + i * m) * NS + r0) * NS;
            tma::mbar_arrive_expect_tx(&full[b], SLAB_BYTES);
            tma::bulk_load(buf[b], src, SLAB_BYTES, &full[b]);
        };
        for (int i = 0; i < 2 && i < n_my; ++i) issue_load(i);
        for (int i = 0; i < n_my; ++i) {
            const int b = i & 1;
            const size_t g = (size_t)(j + i * m) * a.NF + fi;
            tma::mbar_wait(&done[b], (i >> 1) & 1);
            tma::bulk_store(a.out + (g * NS + r0) * NS, buf[b], SLAB_BYTES);
            tma::bulk_commit();
            if (i + 2 < n_my) { tma::bulk_wait_read<0>(); issue_load(i + 2); }
        }
        tma::bulk_wait<0>();
        return;
    }
    // ---------------- compute warps
    constexpr int R0 = ct_plan1(NS).radix[0], Q0 = NS / R0;
    static_assert(Q0 * ROWS <= kTmaComputeThreads, "one first-pass butterfly per compute thread");
    const int row = min(tid / Q0, ROWS - 1), e = tid - (tid / Q0) * Q0;
    T f0[R0];                                                   // this thread's filter values, resident for every image
    if (a.filt) {
        const T* __restrict__ frow = a.filt[fi] + (size_t)(r0 + row) * NS + e;
        static_for<0, R0>([&](auto k_) { constexpr int k = decltype(k_)::value; f0[k] = __ldg(frow + k * Q0) * a.scale; });
    } else {                                                    // unit filter (adjoint row pass of the backward chain)
        static_for<0, R0>([&](auto k_) { constexpr int k = decltype(k_)::value; f0[k] = a.scale; });
    }
    for (int i = 0; i < n_my; ++i) {
        const int b = i & 1;
        cx<T>* s = buf[b];
        tma::mbar_wait(&full[b], (i >> 1) & 1);
        // first radix pass with the filter multiply folded into its loads
        if (tid < Q0 * ROWS) {
            cx<T>* p0 = s + row * NS + e;
            cx<T> v[R0];
            static_for<0, R0>([&](auto k_) { constexpr int k = decltype(k_)::value; v[k] = scal(p0[k * Q0], f0[k]); });
            butterfly_v<R0, false, +1, Q0, T>(v, e, tw);
            static_for<0, R0>([&](auto k_) { constexpr int k = decltype(k_)::value; p0[k * Q0] = v[k]; });
        }
        tma::named_sync<1>(kTmaComputeThreads);
        slab_fft_rm<NS, false, +1, 1, T>(s, ROWS, tw, tid);
        tma::fence_proxy_async();
        tma::named_sync<1>(kTmaComputeThreads);
        if (tid == 0) tma::mbar_arrive(&done[b]);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Hermitian forward row pass (see k2d_rowpass_fwdh): rows v <= n0/2 of `data` (in place) hold scrambled spatial rows;
// forward DIT -> natural-order Fourier rows; row v is written back by a bulk copy, its conjugate mirror row n0 - v by
// the compute warps, and (optionally) both feed the row-folded low-pass product.
// ------------------------------------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(kTmaThreads, 3) k2d_rowfwdh_tma(RowArgs<float> a, int nslabs) {
    using T = float;
    constexpr int ROWS = kTmaRows, n1 = NS;
    constexpr uint32_t ROW_BYTES = NS * sizeof(cx<T>), SLAB_BYTES = ROWS * ROW_BYTES;
    unsigned char* base = dyn_smem<unsigned char>();
    cx<T>* buf[2] = {reinterpret_cast<cx<T>*>(base), reinterpret_cast<cx<T>*>(base + SLAB_BYTES)};
    cx<T>* tw = reinterpret_cast<cx<T>*>(base + 2 * SLAB_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(tw + NS);
    uint64_t* done = full + 2;
    const int tid = threadIdx.x;
    const int H = a.n0 / 2;
    const int SPP = (H + 1 + ROWS - 1) / ROWS;                 // slabs per path (the last one is short)
    if (tid == 0) {
        tma::mbar_init(&full[0], 1); tma::mbar_init(&full[1], 1);
        tma::mbar_init(&done[0], 1); tma::mbar_init(&done[1], 1);
        tma::fence_mbar_init();
    }
    for (int i = tid; i < NS; i += kTmaThreads) tw[i] = a.tw[i];
    __syncthreads();
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_my = first < nslabs ? (nslabs - first + stride - 1) / stride : 0;

    if (tid >= kTmaComputeThreads) {
        if (tid != kTmaComputeThreads) return;
        auto issue_load = [&](int i) {
            const int t = first + i * stride, g = t / SPP, r0 = (t - g * SPP) * ROWS, b = i & 1;
            const uint32_t bytes = (uint32_t)min(ROWS, H + 1 - r0) * ROW_BYTES;
            tma::mbar_arrive_expect_tx(&full[b], bytes);
            tma::bulk_load(buf[b], a.in + ((size_t)g * a.n0 + r0) * n1, bytes, &full[b]);
        };
        for (int i = 0; i < 2 && i < n_my; ++i) issue_load(i);
        for (int i = 0; i < n_my; ++i) {
            const int t = first + i * stride, g = t / SPP, r0 = (t - g * SPP) * ROWS, b = i & 1;
            const uint32_t bytes = (uint32_t)min(ROWS, H + 1 - r0) * ROW_BYTES;
            tma::mbar_wait(&done[b], (i >> 1) & 1);
            tma::bulk_store(a.out + ((size_t)g * a.n0 + r0) * n1, buf[b], bytes);
            tma::bulk_commit();
            if (i + 2 < n_my) { tma::bulk_wait_read<0>(); issue_load(i + 2); }
        }
        tma::bulk_wait<0>();
        return;
    }
    constexpr int half = n1 / 2;
    for (int i = 0; i < n_my; ++i) {
        const int t = first + i * stride, g = t / SPP, r0 = (t - g * SPP) * ROWS, b = i & 1;
        const int nl = min(ROWS, H + 1 - r0);
        cx<T>* s = buf[b];
        cx<T>* ob = a.out + (size_t)g * a.n0 * n1;
        tma::mbar_wait(&full[b], (i >> 1) & 1);
        slab_fft_rm<NS, true, -1, 0, T>(s, nl, tw, tid);
        // the bulk write-back of rows v may start while the mirrors are built: nothing below writes shared memory
        tma::fence_proxy_async();
        // conjugate mirror rows: U[n0 - v][w] = conj(U[v][(n1 - w) % n1]), two columns per thread
        for (int idx = tid; idx < nl * half; idx += kTmaComputeThreads) {
            const int l = idx / half, e = 2 * (idx - l * half);
            const int v = r0 + l;
            if (v > 0 && v < H) {
                const cx<T> m0 = s[l * NS + (e == 0 ? 0 : n1 - e)], m1 = s[l * NS + (n1 - e - 1)];
                cxpair<T> om; om.a = mk<T>(m0.x, -m0.y); om.b = mk<T>(m1.x, -m1.y);
                *reinterpret_cast<cxpair<T>*>(ob + (size_t)(a.n0 - v) * n1 + e) = om;
            }
        }
        if (a.low_out) {
            // row-folded low-pass product: low_out[g][u][e] = sum_d U[u][e + d*m1] * phi[u][e + d*m1] for u = v and n0 - v
            const int m1 = a.low_m1, kf = n1 / m1;
            for (int idx = tid; idx < 2 * nl * m1; idx += kTmaComputeThreads) {
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
                            const cx<T> tv = s[l * NS + (mir ? (C == 0 ? 0 : n1 - C) : C)];
                            const T f = fr[C];
                            ax += tv.x * f; ay += (mir ? -tv.y : tv.y) * f;
                        }
                    }
                }
                a.low_out[((size_t)g * a.n0 + u) * m1 + e] = mk<T>(ax, ay);
            }
        }
        tma::named_sync<1>(kTmaComputeThreads);
        if (tid == 0) tma::mbar_arrive(&done[b]);
    }
}

// lookup / opt-in (instances in tma_inst.cu); null when there is no instance for the line length
using RowProdTmaKernel = void (*)(RowProdArgs<float>, int, int);
using RowFwdhTmaKernel = void (*)(RowArgs<float>, int);
RowProdTmaKernel rowprod_tma_lookup(int n);
RowFwdhTmaKernel rowfwdh_tma_lookup(int n);
size_t tma_row_smem(int n);
void tma_kernels_enable_smem();

}  // namespace sb
