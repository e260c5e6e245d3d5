// tmap_inst.cu - host side of the tensor-map (2-D TMA, hardware swizzle) row pass: descriptor encoding + launch.
#include <cuda.h>
#include "kernels2d_tmap.cuh"
#include "common.cuh"

namespace sb {

namespace tma {
// tensor copy global -> shared (coordinates: c0 = element index along the row, c1 = row), completion on `bar`
__device__ __forceinline__ void tensor_load_2d(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     saddr(dst_smem)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(saddr(bar))
                 : "memory");
}
__device__ __forceinline__ void tensor_store_2d(const CUtensorMap* map, int c0, int c1, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(saddr(src_smem))
                 : "memory");
}
}  // namespace tma

// byte offset of complex element e (< 16) of row r (< 16) inside a swizzled box
__device__ __forceinline__ uint32_t tmap_off(int r, int e) { return (uint32_t)(r * 128 + ((((e >> 1) ^ (r & 7)) << 4) | ((e & 1) << 3))); }

// out[g][r][:] = inverse row transform (DIF: natural in, scrambled out) of parent[g / NF][r][:] * filt[g % NF][r][:] * scale
// for 256 x 256 fields.  Work split as k2d_rowprod_tma: a CTA owns one (filter, 16-row block) pair and walks over the images
// j, j + m, ...; its 16 filter values per thread stay in registers.
__global__ void __launch_bounds__(kTmapThreads, 2)
k2d_rowprod_tmap256(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map, RowProdArgs<float> a, int Bp,
                    int m) {
    using T = float;
    constexpr int NS = 256, ROWS = kTmapRows, SPP = NS / ROWS;
    extern __shared__ unsigned char tmap_smem_raw[];
    unsigned char* base = tmap_smem_raw + ((1024u - (tma::saddr(tmap_smem_raw) & 1023u)) & 1023u);     // swizzle atom alignment
    cx<T>* tw = reinterpret_cast<cx<T>*>(base + 2 * kTmapSlabBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(tw + NS);
    uint64_t* done = full + 2;
    const int tid = threadIdx.x;
    if (tid == 0) {
        tma::mbar_init(&full[0], 1); tma::mbar_init(&full[1], 1);
        tma::mbar_init(&done[0], 1); tma::mbar_init(&done[1], 1);
        tma::fence_mbar_init();
    }
    for (int i = tid; i < NS; i += kTmapThreads) tw[i] = a.tw[i];
    __syncthreads();
    const int npairs = a.NF * SPP;
    const int pair = blockIdx.x % npairs, j = blockIdx.x / npairs;
    const int fi = pair / SPP, r0 = (pair - fi * SPP) * ROWS;
    const int n_my = j < Bp ? (Bp - j + m - 1) / m : 0;       // images j, j + m, ...

    if (tid >= kTmapComputeThreads) {
        // ---------------- producer warp: one lane issues the 16 box copies of every slab
        if (tid != kTmapComputeThreads) return;
        auto issue_load = [&](int i) {
            const int b = i & 1;
            const int row = (j + i * m) * NS + r0;
            tma::mbar_arrive_expect_tx(&full[b], kTmapSlabBytes);
#pragma unroll
            for (int k = 0; k < 16; ++k) tma::tensor_load_2d(base + b * kTmapSlabBytes + k * kTmapBoxBytes, &in_map, 32 * k, row, &full[b]);
        };
        for (int i = 0; i < 2 && i < n_my; ++i) issue_load(i);
        for (int i = 0; i < n_my; ++i) {
            const int b = i & 1;
            const int row = ((j + i * m) * a.NF + fi) * NS + r0;
            tma::mbar_wait(&done[b], (i >> 1) & 1);
#pragma unroll
            for (int k = 0; k < 16; ++k) tma::tensor_store_2d(&out_map, 32 * k, row, base + b * kTmapSlabBytes + k * kTmapBoxBytes);
            tma::bulk_commit();
            if (i + 2 < n_my) { tma::bulk_wait_read<0>(); issue_load(i + 2); }
        }
        tma::bulk_wait<0>();
        return;
    }
    // ---------------- compute warps
    // pass 0: thread (row, e) owns elements e + 16 k, k < 16: the same offset in each of the 16 boxes
    const int rowA = tid >> 4, eA = tid & 15;
    const uint32_t offA = tmap_off(rowA, eA);
    T f0[16];
    if (a.filt) {
        const T* __restrict__ frow = a.filt[fi] + (size_t)(r0 + rowA) * NS + eA;
#pragma unroll
        for (int k = 0; k < 16; ++k) f0[k] = __ldg(frow + 16 * k) * a.scale;
    } else {                                                    // unit filter (adjoint row pass of the backward chain)
#pragma unroll
        for (int k = 0; k < 16; ++k) f0[k] = a.scale;
    }
    // pass 1: thread (box, row) with the row fastest across lanes
    const int boxB = tid >> 4, rowB = tid & 15;
    for (int i = 0; i < n_my; ++i) {
        const int b = i & 1;
        unsigned char* s = base + b * kTmapSlabBytes;
        tma::mbar_wait(&full[b], (i >> 1) & 1);
        {
            cx<T> v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = scal(*reinterpret_cast<const cx<T>*>(s + k * kTmapBoxBytes + offA), f0[k]);
            butterfly_v<16, false, +1, 16, T>(v, eA, tw);
#pragma unroll
            for (int k = 0; k < 16; ++k) *reinterpret_cast<cx<T>*>(s + k * kTmapBoxBytes + offA) = v[k];
        }
        tma::named_sync<1>(kTmapComputeThreads);
        {
            cx<T> v[16];
            unsigned char* bx = s + boxB * kTmapBoxBytes + rowB * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 q = *reinterpret_cast<const float4*>(bx + ((c ^ (rowB & 7)) << 4));
                v[2 * c] = mk<T>(q.x, q.y); v[2 * c + 1] = mk<T>(q.z, q.w);
            }
            butterfly_v<16, false, +1, 1, T>(v, 0, tw);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 q; q.x = v[2 * c].x; q.y = v[2 * c].y; q.z = v[2 * c + 1].x; q.w = v[2 * c + 1].y;
                *reinterpret_cast<float4*>(bx + ((c ^ (rowB & 7)) << 4)) = q;
            }
        }
        tma::fence_proxy_async();
        tma::named_sync<1>(kTmapComputeThreads);
        if (tid == 0) tma::mbar_arrive(&done[b]);
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Hermitian forward row pass for 256 x 256 fields (see k2d_rowfwdh_tma, kernels2d_tma.cuh): rows v <= 128 of `data` hold
// scrambled spatial rows; forward DIT -> natural-order Fourier rows; row v goes back through the tensor map, its conjugate
// mirror row 256 - v is written by the compute warps, and (optionally) both feed the row-folded low-pass product.
// A path has 129 such rows = 8 full slabs + one slab with a single valid row, which is stored through a second tensor map
// whose box is one row high (rows 129.. of the slab are other slabs' mirror rows and must not be overwritten).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tmap_pos(int l, int col) { return (uint32_t)((col >> 4) * kTmapBoxBytes) + tmap_off(l, col & 15); }

__global__ void __launch_bounds__(kTmapThreads, 3)
k2d_rowfwdh_tmap256(const __grid_constant__ CUtensorMap map16, const __grid_constant__ CUtensorMap map1, RowArgs<float> a, int nslabs) {
    using T = float;
    constexpr int NS = 256, ROWS = kTmapRows, n1 = NS, H = NS / 2;
    constexpr int SPP = (H + 1 + ROWS - 1) / ROWS;             // 9 slabs per path (the last one holds row 128 only)
    extern __shared__ unsigned char tmap_smem_raw[];
    unsigned char* base = tmap_smem_raw + ((1024u - (tma::saddr(tmap_smem_raw) & 1023u)) & 1023u);
    cx<T>* tw = reinterpret_cast<cx<T>*>(base + 2 * kTmapSlabBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(tw + NS);
    uint64_t* done = full + 2;
    const int tid = threadIdx.x;
    if (tid == 0) {
        tma::mbar_init(&full[0], 1); tma::mbar_init(&full[1], 1);
        tma::mbar_init(&done[0], 1); tma::mbar_init(&done[1], 1);
        tma::fence_mbar_init();
    }
    for (int i = tid; i < NS; i += kTmapThreads) tw[i] = a.tw[i];
    __syncthreads();
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_my = first < nslabs ? (nslabs - first + stride - 1) / stride : 0;

    if (tid >= kTmapComputeThreads) {
        if (tid != kTmapComputeThreads) return;
        auto issue_load = [&](int i) {
            const int t = first + i * stride, g = t / SPP, r0 = (t - g * SPP) * ROWS, b = i & 1;
            // the short slab also loads a full box (rows 129.. are in bounds and ignored)
            tma::mbar_arrive_expect_tx(&full[b], kTmapSlabBytes);
#pragma unroll
            for (int k = 0; k < 16; ++k)
                tma::tensor_load_2d(base + b * kTmapSlabBytes + k * kTmapBoxBytes, &map16, 32 * k, g * NS + r0, &full[b]);
        };
        for (int i = 0; i < 2 && i < n_my; ++i) issue_load(i);
        for (int i = 0; i < n_my; ++i) {
            const int t = first + i * stride, g = t / SPP, r0 = (t - g * SPP) * ROWS, b = i & 1;
            tma::mbar_wait(&done[b], (i >> 1) & 1);
            if (r0 + ROWS <= H + 1) {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    tma::tensor_store_2d(&map16, 32 * k, g * NS + r0, base + b * kTmapSlabBytes + k * kTmapBoxBytes);
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    tma::tensor_store_2d(&map1, 32 * k, g * NS + r0, base + b * kTmapSlabBytes + k * kTmapBoxBytes);
            }
            tma::bulk_commit();
            if (i + 2 < n_my) { tma::bulk_wait_read<0>(); issue_load(i + 2); }
        }
        tma::bulk_wait<0>();
        return;
    }
    const int rowA = tid >> 4, eA = tid & 15;                   // stride-16 pass: the same offset in each box
    const uint32_t offA = tmap_off(rowA, eA);
    const int boxB = tid >> 4, rowB = tid & 15;                 // contiguous pass: (box, row), row fastest across lanes
    constexpr int half = n1 / 2;
    for (int i = 0; i < n_my; ++i) {
        const int t = first + i * stride, g = t / SPP, r0 = (t - g * SPP) * ROWS, b = i & 1;
        const int nl = min(ROWS, H + 1 - r0);
        unsigned char* s = base + b * kTmapSlabBytes;
        cx<T>* ob = a.out + (size_t)g * a.n0 * n1;
        tma::mbar_wait(&full[b], (i >> 1) & 1);
        {   // forward DIT, first pass: radix 16 on the 16 contiguous elements of a box row (no twiddles)
            cx<T> v[16];
            unsigned char* bx = s + boxB * kTmapBoxBytes + rowB * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 q = *reinterpret_cast<const float4*>(bx + ((c ^ (rowB & 7)) << 4));
                v[2 * c] = mk<T>(q.x, q.y); v[2 * c + 1] = mk<T>(q.z, q.w);
            }
            butterfly_v<16, true, -1, 1, T>(v, 0, tw);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 q; q.x = v[2 * c].x; q.y = v[2 * c].y; q.z = v[2 * c + 1].x; q.w = v[2 * c + 1].y;
                *reinterpret_cast<float4*>(bx + ((c ^ (rowB & 7)) << 4)) = q;
            }
        }
        tma::named_sync<1>(kTmapComputeThreads);
        {   // second pass: radix 16 at stride 16 with the twiddles
            cx<T> v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = *reinterpret_cast<const cx<T>*>(s + k * kTmapBoxBytes + offA);
            butterfly_v<16, true, -1, 16, T>(v, eA, tw);
#pragma unroll
            for (int k = 0; k < 16; ++k) *reinterpret_cast<cx<T>*>(s + k * kTmapBoxBytes + offA) = v[k];
        }
        tma::named_sync<1>(kTmapComputeThreads);
        // the write-back of rows v may start while the mirrors are built: nothing below writes shared memory
        tma::fence_proxy_async();
        auto at = [&](int l, int col) { return *reinterpret_cast<const cx<T>*>(s + tmap_pos(l, col)); };
        // conjugate mirror rows: U[n0 - v][w] = conj(U[v][(n1 - w) % n1]), two columns per thread
        for (int idx = tid; idx < nl * half; idx += kTmapComputeThreads) {
            const int l = idx / half, e = 2 * (idx - l * half);
            const int v = r0 + l;
            if (v > 0 && v < H) {
                const cx<T> m0 = at(l, e == 0 ? 0 : n1 - e), m1 = at(l, n1 - e - 1);
                cxpair<T> om; om.a = mk<T>(m0.x, -m0.y); om.b = mk<T>(m1.x, -m1.y);
                *reinterpret_cast<cxpair<T>*>(ob + (size_t)(a.n0 - v) * n1 + e) = om;
            }
        }
        if (a.low_out) {
            // row-folded low-pass product: low_out[g][u][e] = sum_d U[u][e + d*m1] * phi[u][e + d*m1] for u = v and n0 - v
            const int m1 = a.low_m1, kf = n1 / m1;
            for (int idx = tid; idx < 2 * nl * m1; idx += kTmapComputeThreads) {
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
                            const cx<T> tv = at(l, mir ? (C == 0 ? 0 : n1 - C) : C);
                            const T f = fr[C];
                            ax += tv.x * f; ay += (mir ? -tv.y : tv.y) * f;
                        }
                    }
                }
                a.low_out[((size_t)g * a.n0 + u) * m1 + e] = mk<T>(ax, ay);
            }
        }
        tma::named_sync<1>(kTmapComputeThreads);
        if (tid == 0) tma::mbar_arrive(&done[b]);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Column pass of the full-resolution chain (k2d_colpass_imrf, kernels2d.cuh: column inverse, modulus, one complex forward
// transform per pair of real columns, Hermitian half stored) for 272 x 272 and 256 x 256 fields, staged by TMA tensor copies:
// a slab = 16 adjacent columns x NS rows = two boxes {16 complex = 128 B, NS/2 rows}, dense 128-byte rows in shared memory
// (the butterflies run with the 16 columns across a half-warp: conflict-free without padding, also for NS = 256); the
// NS/2 + 1 stored rows go back as one box.  Persistent CTAs, two buffers; thread 0 issues the copies of the NEXT slab before the transforms of the
// current one - no staging loops, no per-slab twiddle staging.
// ------------------------------------------------------------------------------------------------------------------
template <int NS> constexpr size_t imrf_tmap_smem_bytes() {
    return 128 + 2 * (size_t)NS * 16 * sizeof(cx<float>) + NS * sizeof(cx<float>) + 2 * sizeof(uint64_t);
}

template <int NS>
__global__ void __launch_bounds__(288, 3)
k2d_colpass_imrf_tmap(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map, ColArgs<float> a,
                      int nslabs) {
    using T = float;
    constexpr int n0 = NS, LP = 16, H = NS / 2;
    constexpr uint32_t kImrfSlabBytes = NS * 16 * sizeof(cx<float>);            // 34 816 at 272
    extern __shared__ unsigned char tmap_smem_raw[];
    unsigned char* base = tmap_smem_raw + ((128u - (tma::saddr(tmap_smem_raw) & 127u)) & 127u);
    cx<T>* tw = reinterpret_cast<cx<T>*>(base + 2 * kImrfSlabBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(tw + NS);
    const int tid = flat_tid(), nt = flat_nt();
    if (tid == 0) { tma::mbar_init(&full[0], 1); tma::mbar_init(&full[1], 1); tma::fence_mbar_init(); }
    for (int i = tid; i < NS; i += nt) tw[i] = a.tw[i];
    __syncthreads();
    const int nsl = a.n1 / kSLines;
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_my = first < nslabs ? (nslabs - first + stride - 1) / stride : 0;
    auto issue_load = [&](int i) {                              // thread 0 only
        const int t = first + i * stride, g = t / nsl, c0 = (t - g * nsl) * kSLines, b = i & 1;
        unsigned char* dst = base + b * kImrfSlabBytes;
        tma::mbar_arrive_expect_tx(&full[b], kImrfSlabBytes);
        tma::tensor_load_2d(dst, &in_map, 2 * c0, g * n0, &full[b]);
        tma::tensor_load_2d(dst + kImrfSlabBytes / 2, &in_map, 2 * c0, g * n0 + H, &full[b]);
    };
    if (tid == 0 && n_my > 0) issue_load(0);
    for (int i = 0; i < n_my; ++i) {
        const int t = first + i * stride, g = t / nsl, c0 = (t - g * nsl) * kSLines, b = i & 1;
        cx<T>* s = reinterpret_cast<cx<T>*>(base + b * kImrfSlabBytes);
        if (tid == 0 && i + 1 < n_my) {
            tma::bulk_wait_read<0>();                           // the store of slab i-1 has finished reading buffer b^1
            issue_load(i + 1);
        }
        tma::mbar_wait(&full[b], (i >> 1) & 1);
        slab_fft_s<NS, false, +1, 1, LP, T, 1, false>(s, kSLines, tw);          // inverse + modulus -> (|u|, 0), rows scrambled
        // pack column pairs: (|u|_{2m}, |u|_{2m+1}) -> one complex column at lane 2m
        for (int idx = tid; idx < n0 * (kSLines / 2); idx += nt) {
            const int e = idx / (kSLines / 2), l = 2 * (idx - e * (kSLines / 2));
            s[e * LP + l].y = s[e * LP + l + 1].x;
        }
        __syncthreads();
        slab_fft_s<NS, true, -1, 2, LP, T, 0, false>(s, kSLines / 2, tw);       // forward on the 8 packed columns
        // untangle in place: A[v] = (Z[v] + conj Z[n-v]) / 2,  B[v] = (Z[v] - conj Z[n-v]) / (2i);  rows v = 0..n0/2
        // (row v is read only by the thread that rewrites it; rows n0 - v > n0/2 are never written)
        for (int idx = tid; idx < (H + 1) * (kSLines / 2); idx += nt) {
            const int v = idx / (kSLines / 2), l = 2 * (idx - v * (kSLines / 2));
            const int vm = v == 0 ? 0 : n0 - v;
            const cx<T> z = s[v * LP + l], zm = s[vm * LP + l];
            cxpair<T> o;
            o.a = mk<T>(T(0.5) * (z.x + zm.x), T(0.5) * (z.y - zm.y));
            o.b = mk<T>(T(0.5) * (z.y + zm.y), T(0.5) * (zm.x - z.x));
            *reinterpret_cast<cxpair<T>*>(s + v * LP + l) = o;
        }
        tma::fence_proxy_async();
        __syncthreads();
        if (tid == 0) { tma::tensor_store_2d(&out_map, 2 * c0, g * n0, s); tma::bulk_commit(); }
    }
    if (tid == 0) tma::bulk_wait<0>();
}

namespace {
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
EncodeFn encode_fn() {
    static EncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeFn>(p);
    }();
    return fn;
}
// [rows][256 complex] float array as a 2-D tensor of floats, box = 16 complex x 16 rows, 128-byte swizzle
// [rows][n complex] float array, box = 16 complex x box_rows rows, dense (no swizzle)
bool encode_cols(CUtensorMap* map, const void* base, size_t rows, int n, unsigned box_rows) {
    EncodeFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)2 * n, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)2 * n * sizeof(float)};
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t es[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool encode_rows256(CUtensorMap* map, const void* base, size_t rows, unsigned box_rows = 16) {
    EncodeFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {512, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {512 * sizeof(float)};
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t es[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

bool rowprod_tmap256_launch(const RowProdArgs<float>& a, int Bp, int m, int npairs, cudaStream_t st) {
    if (a.n0 != 256 || a.n1 != 256 || a.k != 1) return false;
    if ((reinterpret_cast<uintptr_t>(a.parent) | reinterpret_cast<uintptr_t>(a.out)) & 15) return false;
    CUtensorMap in_map, out_map;
    if (!encode_rows256(&in_map, a.parent, (size_t)Bp * 256)) return false;
    if (!encode_rows256(&out_map, a.out, (size_t)Bp * a.NF * 256)) return false;
    k2d_rowprod_tmap256<<<(unsigned)(npairs * m), kTmapThreads, tmap_row_smem_bytes(), st>>>(in_map, out_map, a, Bp, m);
    return true;
}

bool rowfwdh_tmap256_launch(const RowArgs<float>& a, int G, int grid, cudaStream_t st) {
    if (a.n0 != 256 || a.n1 != 256 || a.in != a.out) return false;
    if (reinterpret_cast<uintptr_t>(a.in) & 15) return false;
    CUtensorMap map16, map1;
    if (!encode_rows256(&map16, a.in, (size_t)G * 256, 16)) return false;
    if (!encode_rows256(&map1, a.in, (size_t)G * 256, 1)) return false;
    k2d_rowfwdh_tmap256<<<(unsigned)grid, kTmapThreads, tmap_row_smem_bytes(), st>>>(map16, map1, a, G * 9);
    return true;
}

bool colpass_imrf_tmap_launch(const ColArgs<float>& a, int G, int ctas_per_sm, int num_sms, cudaStream_t st) {
    const int n = a.n0;
    if ((n != 272 && n != 256) || a.n1 != n || a.in != a.out) return false;
    if (reinterpret_cast<uintptr_t>(a.in) & 15) return false;
    CUtensorMap in_map, out_map;
    if (!encode_cols(&in_map, a.in, (size_t)G * n, n, n / 2)) return false;
    if (!encode_cols(&out_map, a.out, (size_t)G * n, n, n / 2 + 1)) return false;
    const int nslabs = G * (n / kSLines);
    const int grid = std::max(1, std::min(nslabs, ctas_per_sm * num_sms));
    if (n == 272)
        k2d_colpass_imrf_tmap<272><<<(unsigned)grid, dim3(16, 17), imrf_tmap_smem_bytes<272>(), st>>>(in_map, out_map, a, nslabs);
    else
        k2d_colpass_imrf_tmap<256><<<(unsigned)grid, dim3(16, 16), imrf_tmap_smem_bytes<256>(), st>>>(in_map, out_map, a, nslabs);
    return true;
}

void tmap_kernels_enable_smem() {
    enable_big_smem(k2d_colpass_imrf_tmap<272>); enable_big_smem(k2d_colpass_imrf_tmap<256>);
    enable_big_smem(k2d_rowprod_tmap256); enable_big_smem(k2d_rowfwdh_tmap256); }

}  // namespace sb
