// prims.cuh - the per-primitive backend protocol as stand-alone kernels (eager path).
//
// These mirror, one for one, the primitives the reference's core calls on its backend object
// (kymatio/scattering2d/core/scattering2d.py:3-9): Pad, rfft/ifft/irfft, cdgmm,
// subsample_fourier, modulus.  The fused engine does not use them; they exist so that the
// `torch_b200` backend object satisfies the whole protocol (kymatio_b200/kymatio_plugin.py) and so
// that the reference's primitive-level tests can be pointed at this library.
#pragma once
#include "common.cuh"
#include "kernels2d.cuh"
#include "plan_host.h"

namespace sb {

// reflect pad (kymatio/scattering2d/backend/torch_backend.py:36-86), real in -> real out
template <typename T>
__global__ void kp_pad2d(const T* __restrict__ x, T* __restrict__ out, int M, int N, int top, int left, int P0, int P1) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    const size_t b = blockIdx.z;
    if (c >= P1) return;
    out[(b * P0 + r) * P1 + c] = x[(b * M + reflect_idx(r - top, M)) * N + reflect_idx(c - left, N)];
}

// C[b][i] = A[b][i] * B[i]   (kymatio/backend/torch_backend.py:148-219); B real or complex
template <typename T>
__global__ void kp_cdgmm(const cx<T>* __restrict__ A, const T* __restrict__ B, cx<T>* __restrict__ out, size_t n,
                         size_t total, int b_complex) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t j = i % n;
    const cx<T> a = A[i];
    if (b_complex) {
        const cx<T> w = reinterpret_cast<const cx<T>*>(B)[j];
        out[i] = b_complex == 2 ? cmulc(a, w) : cmul(a, w);       // 2: multiply by conj(B) (adjoint of the complex product)
    } else {
        const T w = B[j];
        out[i] = mk<T>(a.x * w, a.y * w);
    }
}

// out[g][u][v] = mean_{a,b<k} in[g][u + a*n0/k][v + b*n1/k]   (torch_backend.py:93-129)
template <typename T>
__global__ void kp_periodize2d(const cx<T>* __restrict__ in, cx<T>* __restrict__ out, int n0, int n1, int k) {
    const int m0 = n0 / k, m1 = n1 / k;
    const int v = blockIdx.x * blockDim.x + threadIdx.x, u = blockIdx.y;
    const size_t g = blockIdx.z;
    if (v >= m1) return;
    T ax = T(0), ay = T(0);
    for (int a = 0; a < k; ++a)
        for (int b = 0; b < k; ++b) {
            const cx<T> t = in[(g * n0 + u + a * m0) * n1 + v + b * m1];
            ax += t.x; ay += t.y;
        }
    const T s = T(1) / (T(k) * T(k));
    out[(g * m0 + u) * m1 + v] = mk<T>(ax * s, ay * s);
}

// 1-D analogue: out[g][t] = mean_{a<k} in[g][t + a*n/k]   (kymatio/scattering1d/backend/torch_backend.py:19-48)
template <typename T>
__global__ void kp_periodize1d(const cx<T>* __restrict__ in, cx<T>* __restrict__ out, int n, int k) {
    const int m = n / k;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t g = blockIdx.y;
    if (t >= m) return;
    T ax = T(0), ay = T(0);
    for (int a = 0; a < k; ++a) { const cx<T> v = in[g * n + t + a * m]; ax += v.x; ay += v.y; }
    out[g * m + t] = mk<T>(ax / T(k), ay / T(k));
}

template <typename T> __global__ void kp_modulus(const cx<T>* __restrict__ in, T* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const cx<T> v = in[i]; out[i] = sqrt(v.x * v.x + v.y * v.y); }
}
template <typename T> __global__ void kp_from_real(const T* __restrict__ in, cx<T>* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = mk<T>(in[i], T(0));
}
template <typename T> __global__ void kp_real_scaled(const cx<T>* __restrict__ in, T* __restrict__ out, size_t n, T s) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i].x * s;
}
template <typename T> __global__ void kp_scale(cx<T>* __restrict__ x, size_t n, T s) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = scal(x[i], s);
}

// ---- adjoint / backward kernels ------------------------------------------------
// broadcast filter multiply: out[b][f][i] = A[b][i] * W[f][i]   (W real), i < n
template <typename T>
__global__ void kp_cdgmm_bcast(const cx<T>* __restrict__ A, const T* __restrict__ W, cx<T>* __restrict__ out,
                               size_t nb, int nf, size_t n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nb * n) return;
    const size_t b = idx / n, i = idx - b * n;
    const cx<T> a = A[idx];
    for (int f = 0; f < nf; ++f) {
        const T w = W[(size_t)f * n + i];
        out[(b * nf + f) * n + i] = mk<T>(a.x * w, a.y * w);
    }
}
// its adjoint w.r.t. A: gA[b][i] = sum_f g[b][f][i] * W[f][i]
template <typename T>
__global__ void kp_cdgmm_bcast_bwd(const cx<T>* __restrict__ g, const T* __restrict__ W, cx<T>* __restrict__ gA,
                                   size_t nb, int nf, size_t n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nb * n) return;
    const size_t b = idx / n, i = idx - b * n;
    T ax = T(0), ay = T(0);
    for (int f = 0; f < nf; ++f) {
        const T w = W[(size_t)f * n + i];
        const cx<T> v = g[(b * nf + f) * n + i];
        ax += v.x * w; ay += v.y * w;
    }
    gA[idx] = mk<T>(ax, ay);
}
// adjoint of the Fourier periodisation: gin[g][u + a*m0][v + b*m1] = gout[g][u][v] / k^2
template <typename T>
__global__ void kp_periodize2d_bwd(const cx<T>* __restrict__ gout, cx<T>* __restrict__ gin, int n0, int n1, int k) {
    const int m0 = n0 / k, m1 = n1 / k;
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    const size_t g = blockIdx.z;
    if (c >= n1) return;
    const T s = T(1) / (T(k) * T(k));
    gin[(g * n0 + r) * n1 + c] = scal(gout[(g * m0 + r % m0) * m1 + c % m1], s);
}
// modulus backward (kymatio/backend/torch_backend.py:85-96): gx = x * g / |x|, 0 where |x| = 0
template <typename T>
__global__ void kp_modulus_bwd(const cx<T>* __restrict__ x, const T* __restrict__ g, cx<T>* __restrict__ gx, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cx<T> v = x[i];
    const T m = sqrt(v.x * v.x + v.y * v.y);
    const T s = m > T(0) ? g[i] / m : T(0);
    gx[i] = mk<T>(v.x * s, v.y * s);
}
// adjoint of reflect padding: fold-add every padded sample onto its source pixel
template <typename T>
__global__ void kp_pad2d_bwd(const T* __restrict__ gout, T* __restrict__ gx, int M, int N, int top, int left, int P0,
                             int P1) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    const size_t b = blockIdx.z;
    if (c >= P1) return;
    atomicAdd(&gx[(b * M + reflect_idx(r - top, M)) * N + reflect_idx(c - left, N)], gout[(b * P0 + r) * P1 + c]);
}

// adjoint of the 1-D Fourier periodisation: gin[g][t] = gout[g][t mod (n/k)] / k
template <typename T>
__global__ void kp_periodize1d_bwd(const cx<T>* __restrict__ gout, cx<T>* __restrict__ gin, int n, int k) {
    const int m = n / k;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t g = blockIdx.y;
    if (t >= n) return;
    gin[g * n + t] = scal(gout[g * m + t % m], T(1) / T(k));
}
// adjoint of out = sqrt(prev^2 + |x|^2): gx = x g / out, gprev = prev g / out (0 where out = 0); prev / gprev may be null
template <typename T>
__global__ void kp_modrot_bwd(const cx<T>* __restrict__ x, const T* __restrict__ prev, const T* __restrict__ out,
                              const T* __restrict__ g, cx<T>* __restrict__ gx, T* __restrict__ gprev, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T o = out[i];
    const T s = o > T(0) ? g[i] / o : T(0);
    const cx<T> v = x[i];
    gx[i] = mk<T>(v.x * s, v.y * s);
    if (gprev) gprev[i] = prev[i] * s;
}
// adjoint of the integrals out[b][p] = sum_i x[b][i]^q_p: gx[b][i] = sum_p g[b][p] q_p x^(q_p - 1)
template <typename T>
__global__ void kp_integrals_bwd(const T* __restrict__ x, const T* __restrict__ g, T* __restrict__ gx, size_t n,
                                 const float* __restrict__ powers, int P) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t b = blockIdx.y;
    if (i >= n) return;
    const T v = x[b * n + i];
    T acc = T(0);
    for (int p = 0; p < P; ++p) {
        const T q = T(powers[p]);
        const T d = q == T(1) ? T(1) : q == T(2) ? T(2) * v : (v > T(0) ? q * pow(v, q - T(1)) : T(0));
        acc += g[b * P + p] * d;
    }
    gx[b * n + i] = acc;
}

inline unsigned blocks_for(size_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

// natural-order 2-D complex FFT tables: [tw0 | pos0 | tw1 | pos1]
template <typename T> struct Fft2dTables {
    Plan1 p0, p1;
    size_t tw0, pos0, tw1, pos1, bytes;
    Fft2dTables(int n0, int n1) {
        p0 = make_plan1(n0); p1 = make_plan1(n1);
        size_t off = 0;
        auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
        tw0 = take((size_t)n0 * sizeof(cx<T>)); pos0 = take((size_t)n0 * sizeof(int));
        tw1 = take((size_t)n1 * sizeof(cx<T>)); pos1 = take((size_t)n1 * sizeof(int));
        bytes = off;
    }
};

template <typename T> void fft2d_init(void* const_dev, int n0, int n1, cudaStream_t st) {
    Fft2dTables<T> t(n0, n1);
    std::vector<unsigned char> h(t.bytes, 0);
    auto tw0 = twiddle_table<T>(n0); auto tw1 = twiddle_table<T>(n1);
    auto p0 = scramble_table(t.p0); auto p1 = scramble_table(t.p1);
    memcpy(h.data() + t.tw0, tw0.data(), (size_t)n0 * sizeof(cx<T>));
    memcpy(h.data() + t.pos0, p0.data(), (size_t)n0 * sizeof(int));
    memcpy(h.data() + t.tw1, tw1.data(), (size_t)n1 * sizeof(cx<T>));
    memcpy(h.data() + t.pos1, p1.data(), (size_t)n1 * sizeof(int));
    SB_CUDA(cudaMemcpyAsync(const_dev, h.data(), t.bytes, cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaStreamSynchronize(st));
}

// in/out: (G, n0, n1) complex, natural order both sides; mode 0 = forward, 1 = inverse 1/(n0 n1)-normalised
// (torch.fft.fft2 / ifft2 conventions, kymatio/scattering2d/backend/torch_backend.py:10-12), 2 = inverse without the
// normalisation (the adjoint of the forward transform, used by the autograd graph)
template <typename T>
void fft2d_exec(const void* const_dev, const void* in, void* out, int64_t G, int n0, int n1, int mode, cudaStream_t st) {
    const bool inverse = mode != 0;
    Fft2dTables<T> t(n0, n1);
    const unsigned char* cb = static_cast<const unsigned char*>(const_dev);
    const SlabCfg rc = slab_cfg(t.p1, n0, sizeof(cx<T>), sizeof(int));
    const SlabCfg cc = slab_cfg(t.p0, n1, sizeof(cx<T>), sizeof(int));
    // compile-time specialised instances keep natural order on both sides for the column passes and the inverse
    // row pass (only the static FORWARD row pass belongs to the scrambled-row chain of plan2d.cuh)
    const StreamKernels<T> kr = stream_kernels_lookup<T>(n1, inverse && rc.lines == kSLines && rc.LP == kSLP);
    const StreamKernels<T> kc = stream_kernels_lookup<T>(n0, cc.lines == kSLines && cc.LP == kSLP);
    RowArgs<T> ra{};
    ra.in = static_cast<const cx<T>*>(in); ra.out = static_cast<cx<T>*>(out); ra.n0 = n0; ra.n1 = n1;
    ra.lines = rc.lines; ra.LP = rc.LP; ra.plan = t.p1;
    ra.tw = reinterpret_cast<const cx<T>*>(cb + t.tw1); ra.pos = reinterpret_cast<const int*>(cb + t.pos1);
    dim3 gr((unsigned)G, ceil_div(n0, rc.lines));
    launch(inverse ? "prim_rowpass_inv" : "prim_rowpass_fwd", 2.0 * G * n0 * n1 * sizeof(cx<T>), st,
           [&] { (inverse ? kr.row_inv : kr.row_fwd)<<<gr, rc.block, rc.smem, st>>>(ra); });
    ColArgs<T> ca{};
    ca.in = static_cast<cx<T>*>(out); ca.out = static_cast<cx<T>*>(out); ca.n0 = n0; ca.n1 = n1;
    ca.lines = cc.lines; ca.LP = cc.LP; ca.plan = t.p0;
    ca.tw = reinterpret_cast<const cx<T>*>(cb + t.tw0); ca.pos = reinterpret_cast<const int*>(cb + t.pos0);
    dim3 gc((unsigned)G, ceil_div(n1, cc.lines));
    launch(inverse ? "prim_colpass_inv" : "prim_colpass_fwd", 2.0 * G * n0 * n1 * sizeof(cx<T>), st,
           [&] { (inverse ? kc.col_inv : kc.col_fwd)<<<gc, cc.block, cc.smem, st>>>(ca); });
    if (mode == 1) {
        const size_t n = (size_t)G * n0 * n1;
        launch("prim_scale", 2.0 * n * sizeof(cx<T>), st,
               [&] { kp_scale<T><<<blocks_for(n), 256, 0, st>>>(static_cast<cx<T>*>(out), n, T(1) / (T(n0) * T(n1))); });
    }
}

// ---------------------------------------------------------------------------------------------------
// 1-D primitives (kymatio/scattering1d/backend/torch_backend.py)
// ---------------------------------------------------------------------------------------------------
// reflect pad along the last axis, (G, N) -> (G, N + pl + pr)   (torch_backend.py:51-82)
template <typename T>
__global__ void kp_pad1d(const T* __restrict__ x, T* __restrict__ out, int N, int pl, int P) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t g = blockIdx.y;
    if (c < P) out[g * P + c] = x[g * N + reflect_idx(c - pl, N)];
}

// Second half of the four-step 1-D FFT of length N = Na * Nb (first half: column transform of the
// [Na][Nb] view): multiply row ka by w_N^{-+ ka*jb}, transform along jb, store X[ka + Na*kb].
template <typename T> struct RowTwArgs {
    const cx<T>* in; cx<T>* out;
    int Na, Nb;
    int lines, LP;
    T scale;
    Plan1 plan; const cx<T>* tw; const int* pos;
};
__device__ __forceinline__ void sincos2pi(float t, float* s, float* c) { sincospif(2.0f * t, s, c); }
__device__ __forceinline__ void sincos2pi(double t, double* s, double* c) { sincospi(2.0 * t, s, c); }

template <typename T, bool INV> __global__ void __launch_bounds__(kMaxThreads) k1d_rowpass_tw(RowTwArgs<T> a) {
    cx<T>* s = dyn_smem<cx<T>>();
    cx<T>* tw = s + (size_t)a.Nb * a.LP;
    int* pos = reinterpret_cast<int*>(tw + a.Nb);
    const size_t g = blockIdx.x;
    const int r0 = blockIdx.y * a.lines;
    const int nl = min(a.lines, a.Na - r0);
    const size_t N = (size_t)a.Na * a.Nb;
    stage(tw, a.tw, a.Nb);
    stage(pos, a.pos, a.Nb);
    __syncthreads();
    const cx<T>* ib = a.in + g * N + (size_t)r0 * a.Nb;
    for (int idx = flat_tid(); idx < nl * a.Nb; idx += flat_nt()) {
        const int l = idx / a.Nb, e = idx - l * a.Nb;
        cx<T> v = ib[(size_t)l * a.Nb + e];
        T sn, cs;
        sincos2pi(T((long long)(r0 + l) * e) / T(N), &sn, &cs);
        v = INV ? cmul(v, mk<T>(cs, sn)) : cmul(v, mk<T>(cs, -sn));
        s[(INV ? pos[e] : e) * a.LP + l] = scal(v, a.scale);
    }
    __syncthreads();
    slab_fft<INV, T>(s, nl, 1, a.LP, a.plan, tw);
    cx<T>* ob = a.out + g * N + r0;
    for (int idx = flat_tid(); idx < nl * a.Nb; idx += flat_nt()) {
        const int e = idx / nl, l = idx - e * nl;
        ob[(size_t)e * a.Na + l] = s[(INV ? e : pos[e]) * a.LP + l];
    }
}

// tables: [twA | posA | twB | posB] for N = Na * Nb (balanced power-of-two or general split)
template <typename T> struct Fft1dTables {
    int Na, Nb;
    Plan1 pa, pb;
    size_t twa, posa, twb, posb, bytes;
    explicit Fft1dTables(int N) {
        // Na = largest divisor of N not above sqrt(N) (1 for primes); both halves must fit a slab line
        Na = 1;
        for (int d = 1; (long long)d * d <= N; ++d) if (N % d == 0) Na = d;
        Nb = N / Na;
        pa = make_plan1(Na); pb = make_plan1(Nb);
        size_t off = 0;
        auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
        twa = take((size_t)Na * sizeof(cx<T>)); posa = take((size_t)Na * sizeof(int));
        twb = take((size_t)Nb * sizeof(cx<T>)); posb = take((size_t)Nb * sizeof(int));
        bytes = off;
    }
};
template <typename T> void fft1d_init(void* const_dev, int N, cudaStream_t st) {
    Fft1dTables<T> t(N);
    std::vector<unsigned char> h(t.bytes, 0);
    auto twa = twiddle_table<T>(t.Na); auto twb = twiddle_table<T>(t.Nb);
    auto pa = scramble_table(t.pa); auto pb = scramble_table(t.pb);
    memcpy(h.data() + t.twa, twa.data(), (size_t)t.Na * sizeof(cx<T>));
    if (t.Na > 1) memcpy(h.data() + t.posa, pa.data(), (size_t)t.Na * sizeof(int));
    memcpy(h.data() + t.twb, twb.data(), (size_t)t.Nb * sizeof(cx<T>));
    if (t.Nb > 1) memcpy(h.data() + t.posb, pb.data(), (size_t)t.Nb * sizeof(int));
    SB_CUDA(cudaMemcpyAsync(const_dev, h.data(), t.bytes, cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaStreamSynchronize(st));
}
// in, tmp, out: (G, N) complex; natural order both sides; inverse is 1/N-normalised
// (torch.fft.fft / ifft, kymatio/scattering1d/backend/torch_backend.py:8-10)
template <typename T>
void fft1d_exec(const void* const_dev, const void* in, void* tmp, void* out, int64_t G, int N, bool inverse,
                cudaStream_t st) {
    once_per_device("fft1d", [] {
        enable_big_smem(k1d_rowpass_tw<float, false>); enable_big_smem(k1d_rowpass_tw<float, true>);
        enable_big_smem(k1d_rowpass_tw<double, false>); enable_big_smem(k1d_rowpass_tw<double, true>);
    });
    Fft1dTables<T> t(N);
    const unsigned char* cb = static_cast<const unsigned char*>(const_dev);
    const cx<T>* src = static_cast<const cx<T>*>(in);
    if (t.Na > 1) {
        const SlabCfg cc = slab_cfg(t.pa, t.Nb, sizeof(cx<T>), sizeof(int));
        const StreamKernels<T> kc = stream_kernels_lookup<T>(t.Na, false);
        ColArgs<T> ca{};
        ca.in = src; ca.out = static_cast<cx<T>*>(tmp); ca.n0 = t.Na; ca.n1 = t.Nb;
        ca.lines = cc.lines; ca.LP = cc.LP; ca.plan = t.pa;
        ca.tw = reinterpret_cast<const cx<T>*>(cb + t.twa); ca.pos = reinterpret_cast<const int*>(cb + t.posa);
        dim3 gc((unsigned)G, ceil_div(t.Nb, cc.lines));
        launch(inverse ? "prim1d_colpass_inv" : "prim1d_colpass_fwd", 2.0 * G * N * sizeof(cx<T>), st,
               [&] { (inverse ? kc.col_inv : kc.col_fwd)<<<gc, cc.block, cc.smem, st>>>(ca); });
        src = static_cast<const cx<T>*>(tmp);
    }
    const SlabCfg rc = slab_cfg(t.pb, t.Na, sizeof(cx<T>), sizeof(int));
    RowTwArgs<T> ra{};
    ra.in = src; ra.out = static_cast<cx<T>*>(out); ra.Na = t.Na; ra.Nb = t.Nb;
    ra.lines = rc.lines; ra.LP = rc.LP; ra.scale = inverse ? T(1) / T(N) : T(1); ra.plan = t.pb;
    ra.tw = reinterpret_cast<const cx<T>*>(cb + t.twb); ra.pos = reinterpret_cast<const int*>(cb + t.posb);
    dim3 gr((unsigned)G, ceil_div(t.Na, rc.lines));
    launch(inverse ? "prim1d_rowpass_tw_inv" : "prim1d_rowpass_tw_fwd", 2.0 * G * N * sizeof(cx<T>), st, [&] {
        if (inverse) k1d_rowpass_tw<T, true><<<gr, rc.block, rc.smem, st>>>(ra);
        else k1d_rowpass_tw<T, false><<<gr, rc.block, rc.smem, st>>>(ra);
    });
}

// ---------------------------------------------------------------------------------------------------
// 3-D primitives (kymatio/scattering3d/backend/torch_backend.py)
// ---------------------------------------------------------------------------------------------------
// sqrt(prev^2 + |x|^2): the running rotation-covariant modulus (torch_backend.py:102-124); prev may be null
template <typename T>
__global__ void kp_modulus_rotation(const cx<T>* __restrict__ x, const T* __restrict__ prev, T* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cx<T> v = x[i];
    T m2 = v.x * v.x + v.y * v.y;
    if (prev) { const T p = prev[i]; m2 += p * p; }
    out[i] = sqrt(m2);
}
// integrals[b][p] += sum_i x[b][i]^q_p  (torch_backend.py:127-151); out must be zeroed by the caller
template <typename T>
__global__ void kp_integrals(const T* __restrict__ x, double* __restrict__ out, size_t n, const float* __restrict__ powers,
                             int P) {
    __shared__ double red[8][32];
    const size_t b = blockIdx.y;
    double acc[8];
    for (int p = 0; p < 8; ++p) acc[p] = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const T v = x[b * n + i];
        for (int p = 0; p < P; ++p) {
            const float q = powers[p];
            acc[p] += q == 1.f ? (double)v : q == 2.f ? (double)v * (double)v : pow((double)v, (double)q);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int p = 0; p < P; ++p) {
        double a = acc[p];
        for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
        if (lane == 0) red[p][warp] = a;
    }
    __syncthreads();
    if (warp == 0) {
        for (int p = 0; p < P; ++p) {
            double a = lane < (int)(blockDim.x >> 5) ? red[p][lane] : 0.0;
            for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
            if (lane == 0) atomicAdd(&out[b * P + p], a);
        }
    }
}

// natural-order complex 3-D FFT on (G, M, N, O): one row pass (O) and two column passes (N, then M)
template <typename T> struct Fft3dTables {
    int n[3]; Plan1 p[3]; size_t tw[3], pos[3], bytes;
    Fft3dTables(int M, int N, int O) {
        n[0] = M; n[1] = N; n[2] = O;
        size_t off = 0;
        auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
        for (int a = 0; a < 3; ++a) {
            p[a] = make_plan1(n[a]);
            tw[a] = take((size_t)n[a] * sizeof(cx<T>)); pos[a] = take((size_t)n[a] * sizeof(int));
        }
        bytes = off;
    }
};
template <typename T> void fft3d_init(void* const_dev, int M, int N, int O, cudaStream_t st) {
    Fft3dTables<T> t(M, N, O);
    std::vector<unsigned char> h(t.bytes, 0);
    for (int a = 0; a < 3; ++a) {
        auto tw = twiddle_table<T>(t.n[a]); auto ps = scramble_table(t.p[a]);
        memcpy(h.data() + t.tw[a], tw.data(), (size_t)t.n[a] * sizeof(cx<T>));
        if (t.n[a] > 1) memcpy(h.data() + t.pos[a], ps.data(), (size_t)t.n[a] * sizeof(int));
    }
    SB_CUDA(cudaMemcpyAsync(const_dev, h.data(), t.bytes, cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaStreamSynchronize(st));
}
template <typename T>
void fft3d_exec(const void* const_dev, const void* in, void* out, int64_t G, int M, int N, int O, bool inverse, cudaStream_t st) {
    Fft3dTables<T> t(M, N, O);
    const unsigned char* cb = static_cast<const unsigned char*>(const_dev);
    auto twp = [&](int a) { return reinterpret_cast<const cx<T>*>(cb + t.tw[a]); };
    auto posp = [&](int a) { return reinterpret_cast<const int*>(cb + t.pos[a]); };
    // along O: rows of the (G*M, N, O) view
    {
        const SlabCfg rc = slab_cfg(t.p[2], N, sizeof(cx<T>), sizeof(int));
        const StreamKernels<T> k = stream_kernels_lookup<T>(O, false);
        RowArgs<T> ra{};
        ra.in = static_cast<const cx<T>*>(in); ra.out = static_cast<cx<T>*>(out); ra.n0 = N; ra.n1 = O;
        ra.lines = rc.lines; ra.LP = rc.LP; ra.plan = t.p[2]; ra.tw = twp(2); ra.pos = posp(2);
        dim3 g((unsigned)(G * M), ceil_div(N, rc.lines));
        launch("prim3d_rowpass", 2.0 * G * M * N * O * sizeof(cx<T>), st,
               [&] { (inverse ? k.row_inv : k.row_fwd)<<<g, rc.block, rc.smem, st>>>(ra); });
    }
    // along N: columns of the (G*M, N, O) view; along M: columns of the (G, M, N*O) view
    for (int pass = 0; pass < 2; ++pass) {
        const int ax = pass == 0 ? 1 : 0;
        const int n0 = t.n[ax], n1 = pass == 0 ? O : N * O;
        const int64_t g = pass == 0 ? G * M : G;
        const SlabCfg cc = slab_cfg(t.p[ax], n1, sizeof(cx<T>), sizeof(int));
        const StreamKernels<T> k = stream_kernels_lookup<T>(n0, false);
        ColArgs<T> ca{};
        ca.in = static_cast<cx<T>*>(out); ca.out = static_cast<cx<T>*>(out); ca.n0 = n0; ca.n1 = n1;
        ca.lines = cc.lines; ca.LP = cc.LP; ca.plan = t.p[ax]; ca.tw = twp(ax); ca.pos = posp(ax);
        dim3 grid((unsigned)g, ceil_div(n1, cc.lines));
        launch("prim3d_colpass", 2.0 * G * M * N * O * sizeof(cx<T>), st,
               [&] { (inverse ? k.col_inv : k.col_fwd)<<<grid, cc.block, cc.smem, st>>>(ca); });
    }
    if (inverse) {
        const size_t n = (size_t)G * M * N * O;
        launch("prim_scale", 2.0 * n * sizeof(cx<T>), st, [&] {
            kp_scale<T><<<blocks_for(n), 256, 0, st>>>(static_cast<cx<T>*>(out), n, T(1) / (T(M) * T(N) * T(O)));
        });
    }
}

}  // namespace sb
