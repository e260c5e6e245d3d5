// scat_b200.cu - extern "C" entry points of libscat_b200.so (see include/scat_b200.h).
#include <cstdio>
#include <cstring>
#include "plan2d.cuh"
#include "prims.cuh"
#include "plan1d.cuh"
#include "plan3d.cuh"

using namespace sb;

namespace {
template <typename F> int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { last_error() = e.what(); return 1; }
    catch (...) { last_error() = "unknown error"; return 2; }
}
}  // namespace

// grid.y / grid.z are capped at 65535: entry points that put the batch there split larger batches into slices
constexpr int64_t kMaxGridYZ = 65535;
static inline size_t esize(int32_t dtype) { return dtype == 1 ? 8 : 4; }
static inline const void* at(const void* p, size_t bytes) { return p ? static_cast<const char*>(p) + bytes : nullptr; }
static inline void* at(void* p, size_t bytes) { return p ? static_cast<char*>(p) + bytes : nullptr; }
#define SB_SPLIT_BATCH(B, CALL)                                                        \
    if ((B) > kMaxGridYZ) {                                                            \
        for (int64_t b0 = 0; b0 < (B); b0 += kMaxGridYZ) {                             \
            const int64_t nb = std::min<int64_t>(kMaxGridYZ, (B) - b0);                \
            const int rc = (CALL);                                                     \
            if (rc) return rc;                                                         \
        }                                                                              \
        return 0;                                                                      \
    }

extern "C" {

int scat_version(void) { return 100; }
const char* scat_last_error(void) { return last_error().c_str(); }
uint64_t scat_launch_count(void) { return launch_counter().load(); }

// profiling build only (libscat_b200_prof.so): per-phase cycle counters of the 2-D tile kernels, [24 kinds][8 phases];
// returns the number of slots written (0 in the production build).  Synchronises the device.
int scat_phase_prof_read(unsigned long long* out, int max_n, int reset) {
    int n = 0;
    guarded([&] { n = phase_prof_read(out, max_n, reset != 0); });
    return n;
}

void scat_timing_enable(int on) {
    timing_on() = on != 0;
    if (!on) {
        for (auto& r : timing_recs()) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        timing_recs().clear();
    }
}

// Synchronises the recorded events and writes "label\tcount\ttotal_ms\ttotal_bytes\n" lines
// (aggregated per label) into buf; returns the number of bytes needed (excluding the NUL).
size_t scat_timing_report(char* buf, size_t buflen) {
    std::vector<std::string> labels; std::vector<double> ms, bytes; std::vector<long> cnt;
    for (auto& r : timing_recs()) {
        float t = 0.f;
        cudaEventSynchronize(r.e1);
        cudaEventElapsedTime(&t, r.e0, r.e1);
        size_t i = 0;
        for (; i < labels.size(); ++i) if (labels[i] == r.label) break;
        if (i == labels.size()) { labels.push_back(r.label); ms.push_back(0); bytes.push_back(0); cnt.push_back(0); }
        ms[i] += t; bytes[i] += r.bytes; cnt[i] += 1;
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    timing_recs().clear();
    std::string out;
    for (size_t i = 0; i < labels.size(); ++i) {
        char line[512];
        snprintf(line, sizeof line, "%s\t%ld\t%.6f\t%.0f\n", labels[i].c_str(), cnt[i], ms[i], bytes[i]);
        out += line;
    }
    if (buf && buflen) {
        size_t n = std::min(buflen - 1, out.size());
        memcpy(buf, out.data(), n); buf[n] = 0;
    }
    return out.size();
}

int scat_plan2d_create(const scat_plan2d_desc* desc, scat_plan2d** out_plan) {
    return guarded([&] {
        if (!desc || !out_plan) throw std::runtime_error("null argument");
        if (desc->dtype == 0) *out_plan = new Plan2D<float>(*desc);
        else if (desc->dtype == 1) *out_plan = new Plan2D<double>(*desc);
        else throw std::runtime_error("dtype must be 0 (float32) or 1 (float64)");
    });
}
void scat_plan2d_destroy(scat_plan2d* plan) { delete plan; }

int scat_plan2d_info(const scat_plan2d* plan, int32_t* Mp, int32_t* Np, int32_t* out_h, int32_t* out_w,
                     int32_t* K) {
    return guarded([&] {
        if (!plan) throw std::runtime_error("null plan");
        plan->info(Mp, Np, out_h, out_w, K);
    });
}
size_t scat_plan2d_const_bytes(const scat_plan2d* plan) { return plan ? plan->const_bytes() : 0; }

int scat_plan2d_bind(scat_plan2d* plan, void* const_dev, const void* const* phi_dev, int32_t n_phi,
                     const void* const* psi_dev, int32_t n_psi, void* stream) {
    return guarded([&] {
        if (!plan) throw std::runtime_error("null plan");
        plan->bind(const_dev, phi_dev, n_phi, psi_dev, n_psi, static_cast<cudaStream_t>(stream));
    });
}
size_t scat_plan2d_workspace_bytes(const scat_plan2d* plan, int64_t batch) {
    return plan ? plan->workspace_bytes(batch) : 0;
}
int scat_plan2d_forward(scat_plan2d* plan, const void* x_dev, void* out_dev, void* workspace_dev,
                        size_t workspace_bytes, int64_t batch, void* stream) {
    return guarded([&] {
        if (!plan) throw std::runtime_error("null plan");
        plan->forward(x_dev, out_dev, workspace_dev, workspace_bytes, batch, static_cast<cudaStream_t>(stream));
    });
}
int scat_plan2d_forward_peers(scat_plan2d* plan, const void* x_dev, void* out_dev, void* const* peer_out_dev, int32_t n_peers,
                              void* ws_dev, size_t ws_bytes, int64_t batch, void* stream) {
    return guarded([&] {
        if (!plan) throw std::runtime_error("null plan");
        plan->forward_peers(x_dev, out_dev, peer_out_dev, n_peers, ws_dev, ws_bytes, batch, static_cast<cudaStream_t>(stream));
    });
}

int scat_plan2d_forward_save(scat_plan2d* plan, const void* x_dev, void* out_dev, void* const* saved_u1_dev, void* ws_dev,
                             size_t ws_bytes, int64_t batch, void* stream) {
    return guarded([&] {
        if (!plan) throw std::runtime_error("null plan");
        if (!saved_u1_dev) throw std::runtime_error("forward_save: null table of U1 buffers");
        plan->forward_save(x_dev, out_dev, saved_u1_dev, ws_dev, ws_bytes, batch, static_cast<cudaStream_t>(stream));
    });
}

int32_t scat_plan2d_order2_channels(const scat_plan2d* plan, int32_t j1) { return plan ? plan->order2_channels(j1) : 0; }
int scat_plan2d_order2_forward(scat_plan2d* plan, int32_t j1, const void* u1_dev, void* out_dev, int64_t batch, void* stream) {
    return guarded([&] {
        if (!plan) throw std::runtime_error("null plan");
        plan->order2_forward(j1, u1_dev, out_dev, batch, static_cast<cudaStream_t>(stream));
    });
}
int32_t scat_plan2d_order1_mode(const scat_plan2d* plan, int32_t j1) { return plan ? plan->order1_mode(j1) : 0; }
size_t scat_plan2d_order1_workspace_bytes(const scat_plan2d* plan, int32_t j1, int64_t batch) {
    return plan ? plan->order1_workspace_bytes(j1, batch) : 0;
}
int scat_plan2d_order1_forward(scat_plan2d* plan, int32_t j1, const void* u0_dev, void* s1_dev, void* u1_dev, int64_t batch,
                               void* stream) {
    return guarded([&] {
        if (!plan) throw std::runtime_error("null plan");
        plan->order1_forward(j1, u0_dev, s1_dev, u1_dev, batch, static_cast<cudaStream_t>(stream));
    });
}
int scat_plan2d_order1_backward(scat_plan2d* plan, int32_t j1, const void* u0_dev, const void* gs1_dev, const void* gu1_dev,
                                void* gu0_dev, void* ws_dev, size_t ws_bytes, int64_t batch, void* stream) {
    return guarded([&] {
        if (!plan) throw std::runtime_error("null plan");
        plan->order1_backward(j1, u0_dev, gs1_dev, gu1_dev, gu0_dev, ws_dev, ws_bytes, batch, static_cast<cudaStream_t>(stream));
    });
}
int scat_plan2d_order2_backward(scat_plan2d* plan, int32_t j1, const void* u1_dev, const void* gout_dev, void* gu1_dev,
                                int64_t batch, void* stream) {
    return guarded([&] {
        if (!plan) throw std::runtime_error("null plan");
        plan->order2_backward(j1, u1_dev, gout_dev, gu1_dev, batch, static_cast<cudaStream_t>(stream));
    });
}

// ---------------------------------------------------------------- eager primitives
#define SB_DISPATCH(dtype, CALL)                                                     \
    do {                                                                             \
        if ((dtype) == 0) { typedef float T; CALL; }                                 \
        else if ((dtype) == 1) { typedef double T; CALL; }                           \
        else throw std::runtime_error("dtype must be 0 (float32) or 1 (float64)");   \
    } while (0)

size_t scat_fft2d_const_bytes(int32_t n0, int32_t n1, int32_t dtype) {
    try {
        return dtype == 1 ? Fft2dTables<double>(n0, n1).bytes : Fft2dTables<float>(n0, n1).bytes;
    } catch (const std::exception& e) { last_error() = e.what(); return 0; }
}
int scat_fft2d_init(void* const_dev, int32_t n0, int32_t n1, int32_t dtype, void* stream) {
    return guarded([&] { SB_DISPATCH(dtype, fft2d_init<T>(const_dev, n0, n1, static_cast<cudaStream_t>(stream))); });
}
int scat_fft2d_exec(const void* const_dev, const void* in_dev, void* out_dev, int64_t G, int32_t n0, int32_t n1,
                    int32_t inverse, int32_t dtype, void* stream) {
    return guarded([&] {
        once_per_device("stream", [] { stream_kernels_enable_smem<float>(); stream_kernels_enable_smem<double>(); });
        SB_DISPATCH(dtype, fft2d_exec<T>(const_dev, in_dev, out_dev, G, n0, n1, inverse, static_cast<cudaStream_t>(stream)));
    });
}
int scat_pad2d(const void* x_dev, void* out_dev, int64_t B, int32_t M, int32_t N, int32_t top, int32_t bottom,
               int32_t left, int32_t right, int32_t dtype, void* stream) {
    SB_SPLIT_BATCH(B, scat_pad2d(at(x_dev, (size_t)b0 * M * N * esize(dtype)),
                                 at(out_dev, (size_t)b0 * (M + top + bottom) * (N + left + right) * esize(dtype)), nb, M, N, top,
                                 bottom, left, right, dtype, stream))
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int P0 = M + top + bottom, P1 = N + left + right;
        if (B <= 0) return;
        dim3 grid(ceil_div(P1, 128), P0, (unsigned)B);
        SB_DISPATCH(dtype, launch("prim_pad2d", (double)B * (M * N + P0 * P1) * sizeof(T), st, [&] {
            kp_pad2d<T><<<grid, 128, 0, st>>>(static_cast<const T*>(x_dev), static_cast<T*>(out_dev), M, N, top, left, P0, P1);
        }));
    });
}
int scat_cdgmm(const void* a_dev, const void* b_dev, void* out_dev, int64_t batch, int64_t n, int32_t b_is_complex,
               int32_t dtype, void* stream) {
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const size_t total = (size_t)batch * n;
        if (!total) return;
        SB_DISPATCH(dtype, launch("prim_cdgmm", 2.0 * total * sizeof(cx<T>), st, [&] {
            kp_cdgmm<T><<<blocks_for(total), 256, 0, st>>>(static_cast<const cx<T>*>(a_dev), static_cast<const T*>(b_dev),
                                                          static_cast<cx<T>*>(out_dev), (size_t)n, total, b_is_complex);
        }));
    });
}
int scat_subsample_fourier2d(const void* in_dev, void* out_dev, int64_t G, int32_t n0, int32_t n1, int32_t k,
                             int32_t dtype, void* stream) {
    SB_SPLIT_BATCH(G, scat_subsample_fourier2d(at(in_dev, (size_t)b0 * n0 * n1 * 2 * esize(dtype)),
                                               at(out_dev, (size_t)b0 * (n0 / std::max(k, 1)) * (n1 / std::max(k, 1)) * 2 * esize(dtype)),
                                               nb, n0, n1, k, dtype, stream))
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (k < 1 || n0 % k || n1 % k) throw std::runtime_error("subsample_fourier: k must divide both sizes");
        if (G <= 0) return;
        dim3 grid(ceil_div(n1 / k, 128), n0 / k, (unsigned)G);
        SB_DISPATCH(dtype, launch("prim_periodize2d", (double)G * n0 * n1 * sizeof(cx<T>), st, [&] {
            kp_periodize2d<T><<<grid, 128, 0, st>>>(static_cast<const cx<T>*>(in_dev), static_cast<cx<T>*>(out_dev), n0, n1, k);
        }));
    });
}
int scat_modulus(const void* in_dev, void* out_dev, int64_t n, int32_t dtype, void* stream) {
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (n <= 0) return;
        SB_DISPATCH(dtype, launch("prim_modulus", (double)n * 3 * sizeof(T), st, [&] {
            kp_modulus<T><<<blocks_for((size_t)n), 256, 0, st>>>(static_cast<const cx<T>*>(in_dev), static_cast<T*>(out_dev), (size_t)n);
        }));
    });
}
int scat_complex_from_real(const void* in_dev, void* out_dev, int64_t n, int32_t dtype, void* stream) {
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (n <= 0) return;
        SB_DISPATCH(dtype, launch("prim_from_real", (double)n * 3 * sizeof(T), st, [&] {
            kp_from_real<T><<<blocks_for((size_t)n), 256, 0, st>>>(static_cast<const T*>(in_dev), static_cast<cx<T>*>(out_dev), (size_t)n);
        }));
    });
}
int scat_real_part(const void* in_dev, void* out_dev, int64_t n, int32_t dtype, void* stream) {
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (n <= 0) return;
        SB_DISPATCH(dtype, launch("prim_real_part", (double)n * 3 * sizeof(T), st, [&] {
            kp_real_scaled<T><<<blocks_for((size_t)n), 256, 0, st>>>(static_cast<const cx<T>*>(in_dev), static_cast<T*>(out_dev), (size_t)n, T(1));
        }));
    });
}

// ---- 1-D primitives -----------------------------------------------------------------------------------
size_t scat_fft1d_const_bytes(int32_t N, int32_t dtype) {
    try {
        return dtype == 1 ? Fft1dTables<double>(N).bytes : Fft1dTables<float>(N).bytes;
    } catch (const std::exception& e) { last_error() = e.what(); return 0; }
}
int scat_fft1d_init(void* const_dev, int32_t N, int32_t dtype, void* stream) {
    return guarded([&] { SB_DISPATCH(dtype, fft1d_init<T>(const_dev, N, static_cast<cudaStream_t>(stream))); });
}
int scat_fft1d_exec(const void* const_dev, const void* in_dev, void* tmp_dev, void* out_dev, int64_t G, int32_t N,
                    int32_t inverse, int32_t dtype, void* stream) {
    return guarded([&] {
        once_per_device("stream", [] { stream_kernels_enable_smem<float>(); stream_kernels_enable_smem<double>(); });
        if (G <= 0) return;
        SB_DISPATCH(dtype, fft1d_exec<T>(const_dev, in_dev, tmp_dev, out_dev, G, N, inverse != 0, static_cast<cudaStream_t>(stream)));
    });
}
int scat_pad1d(const void* x_dev, void* out_dev, int64_t G, int32_t N, int32_t pad_left, int32_t pad_right,
               int32_t dtype, void* stream) {
    SB_SPLIT_BATCH(G, scat_pad1d(at(x_dev, (size_t)b0 * N * esize(dtype)),
                                 at(out_dev, (size_t)b0 * (N + pad_left + pad_right) * esize(dtype)), nb, N, pad_left, pad_right,
                                 dtype, stream))
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int P = N + pad_left + pad_right;
        if (G <= 0) return;
        dim3 grid(ceil_div(P, 256), (unsigned)G);
        SB_DISPATCH(dtype, launch("prim_pad1d", (double)G * (N + P) * sizeof(T), st, [&] {
            kp_pad1d<T><<<grid, 256, 0, st>>>(static_cast<const T*>(x_dev), static_cast<T*>(out_dev), N, pad_left, P);
        }));
    });
}
int scat_subsample_fourier1d(const void* in_dev, void* out_dev, int64_t G, int32_t N, int32_t k, int32_t dtype,
                             void* stream) {
    SB_SPLIT_BATCH(G, scat_subsample_fourier1d(at(in_dev, (size_t)b0 * N * 2 * esize(dtype)),
                                               at(out_dev, (size_t)b0 * (N / std::max(k, 1)) * 2 * esize(dtype)), nb, N, k, dtype,
                                               stream))
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (k < 1 || N % k) throw std::runtime_error("subsample_fourier: k must divide the length");
        if (G <= 0) return;
        dim3 grid(ceil_div(N / k, 256), (unsigned)G);
        SB_DISPATCH(dtype, launch("prim_periodize1d", (double)G * N * sizeof(cx<T>), st, [&] {
            kp_periodize1d<T><<<grid, 256, 0, st>>>(static_cast<const cx<T>*>(in_dev), static_cast<cx<T>*>(out_dev), N, k);
        }));
    });
}

// ---- 3-D primitives -----------------------------------------------------------------------------------
size_t scat_fft3d_const_bytes(int32_t M, int32_t N, int32_t O, int32_t dtype) {
    try {
        return dtype == 1 ? Fft3dTables<double>(M, N, O).bytes : Fft3dTables<float>(M, N, O).bytes;
    } catch (const std::exception& e) { last_error() = e.what(); return 0; }
}
int scat_fft3d_init(void* const_dev, int32_t M, int32_t N, int32_t O, int32_t dtype, void* stream) {
    return guarded([&] { SB_DISPATCH(dtype, fft3d_init<T>(const_dev, M, N, O, static_cast<cudaStream_t>(stream))); });
}
int scat_fft3d_exec(const void* const_dev, const void* in_dev, void* out_dev, int64_t G, int32_t M, int32_t N, int32_t O,
                    int32_t inverse, int32_t dtype, void* stream) {
    return guarded([&] {
        once_per_device("stream", [] { stream_kernels_enable_smem<float>(); stream_kernels_enable_smem<double>(); });
        if (G <= 0) return;
        SB_DISPATCH(dtype, fft3d_exec<T>(const_dev, in_dev, out_dev, G, M, N, O, inverse != 0, static_cast<cudaStream_t>(stream)));
    });
}
int scat_modulus_rotation(const void* x_dev, const void* prev_dev, void* out_dev, int64_t n, int32_t dtype, void* stream) {
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (n <= 0) return;
        SB_DISPATCH(dtype, launch("prim_modulus_rotation", (double)n * 4 * sizeof(T), st, [&] {
            kp_modulus_rotation<T><<<blocks_for((size_t)n), 256, 0, st>>>(static_cast<const cx<T>*>(x_dev), static_cast<const T*>(prev_dev),
                                                                           static_cast<T*>(out_dev), (size_t)n);
        }));
    });
}
int scat_compute_integrals(const void* x_dev, void* out_f64_dev, int64_t B, int64_t n, const void* powers_f32_dev,
                           int32_t P, int32_t dtype, void* stream) {
    SB_SPLIT_BATCH(B, scat_compute_integrals(at(x_dev, (size_t)b0 * n * esize(dtype)), at(out_f64_dev, (size_t)b0 * P * sizeof(double)),
                                             nb, n, powers_f32_dev, P, dtype, stream))
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (P < 1 || P > 8) throw std::runtime_error("compute_integrals supports 1..8 powers");
        if (B <= 0 || n <= 0) return;
        SB_CUDA(cudaMemsetAsync(out_f64_dev, 0, (size_t)B * P * sizeof(double), st));
        dim3 grid((unsigned)std::min<int64_t>(1024, (n + 1023) / 1024), (unsigned)B);
        SB_DISPATCH(dtype, launch("prim_integrals", (double)B * n * sizeof(T), st, [&] {
            kp_integrals<T><<<grid, 256, 0, st>>>(static_cast<const T*>(x_dev), static_cast<double*>(out_f64_dev), (size_t)n,
                                                   static_cast<const float*>(powers_f32_dev), P);
        }));
    });
}

// ---- adjoints used by the autograd graph ------------------------------------------------------------
int scat_cdgmm_bcast(const void* a_dev, const void* w_dev, void* out_dev, int64_t nb, int32_t nf, int64_t n,
                     int32_t adjoint, int32_t dtype, void* stream) {
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const size_t total = (size_t)nb * n;
        if (!total) return;
        SB_DISPATCH(dtype, launch(adjoint ? "prim_cdgmm_bcast_bwd" : "prim_cdgmm_bcast", (double)total * (nf + 1) * sizeof(cx<T>), st, [&] {
            if (adjoint)
                kp_cdgmm_bcast_bwd<T><<<blocks_for(total), 256, 0, st>>>(static_cast<const cx<T>*>(a_dev), static_cast<const T*>(w_dev),
                                                                          static_cast<cx<T>*>(out_dev), (size_t)nb, nf, (size_t)n);
            else
                kp_cdgmm_bcast<T><<<blocks_for(total), 256, 0, st>>>(static_cast<const cx<T>*>(a_dev), static_cast<const T*>(w_dev),
                                                                      static_cast<cx<T>*>(out_dev), (size_t)nb, nf, (size_t)n);
        }));
    });
}
int scat_subsample_fourier2d_bwd(const void* gout_dev, void* gin_dev, int64_t G, int32_t n0, int32_t n1, int32_t k,
                                 int32_t dtype, void* stream) {
    SB_SPLIT_BATCH(G, scat_subsample_fourier2d_bwd(at(gout_dev, (size_t)b0 * (n0 / std::max(k, 1)) * (n1 / std::max(k, 1)) * 2 * esize(dtype)),
                                                   at(gin_dev, (size_t)b0 * n0 * n1 * 2 * esize(dtype)), nb, n0, n1, k, dtype, stream))
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (k < 1 || n0 % k || n1 % k) throw std::runtime_error("subsample_fourier: k must divide both sizes");
        if (G <= 0) return;
        dim3 grid(ceil_div(n1, 128), n0, (unsigned)G);
        SB_DISPATCH(dtype, launch("prim_periodize2d_bwd", (double)G * n0 * n1 * sizeof(cx<T>), st, [&] {
            kp_periodize2d_bwd<T><<<grid, 128, 0, st>>>(static_cast<const cx<T>*>(gout_dev), static_cast<cx<T>*>(gin_dev), n0, n1, k);
        }));
    });
}
int scat_modulus_bwd(const void* x_dev, const void* g_dev, void* gx_dev, int64_t n, int32_t dtype, void* stream) {
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (n <= 0) return;
        SB_DISPATCH(dtype, launch("prim_modulus_bwd", (double)n * 5 * sizeof(T), st, [&] {
            kp_modulus_bwd<T><<<blocks_for((size_t)n), 256, 0, st>>>(static_cast<const cx<T>*>(x_dev), static_cast<const T*>(g_dev),
                                                                      static_cast<cx<T>*>(gx_dev), (size_t)n);
        }));
    });
}
int scat_pad2d_bwd(const void* gout_dev, void* gx_dev, int64_t B, int32_t M, int32_t N, int32_t top, int32_t bottom,
                   int32_t left, int32_t right, int32_t dtype, void* stream) {
    SB_SPLIT_BATCH(B, scat_pad2d_bwd(at(gout_dev, (size_t)b0 * (M + top + bottom) * (N + left + right) * esize(dtype)),
                                     at(gx_dev, (size_t)b0 * M * N * esize(dtype)), nb, M, N, top, bottom, left, right, dtype, stream))
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int P0 = M + top + bottom, P1 = N + left + right;
        if (B <= 0) return;
        dim3 grid(ceil_div(P1, 128), P0, (unsigned)B);
        SB_DISPATCH(dtype, {
            SB_CUDA(cudaMemsetAsync(gx_dev, 0, (size_t)B * M * N * sizeof(T), st));
            launch("prim_pad2d_bwd", (double)B * (M * N + P0 * P1) * sizeof(T), st, [&] {
                kp_pad2d_bwd<T><<<grid, 128, 0, st>>>(static_cast<const T*>(gout_dev), static_cast<T*>(gx_dev), M, N, top, left, P0, P1);
            });
        });
    });
}

// ---------------------------------------------------------------- fused 1-D kernels (engine1d.py drives the cascade)
int scat1d_split(int32_t N, int32_t* Na, int32_t* Nb) {
    return guarded([&] { Split1d s = split1d(N); if (Na) *Na = s.Na; if (Nb) *Nb = s.Nb; });
}
size_t scat1d_tables_bytes(int32_t N) {
    try { return Tables1d(N).bytes; } catch (const std::exception& e) { last_error() = e.what(); return 0; }
}
int scat1d_tables_init(void* tables_dev, int32_t N, void* stream) {
    return guarded([&] { tables1d_init(tables_dev, N, static_cast<cudaStream_t>(stream)); });
}
size_t scat1d_fin_tables_bytes(int32_t M) {
    try { return FinTables1d(M).bytes; } catch (const std::exception& e) { last_error() = e.what(); return 0; }
}
int scat1d_fin_tables_init(void* tables_dev, int32_t M, void* stream) {
    return guarded([&] { fin_tables1d_init(tables_dev, M, static_cast<cudaStream_t>(stream)); });
}
int scat1d_col_prod(const void* tables_dev, const void* parent_dev, int64_t ps_b, int64_t ps_i, const void* filt_ptrs_dev,
                    const void* supp_dev, void* y_dev, int64_t G, int32_t NI, int32_t Npar, int32_t N, double algo_bytes,
                    void* stream) {
    return guarded([&] {
        col_prod1d(tables_dev, parent_dev, ps_b, ps_i, filt_ptrs_dev, supp_dev, y_dev, G, NI, Npar, N, algo_bytes,
                   static_cast<cudaStream_t>(stream));
    });
}
int scat1d_row_mod(const void* tables_dev, void* y_dev, int64_t G, int32_t N, void* part_dev, int32_t Fc, double algo_bytes,
                   void* stream) {
    return guarded([&] { row_mod1d(tables_dev, y_dev, G, N, part_dev, Fc, algo_bytes, static_cast<cudaStream_t>(stream)); });
}
int scat1d_row_mod_t0(const void* tables_dev, void* y_dev, int64_t G, int32_t N, void* mod_dev, int32_t is_leaf,
                      double algo_bytes, void* stream) {
    return guarded([&] {
        if (!mod_dev) throw std::runtime_error("scat1d_row_mod_t0: mod_dev is null");
        row_mod1d(tables_dev, y_dev, G, N, nullptr, 0, algo_bytes, static_cast<cudaStream_t>(stream), mod_dev, is_leaf != 0);
    });
}
int scat1d_col_fwd(const void* tables_dev, const void* z_dev, void* out_dev, int64_t G, int32_t N, double algo_bytes,
                   void* stream) {
    return guarded([&] { col_fwd1d(tables_dev, z_dev, out_dev, G, N, algo_bytes, static_cast<cudaStream_t>(stream)); });
}
int scat1d_rfft(const void* tables_dev, const void* x_dev, void* z_dev, void* out_dev, int64_t G, int32_t N, void* stream) {
    return guarded([&] { rfft1d(tables_dev, x_dev, z_dev, out_dev, G, N, static_cast<cudaStream_t>(stream)); });
}
int scat1d_tile_max(void) { return k1TileMaxN; }
int scat1d_tile(const void* tables_dev, const void* parent_dev, int64_t ps_b, int64_t ps_i, const void* filt_ptrs_dev,
                const void* supp_dev, void* spec_dev, void* part_dev, int32_t Fc, int64_t G, int32_t NI, int32_t Npar,
                int32_t N, double algo_bytes, void* stream) {
    return guarded([&] {
        tile1d(tables_dev, parent_dev, ps_b, ps_i, filt_ptrs_dev, supp_dev, spec_dev, part_dev, Fc, G, NI, Npar, N, algo_bytes,
               static_cast<cudaStream_t>(stream));
    });
}
int scat1d_tile_t0(const void* tables_dev, const void* parent_dev, int64_t ps_b, int64_t ps_i, const void* filt_ptrs_dev,
                   const void* supp_dev, void* spec_dev, void* mod_dev, int64_t G, int32_t NI, int32_t Npar, int32_t N,
                   double algo_bytes, void* stream) {
    return guarded([&] {
        if (!mod_dev) throw std::runtime_error("scat1d_tile_t0: mod_dev is null");
        tile1d(tables_dev, parent_dev, ps_b, ps_i, filt_ptrs_dev, supp_dev, spec_dev, nullptr, 0, G, NI, Npar, N, algo_bytes,
               static_cast<cudaStream_t>(stream), mod_dev);
    });
}
int scat1d_finish(const void* fin_tables_dev, const void* u0_dev, const void* u1_dev, const void* part_dev,
                  const void* segs_dev, int32_t nseg, int64_t total_lines, int32_t M, void* out_dev, int64_t os_b, int32_t i0,
                  int32_t W, double algo_bytes, void* stream) {
    return guarded([&] {
        finish1d(fin_tables_dev, u0_dev, u1_dev, part_dev, segs_dev, nseg, total_lines, M, out_dev, os_b, i0, W, algo_bytes,
                 static_cast<cudaStream_t>(stream));
    });
}
int scat1d_finish_global(const void* u0_dev, const void* u1_dev, const void* part_dev, const void* segs_dev, int32_t nseg,
                         int64_t total_lines, void* out_dev, int64_t os_b, void* stream) {
    return guarded([&] {
        finish1d_global(u0_dev, u1_dev, part_dev, segs_dev, nseg, total_lines, out_dev, os_b, static_cast<cudaStream_t>(stream));
    });
}
size_t scat1d_finseg_bytes(void) { return sizeof(FinSeg<float>); }

// ---------------------------------------------------------------- fused 3-D kernels (engine3d.py drives the cascade)
int scat3d_supported(int32_t M, int32_t N, int32_t O) { return fused3d_supported(M, N, O) ? 1 : 0; }
size_t scat3d_tables_bytes(int32_t M, int32_t N, int32_t O) {
    try { return Tables3d(M, N, O).bytes; } catch (const std::exception& e) { last_error() = e.what(); return 0; }
}
int scat3d_tables_init(void* tables_dev, int32_t M, int32_t N, int32_t O, void* stream) {
    return guarded([&] { tables3d_init(tables_dev, M, N, O, static_cast<cudaStream_t>(stream)); });
}
int scat3d_rfft(const void* tables_dev, const void* x_dev, void* out_dev, int64_t B, int32_t M, int32_t N, int32_t O,
                void* stream) {
    return guarded([&] { rfft3d(tables_dev, x_dev, out_dev, B, M, N, O, static_cast<cudaStream_t>(stream)); });
}
int scat3d_col_prod(const void* tables_dev, const void* u_dev, const void* filt_dev, void* y_dev, int64_t B, int32_t nm,
                    int32_t M, int32_t N, int32_t O, void* stream) {
    return guarded([&] { col_prod3d(tables_dev, u_dev, filt_dev, y_dev, B, nm, M, N, O, static_cast<cudaStream_t>(stream)); });
}
int scat3d_plane(const void* tables_dev, const void* y_dev, void* spec_dev, void* integ_f64_dev, int64_t istride, int32_t ioff,
                 const void* powers_f32_dev, int32_t P, int64_t B, int32_t nm, int32_t M, int32_t N, int32_t O, void* stream) {
    return guarded([&] {
        plane3d(tables_dev, y_dev, spec_dev, integ_f64_dev, istride, ioff, powers_f32_dev, P, B, nm, M, N, O,
                static_cast<cudaStream_t>(stream));
    });
}
int scat3d_col_fwd(const void* tables_dev, const void* z_dev, void* out_dev, int64_t B, int32_t M, int32_t N, int32_t O,
                   void* stream) {
    return guarded([&] { col_fwd3d(tables_dev, z_dev, out_dev, B, M, N, O, static_cast<cudaStream_t>(stream)); });
}

// ---------------------------------------------------------------- adjoints of the 1-D / 3-D eager primitives
int scat_subsample_fourier1d_bwd(const void* gout_dev, void* gin_dev, int64_t G, int32_t N, int32_t k, int32_t dtype,
                                 void* stream) {
    SB_SPLIT_BATCH(G, scat_subsample_fourier1d_bwd(at(gout_dev, (size_t)b0 * (N / std::max(k, 1)) * 2 * esize(dtype)),
                                                   at(gin_dev, (size_t)b0 * N * 2 * esize(dtype)), nb, N, k, dtype, stream))
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (k < 1 || N % k) throw std::runtime_error("subsample_fourier: k must divide the length");
        if (G <= 0) return;
        dim3 grid(ceil_div(N, 256), (unsigned)G);
        SB_DISPATCH(dtype, launch("prim_periodize1d_bwd", (double)G * N * sizeof(cx<T>), st, [&] {
            kp_periodize1d_bwd<T><<<grid, 256, 0, st>>>(static_cast<const cx<T>*>(gout_dev), static_cast<cx<T>*>(gin_dev), N, k);
        }));
    });
}
int scat_modulus_rotation_bwd(const void* x_dev, const void* prev_dev, const void* out_dev, const void* g_dev, void* gx_dev,
                              void* gprev_dev, int64_t n, int32_t dtype, void* stream) {
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (n <= 0) return;
        SB_DISPATCH(dtype, launch("prim_modulus_rotation_bwd", (double)n * 8 * sizeof(T), st, [&] {
            kp_modrot_bwd<T><<<blocks_for((size_t)n), 256, 0, st>>>(
                static_cast<const cx<T>*>(x_dev), static_cast<const T*>(prev_dev), static_cast<const T*>(out_dev),
                static_cast<const T*>(g_dev), static_cast<cx<T>*>(gx_dev), static_cast<T*>(gprev_dev), (size_t)n);
        }));
    });
}
int scat_compute_integrals_bwd(const void* x_dev, const void* g_dev, void* gx_dev, int64_t B, int64_t n,
                               const void* powers_f32_dev, int32_t P, int32_t dtype, void* stream) {
    SB_SPLIT_BATCH(B, scat_compute_integrals_bwd(at(x_dev, (size_t)b0 * n * esize(dtype)), at(g_dev, (size_t)b0 * P * esize(dtype)),
                                                 at(gx_dev, (size_t)b0 * n * esize(dtype)), nb, n, powers_f32_dev, P, dtype, stream))
    return guarded([&] {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (B <= 0 || n <= 0) return;
        dim3 grid(blocks_for((size_t)n), (unsigned)B);
        SB_DISPATCH(dtype, launch("prim_integrals_bwd", (double)B * n * 2 * sizeof(T), st, [&] {
            kp_integrals_bwd<T><<<grid, 256, 0, st>>>(static_cast<const T*>(x_dev), static_cast<const T*>(g_dev),
                                                      static_cast<T*>(gx_dev), (size_t)n,
                                                      static_cast<const float*>(powers_f32_dev), P);
        }));
    });
}

}  // extern "C"
