// common.cuh - host-side helpers: error reporting, launch accounting, slab configuration.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstdint>
#include <mutex>
#include <set>
#include <utility>
#include <stdexcept>
#include <string>
#include <vector>
#include "fft_core.cuh"

namespace sb {

inline std::string& last_error() { static thread_local std::string e; return e; }
inline std::atomic<uint64_t>& launch_counter() { static std::atomic<uint64_t> c{0}; return c; }

#define SB_CUDA(expr)                                                                           \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(_e));       \
    } while (0)

// Optional per-launch timing (CUDA events on the launching stream), switched on by
// scat_timing_enable(1); bench.py uses it for the live per-kernel roofline figures.
struct TimingRec { std::string label; cudaEvent_t e0, e1; double bytes; };
inline bool& timing_on() { static bool on = false; return on; }
inline std::vector<TimingRec>& timing_recs() { static std::vector<TimingRec> r; return r; }

inline void check_launch(const char* what) {
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw std::runtime_error(std::string("launch ") + what + ": " + cudaGetErrorString(e));
}

// Run one kernel launch `f` on stream `st`; `bytes` = algorithmic bytes (reads + writes) of the launch.
// SCAT_B200_NVTX=1: every launch is wrapped in an NVTX range named by its label (visible in Nsight timelines)
inline bool nvtx_on() {
    static const bool on = [] { const char* v = std::getenv("SCAT_B200_NVTX"); return v && v[0] == '1'; }();
    return on;
}
struct NvtxRange {
    bool on;
    explicit NvtxRange(const std::string& label) : on(nvtx_on()) { if (on) nvtxRangePushA(label.c_str()); }
    ~NvtxRange() { if (on) nvtxRangePop(); }
};

template <typename F> inline void launch(const std::string& label, double bytes, cudaStream_t st, F&& f) {
    NvtxRange range(label);
    if (timing_on()) {
        TimingRec r; r.label = label; r.bytes = bytes;
        SB_CUDA(cudaEventCreate(&r.e0)); SB_CUDA(cudaEventCreate(&r.e1));
        SB_CUDA(cudaEventRecord(r.e0, st));
        f();
        SB_CUDA(cudaEventRecord(r.e1, st));
        timing_recs().push_back(r);
    } else {
        f();
    }
    check_launch(label.c_str());
}

constexpr size_t kMaxDynSmem = 227 * 1024;

// opt in to large dynamic shared memory once per kernel instantiation
template <typename K> inline void enable_big_smem(K kernel) {
    SB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
}

// The > 48 KB dynamic shared-memory opt-in (cudaFuncSetAttribute) is a per-DEVICE attribute: run `f` once per
// (tag, current device), so a process that drives several GPUs opts every one of them in.
template <typename F> inline void once_per_device(const char* tag, F&& f) {
    static std::mutex mu;
    static std::set<std::pair<std::string, int>> done;
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (done.insert(std::make_pair(std::string(tag), dev)).second) {
        try { f(); } catch (...) { done.erase(std::make_pair(std::string(tag), dev)); throw; }
    }
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Slab launch configuration for a batched 1-D transform of length P.n over `total_lines`.
struct SlabCfg {
    int lines;   // lines per CTA
    int LP;      // shared-memory pitch between consecutive elements (odd, >= lines)
    dim3 block;  // (lanes over lines, butterflies in flight)
    size_t smem; // dynamic shared memory bytes (slab + twiddle table)
};

inline int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Defaults: 16 lines x up to 18 butterflies in flight (<= 288 threads, two CTAs per SM at
// <= 112 registers).  SCAT_B200_LINES / SCAT_B200_THREADS override for tuning runs.
inline SlabCfg slab_cfg(const Plan1& P, int total_lines, size_t elem_bytes, size_t table_bytes_per_elem = 0,
                        size_t slab_budget = 72 * 1024) {
    SlabCfg c{};
    int lines = env_int("SCAT_B200_LINES", 16);
    const int max_threads = std::min(576, env_int("SCAT_B200_THREADS", 288));
    if (lines < 1 || lines > 32 || (lines & (lines - 1))) lines = 16;
    while (lines > 1 && (size_t)(lines | 1) * P.n * elem_bytes > slab_budget) lines /= 2;
    while (lines > 1 && lines / 2 >= total_lines) lines /= 2;
    c.lines = lines;
    c.LP = lines | 1;
    int max_bf = 1;
    for (int p = 0; p < P.npass; ++p) max_bf = std::max(max_bf, P.n / P.radix[p]);
    int bx = lines;
    int by = std::max(1, std::min(max_bf, max_threads / bx));
    // keep at least 64 threads for the staging loops
    while (bx * by < 64 && by < 64) ++by;
    c.block = dim3(bx, by, 1);
    c.smem = ((size_t)c.LP * P.n + P.n) * elem_bytes + (size_t)P.n * table_bytes_per_elem;
    if (c.smem > kMaxDynSmem)
        throw std::runtime_error("FFT line of length " + std::to_string(P.n) + " does not fit in shared memory");
    return c;
}

}  // namespace sb
