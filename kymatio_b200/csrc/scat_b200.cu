// scat_b200.cu - extern "C" entry points of libscat_b200.so (see include/scat_b200.h).
#include <cstdio>
#include <cstring>
#include "plan2d.cuh"

using namespace sb;

namespace {
template <typename F> int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { last_error() = e.what(); return 1; }
    catch (...) { last_error() = "unknown error"; return 2; }
}
}  // namespace

extern "C" {

int scat_version(void) { return 100; }
const char* scat_last_error(void) { return last_error().c_str(); }
uint64_t scat_launch_count(void) { return launch_counter().load(); }

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

}  // extern "C"
