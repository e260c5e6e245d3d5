// Device-side filter-bank synthesis: launchers + C ABI (include/scat_b200.h, "filter synthesis").
#include <algorithm>
#include <stdexcept>
#include "filters.cuh"
#include "../../include/scat_b200.h"

using namespace sb;

namespace {
template <typename F> int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { last_error() = e.what(); return 1; }
    catch (...) { last_error() = "unknown error"; return 2; }
}
inline unsigned grid_for(size_t n) { return (unsigned)std::min<size_t>((n + 255) / 256, 4096); }
}

extern "C" {

int scat_filters2d_spatial(const void* params_dev, int32_t n_filters, int32_t M, int32_t N, void* carrier_dev,
                           void* envelope_dev, void* sums_dev, void* stream) {
    return guarded([&] {
        if (n_filters <= 0 || M <= 0 || N <= 0) throw std::runtime_error("filters2d: empty request");
        if (n_filters > 65535) throw std::runtime_error("filters2d: more than 65535 filters per call");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const auto* prm = static_cast<const Gabor2dParams*>(params_dev);
        SB_CUDA(cudaMemsetAsync(sums_dev, 0, (size_t)n_filters * 3 * sizeof(double), st));
        dim3 grid(grid_for((size_t)M * N), (unsigned)n_filters);
        launch("filters2d_gabor", (double)n_filters * M * N * 24.0, st, [&] {
            kf_gabor2d<<<grid, 256, 0, st>>>(prm, static_cast<double2*>(carrier_dev), static_cast<double*>(envelope_dev),
                                             static_cast<double*>(sums_dev), M, N);
        });
        launch("filters2d_zero_mean", (double)n_filters * M * N * 40.0, st, [&] {
            kf_zero_mean2d<<<grid, 256, 0, st>>>(prm, static_cast<double2*>(carrier_dev), static_cast<const double*>(envelope_dev),
                                                 static_cast<const double*>(sums_dev), M, N);
        });
    });
}

int scat_filters2d_fold(const void* spec_dev, void* out_dev, int32_t M, int32_t N, int32_t res, void* stream) {
    return guarded([&] {
        if (res < 0 || res > 30 || M % (1 << res) || N % (1 << res)) throw std::runtime_error("filters2d_fold: size not divisible by 2^res");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const size_t total = (size_t)(M >> res) * (N >> res);
        launch("filters2d_fold", (double)M * N * 16.0, st, [&] {
            kf_fold2d<<<grid_for(total), 256, 0, st>>>(static_cast<const double2*>(spec_dev), static_cast<float*>(out_dev), M, N, res);
        });
    });
}

int scat_filters3d_solid_harmonic(void* out_dev, const void* sigmas_dev, int32_t n_scales, int32_t l, double norm, int32_t M,
                                  int32_t N, int32_t O, void* stream) {
    return guarded([&] {
        if (n_scales <= 0 || l < 0 || M <= 0 || N <= 0 || O <= 0) throw std::runtime_error("filters3d: bad request");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        dim3 grid(grid_for((size_t)M * N * O), (unsigned)n_scales);
        launch("filters3d_solid_harmonic", (double)n_scales * (2 * l + 1) * M * N * O * 8.0, st, [&] {
            kf_solid_harmonic3d<<<grid, 256, 0, st>>>(static_cast<float2*>(out_dev), static_cast<const double*>(sigmas_dev), n_scales,
                                                      l, norm, M, N, O);
        });
    });
}

int scat_filters3d_gaussian(void* out_dev, const void* sigmas_dev, int32_t n_scales, int32_t M, int32_t N, int32_t O,
                            void* stream) {
    return guarded([&] {
        if (n_scales <= 0 || M <= 0 || N <= 0 || O <= 0) throw std::runtime_error("filters3d: bad request");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        dim3 grid(grid_for((size_t)M * N * O), (unsigned)n_scales);
        launch("filters3d_gaussian", (double)n_scales * M * N * O * 8.0, st, [&] {
            kf_gaussian3d<<<grid, 256, 0, st>>>(static_cast<float2*>(out_dev), static_cast<const double*>(sigmas_dev), M, N, O);
        });
    });
}

}  // extern "C"
