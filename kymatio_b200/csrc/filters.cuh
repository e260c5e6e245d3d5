// Filter-bank synthesis on the device.
//
// 2-D: Morlet / Gabor filters as the reference builds them - the Gaussian envelope times a plane wave, summed over the
// 5x5 neighbouring periods in the SPATIAL domain, zero-mean correction beta = sum(carrier) / sum(envelope), then fft2,
// real part, band-limit + alias fold per resolution (kymatio/scattering2d/filter_bank.py:5-53 bank layout, :56-91 fold,
// :94-128 Morlet, :131-175 Gabor).  The reference accumulates in complex64 and transforms in single precision; here the
// spatial sums, the transform (the library's own double-precision fft2) and the fold run in float64 and only the final
// filter is rounded to float32, so the two agree to float32 rounding of the filter's peak (tests/test_filters_gpu.py).
//
// 3-D: solid harmonic wavelets and Gaussians in closed form in the Fourier domain,
//     psi_{l,m}(w) = c_l (-i)^l (|w|/s)^l exp(-|w|^2 / (2 s^2)) Y_l^m(polar, azimuth),   s = 1/sigma
// (kymatio/scattering3d/filter_bank.py:5-36 bank, :64-97 Gaussian, :100-166 solid harmonics; angles as
// kymatio/scattering3d/utils.py get_3d_angles: polar = atan2(z, |xy|) + pi/2, azimuth = atan2(y, x)).
#pragma once
#include "common.cuh"

namespace sb {

// one 2-D filter: exp(-(c00 u^2 + cross u v + c11 v^2)) * exp(i (fu u + fv v)) / norm, periodised 5x5
struct Gabor2dParams {
    double c00, cross, c11, fu, fv, inv_norm;
    double zero_mean;  // != 0: Morlet (subtract beta * envelope); 0: plain Gabor (phi)
    double reserved;
};

// carrier[f][u][v] (complex) and envelope[f][u][v] (real, the same Gaussian without the plane wave), plus their sums
__global__ void __launch_bounds__(256) kf_gabor2d(const Gabor2dParams* __restrict__ prm, double2* __restrict__ carrier,
                                                  double* __restrict__ envelope, double* __restrict__ sums, int M, int N) {
    const int f = blockIdx.y;
    const Gabor2dParams p = prm[f];
    const size_t MN = (size_t)M * N;
    double sr = 0, si = 0, se = 0;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < MN; idx += (size_t)gridDim.x * blockDim.x) {
        const int u0 = (int)(idx / N), v0 = (int)(idx - (size_t)u0 * N);
        double cr = 0, ci = 0, en = 0;
        for (int pu = -2; pu <= 2; ++pu) {
            const double u = (double)(u0 + pu * M);
            for (int pv = -2; pv <= 2; ++pv) {
                const double v = (double)(v0 + pv * N);
                const double e = exp(-(p.c00 * u * u + p.cross * u * v + p.c11 * v * v));
                double s, c;
                sincos(p.fu * u + p.fv * v, &s, &c);
                cr += e * c; ci += e * s; en += e;
            }
        }
        cr *= p.inv_norm; ci *= p.inv_norm; en *= p.inv_norm;
        carrier[(size_t)f * MN + idx] = make_double2(cr, ci);
        envelope[(size_t)f * MN + idx] = en;
        sr += cr; si += ci; se += en;
    }
    __shared__ double red[3][8];
    for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_down_sync(0xffffffffu, sr, o);
        si += __shfl_down_sync(0xffffffffu, si, o);
        se += __shfl_down_sync(0xffffffffu, se, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = sr; red[1][w] = si; red[2][w] = se; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += red[threadIdx.x][i];
        atomicAdd(&sums[3 * f + threadIdx.x], a);
    }
}

// carrier -= beta * envelope, beta = sum(carrier) / sum(envelope)   (filter_bank.py:122-127)
__global__ void __launch_bounds__(256) kf_zero_mean2d(const Gabor2dParams* __restrict__ prm, double2* __restrict__ carrier,
                                                      const double* __restrict__ envelope, const double* __restrict__ sums,
                                                      int M, int N) {
    const int f = blockIdx.y;
    if (prm[f].zero_mean == 0.0) return;
    const double br = sums[3 * f] / sums[3 * f + 2], bi = sums[3 * f + 1] / sums[3 * f + 2];
    const size_t MN = (size_t)M * N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < MN; idx += (size_t)gridDim.x * blockDim.x) {
        double2 c = carrier[(size_t)f * MN + idx];
        const double e = envelope[(size_t)f * MN + idx];
        c.x -= br * e; c.y -= bi * e;
        carrier[(size_t)f * MN + idx] = c;
    }
}

// out[k][l] = float( sum_{i,j < 2^res} keep(k + i m, l + j n) * Re spec[k + i m][l + j n] ),  m = M / 2^res, n = N / 2^res;
// keep zeroes the rows [M/2^(res+1), M/2^(res+1) + M(1 - 2^-res)) and the same band of columns   (filter_bank.py:56-91)
__global__ void __launch_bounds__(256) kf_fold2d(const double2* __restrict__ spec, float* __restrict__ out, int M, int N, int res) {
    const int k2 = 1 << res, m = M / k2, n = N / k2;
    const int r0 = (int)((double)M / (double)(2 << res)), rl = (int)((double)M * (1.0 - 1.0 / (double)k2));
    const int c0 = (int)((double)N / (double)(2 << res)), cl = (int)((double)N * (1.0 - 1.0 / (double)k2));
    const size_t total = (size_t)m * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx / n), l = (int)(idx - (size_t)k * n);
        double acc = 0;
        for (int i = 0; i < k2; ++i) {
            const int r = k + i * m;
            if (r >= r0 && r < r0 + rl) continue;
            for (int j = 0; j < k2; ++j) {
                const int c = l + j * n;
                if (c >= c0 && c < c0 + cl) continue;
                acc += spec[(size_t)r * N + c].x;
            }
        }
        out[idx] = (float)acc;
    }
}

// ------------------------------------------------------------------------------------------------ 3-D
// centred, ifftshifted frequency of index i on an axis of length n: (i < ceil(n/2) ? i : i - n) * 2 pi / n
// (np.mgrid[-n//2 : -n//2 + n] ifftshifted: the grid starts at -ceil(n/2)... for even n the Nyquist bin is -n/2)
__device__ __forceinline__ double axis_freq(int i, int n) {
    // the centred grid starts at (-n) // 2 = -ceil(n/2); ifftshift rotates it left by n/2 (odd n: the origin lands on index 1,
    // as in the reference)
    const int v = -((n + 1) / 2) + ((i + n / 2) % n);
    return (double)v * (6.283185307179586476925286766559 / (double)n);
}

// out: [nj][2l+1][M][N][O] complex64;  sigmas[j] = sigma_0 2^j;  norm = c_l (2 pi)^{3/2} (times (-i)^l applied here)
__global__ void __launch_bounds__(256) kf_solid_harmonic3d(float2* __restrict__ out, const double* __restrict__ sigmas, int nj,
                                                           int l, double norm, int M, int N, int O) {
    const size_t vol = (size_t)M * N * O;
    const int j = blockIdx.y;
    const double sigma = sigmas[j];                               // Fourier width 1/sigma: r/_sigma = r sigma
    const int nm = 2 * l + 1;
    // (-i)^l
    const int ph = l & 3;
    const double pr = ph == 0 ? 1.0 : (ph == 2 ? -1.0 : 0.0), pi_ = ph == 1 ? -1.0 : (ph == 3 ? 1.0 : 0.0);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < vol; idx += (size_t)gridDim.x * blockDim.x) {
        const int a = (int)(idx / ((size_t)N * O));
        const int rem = (int)(idx - (size_t)a * N * O);
        const int b = rem / O, c = rem - b * O;
        // the reference names the axes (z, y, x) = (grid[0], grid[1], grid[2])
        const double z = axis_freq(a, M), y = axis_freq(b, N), x = axis_freq(c, O);
        const double r2 = x * x + y * y + z * z;
        const double gauss = exp(-0.5 * r2 * sigma * sigma);
        float2* o = out + ((size_t)j * nm) * vol + idx;
        if (l == 0) { o[0] = make_float2((float)gauss, 0.f); continue; }
        const double r = sqrt(r2), rxy = sqrt(x * x + y * y);
        const double radial = pow(r * sigma, (double)l) * gauss * norm;
        // polar = atan2(z, rxy) + pi/2  ->  cos(polar) = -z / r, sin(polar) = rxy / r  (r = 0: polar = pi/2)
        const double ct = r > 0 ? -z / r : 0.0, st = r > 0 ? rxy / r : 1.0;
        const double az = atan2(y, x);
        for (int m = 0; m <= l; ++m) {
            // associated Legendre P_l^m(ct) with the Condon-Shortley phase
            double pmm = 1.0;
            for (int i = 1; i <= m; ++i) pmm *= -(2.0 * i - 1.0) * st;
            double plm = pmm;
            if (l > m) {
                double p1 = ct * (2.0 * m + 1.0) * pmm;
                plm = p1;
                double p0 = pmm;
                for (int ll = m + 2; ll <= l; ++ll) {
                    plm = ((2.0 * ll - 1.0) * ct * p1 - (ll + m - 1.0) * p0) / (double)(ll - m);
                    p0 = p1; p1 = plm;
                }
            }
            // sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!)
            double ratio = 1.0;
            for (int i = l - m + 1; i <= l + m; ++i) ratio /= (double)i;
            const double nlm = sqrt((2.0 * l + 1.0) / (4.0 * 3.14159265358979323846) * ratio);
            double s, cs;
            sincos((double)m * az, &s, &cs);
            const double yr = nlm * plm * cs, yi = nlm * plm * s;          // Y_l^m
            // times radial * (-i)^l
            const double vr = radial * (yr * pr - yi * pi_), vi = radial * (yr * pi_ + yi * pr);
            o[(size_t)(l + m) * vol] = make_float2((float)vr, (float)vi);
            if (m > 0) {
                // Y_l^{-m} = (-1)^m conj(Y_l^m)
                const double sg = (m & 1) ? -1.0 : 1.0;
                const double wr = sg * yr, wi = -sg * yi;
                o[(size_t)(l - m) * vol] = make_float2((float)(radial * (wr * pr - wi * pi_)), (float)(radial * (wr * pi_ + wi * pr)));
            }
        }
    }
}

// Gaussian low-pass bank: out[j][M][N][O] complex64 = exp(-|w|^2 sigma_j^2 / 2)
__global__ void __launch_bounds__(256) kf_gaussian3d(float2* __restrict__ out, const double* __restrict__ sigmas, int M, int N, int O) {
    const size_t vol = (size_t)M * N * O;
    const int j = blockIdx.y;
    const double sigma = sigmas[j];
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < vol; idx += (size_t)gridDim.x * blockDim.x) {
        const int a = (int)(idx / ((size_t)N * O));
        const int rem = (int)(idx - (size_t)a * N * O);
        const int b = rem / O, c = rem - b * O;
        const double z = axis_freq(a, M), y = axis_freq(b, N), x = axis_freq(c, O);
        out[(size_t)j * vol + idx] = make_float2((float)exp(-0.5 * (x * x + y * y + z * z) * sigma * sigma), 0.f);
    }
}

}  // namespace sb
