// CPU emulation of the four-step 1-D path of kernels1d.cuh (index maps, twiddles, DIF-inverse / DIT-forward pairing with
// scrambled intermediates) on the host+device butterflies of fft_core.cuh:
//   P1  columns: inverse DIF over f1 (natural in, scrambled rows p out), twiddle exp(+2 pi i f2 t1 / N)
//   P2  rows:    inverse DIF over f2 + modulus, forward DIT over t2 (natural f2' out), twiddle exp(-2 pi i t1 f2' / N)
//   P3  columns: forward DIT over t1 (scrambled in, natural f1' out)
// checked against a long-double O(N^2) evaluation of  fft(|ifft(X)|); also the pruned column DFT of the leaf kernel
// and the real-input forward transform (rows staged at posB, columns at posA).  Built and run by
// tests/test_fft_core_cpu.py (g++, no GPU).  Exit code 0 on success.
#include <cstdio>
#include <cstdlib>
#include <complex>
#include <vector>
#include "../../kymatio_b200/csrc/plan_host.h"

using namespace sb;
typedef float T;
typedef std::complex<long double> cld;
static const long double TAU = 2.0L * 3.14159265358979323846264338327950288L;

// one line of length N at s[e] (element stride 1): all passes of the static plan, sequentially
template <int N, bool DIT, int SIGN, bool MOD> static void fft_line(cx<T>* s, const cx<T>* tw) {
    constexpr int NP = ct_plan1(N).npass;
    static_for<0, NP>([&](auto pp_) {
        constexpr int pp = decltype(pp_)::value;
        constexpr int p = DIT ? NP - 1 - pp : pp;
        constexpr int r = ct_plan1(N).radix[p], m = ct_plan1(N).blen[p];
        constexpr int q = m / r, nbf = N / r, tws = N / m;
        for (int bf = 0; bf < nbf; ++bf) {
            const int blk = bf / q, i = bf - blk * q;
            butterfly_s<r, DIT, SIGN, q, 1, (MOD && pp == NP - 1), T>(s + blk * m + i, i * tws, tw);
        }
    });
}

static cx<T> wN(int j, int N) { long double a = -TAU * (long double)(j % N) / (long double)N; return mk<T>((T)cosl(a), (T)sinl(a)); }

template <int NA, int NB> static double run(int Fc) {
    constexpr int N = NA * NB;
    auto twA = twiddle_table<T>(NA); auto twB = twiddle_table<T>(NB);
    auto posA = scramble_table(ct_plan1(NA)); auto posB = scramble_table(ct_plan1(NB));
    std::vector<int> invA(NA);
    for (int f = 0; f < NA; ++f) invA[posA[f]] = f;
    srand(7 + N);
    std::vector<cld> X(N), u(N), Xref(N);
    for (auto& v : X) v = cld((long double)rand() / RAND_MAX - 0.5L, (long double)rand() / RAND_MAX - 0.5L);
    std::vector<long double> U(N);
    for (int t = 0; t < N; ++t) {
        cld acc = 0;
        for (int f = 0; f < N; ++f) acc += X[f] * std::polar(1.0L, TAU * (long double)((long long)f * t % N) / N);
        u[t] = acc / (long double)N; U[t] = std::abs(u[t]);
    }
    for (int f = 0; f < N; ++f) {
        cld acc = 0;
        for (int t = 0; t < N; ++t) acc += U[t] * std::polar(1.0L, -TAU * (long double)((long long)f * t % N) / N);
        Xref[f] = acc;
    }
    // P1
    std::vector<cx<T>> Y((size_t)N), line(std::max(NA, NB));
    for (int f2 = 0; f2 < NB; ++f2) {
        for (int f1 = 0; f1 < NA; ++f1) line[f1] = mk<T>((T)(X[NB * f1 + f2].real() / N), (T)(X[NB * f1 + f2].imag() / N));
        fft_line<NA, false, +1, false>(line.data(), twA.data());
        for (int p = 0; p < NA; ++p) Y[(size_t)p * NB + f2] = cmulc(line[p], wN(f2 * invA[p], N));
    }
    // P2 (parent variant) and the leaf partial sums over ALL rows
    std::vector<cx<T>> Z((size_t)N);
    std::vector<cld> part(Fc, cld(0));
    double err_mod = 0;
    for (int p = 0; p < NA; ++p) {
        for (int e = 0; e < NB; ++e) line[e] = Y[(size_t)p * NB + e];
        fft_line<NB, false, +1, true>(line.data(), twB.data());
        for (int q = 0; q < NB; ++q) {                                   // modulus field at scrambled t2: check it
            int t2 = 0; for (int f = 0; f < NB; ++f) if (posB[f] == q) t2 = f;
            err_mod = std::max(err_mod, (double)fabsl((long double)line[q].x - U[invA[p] + NA * t2]));
        }
        fft_line<NB, true, -1, false>(line.data(), twB.data());
        for (int e = 0; e < NB; ++e) Z[(size_t)p * NB + e] = cmul(line[e], wN(invA[p] * e, N));
        for (int f = 0; f < Fc; ++f) {
            const cx<T> w = wN((int)((long long)invA[p] * f % N), N);
            const cx<T> v = cmul(line[f % NB], w);
            part[f] += cld(v.x, v.y);
        }
    }
    // P3
    double err = err_mod, ref_max = 0;
    for (int f2 = 0; f2 < NB; ++f2) {
        for (int p = 0; p < NA; ++p) line[p] = Z[(size_t)p * NB + f2];
        fft_line<NA, true, -1, false>(line.data(), twA.data());
        for (int f1 = 0; f1 < NA; ++f1) {
            const cld want = Xref[NB * f1 + f2];
            ref_max = std::max(ref_max, (double)std::abs(want));
            err = std::max(err, (double)std::abs(cld(line[f1].x, line[f1].y) - want));
        }
    }
    for (int f = 0; f < Fc; ++f) err = std::max(err, (double)std::abs(part[f] - Xref[f]));
    // the leaf kernel's formulation: blocks of 16 scrambled rows hold t1 = base + stride*k; one twiddle and a power ladder
    // (z, z^2, z^4, z^8) per bin.  R rows are recomputed from Z (divide the four-step twiddle back out).
    {
        const int L = std::min(16, NA), stride = std::max(1, NA / 16);
        std::vector<cld> part2(Fc, cld(0));
        for (int p0 = 0; p0 < NA; p0 += L) {
            int base = invA[p0];
            for (int l = 1; l < L; ++l) base = std::min(base, invA[p0 + l]);
            std::vector<int> lk(L);
            for (int l = 0; l < L; ++l) lk[(invA[p0 + l] - base) / stride] = l;
            for (int f = 0; f < Fc; ++f) {
                const cx<T> z1 = wN((int)((long long)stride * f % N), N);
                const cx<T> z2 = cmul(z1, z1), z4 = cmul(z2, z2), z8 = cmul(z4, z4);
                cx<T> acc = mk<T>(0, 0);
                for (int k = 0; k < L; ++k) {
                    cx<T> pk = mk<T>(1, 0);
                    if (k & 1) pk = cmul(pk, z1);
                    if (k & 2) pk = cmul(pk, z2);
                    if (k & 4) pk = cmul(pk, z4);
                    if (k & 8) pk = cmul(pk, z8);
                    const int p = p0 + lk[k], e = f % NB;
                    const cx<T> r = cmulc(Z[(size_t)p * NB + e], wN(invA[p] * e, N));     // R[p][e]
                    acc = acc + cmul(r, pk);
                }
                const cx<T> v = cmul(acc, wN((int)((long long)base * f % N), N));
                part2[f] += cld(v.x, v.y);
            }
        }
        for (int f = 0; f < Fc; ++f) err = std::max(err, (double)std::abs(part2[f] - Xref[f]));
    }
    // Hermitian index map of the [NA][NB] layout (next step: half spectra): X[NB f1 + f2] = conj X[NB f1m + f2m] with
    // (f1m, f2m) = (NA-1-f1, NB-f2) for f2 != 0 and ((NA-f1) % NA, 0) for f2 == 0
    for (int f1 = 0; f1 < NA; ++f1) for (int f2 = 0; f2 < NB; ++f2) {
        const int f1m = f2 ? NA - 1 - f1 : (NA - f1) % NA, f2m = f2 ? NB - f2 : 0;
        err = std::max(err, (double)std::abs(Xref[NB * f1 + f2] - std::conj(Xref[NB * f1m + f2m])));
    }
    // real-input forward transform (k1d_row_real + k1d_col_fwd with posA staging) of U
    double err_r = 0;
    for (int t1 = 0; t1 < NA; ++t1) {
        for (int t2 = 0; t2 < NB; ++t2) line[posB[t2]] = mk<T>((T)U[t1 + NA * t2], T(0));
        fft_line<NB, true, -1, false>(line.data(), twB.data());
        for (int e = 0; e < NB; ++e) Z[(size_t)t1 * NB + e] = cmul(line[e], wN(t1 * e, N));
    }
    for (int f2 = 0; f2 < NB; ++f2) {
        for (int t1 = 0; t1 < NA; ++t1) line[posA[t1]] = Z[(size_t)t1 * NB + f2];
        fft_line<NA, true, -1, false>(line.data(), twA.data());
        for (int f1 = 0; f1 < NA; ++f1)
            err_r = std::max(err_r, (double)std::abs(cld(line[f1].x, line[f1].y) - Xref[NB * f1 + f2]));
    }
    const double rel = std::max(err, err_r) / ref_max;
    printf("four-step N=%5d (NA=%3d NB=%3d) Fc=%3d: rel err %.2e (modulus field abs %.2e)\n", N, NA, NB, Fc, rel, err_mod);
    return rel;
}

int main() {
    double worst = 0;
    worst = std::max(worst, run<1, 16>(9));
    worst = std::max(worst, run<2, 16>(17));
    worst = std::max(worst, run<8, 16>(40));
    worst = std::max(worst, run<16, 16>(100));
    worst = std::max(worst, run<16, 32>(64));
    worst = std::max(worst, run<32, 64>(336));
    if (worst < 2e-6) { printf("ALL OK\n"); return 0; }
    printf("FAILED (worst %.3e)\n", worst);
    return 1;
}
