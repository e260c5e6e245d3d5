// CPU harness for the host+device FFT building blocks (fft_core.cuh / plan_host.h).
// Emulates the per-pass butterfly loops of the CUDA slab transform sequentially and checks
//   (1) DIF forward == naive DFT at the canonical scrambled positions,
//   (2) DIT inverse of that == n * original (natural order),
//   (3) adjacency of periodisation aliases in the scrambled order.
// Exit code 0 on success.  Built and run by tests/test_fft_core_cpu.py (g++, no GPU).
#include <cstdio>
#include <cstdlib>
#include <complex>
#include <vector>
#include "../../kymatio_b200/csrc/plan_host.h"

using namespace sb;

template <bool INV, typename T>
static void run_passes(std::vector<cx<T>>& line, const Plan1& P, const std::vector<cx<T>>& tw, int estride) {
    for (int pp = 0; pp < P.npass; ++pp) {
        int p = INV ? P.npass - 1 - pp : pp;
        int r = P.radix[p], m = P.blen[p], q = m / r, nbf = P.n / r, tws = P.n / m;
        for (int bf = 0; bf < nbf; ++bf) {
            int blk = bf / q, i = bf - blk * q;
            butterfly_dispatch<INV, T>(r, line.data(), estride, blk * m + i, q, i * tws, tw.data(), q > 1, P.n);
        }
    }
}

template <typename T> static double check(int n, int max_pow2, double tol, bool pow2_first) {
    Plan1 P = make_plan1(n, max_pow2, pow2_first);
    auto tw = twiddle_table<T>(n);
    auto pos = scramble_table(P);
    const int estride = 3;
    std::vector<cx<T>> line((size_t)n * estride);
    std::vector<std::complex<long double>> x(n), X(n);
    srand(1234 + n);
    for (int i = 0; i < n; ++i) {
        x[i] = {(long double)rand() / RAND_MAX - 0.5L, (long double)rand() / RAND_MAX - 0.5L};
        line[(size_t)i * estride] = mk<T>((T)x[i].real(), (T)x[i].imag());
    }
    const long double tau = 2.0L * 3.14159265358979323846264338327950288L;
    for (int f = 0; f < n; ++f) {
        std::complex<long double> acc = 0;
        for (int t = 0; t < n; ++t) {
            long double a = -tau * (long double)((long long)f * t % n) / n;
            acc += x[t] * std::complex<long double>(cosl(a), sinl(a));
        }
        X[f] = acc;
    }
    run_passes<false, T>(line, P, tw, estride);
    double err = 0, nrm = 0;
    for (int f = 0; f < n; ++f) {
        cx<T> v = line[(size_t)pos[f] * estride];
        err = std::max(err, (double)std::abs(std::complex<long double>(v.x, v.y) - X[f]));
        nrm = std::max(nrm, (double)std::abs(X[f]));
    }
    double e1 = err / nrm;
    // periodisation adjacency: for k = 2, 4 (if they divide the 2-part) aliases f + c*n/k are adjacent
    for (int k = 2; k <= 8 && !pow2_first; k *= 2) {
        if (n % k) break;
        bool pow2part_ok = ((n / k) * k == n);
        Plan1 Pc = make_plan1(n / k, max_pow2, false);
        auto posc = scramble_table(Pc);
        for (int u = 0; u < n / k && pow2part_ok; ++u)
            for (int c = 0; c < k; ++c) {
                int pp = pos[u + c * (n / k)];
                if (pp / k != posc[u]) { printf("adjacency violated n=%d k=%d u=%d c=%d\n", n, k, u, c); return 1e9; }
            }
    }
    run_passes<true, T>(line, P, tw, estride);
    double err2 = 0;
    for (int i = 0; i < n; ++i) {
        cx<T> v = line[(size_t)i * estride];
        err2 = std::max(err2, (double)std::abs(std::complex<long double>(v.x / (T)n, v.y / (T)n) - x[i]));
    }
    double e = std::max(e1, err2);
    printf("n=%5d max_pow2=%2d passes=", n, max_pow2);
    for (int p = 0; p < P.npass; ++p) printf("%d ", P.radix[p]);
    printf(" fwd_rel=%.3g inv_abs=%.3g %s\n", e1, err2, e < tol ? "ok" : "FAIL");
    return e;
}

// static (compile-time plan) butterflies: every (flow, sign) combination against the naive DFT
template <int N, bool DIT, int SIGN, typename T> static void run_static(std::vector<cx<T>>& line, const std::vector<cx<T>>& tw) {
    constexpr int NP = ct_plan1(N).npass;
    static_for<0, NP>([&](auto pp_) {
        constexpr int pp = decltype(pp_)::value;
        constexpr int p = DIT ? NP - 1 - pp : pp;
        constexpr int r = ct_plan1(N).radix[p], m = ct_plan1(N).blen[p];
        constexpr int q = m / r, nbf = N / r, tws = N / m;
        for (int bf = 0; bf < nbf; ++bf) {
            const int blk = bf / q, i = bf - blk * q;
            butterfly_s<r, DIT, SIGN, q, 1, false, T>(line.data() + blk * m + i, i * tws, tw.data());
        }
    });
}
template <int N> static int check_static() {
    typedef float T;
    Plan1 P = make_plan1(N);
    auto tw = twiddle_table<T>(N);
    auto pos = scramble_table(P);
    std::vector<std::complex<long double>> x(N);
    srand(77 + N);
    for (int i = 0; i < N; ++i) x[i] = {(long double)rand() / RAND_MAX - 0.5L, (long double)rand() / RAND_MAX - 0.5L};
    const long double tau = 2.0L * 3.14159265358979323846264338327950288L;
    int bad = 0;
    for (int sign = -1; sign <= 1; sign += 2) {
        std::vector<std::complex<long double>> X(N);
        for (int f = 0; f < N; ++f) {
            std::complex<long double> acc = 0;
            for (int t = 0; t < N; ++t) {
                long double a = sign * tau * (long double)((long long)f * t % N) / N;
                acc += x[t] * std::complex<long double>(cosl(a), sinl(a));
            }
            X[f] = acc;
        }
        // DIF: natural in -> scrambled out
        std::vector<cx<T>> line(N);
        for (int i = 0; i < N; ++i) line[i] = mk<T>((T)x[i].real(), (T)x[i].imag());
        if (sign < 0) run_static<N, false, -1, T>(line, tw); else run_static<N, false, +1, T>(line, tw);
        double e1 = 0, nrm = 0;
        for (int f = 0; f < N; ++f) {
            e1 = std::max(e1, (double)std::abs(std::complex<long double>(line[pos[f]].x, line[pos[f]].y) - X[f]));
            nrm = std::max(nrm, (double)std::abs(X[f]));
        }
        // DIT: scrambled in -> natural out
        for (int i = 0; i < N; ++i) line[pos[i]] = mk<T>((T)x[i].real(), (T)x[i].imag());
        if (sign < 0) run_static<N, true, -1, T>(line, tw); else run_static<N, true, +1, T>(line, tw);
        double e2 = 0;
        for (int f = 0; f < N; ++f)
            e2 = std::max(e2, (double)std::abs(std::complex<long double>(line[f].x, line[f].y) - X[f]));
        printf("static N=%d sign=%+d DIF rel=%.3g DIT rel=%.3g\n", N, sign, e1 / nrm, e2 / nrm);
        if (e1 / nrm > 2e-5 || e2 / nrm > 2e-5) ++bad;
    }
    return bad;
}

// prime-factor variant (no twiddles between the pow2 and the odd-prime pass): DIF with pfa_in / pfa_out positions and the
// transposed DIT flow, both signs, against the naive DFT
template <int N, bool DIT, int SIGN, typename T> static void run_static_pfa(std::vector<cx<T>>& line, const std::vector<cx<T>>& tw) {
    constexpr int NP = ct_plan1(N).npass;
    static_for<0, NP>([&](auto pp_) {
        constexpr int pp = decltype(pp_)::value;
        constexpr int p = DIT ? NP - 1 - pp : pp;
        constexpr int r = ct_plan1(N).radix[p], m = ct_plan1(N).blen[p];
        constexpr int q = m / r, nbf = N / r, tws = N / m;
        for (int bf = 0; bf < nbf; ++bf) {
            const int blk = bf / q, i = bf - blk * q;
            butterfly_s<r, DIT, SIGN, q, 1, false, T, false>(line.data() + blk * m + i, i * tws, tw.data());
        }
    });
}
template <int N> static int check_pfa() {
    typedef float T;
    static_assert(ct_pfa_ok(N), "not a prime-factor length");
    auto tw = twiddle_table<T>(N);
    std::vector<int> pin, pout;
    pfa_tables(N, pin, pout);
    std::vector<char> seen(N, 0);
    for (int i = 0; i < N; ++i) { if (seen[pin[i]]) return 1; seen[pin[i]] = 1; }
    std::vector<std::complex<long double>> x(N);
    srand(99 + N);
    for (int i = 0; i < N; ++i) x[i] = {(long double)rand() / RAND_MAX - 0.5L, (long double)rand() / RAND_MAX - 0.5L};
    const long double tau = 2.0L * 3.14159265358979323846264338327950288L;
    int bad = 0;
    for (int sign = -1; sign <= 1; sign += 2) {
        std::vector<std::complex<long double>> X(N);
        double nrm = 0;
        for (int f = 0; f < N; ++f) {
            std::complex<long double> acc = 0;
            for (int t = 0; t < N; ++t) {
                long double a = sign * tau * (long double)((long long)f * t % N) / N;
                acc += x[t] * std::complex<long double>(cosl(a), sinl(a));
            }
            X[f] = acc; nrm = std::max(nrm, (double)std::abs(acc));
        }
        std::vector<cx<T>> line(N);
        for (int i = 0; i < N; ++i) line[pin[i]] = mk<T>((T)x[i].real(), (T)x[i].imag());
        if (sign < 0) run_static_pfa<N, false, -1, T>(line, tw); else run_static_pfa<N, false, +1, T>(line, tw);
        double e1 = 0, e2 = 0;
        for (int f = 0; f < N; ++f)
            e1 = std::max(e1, (double)std::abs(std::complex<long double>(line[pout[f]].x, line[pout[f]].y) - X[f]));
        for (int i = 0; i < N; ++i) line[pout[i]] = mk<T>((T)x[i].real(), (T)x[i].imag());
        if (sign < 0) run_static_pfa<N, true, -1, T>(line, tw); else run_static_pfa<N, true, +1, T>(line, tw);
        for (int f = 0; f < N; ++f)
            e2 = std::max(e2, (double)std::abs(std::complex<long double>(line[pin[f]].x, line[pin[f]].y) - X[f]));
        printf("pfa N=%d sign=%+d DIF rel=%.3g DIT rel=%.3g\n", N, sign, e1 / nrm, e2 / nrm);
        if (e1 / nrm > 2e-5 || e2 / nrm > 2e-5) ++bad;
    }
    return bad;
}

int main() {
    int sizes[] = {1, 2, 3, 4, 5, 6, 8, 9, 10, 12, 16, 17, 18, 20, 24, 30, 32, 34, 36, 40, 48, 60, 64, 66, 68, 96,
                   120, 128, 136, 138, 240, 256, 272, 286, 290, 442, 512, 1016, 1024, 4096};
    int bad = 0;
    for (int n : sizes) {
        for (int mp : {16, 8, 4, 2}) {
            if (n > 300 && mp < 16) continue;
            if (check<float>(n, mp, 2e-5, false) > 2e-5) ++bad;
            if (check<float>(n, mp, 2e-5, true) > 2e-5) ++bad;
        }
        if (check<double>(n, 16, 1e-12, true) > 1e-12) ++bad;
    }
    bad += check_static<136>() + check_static<68>() + check_static<272>() + check_static<128>() +
           check_static<240>() + check_static<34>();
    bad += check_pfa<136>() + check_pfa<68>() + check_pfa<272>() + check_pfa<34>();
    printf(bad ? "FAILED %d\n" : "ALL OK\n", bad);
    return bad ? 1 : 0;
}
