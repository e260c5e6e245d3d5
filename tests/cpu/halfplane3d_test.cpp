// CPU emulation of the two-pass 3-D transform of kernels3d.cuh on the host+device butterflies of fft_core.cuh:
//   k3d_col_prod : product, radix-2 DIF along O (top = a + b, bottom = (a - b) e^{+2 pi i o / O}), inverse DIF along M
//   k3d_plane    : per half plane h, inverse DIF along o' (length O/2) and along n -> the field at scrambled positions
//                  (p, qn, qo, h) <-> voxel (invM[p], invN[qn], 2 invH[qo] + h); modulus; forward DIT along n and o'
//   k3d_col_fwd  : radix-2 DIT along O (X[f] = E[f] + w^f D[f], X[f + O/2] = E[f] - w^f D[f]), forward DIT along M
// checked against long-double O(N^2) evaluations of ifftn(U_hat * Psi) and fftn(|.|).  Exit code 0 on success.
#include <cstdio>
#include <cstdlib>
#include <complex>
#include <vector>
#include "../../kymatio_b200/csrc/plan_host.h"

using namespace sb;
typedef float T;
typedef std::complex<long double> cld;
static const long double TAU = 2.0L * 3.14159265358979323846264338327950288L;

template <int N, bool DIT, int SIGN> static void fft_line(cx<T>* s, const cx<T>* tw) {
    constexpr int NP = ct_plan1(N).npass;
    static_for<0, NP>([&](auto pp_) {
        constexpr int pp = decltype(pp_)::value;
        constexpr int p = DIT ? NP - 1 - pp : pp;
        constexpr int r = ct_plan1(N).radix[p], m = ct_plan1(N).blen[p];
        constexpr int q = m / r, nbf = N / r, tws = N / m;
        for (int bf = 0; bf < nbf; ++bf) {
            const int blk = bf / q, i = bf - blk * q;
            butterfly_s<r, DIT, SIGN, q, 1, false, T>(s + blk * m + i, i * tws, tw);
        }
    });
}
static std::vector<int> inverse_of(const std::vector<int>& pos) {
    std::vector<int> inv(pos.size());
    for (size_t f = 0; f < pos.size(); ++f) inv[pos[f]] = (int)f;
    return inv;
}

template <int M, int N, int O> static double run() {
    constexpr int OH = O / 2, V = M * N * O;
    auto twM = twiddle_table<T>(M); auto twN = twiddle_table<T>(N); auto twO = twiddle_table<T>(O); auto twH = twiddle_table<T>(OH);
    auto invM = inverse_of(scramble_table(ct_plan1(M))), invN = inverse_of(scramble_table(ct_plan1(N))),
         invH = inverse_of(scramble_table(ct_plan1(OH)));
    auto idx = [](int m, int n, int o) { return (m * N + n) * O + o; };
    srand(11 + V);
    std::vector<cld> Uh(V), Psi(V), P(V), u(V), Xref(V);
    for (int i = 0; i < V; ++i) {
        Uh[i] = cld((long double)rand() / RAND_MAX - 0.5L, (long double)rand() / RAND_MAX - 0.5L);
        Psi[i] = cld((long double)rand() / RAND_MAX - 0.5L, (long double)rand() / RAND_MAX - 0.5L);
        P[i] = Uh[i] * Psi[i];
    }
    auto dft3 = [&](const std::vector<cld>& in, std::vector<cld>& out, int sign, long double scale) {
        std::vector<cld> a(V), b(V);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) for (int f = 0; f < O; ++f) {       // along o
            cld acc = 0; for (int o = 0; o < O; ++o) acc += in[idx(m, n, o)] * std::polar(1.0L, sign * TAU * (f * o % O) / O);
            a[idx(m, n, f)] = acc; }
        for (int m = 0; m < M; ++m) for (int f = 0; f < N; ++f) for (int o = 0; o < O; ++o) {       // along n
            cld acc = 0; for (int n = 0; n < N; ++n) acc += a[idx(m, n, o)] * std::polar(1.0L, sign * TAU * (f * n % N) / N);
            b[idx(m, f, o)] = acc; }
        for (int f = 0; f < M; ++f) for (int n = 0; n < N; ++n) for (int o = 0; o < O; ++o) {       // along m
            cld acc = 0; for (int m = 0; m < M; ++m) acc += b[idx(m, n, o)] * std::polar(1.0L, sign * TAU * (f * m % M) / M);
            out[idx(f, n, o)] = acc * scale; }
    };
    dft3(P, u, +1, 1.0L / V);
    std::vector<cld> Uabs(V);
    for (int i = 0; i < V; ++i) Uabs[i] = std::abs(u[i]);
    dft3(Uabs, Xref, -1, 1.0L);

    // k3d_col_prod
    std::vector<cx<T>> Y(V), line(std::max(M, std::max(N, O)));
    for (int n = 0; n < N; ++n) for (int o = 0; o < OH; ++o) {
        std::vector<cx<T>> top(M), bot(M);
        for (int m = 0; m < M; ++m) {
            const cld a = P[idx(m, n, o)] / (long double)V, b = P[idx(m, n, o + OH)] / (long double)V;
            const cx<T> A = mk<T>((T)a.real(), (T)a.imag()), B = mk<T>((T)b.real(), (T)b.imag());
            top[m] = A + B; bot[m] = cmulc(A - B, twO[o]);
        }
        fft_line<M, false, +1>(top.data(), twM.data());
        fft_line<M, false, +1>(bot.data(), twM.data());
        for (int p = 0; p < M; ++p) { Y[idx(p, n, o)] = top[p]; Y[idx(p, n, OH + o)] = bot[p]; }
    }
    // k3d_plane (inverse, modulus, forward) per half plane
    std::vector<cx<T>> S(V);
    double err_u = 0;
    for (int p = 0; p < M; ++p) for (int h = 0; h < 2; ++h) {
        std::vector<cx<T>> s((size_t)N * OH);
        for (int n = 0; n < N; ++n) for (int o = 0; o < OH; ++o) s[n * OH + o] = Y[idx(p, n, h * OH + o)];
        for (int n = 0; n < N; ++n) fft_line<OH, false, +1>(&s[n * OH], twH.data());
        for (int o = 0; o < OH; ++o) {
            for (int n = 0; n < N; ++n) line[n] = s[n * OH + o];
            fft_line<N, false, +1>(line.data(), twN.data());
            for (int n = 0; n < N; ++n) s[n * OH + o] = line[n];
        }
        for (int qn = 0; qn < N; ++qn) for (int qo = 0; qo < OH; ++qo) {
            const cx<T> v = s[qn * OH + qo];
            const long double mod = sqrtl((long double)v.x * v.x + (long double)v.y * v.y);
            err_u = std::max(err_u, (double)fabsl(mod - Uabs[idx(invM[p], invN[qn], 2 * invH[qo] + h)].real()));
            s[qn * OH + qo] = mk<T>((T)mod, T(0));
        }
        for (int o = 0; o < OH; ++o) {
            for (int n = 0; n < N; ++n) line[n] = s[n * OH + o];
            fft_line<N, true, -1>(line.data(), twN.data());
            for (int n = 0; n < N; ++n) s[n * OH + o] = line[n];
        }
        for (int n = 0; n < N; ++n) fft_line<OH, true, -1>(&s[n * OH], twH.data());
        for (int n = 0; n < N; ++n) for (int o = 0; o < OH; ++o) S[idx(p, n, h * OH + o)] = s[n * OH + o];
    }
    // k3d_col_fwd
    double err = 0, ref_max = 0;
    for (int n = 0; n < N; ++n) for (int o = 0; o < OH; ++o) {
        std::vector<cx<T>> lo(M), hi(M);
        for (int p = 0; p < M; ++p) {
            const cx<T> e = S[idx(p, n, o)], t = cmul(S[idx(p, n, OH + o)], twO[o]);
            lo[p] = e + t; hi[p] = e - t;
        }
        fft_line<M, true, -1>(lo.data(), twM.data());
        fft_line<M, true, -1>(hi.data(), twM.data());
        for (int f = 0; f < M; ++f) {
            const cld w0 = Xref[idx(f, n, o)], w1 = Xref[idx(f, n, o + OH)];
            ref_max = std::max(ref_max, (double)std::max(std::abs(w0), std::abs(w1)));
            err = std::max(err, (double)std::abs(cld(lo[f].x, lo[f].y) - w0));
            err = std::max(err, (double)std::abs(cld(hi[f].x, hi[f].y) - w1));
        }
    }
    printf("half-plane 3-D %dx%dx%d: spectrum rel err %.2e, modulus field abs err %.2e\n", M, N, O, err / ref_max, err_u);
    return std::max(err / ref_max, err_u);
}

int main() {
    double worst = 0;
    worst = std::max(worst, run<8, 8, 16>());
    worst = std::max(worst, run<4, 16, 32>());
    worst = std::max(worst, run<12, 8, 48>());
    if (worst < 2e-6) { printf("ALL OK\n"); return 0; }
    printf("FAILED (worst %.3e)\n", worst);
    return 1;
}
