// plan_host.h - host-side construction of 1-D FFT plans, scramble tables, twiddles.
#pragma once
#include <cmath>
#include <vector>
#include <stdexcept>
#include <string>
#include "fft_core.cuh"

namespace sb {

inline bool radix_compiled(int r) {
    switch (r) { case 2: case 3: case 4: case 5: case 7: case 8: case 11: case 13: case 16: case 17: return true; }
    return false;
}

// Host wrapper around ct_plan1 (fft_core.cuh) with validation.  Throws for unsupported sizes.
inline Plan1 make_plan1(int n, int max_pow2 = 16, bool pow2_first = true) {
    if (n < 1) throw std::runtime_error("FFT length must be >= 1");
    Plan1 P = ct_plan1(n, max_pow2, pow2_first);
    if (P.npass < 0) throw std::runtime_error("too many FFT passes for length " + std::to_string(n));
    for (int p = 0; p < P.npass; ++p)
        if (!radix_compiled(P.radix[p]) && P.radix[p] > kMaxGenericRadix)
            throw std::runtime_error("FFT length " + std::to_string(n) + " has prime factor " +
                                     std::to_string(P.radix[p]) + " > " + std::to_string(kMaxGenericRadix) +
                                     " (unsupported)");
    return P;
}

// pos[f] = canonical scrambled position of natural frequency f.
inline std::vector<int> scramble_table(const Plan1& P) {
    std::vector<int> pos(P.n);
    for (int f0 = 0; f0 < P.n; ++f0) {
        int f = f0, position = 0;
        for (int p = 0; p < P.npass; ++p) {
            int r = P.radix[p], q = P.blen[p] / r;
            int d = f % r; f /= r;
            int slot = ct_is_pow2(r) ? rt_bitrev(d, ct_log2(r)) : d;
            position += slot * q;
        }
        pos[f0] = position;
    }
    return pos;
}

// prime-factor tables (fft_core.cuh): pin[n] = position of natural INPUT index n of the DIF flow (= where the DIT flow
// leaves natural output n), pout[k] = position of natural OUTPUT index k of the DIF flow (= DIT input).  Identity /
// scramble_table for lengths without a prime-factor plan.
inline bool pfa_ok(int n) { return n >= 2 && ct_pfa_ok(n); }
inline void pfa_tables(int n, std::vector<int>& pin, std::vector<int>& pout) {
    pin.resize(n); pout.resize(n);
    const Plan1 P = ct_plan1(n);
    if (!pfa_ok(n)) {
        const std::vector<int> pos = scramble_table(P);
        for (int i = 0; i < n; ++i) { pin[i] = i; pout[i] = pos[i]; }
        return;
    }
    const int N1 = P.radix[0], N2 = P.radix[1];
    const int inv21 = ct_modinv(N2, N1), inv12 = ct_modinv(N1, N2);
    for (int i = 0; i < n; ++i) { pin[i] = pfa_in(i, N1, N2, inv21, inv12); pout[i] = pfa_out(i, N1, N2); }
}

// tw[j] = exp(-2*pi*i*j/n)
template <typename T> inline std::vector<cx<T>> twiddle_table(int n) {
    std::vector<cx<T>> tw(n > 0 ? n : 1);
    for (int j = 0; j < n; ++j) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)n;
        tw[j].x = (T)cosl(a); tw[j].y = (T)sinl(a);
    }
    if (n <= 0) { tw[0].x = 1; tw[0].y = 0; }
    return tw;
}

}  // namespace sb
