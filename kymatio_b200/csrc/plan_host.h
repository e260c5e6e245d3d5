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

// Factor n into DIF pass radices: the power-of-two part split into radices <= max_pow2 and
// the odd primes (descending).  pow2_first = true puts the power-of-two passes first, which
// makes the natural -> scrambled scatter of consecutive frequencies hit distinct shared-memory
// banks (pos jumps by odd multiples); false gives the "odd first" order whose scrambled
// layout nests across resolutions.  Throws for unsupported sizes.
inline Plan1 make_plan1(int n, int max_pow2 = 16, bool pow2_first = true) {
    if (n < 1) throw std::runtime_error("FFT length must be >= 1");
    Plan1 P{};
    P.n = n;
    int two = 0, m = n;
    while (m % 2 == 0) { m /= 2; ++two; }
    std::vector<int> odd;
    for (int p = 3; (long long)p * p <= m; p += 2)
        while (m % p == 0) { odd.push_back(p); m /= p; }
    if (m > 1) odd.push_back(m);
    // descending
    for (size_t i = 0; i < odd.size(); ++i)
        for (size_t j = i + 1; j < odd.size(); ++j)
            if (odd[j] > odd[i]) std::swap(odd[i], odd[j]);
    std::vector<int> rad, rad2;
    for (int p : odd) {
        if (!radix_compiled(p) && p > kMaxGenericRadix)
            throw std::runtime_error("FFT length " + std::to_string(n) + " has prime factor " +
                                     std::to_string(p) + " > " + std::to_string(kMaxGenericRadix) +
                                     " (unsupported)");
        rad.push_back(p);
    }
    int lgmax = 0; while ((1 << (lgmax + 1)) <= max_pow2) ++lgmax;
    // balanced split of the 2-part into ceil(two/lgmax) passes
    if (two > 0) {
        int np = (two + lgmax - 1) / lgmax;
        int base = two / np, extra = two % np;
        for (int i = 0; i < np; ++i) rad2.push_back(1 << (base + (i < extra ? 1 : 0)));
    }
    if (pow2_first) rad.insert(rad.begin(), rad2.begin(), rad2.end());
    else rad.insert(rad.end(), rad2.begin(), rad2.end());
    if (rad.empty()) rad.push_back(1);  // n == 1: degenerate
    if ((int)rad.size() > kMaxPass) throw std::runtime_error("too many FFT passes");
    if (n == 1) { P.npass = 0; return P; }
    P.npass = (int)rad.size();
    int bl = n;
    for (int p = 0; p < P.npass; ++p) { P.radix[p] = rad[p]; P.blen[p] = bl; bl /= rad[p]; }
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
