// fft_core.cuh - register-level mixed-radix FFT building blocks (host+device).
//
// A 1-D transform of length n is a sequence of in-place radix passes over a line (Plan1).  Two data
// flows are provided, each for either exponent sign:
//   * DIF (decimation in frequency): natural-order input  -> "scrambled" (digit-reversed) output
//   * DIT (decimation in time)     : scrambled input      -> natural-order output (the exact
//                                    transpose of the DIF flow graph, same plan, same positions)
// scramble_table() (plan_host.h) gives pos[f], the scrambled position of natural index f.
// The kernels pair them so that no permutation pass is needed: inverse transforms run as DIF
// (natural Fourier data in), forward transforms as DIT (natural Fourier data out); the spatial field in
// between stays scrambled and is only consumed by order-agnostic ops (modulus) or through pos[].
// Generic runtime-size kernels use the older pairing (DIT inverse / DIF forward with pos[] gathers).
//
// Radices: 2, 4, 8, 16 (in-register radix-2 networks with compile-time twiddles) and odd primes
// 3, 5, 7, 11, 13, 17 (x_n +- x_{R-n} pairing, compile-time constants); any other odd prime <= 127 goes
// through a generic O(R^2) butterfly.  On sm_100+ complex add / subtract / multiply-by-real use the
// packed fp32 instructions (FADD2 / FFMA2).
#pragma once
#include <type_traits>
#include <utility>

#if defined(__CUDACC__)
#define SB_HD __host__ __device__ __forceinline__
#else
#define SB_HD inline
#endif

namespace sb {

template <typename T> struct alignas(2 * sizeof(T)) cx { T x, y; };

template <typename T> SB_HD cx<T> mk(T a, T b) { cx<T> r; r.x = a; r.y = b; return r; }
template <typename T> SB_HD cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> SB_HD cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
// acc + a * c with a complex and c real (the inner step of the odd-prime DFTs)
template <typename T> SB_HD cx<T> fma_rc(cx<T> acc, cx<T> a, T c) { return mk<T>(acc.x + a.x * c, acc.y + a.y * c); }

// ---------------------------------------------------------------------------
// Blackwell packed fp32 (FADD2 / FFMA2, PTX add/sub/fma.f32x2, sm_100+): a complex number is one
// (re, im) register pair, so complex add / subtract / multiply-by-real each issue ONE instruction
// instead of two.  The kernels are issue-bound on the butterfly arithmetic, so this matters.
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
__device__ __forceinline__ unsigned long long sb_pk(cx<float> a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ cx<float> sb_upk(unsigned long long v) {
    cx<float> r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ cx<float> operator+(cx<float> a, cx<float> b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(sb_pk(a)), "l"(sb_pk(b)));
    return sb_upk(r);
}
__device__ __forceinline__ cx<float> operator-(cx<float> a, cx<float> b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(sb_pk(a)), "l"(sb_pk(b)));
    return sb_upk(r);
}
__device__ __forceinline__ cx<float> fma_rc(cx<float> acc, cx<float> a, float c) {
    unsigned long long r;
    cx<float> cc; cc.x = c; cc.y = c;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(sb_pk(a)), "l"(sb_pk(cc)), "l"(sb_pk(acc)));
    return sb_upk(r);
}
#endif
// element-wise acc + a * b on the (x, y) pair (one FFMA2 for float on sm_100+)
template <typename T> SB_HD cx<T> fma_cc(cx<T> acc, cx<T> a, cx<T> b) { return mk<T>(acc.x + a.x * b.x, acc.y + a.y * b.y); }
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
__device__ __forceinline__ cx<float> fma_cc(cx<float> acc, cx<float> a, cx<float> b) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(sb_pk(a)), "l"(sb_pk(b)), "l"(sb_pk(acc)));
    return sb_upk(r);
}
#endif
template <typename T> SB_HD cx<T> cmul(cx<T> a, cx<T> b) { return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
template <typename T> SB_HD cx<T> cmulc(cx<T> a, cx<T> b) { return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
template <typename T> SB_HD cx<T> scal(cx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }

// |v|: the reference computes sqrt(x^2 + y^2) (kymatio/backend/torch_backend.py:57); on the device the
// float version uses m2 * rsqrt(m2) (<= 2 ulp), exact zero at zero.
template <typename T> SB_HD T cabs_fast(cx<T> v) {
    const T m2 = v.x * v.x + v.y * v.y;
#if defined(__CUDA_ARCH__)
    if constexpr (std::is_same<T, float>::value) return m2 > 0.f ? m2 * rsqrtf(m2) : 0.f;
    else return sqrt(m2);
#else
    return (T)__builtin_sqrt((double)m2);
#endif
}

// ---------------------------------------------------------------------------
// compile-time trigonometry (Taylor series in double; |x| <= pi)
// ---------------------------------------------------------------------------
constexpr double kPi = 3.14159265358979323846264338327950288;

constexpr double ct_sin_series(double x) {
    double term = x, sum = x;
    for (int k = 1; k < 20; ++k) { term *= -x * x / double((2 * k) * (2 * k + 1)); sum += term; }
    return sum;
}
constexpr double ct_cos_series(double x) {
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 20; ++k) { term *= -x * x / double((2 * k - 1) * (2 * k)); sum += term; }
    return sum;
}
// cos / sin of 2*pi*k/n with exact octant reduction on the integer ratio
constexpr double ct_cos2pi(int k, int n) {
    k %= n; if (k < 0) k += n;
    if (2 * k > n) k = n - k;                 // cos(2pi - t) = cos t   -> t in [0, pi]
    if (4 * k > n) return -ct_cos2pi(n - 2 * k, 2 * n);  // cos(t) = -cos(pi - t); pi - t = 2pi (n-2k)/(2n)
    if (8 * k > n) return ct_sin_series(2.0 * kPi * double(n - 4 * k) / double(4 * n)); // cos t = sin(pi/2 - t)
    return ct_cos_series(2.0 * kPi * double(k) / double(n));
}
constexpr double ct_sin2pi(int k, int n) {
    k %= n; if (k < 0) k += n;
    if (2 * k > n) return -ct_sin2pi(n - k, n);
    if (4 * k > n) return ct_sin2pi(n - 2 * k, 2 * n);      // sin(t) = sin(pi - t)
    if (8 * k > n) return ct_cos_series(2.0 * kPi * double(n - 4 * k) / double(4 * n)); // sin t = cos(pi/2 - t)
    return ct_sin_series(2.0 * kPi * double(k) / double(n));
}
template <int K, int N> struct Root {            // exp(+2*pi*i*K/N)
    static constexpr double c = ct_cos2pi(K, N);
    static constexpr double s = ct_sin2pi(K, N);
};

template <int B, int E, typename F> SB_HD void static_for(F&& f) {
    if constexpr (B < E) { f(std::integral_constant<int, B>{}); static_for<B + 1, E>(f); }
}

constexpr int ct_log2(int r) { return r <= 1 ? 0 : 1 + ct_log2(r / 2); }
constexpr int ct_bitrev(int v, int bits) { int r = 0; for (int i = 0; i < bits; ++i) { r = (r << 1) | ((v >> i) & 1); } return r; }
SB_HD int rt_bitrev(int v, int bits) { int r = 0; for (int i = 0; i < bits; ++i) { r = (r << 1) | ((v >> i) & 1); } return r; }

// d * exp(SIGN * 2*pi*i * J / N) with the trivial cases folded at compile time
template <int J, int N, int SIGN, typename T> SB_HD cx<T> mul_root(cx<T> d) {
    constexpr int j = ((J % N) + N) % N;
    if constexpr (j == 0) return d;
    else if constexpr (2 * j == N) return mk<T>(-d.x, -d.y);
    else if constexpr (4 * j == N) return SIGN > 0 ? mk<T>(-d.y, d.x) : mk<T>(d.y, -d.x);       // * (+-i)
    else if constexpr (4 * j == 3 * N) return SIGN > 0 ? mk<T>(d.y, -d.x) : mk<T>(-d.y, d.x);   // * (-+i)
    else {
        constexpr T c = T(Root<j, N>::c);
        constexpr T s = T(SIGN > 0 ? Root<j, N>::s : -Root<j, N>::s);
        return mk<T>(d.x * c - d.y * s, d.x * s + d.y * c);
    }
}

// ---------------------------------------------------------------------------
// in-register power-of-two transforms.
//   dif_pow2: natural v[c] in  -> v[k] = y_{bitrev(k)} out
//   dit_pow2: v[k] = y_{bitrev(k)} in -> natural x[c] out   (transpose network)
// y_f = sum_c x_c exp(SIGN*2*pi*i*f*c/R)
// ---------------------------------------------------------------------------
template <int R, int SIGN, typename T> SB_HD void dif_pow2(cx<T>* v) {
    constexpr int LG = ct_log2(R);
    static_for<0, LG>([&](auto s_) {
        constexpr int s = decltype(s_)::value;
        constexpr int half = R >> (s + 1);
        static_for<0, R / (2 * half)>([&](auto b_) {
            constexpr int b = decltype(b_)::value;
            static_for<0, half>([&](auto j_) {
                constexpr int j = decltype(j_)::value;
                constexpr int i0 = b * 2 * half + j, i1 = i0 + half;
                cx<T> a = v[i0], c = v[i1];
                v[i0] = a + c;
                v[i1] = mul_root<j, 2 * half, SIGN, T>(a - c);
            });
        });
    });
}
template <int R, int SIGN, typename T> SB_HD void dit_pow2(cx<T>* v) {
    constexpr int LG = ct_log2(R);
    static_for<0, LG>([&](auto s_) {
        constexpr int s = LG - 1 - decltype(s_)::value;
        constexpr int half = R >> (s + 1);
        static_for<0, R / (2 * half)>([&](auto b_) {
            constexpr int b = decltype(b_)::value;
            static_for<0, half>([&](auto j_) {
                constexpr int j = decltype(j_)::value;
                constexpr int i0 = b * 2 * half + j, i1 = i0 + half;
                cx<T> a = v[i0], c = mul_root<j, 2 * half, SIGN, T>(v[i1]);
                v[i0] = a + c;
                v[i1] = a - c;
            });
        });
    });
}

// ---------------------------------------------------------------------------
// in-register odd-prime DFT (natural in, natural out; the matrix is symmetric so the
// same routine serves DIF and DIT).  Uses the x_n +- x_{R-n} pairing: (R-1)^2/2 real
// FMAs with compile-time constants.
// ---------------------------------------------------------------------------
template <int R, int SIGN, typename T> SB_HD void dft_prime(cx<T>* v) {
    constexpr int H = (R - 1) / 2;
    cx<T> a[H], bi[H];
    static_for<0, H>([&](auto n_) {
        constexpr int n = decltype(n_)::value;
        a[n] = v[n + 1] + v[R - 1 - n];
        const cx<T> b = v[n + 1] - v[R - 1 - n];
        bi[n] = mk<T>(-b.y, b.x);                       // i * (x_n - x_{R-n})
    });
    const cx<T> x0 = v[0];
    cx<T> s0 = x0;
    static_for<0, H>([&](auto n_) { s0 = s0 + a[decltype(n_)::value]; });
    v[0] = s0;
    static_for<1, H + 1>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        cx<T> p = x0, qi = mk<T>(T(0), T(0));
        static_for<0, H>([&](auto n_) {
            constexpr int n = decltype(n_)::value;
            constexpr int idx = ((n + 1) * k) % R;
            constexpr T c = T(Root<idx, R>::c);
            constexpr T sn = T(Root<idx, R>::s);
            p = fma_rc(p, a[n], c);
            qi = fma_rc(qi, bi[n], sn);
        });
        // y_k = p + SIGN * i*q, y_{R-k} = p - SIGN * i*q  with qi = i*q
        if (SIGN > 0) { v[k] = p + qi; v[R - k] = p - qi; }
        else          { v[k] = p - qi; v[R - k] = p + qi; }
    });
}

// ---------------------------------------------------------------------------
// 1-D plan: radices in DIF order (odd primes descending, then powers of two),
// blen[p] = block length the p-th DIF pass works on (n, n/r0, n/(r0 r1), ...).
// ---------------------------------------------------------------------------
constexpr int kMaxPass = 12;
constexpr int kMaxGenericRadix = 128;
struct Plan1 {
    int n;
    int npass;
    int radix[kMaxPass];
    int blen[kMaxPass];
};

constexpr bool ct_is_pow2(int r) { return (r & (r - 1)) == 0; }

// Factor n into DIF pass radices (usable at compile time and on the host):
//   power-of-two part split into balanced radices <= max_pow2, odd primes descending;
//   pow2_first selects which group runs first in the DIF order.
constexpr Plan1 ct_plan1(int n, int max_pow2 = 16, bool pow2_first = true) {
    Plan1 P{};
    P.n = n;
    if (n <= 1) { P.npass = 0; return P; }
    int two = 0, m = n;
    while (m % 2 == 0) { m /= 2; ++two; }
    int odd[kMaxPass] = {}; int nodd = 0;
    for (int p = 3; (long long)p * p <= m; p += 2)
        while (m % p == 0) { if (nodd < kMaxPass) odd[nodd] = p; ++nodd; m /= p; }
    if (m > 1) { if (nodd < kMaxPass) odd[nodd] = m; ++nodd; }
    for (int i = 0; i < nodd && i < kMaxPass; ++i)
        for (int j = i + 1; j < nodd && j < kMaxPass; ++j)
            if (odd[j] > odd[i]) { int t = odd[i]; odd[i] = odd[j]; odd[j] = t; }
    int lgmax = 0; while ((1 << (lgmax + 1)) <= max_pow2) ++lgmax;
    int p2[kMaxPass] = {}; int np2 = 0;
    if (two > 0) {
        np2 = (two + lgmax - 1) / lgmax;
        const int base = two / np2, extra = two % np2;
        for (int i = 0; i < np2 && i < kMaxPass; ++i) p2[i] = 1 << (base + (i < extra ? 1 : 0));
    }
    if (nodd + np2 > kMaxPass) { P.npass = -1; return P; }   // too many passes (reported by the host wrapper)
    int k = 0;
    if (pow2_first) { for (int i = 0; i < np2; ++i) P.radix[k++] = p2[i]; for (int i = 0; i < nodd; ++i) P.radix[k++] = odd[i]; }
    else            { for (int i = 0; i < nodd; ++i) P.radix[k++] = odd[i]; for (int i = 0; i < np2; ++i) P.radix[k++] = p2[i]; }
    P.npass = k;
    int bl = n;
    for (int p = 0; p < k; ++p) { P.blen[p] = bl; bl /= P.radix[p]; }
    return P;
}
template <int N> struct SPlan { static constexpr Plan1 p = ct_plan1(N); };

// One radix-R butterfly of pass (block length m = R*q) on one line.
//   INV=false : DIF forward  (load natural group, DFT, twiddle, store to slots)
//   INV=true  : DIT inverse  (load slots, conj twiddle, DFT(+), store natural group)
// `line` addresses element e at line[e*estride]; base = blk*m + i; twstep = i*(n/m).
template <int R, bool INV, typename T>
SB_HD void butterfly(cx<T>* line, int estride, int base, int q, int twstep, const cx<T>* tw, bool do_tw) {
    constexpr int SIGN = INV ? +1 : -1;
    constexpr bool P2 = ct_is_pow2(R);
    constexpr int LG = ct_log2(R);
    cx<T> v[R];
    static_for<0, R>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        v[k] = line[(base + k * q) * estride];
    });
    if (!INV) {
        if constexpr (P2) dif_pow2<R, SIGN, T>(v); else dft_prime<R, SIGN, T>(v);
        if (do_tw) {
            static_for<1, R>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                constexpr int f = P2 ? ct_bitrev(k, LG) : k;
                v[k] = cmul(v[k], tw[f * twstep]);
            });
        }
    } else {
        if (do_tw) {
            static_for<1, R>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                constexpr int f = P2 ? ct_bitrev(k, LG) : k;
                v[k] = cmulc(v[k], tw[f * twstep]);
            });
        }
        if constexpr (P2) dit_pow2<R, SIGN, T>(v); else dft_prime<R, SIGN, T>(v);
    }
    static_for<0, R>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        line[(base + k * q) * estride] = v[k];
    });
}

// Same butterfly with the group stride Q (in elements) and the element stride ES known at compile
// time: every shared-memory access becomes base + immediate offset.  The flow (DIT: twiddle then
// butterfly, scrambled -> natural; DIF: butterfly then twiddle, natural -> scrambled) and the sign of
// the exponent are independent, so an inverse transform can run as DIF (natural-order input) and a
// forward one as DIT (natural-order output).
// the register part of butterfly_s: v[0..R) holds the group (natural order in for DIF, slot order in for DIT)
template <int R, bool DIT, int SIGN, int Q, typename T, bool TW = (Q > 1)>
SB_HD void butterfly_v(cx<T>* v, int twstep, const cx<T>* tw) {
    constexpr bool P2 = ct_is_pow2(R);
    constexpr int LG = ct_log2(R);
    if (!DIT) {
        if constexpr (P2) dif_pow2<R, SIGN, T>(v); else dft_prime<R, SIGN, T>(v);
        if constexpr (TW) {
            static_for<1, R>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                constexpr int f = P2 ? ct_bitrev(k, LG) : k;
                v[k] = SIGN < 0 ? cmul(v[k], tw[f * twstep]) : cmulc(v[k], tw[f * twstep]);
            });
        }
    } else {
        if constexpr (TW) {
            static_for<1, R>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                constexpr int f = P2 ? ct_bitrev(k, LG) : k;
                v[k] = SIGN < 0 ? cmul(v[k], tw[f * twstep]) : cmulc(v[k], tw[f * twstep]);
            });
        }
        if constexpr (P2) dit_pow2<R, SIGN, T>(v); else dft_prime<R, SIGN, T>(v);
    }
}

// MODULUS: 0 store the outputs; 1 store (|v|, 0); 2 store (|v|, |v|) - the duplicated form lets a consumer that only needs
// the real field fetch a ready-made packed operand pair for FFMA2 with one 64-bit load
template <int R, bool DIT, int SIGN, int Q, int ES, int MODULUS, typename T, bool TW = (Q > 1)>
SB_HD void butterfly_s(cx<T>* p0, int twstep, const cx<T>* tw) {
    cx<T> v[R];
    static_for<0, R>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        v[k] = p0[k * Q * ES];
    });
    butterfly_v<R, DIT, SIGN, Q, T, TW>(v, twstep, tw);
    static_for<0, R>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        if constexpr (MODULUS == 1) p0[k * Q * ES] = mk<T>(cabs_fast<T>(v[k]), T(0));
        else if constexpr (MODULUS == 2) { const T m = cabs_fast<T>(v[k]); p0[k * Q * ES] = mk<T>(m, m); }
        else p0[k * Q * ES] = v[k];
    });
}

// ---------------------------------------------------------------------------------------------------------------
// Good-Thomas / prime-factor variant for N = N1 * N2, N1 = 2^a (one radix pass), N2 an odd prime (one pass): the two
// factors are coprime, so with the index maps
//     input   n  = (N2 n1 + N1 n2) mod N   stored at position  n1 N2 + n2          (pfa_in)
//     output  k  : k1 = k mod N1, k2 = k mod N2   found at      slot(k1) N2 + k2    (pfa_out; slot = bit reversal)
// W_N^{nk} = W_N1^{n1 k1} W_N2^{n2 k2}: a true two-dimensional DFT - the SAME two butterfly passes as the Cooley-Tukey
// plan (same strides, same positions touched) but WITHOUT the N twiddle multiplications between them (a quarter of the
// instructions of a 136- or 272-point transform).  The price is a permuted input order, which is free wherever a kernel
// scatters its input into shared memory anyway (the product/periodise prologue of the tile kernels).
// The transposed (DIT) flow takes pfa_out positions in and leaves natural index n at pfa_in(n).
// ---------------------------------------------------------------------------------------------------------------
constexpr bool ct_pfa_ok(int n) {
    const Plan1 P = ct_plan1(n);
    return P.npass == 2 && ct_is_pow2(P.radix[0]) && P.radix[0] >= 2 && (P.radix[1] & 1) && P.radix[1] > 1 &&
           P.radix[0] * P.radix[1] == n;
}
constexpr int ct_modinv(int a, int m) {          // a^-1 mod m (a, m coprime, small)
    a %= m;
    for (int x = 1; x < m; ++x) if ((a * x) % m == 1) return x;
    return 0;
}
SB_HD int pfa_in(int n, int N1, int N2, int inv21, int inv12) {     // inv21 = N2^-1 mod N1, inv12 = N1^-1 mod N2
    return ((n * inv21) % N1) * N2 + (n * inv12) % N2;
}
SB_HD int pfa_out(int k, int N1, int N2) { return rt_bitrev(k % N1, ct_log2(N1)) * N2 + k % N2; }

constexpr bool ct_radix_compiled(int r) {
    return r == 2 || r == 3 || r == 4 || r == 5 || r == 7 || r == 8 || r == 11 || r == 13 || r == 16 || r == 17;
}
constexpr bool ct_plan_static_ok(const Plan1& P) {
    if (P.npass <= 0) return false;
    for (int p = 0; p < P.npass; ++p) if (!ct_radix_compiled(P.radix[p])) return false;
    return true;
}

// Runtime odd radix (primes not in the compiled list); O(R^2), local-memory array.
template <bool INV, typename T>
SB_HD void butterfly_generic(cx<T>* line, int estride, int base, int q, int twstep, const cx<T>* tw,
                             bool do_tw, int n, int R) {
    cx<T> v[kMaxGenericRadix];
    const int rstep = n / R;
    for (int k = 0; k < R; ++k) {
        cx<T> t = line[(base + k * q) * estride];
        if (INV && do_tw && k) t = cmulc(t, tw[k * twstep]);
        v[k] = t;
    }
    for (int f = 0; f < R; ++f) {
        cx<T> acc = v[0];
        int idx = 0;
        for (int c = 1; c < R; ++c) {
            idx += f; if (idx >= R) idx -= R;
            cx<T> w = tw[idx * rstep];
            acc = acc + (INV ? cmulc(v[c], w) : cmul(v[c], w));
        }
        if (!INV && do_tw && f) acc = cmul(acc, tw[f * twstep]);
        line[(base + f * q) * estride] = acc;     // safe: inputs are all in v[]
    }
}

// Dispatch one butterfly of plan pass p (work item bf in [0, n/r)) on one line.
template <bool INV, typename T>
SB_HD void butterfly_dispatch(int r, cx<T>* line, int estride, int base, int q, int twstep,
                              const cx<T>* tw, bool do_tw, int n) {
    switch (r) {
        case 2:  butterfly<2, INV, T>(line, estride, base, q, twstep, tw, do_tw); break;
        case 3:  butterfly<3, INV, T>(line, estride, base, q, twstep, tw, do_tw); break;
        case 4:  butterfly<4, INV, T>(line, estride, base, q, twstep, tw, do_tw); break;
        case 5:  butterfly<5, INV, T>(line, estride, base, q, twstep, tw, do_tw); break;
        case 7:  butterfly<7, INV, T>(line, estride, base, q, twstep, tw, do_tw); break;
        case 8:  butterfly<8, INV, T>(line, estride, base, q, twstep, tw, do_tw); break;
        case 11: butterfly<11, INV, T>(line, estride, base, q, twstep, tw, do_tw); break;
        case 13: butterfly<13, INV, T>(line, estride, base, q, twstep, tw, do_tw); break;
        case 16: butterfly<16, INV, T>(line, estride, base, q, twstep, tw, do_tw); break;
        case 17: butterfly<17, INV, T>(line, estride, base, q, twstep, tw, do_tw); break;
        default: butterfly_generic<INV, T>(line, estride, base, q, twstep, tw, do_tw, n, r); break;
    }
}

}  // namespace sb
