// plan1d.cuh - host side of the fused 1-D kernels: the N = NA*NB split, the per-length constant tables
// and one launcher per kernel of kernels1d.cuh.  The cascade itself (which paths, which filters, which
// channel) is driven by kymatio_b200/engine1d.py through the scat1d_* entry points of include/scat_b200.h.
#pragma once
#include "common.cuh"
#include "kernels1d.cuh"
#include "plan_host.h"

namespace sb {

struct Split1d { int n, Na, Nb, lb, nhi, nlo; };
inline Split1d split1d(int N) {
    if (N < 16 || N > (1 << 18) || (N & (N - 1))) throw std::runtime_error("fused 1-D transform length must be a power of two in [16, 2^18], got " + std::to_string(N));
    Split1d s{};
    s.n = ct_log2(N);
    const int lgb = std::max(4, (s.n + 1) / 2);
    s.Nb = 1 << lgb; s.Na = N / s.Nb;
    s.lb = (s.n + 1) / 2; s.nlo = 1 << s.lb; s.nhi = N >> s.lb;
    return s;
}

// [twA | twB | hi | lo | invA | posA | posB]
struct Tables1d {
    Split1d sp; size_t twa, twb, hi, lo, inva, posa, posb, bytes;
    explicit Tables1d(int N) : sp(split1d(N)) {
        size_t off = 0;
        auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
        twa = take((size_t)sp.Na * sizeof(cx<float>)); twb = take((size_t)sp.Nb * sizeof(cx<float>));
        hi = take((size_t)sp.nhi * sizeof(cx<float>)); lo = take((size_t)sp.nlo * sizeof(cx<float>));
        inva = take((size_t)sp.Na * sizeof(int));
        posa = take((size_t)sp.Na * sizeof(int)); posb = take((size_t)sp.Nb * sizeof(int));
        bytes = off;
    }
};
inline void tables1d_init(void* dev, int N, cudaStream_t st) {
    Tables1d t(N);
    std::vector<unsigned char> h(t.bytes, 0);
    auto twa = twiddle_table<float>(t.sp.Na); auto twb = twiddle_table<float>(t.sp.Nb);
    memcpy(h.data() + t.twa, twa.data(), (size_t)t.sp.Na * sizeof(cx<float>));
    memcpy(h.data() + t.twb, twb.data(), (size_t)t.sp.Nb * sizeof(cx<float>));
    cx<float>* hi = reinterpret_cast<cx<float>*>(h.data() + t.hi);
    cx<float>* lo = reinterpret_cast<cx<float>*>(h.data() + t.lo);
    const long double tau = -2.0L * 3.14159265358979323846264338327950288L / (long double)N;
    for (int a = 0; a < t.sp.nhi; ++a) { long double x = tau * (long double)((long long)a << t.sp.lb); hi[a].x = (float)cosl(x); hi[a].y = (float)sinl(x); }
    for (int b = 0; b < t.sp.nlo; ++b) { long double x = tau * (long double)b; lo[b].x = (float)cosl(x); lo[b].y = (float)sinl(x); }
    int* inva = reinterpret_cast<int*>(h.data() + t.inva);
    int* posa = reinterpret_cast<int*>(h.data() + t.posa);
    int* posb = reinterpret_cast<int*>(h.data() + t.posb);
    if (t.sp.Na > 1) {
        auto pos = scramble_table(ct_plan1(t.sp.Na));
        for (int f = 0; f < t.sp.Na; ++f) { inva[pos[f]] = f; posa[f] = pos[f]; }
    } else { inva[0] = 0; posa[0] = 0; }
    // k1d_row_mod<LEAF> relies on: the k1L consecutive scrambled rows of one CTA hold t1 = base + (Na/k1L)*k, k < k1L
    for (int p0 = 0; p0 + k1L <= t.sp.Na; p0 += k1L) {
        int base = inva[p0];
        for (int l = 1; l < k1L; ++l) base = std::min(base, inva[p0 + l]);
        unsigned seen = 0;
        for (int l = 0; l < k1L; ++l) {
            const int d = inva[p0 + l] - base, st = t.sp.Na / k1L;
            if (d % st || d / st >= k1L) throw std::runtime_error("row blocks of the column transform are not arithmetic progressions");
            seen |= 1u << (d / st);
        }
        if (seen != (1u << k1L) - 1u) throw std::runtime_error("row blocks of the column transform are not arithmetic progressions");
    }
    {
        auto pos = scramble_table(ct_plan1(t.sp.Nb));
        for (int f = 0; f < t.sp.Nb; ++f) posb[f] = pos[f];
    }
    SB_CUDA(cudaMemcpyAsync(dev, h.data(), t.bytes, cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaStreamSynchronize(st));
}

// [twM | posM]
struct FinTables1d {
    int M; size_t tw, pos, bytes;
    explicit FinTables1d(int M_) : M(M_) {
        if (M < 8 || M > 1024 || (M & (M - 1))) throw std::runtime_error("fused 1-D low-pass length must be a power of two in [8, 1024], got " + std::to_string(M));
        size_t off = 0;
        auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
        tw = take((size_t)M * sizeof(cx<float>)); pos = take((size_t)M * sizeof(int));
        bytes = off;
    }
};
inline void fin_tables1d_init(void* dev, int M, cudaStream_t st) {
    FinTables1d t(M);
    std::vector<unsigned char> h(t.bytes, 0);
    auto tw = twiddle_table<float>(M);
    auto pos = scramble_table(ct_plan1(M));
    memcpy(h.data() + t.tw, tw.data(), (size_t)M * sizeof(cx<float>));
    memcpy(h.data() + t.pos, pos.data(), (size_t)M * sizeof(int));
    SB_CUDA(cudaMemcpyAsync(dev, h.data(), t.bytes, cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaStreamSynchronize(st));
}

inline void enable1d_once() {
    once_per_device("kern1d", [] { kern1d_enable_smem(); });
}
inline TwN<float> twn_of(const Tables1d& t, const unsigned char* cb) {
    TwN<float> w;
    w.hi = reinterpret_cast<const cx<float>*>(cb + t.hi); w.lo = reinterpret_cast<const cx<float>*>(cb + t.lo);
    w.lb = t.sp.lb; w.nhi = t.sp.nhi;
    return w;
}
inline dim3 block1d() { return dim3(k1L, k1Threads / k1L, 1); }

inline void col_prod1d(const void* tables, const void* parent, long long ps_b, long long ps_i, const void* filt_dev,
                       const void* supp_dev, void* Y, long long G, int NI, int Npar, int N, double algo_bytes,
                       cudaStream_t st) {
    if (G <= 0) return;
    enable1d_once();
    Tables1d t(N);
    if (Npar % N || NI < 1) throw std::runtime_error("col_prod1d: bad sizes");
    const unsigned char* cb = static_cast<const unsigned char*>(tables);
    auto k = kern1d_cols<float>(t.sp.Na);
    if (!k.col_prod) throw std::runtime_error("col_prod1d: no instance for NA=" + std::to_string(t.sp.Na));
    ColProd1<float> a{};
    a.parent = static_cast<const cx<float>*>(parent); a.ps_b = ps_b; a.ps_i = ps_i;
    a.filt = static_cast<const float* const*>(filt_dev); a.supp = static_cast<const int2*>(supp_dev);
    a.Y = static_cast<cx<float>*>(Y);
    a.NI = NI; a.Npar = Npar; a.k = Npar / N; a.NB = t.sp.Nb;
    a.scale = 1.0f / ((float)N * (float)a.k);
    a.twA = reinterpret_cast<const cx<float>*>(cb + t.twa); a.invA = reinterpret_cast<const int*>(cb + t.inva);
    a.w = twn_of(t, cb);
    const size_t smem = ((size_t)t.sp.Na * k1LP + t.sp.Na) * sizeof(cx<float>);
    dim3 grid((unsigned)(G * (t.sp.Nb / k1L)));
    launch("1d_col_prod:N" + std::to_string(N) + ":k" + std::to_string(a.k), algo_bytes, st,
           [&] { k.col_prod<<<grid, block1d(), smem, st>>>(a); });
}

inline void row_mod1d(const void* tables, void* Y, long long G, int N, void* part, int Fc, double algo_bytes, cudaStream_t st,
                      void* mod = nullptr, bool skip_fwd = false) {
    if (G <= 0) return;
    enable1d_once();
    Tables1d t(N);
    const unsigned char* cb = static_cast<const unsigned char*>(tables);
    auto k = kern1d_rows<float>(t.sp.Nb);
    if (!k.parent) throw std::runtime_error("row_mod1d: no instance for NB=" + std::to_string(t.sp.Nb));
    if (part && (Fc < 1 || Fc > N / 2 + 1)) throw std::runtime_error("row_mod1d: Fc out of range");
    RowMod1<float> a{};
    a.Y = static_cast<cx<float>*>(Y); a.NA = t.sp.Na; a.N = N;
    a.twB = reinterpret_cast<const cx<float>*>(cb + t.twb); a.invA = reinterpret_cast<const int*>(cb + t.inva);
    a.w = twn_of(t, cb);
    a.part = static_cast<cx<float>*>(part); a.Fc = Fc;
    if (mod && part) throw std::runtime_error("row_mod1d: the unaveraged output has no low-pass partial sums");
    a.mod = static_cast<float*>(mod); a.skip_fwd = skip_fwd ? 1 : 0;
    a.posA = reinterpret_cast<const int*>(cb + t.posa); a.posB = reinterpret_cast<const int*>(cb + t.posb);
    const size_t smem = ((size_t)(t.sp.Nb + 1) * k1L + t.sp.Nb) * sizeof(cx<float>) + 2 * k1L * sizeof(int);
    dim3 grid((unsigned)G, ceil_div(t.sp.Na, k1L));
    launch(std::string(part ? "1d_row_mod_leaf:N" : mod ? "1d_row_mod_t0:N" : "1d_row_mod:N") + std::to_string(N), algo_bytes, st,
           [&] { (part ? k.leaf : k.parent)<<<grid, block1d(), smem, st>>>(a); });
}

inline void col_fwd1d(const void* tables, const void* Z, void* out, long long G, int N, double algo_bytes, cudaStream_t st) {
    if (G <= 0) return;
    enable1d_once();
    Tables1d t(N);
    const unsigned char* cb = static_cast<const unsigned char*>(tables);
    auto k = kern1d_cols<float>(t.sp.Na);
    if (!k.col_fwd) throw std::runtime_error("col_fwd1d: no instance for NA=" + std::to_string(t.sp.Na));
    ColFwd1<float> a{};
    a.Z = static_cast<const cx<float>*>(Z); a.out = static_cast<cx<float>*>(out); a.NB = t.sp.Nb;
    a.twA = reinterpret_cast<const cx<float>*>(cb + t.twa);
    const size_t smem = ((size_t)t.sp.Na * k1LP + t.sp.Na) * sizeof(cx<float>);
    dim3 grid((unsigned)(G * (t.sp.Nb / k1L)));
    launch("1d_col_fwd:N" + std::to_string(N), algo_bytes, st, [&] { k.col_fwd<<<grid, block1d(), smem, st>>>(a); });
}

// U_hat = fft(x) for real x (G, N): k1d_row_real then k1d_col_fwd with scrambled row staging; Z is a (G, N) complex scratch
// (may alias out)
inline void rfft1d(const void* tables, const void* x, void* Z, void* out, long long G, int N, cudaStream_t st) {
    if (G <= 0) return;
    enable1d_once();
    Tables1d t(N);
    const unsigned char* cb = static_cast<const unsigned char*>(tables);
    auto kr = kern1d_rows<float>(t.sp.Nb);
    auto kc = kern1d_cols<float>(t.sp.Na);
    if (!kr.real || !kc.col_fwd) throw std::runtime_error("rfft1d: no instance for N=" + std::to_string(N));
    RowReal1<float> a{};
    a.x = static_cast<const float*>(x); a.Z = static_cast<cx<float>*>(Z); a.NA = t.sp.Na;
    a.twB = reinterpret_cast<const cx<float>*>(cb + t.twb); a.posB = reinterpret_cast<const int*>(cb + t.posb);
    a.w = twn_of(t, cb);
    const size_t smem_r = ((size_t)(t.sp.Nb + 1) * k1L + t.sp.Nb) * sizeof(cx<float>);
    dim3 grid_r((unsigned)G, ceil_div(t.sp.Na, k1L));
    launch("1d_row_real:N" + std::to_string(N), (double)G * N * 12.0, st, [&] { kr.real<<<grid_r, block1d(), smem_r, st>>>(a); });
    ColFwd1<float> c{};
    c.Z = static_cast<const cx<float>*>(Z); c.out = static_cast<cx<float>*>(out); c.NB = t.sp.Nb;
    c.twA = reinterpret_cast<const cx<float>*>(cb + t.twa); c.posA = reinterpret_cast<const int*>(cb + t.posa);
    const size_t smem_c = ((size_t)t.sp.Na * k1LP + t.sp.Na) * sizeof(cx<float>);
    dim3 grid_c((unsigned)(G * (t.sp.Nb / k1L)));
    launch("1d_col_fwd0:N" + std::to_string(N), (double)G * N * 16.0, st, [&] { kc.col_fwd<<<grid_c, block1d(), smem_c, st>>>(c); });
}

inline void tile1d(const void* tables, const void* parent, long long ps_b, long long ps_i, const void* filt_dev,
                   const void* supp_dev, void* spec, void* part, int Fc, long long G, int NI, int Npar, int N,
                   double algo_bytes, cudaStream_t st, void* mod = nullptr) {
    if (G <= 0) return;
    enable1d_once();
    Tables1d t(N);
    if (Npar % N || NI < 1 || N > k1TileMaxN) throw std::runtime_error("tile1d: bad sizes");
    if (part && (Fc < 1 || Fc > N / 2 + 1)) throw std::runtime_error("tile1d: Fc out of range");
    const unsigned char* cb = static_cast<const unsigned char*>(tables);
    auto kern = kern1d_tile<float>(t.sp.Na, t.sp.Nb);
    if (!kern) throw std::runtime_error("tile1d: no instance for N=" + std::to_string(N));
    Tile1<float> a{};
    a.parent = static_cast<const cx<float>*>(parent); a.ps_b = ps_b; a.ps_i = ps_i;
    a.filt = static_cast<const float* const*>(filt_dev); a.supp = static_cast<const int2*>(supp_dev);
    a.spec = static_cast<cx<float>*>(spec); a.part = static_cast<cx<float>*>(part); a.Fc = Fc;
    a.NI = NI; a.Npar = Npar; a.k = Npar / N;
    a.scale = 1.0f / ((float)N * (float)a.k);
    a.twA = reinterpret_cast<const cx<float>*>(cb + t.twa); a.twB = reinterpret_cast<const cx<float>*>(cb + t.twb);
    a.invA = reinterpret_cast<const int*>(cb + t.inva);
    a.w = twn_of(t, cb);
    a.mod = static_cast<float*>(mod);
    a.posA = reinterpret_cast<const int*>(cb + t.posa); a.posB = reinterpret_cast<const int*>(cb + t.posb);
    const size_t smem = ((size_t)t.sp.Na * (t.sp.Nb + 1) + t.sp.Na + t.sp.Nb) * sizeof(cx<float>);
    dim3 grid((unsigned)G);
    launch(std::string(mod ? "1d_tile_t0:N" : spec ? "1d_tile_parent:N" : "1d_tile_leaf:N") + std::to_string(N) + ":k" + std::to_string(a.k), algo_bytes,
           st, [&] { kern<<<grid, block1d(), smem, st>>>(a); });
}

inline void finish1d(const void* fin_tables, const void* base0, const void* base1, const void* base2, const void* segs_dev,
                     int nseg, long long total_lines, int M, void* out, long long os_b, int i0, int W, double algo_bytes,
                     cudaStream_t st) {
    if (total_lines <= 0) return;
    enable1d_once();
    FinTables1d t(M);
    if (i0 < 0 || i0 + W > M || nseg < 1) throw std::runtime_error("finish1d: bad sizes");
    const unsigned char* cb = static_cast<const unsigned char*>(fin_tables);
    auto kern = kern1d_finish<float>(M);
    if (!kern) throw std::runtime_error("finish1d: no instance for M=" + std::to_string(M));
    Finish1<float> a{};
    a.base[0] = static_cast<const cx<float>*>(base0); a.base[1] = static_cast<const cx<float>*>(base1);
    a.base[2] = static_cast<const cx<float>*>(base2);
    a.segs = static_cast<const FinSeg<float>*>(segs_dev); a.nseg = nseg; a.total = (int)total_lines;
    a.out = static_cast<float*>(out); a.os_b = os_b; a.i0 = i0; a.W = W;
    a.twM = reinterpret_cast<const cx<float>*>(cb + t.tw); a.posM = reinterpret_cast<const int*>(cb + t.pos);
    const size_t smem = ((size_t)M * k1FLP + M) * sizeof(cx<float>) + (size_t)(M + k1FL) * sizeof(int);
    dim3 grid((unsigned)((total_lines + k1FL - 1) / k1FL));
    launch("1d_finish:M" + std::to_string(M), algo_bytes, st, [&] { kern<<<grid, block1d(), smem, st>>>(a); });
}

// bin 0 of every path's spectrum -> out[b*os_b + chan] (average='global')
inline void finish1d_global(const void* base0, const void* base1, const void* base2, const void* segs_dev, int nseg,
                            long long total_lines, void* out, long long os_b, cudaStream_t st) {
    if (total_lines <= 0) return;
    if (nseg < 1) throw std::runtime_error("finish1d_global: bad sizes");
    Finish1<float> a{};
    a.base[0] = static_cast<const cx<float>*>(base0); a.base[1] = static_cast<const cx<float>*>(base1);
    a.base[2] = static_cast<const cx<float>*>(base2);
    a.segs = static_cast<const FinSeg<float>*>(segs_dev); a.nseg = nseg; a.total = (int)total_lines;
    a.out = static_cast<float*>(out); a.os_b = os_b;
    const unsigned grid = (unsigned)((total_lines + 127) / 128);
    launch("1d_finish_global", (double)total_lines * 12.0, st, [&] { k1d_finish_global<float><<<grid, 128, 0, st>>>(a); });
}

}  // namespace sb
