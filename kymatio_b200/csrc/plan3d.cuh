// plan3d.cuh - host side of the fused 3-D kernels (kernels3d.cuh): constant tables and one launcher per kernel.
// The cascade is driven by kymatio_b200/engine3d.py through the scat3d_* entry points of include/scat_b200.h.
#pragma once
#include "common.cuh"
#include "kernels3d.cuh"
#include "plan_host.h"

namespace sb {

inline bool fused3d_supported(int M, int N, int O) {
    return M >= 8 && N >= 16 && O >= 16 && O % 16 == 0 && kern3d_col_prod<float>(M) && kern3d_col_fwd<float>(M) &&
           kern3d_plane<float>(N, O) && kern3d_plane_real<float>(N, O);
}

// [twM | twN | twO | twO/2 | posM | posN | posO/2]
struct Tables3d {
    int n[4]; size_t tw[4], pos[3], bytes;
    Tables3d(int M, int N, int O) {
        if (!fused3d_supported(M, N, O))
            throw std::runtime_error("no fused 3-D kernel instance for this volume shape (see scat3d_inst.cu)");
        n[0] = M; n[1] = N; n[2] = O; n[3] = O / 2;
        size_t off = 0;
        for (int a = 0; a < 4; ++a) { tw[a] = off; off = align_up(off + (size_t)n[a] * sizeof(cx<float>), 256); }
        const int pn[3] = {M, N, O / 2};
        for (int a = 0; a < 3; ++a) { pos[a] = off; off = align_up(off + (size_t)pn[a] * sizeof(int), 256); }
        bytes = off;
    }
};
inline void tables3d_init(void* dev, int M, int N, int O, cudaStream_t st) {
    Tables3d t(M, N, O);
    std::vector<unsigned char> h(t.bytes, 0);
    for (int a = 0; a < 4; ++a) {
        auto tw = twiddle_table<float>(t.n[a]);
        memcpy(h.data() + t.tw[a], tw.data(), (size_t)t.n[a] * sizeof(cx<float>));
    }
    const int pn[3] = {M, N, O / 2};
    for (int a = 0; a < 3; ++a) {
        auto pos = scramble_table(ct_plan1(pn[a]));
        memcpy(h.data() + t.pos[a], pos.data(), (size_t)pn[a] * sizeof(int));
    }
    SB_CUDA(cudaMemcpyAsync(dev, h.data(), t.bytes, cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaStreamSynchronize(st));
}
inline void enable3d_once() {
    once_per_device("kern3d", [] { kern3d_enable_smem(); });
}

inline void col_prod3d(const void* tables, const void* U, const void* filt, void* Y, long long B, int nm, int M, int N, int O,
                       cudaStream_t st) {
    if (B <= 0 || nm <= 0) return;
    enable3d_once();
    Tables3d t(M, N, O);
    const unsigned char* cb = static_cast<const unsigned char*>(tables);
    ColProd3<float> a{};
    a.U = static_cast<const cx<float>*>(U); a.filt = static_cast<const cx<float>*>(filt); a.Y = static_cast<cx<float>*>(Y);
    a.B = (int)B; a.nm = nm; a.NO = N * O; a.O = O; a.scale = 1.0f / ((float)M * (float)N * (float)O);
    a.twM = reinterpret_cast<const cx<float>*>(cb + t.tw[0]); a.twO = reinterpret_cast<const cx<float>*>(cb + t.tw[2]);
    const size_t smem = ((size_t)M * k1LP + M) * sizeof(cx<float>);
    const size_t vol = (size_t)M * N * O * sizeof(cx<float>);
    dim3 grid((unsigned)(B * (a.NO / k1L)));
    auto kern = kern3d_col_prod<float>(M);
    launch("3d_col_prod:nm" + std::to_string(nm), (double)B * (nm + 1.0) * vol + (double)nm * vol, st,
           [&] { kern<<<grid, block1d(), smem, st>>>(a); });
}

inline void plane3d(const void* tables, const void* Y, void* spec, void* integ, long long istride, int ioff, const void* powers,
                    int P, long long B, int nm, int M, int N, int O, cudaStream_t st) {
    if (B <= 0 || nm <= 0) return;
    enable3d_once();
    if (P < 1 || P > 8) throw std::runtime_error("plane3d: 1..8 integral powers supported");
    Tables3d t(M, N, O);
    const unsigned char* cb = static_cast<const unsigned char*>(tables);
    Plane3<float> a{};
    a.Y = static_cast<const cx<float>*>(Y); a.spec = static_cast<cx<float>*>(spec);
    a.integ = static_cast<double*>(integ); a.powers = static_cast<const float*>(powers); a.P = P;
    a.istride = istride; a.ioff = ioff; a.nm = nm; a.M = M;
    a.twN = reinterpret_cast<const cx<float>*>(cb + t.tw[1]); a.twH = reinterpret_cast<const cx<float>*>(cb + t.tw[3]);
    const size_t smem = ((size_t)N * (O / 2 + 1) + N + O / 2) * sizeof(cx<float>) + 8 * 32 * sizeof(double);
    const size_t vol = (size_t)M * N * O * sizeof(cx<float>);
    dim3 grid((unsigned)(B * M * 2));
    auto kern = kern3d_plane<float>(N, O);
    launch(std::string(spec ? "3d_plane_parent:nm" : "3d_plane_leaf:nm") + std::to_string(nm),
           (double)B * (nm + (spec ? 1.0 : 0.0)) * vol, st, [&] { kern<<<grid, plane_threads(N * O / 2), smem, st>>>(a); });
}

// parents: radix-2 DIT along O + forward transform along M of the half-plane spectra (in place allowed)
inline void col_fwd3d(const void* tables, const void* Z, void* out, long long B, int M, int N, int O, cudaStream_t st) {
    if (B <= 0) return;
    enable3d_once();
    Tables3d t(M, N, O);
    const unsigned char* cb = static_cast<const unsigned char*>(tables);
    auto kern = kern3d_col_fwd<float>(M);
    ColFwd3<float> a{};
    a.Z = static_cast<const cx<float>*>(Z); a.out = static_cast<cx<float>*>(out); a.B = (int)B; a.NO = N * O; a.O = O;
    a.twM = reinterpret_cast<const cx<float>*>(cb + t.tw[0]); a.twO = reinterpret_cast<const cx<float>*>(cb + t.tw[2]);
    const size_t smem = ((size_t)M * k1LP + M) * sizeof(cx<float>);
    const size_t vol = (size_t)M * N * O * sizeof(cx<float>);
    dim3 grid((unsigned)(B * (a.NO / k1L)));
    launch("3d_col_fwd", (double)B * 2.0 * vol, st, [&] { kern<<<grid, block1d(), smem, st>>>(a); });
}

// U0_hat = rfft(x) for real volumes x (B, M, N, O): half-plane forward transforms, then radix-2 along O + transform along M
inline void rfft3d(const void* tables, const void* x, void* out, long long B, int M, int N, int O, cudaStream_t st) {
    if (B <= 0) return;
    enable3d_once();
    Tables3d t(M, N, O);
    const unsigned char* cb = static_cast<const unsigned char*>(tables);
    auto kp = kern3d_plane_real<float>(N, O);
    auto kc = kern3d_col_fwd<float>(M);
    PlaneReal3<float> a{};
    a.x = static_cast<const float*>(x); a.spec = static_cast<cx<float>*>(out); a.M = M;
    a.twN = reinterpret_cast<const cx<float>*>(cb + t.tw[1]); a.twH = reinterpret_cast<const cx<float>*>(cb + t.tw[3]);
    a.posN = reinterpret_cast<const int*>(cb + t.pos[1]); a.posH = reinterpret_cast<const int*>(cb + t.pos[2]);
    const size_t smem_p = ((size_t)N * (O / 2 + 1) + N + O / 2) * sizeof(cx<float>);
    const size_t vol = (size_t)M * N * O * sizeof(cx<float>);
    dim3 grid_p((unsigned)(B * M * 2));
    launch("3d_plane_real", (double)B * 1.5 * vol, st, [&] { kp<<<grid_p, plane_threads(N * O / 2), smem_p, st>>>(a); });
    ColFwd3<float> c{};
    c.Z = static_cast<const cx<float>*>(out); c.out = static_cast<cx<float>*>(out); c.B = (int)B; c.NO = N * O; c.O = O;
    c.twM = reinterpret_cast<const cx<float>*>(cb + t.tw[0]); c.twO = reinterpret_cast<const cx<float>*>(cb + t.tw[2]);
    c.posM = reinterpret_cast<const int*>(cb + t.pos[0]);
    const size_t smem_c = ((size_t)M * k1LP + M) * sizeof(cx<float>);
    dim3 grid_c((unsigned)(B * (c.NO / k1L)));
    launch("3d_col_fwd0", (double)B * 2.0 * vol, st, [&] { kc<<<grid_c, block1d(), smem_c, st>>>(c); });
}

}  // namespace sb
