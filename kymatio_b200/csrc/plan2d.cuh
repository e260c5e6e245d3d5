// plan2d.cuh - host orchestration of the fused 2-D scattering forward.
//
// Mirrors the loop of kymatio/scattering2d/core/scattering2d.py:14-86, restructured into
// grouped launches: one (row-pass, column-pass, row-pass, low-pass) quartet per first-order
// scale j1 and per (j1, j2) second-order pair, each launch covering batch x angles.
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>
#include "../../include/scat_b200.h"
#include "common.cuh"
#include "kernels2d.cuh"
#include "plan_host.h"

struct scat_plan2d {
    virtual ~scat_plan2d() {}
    virtual void info(int32_t* Mp, int32_t* Np, int32_t* oh, int32_t* ow, int32_t* K) const = 0;
    virtual size_t const_bytes() const = 0;
    virtual void bind(void* const_dev, const void* const* phi, int n_phi, const void* const* psi, int n_psi,
                      cudaStream_t st) = 0;
    virtual size_t workspace_bytes(int64_t batch) const = 0;
    virtual void forward(const void* x, void* out, void* ws, size_t ws_bytes, int64_t batch, cudaStream_t st) = 0;
};

namespace sb {

struct Axis {
    int n = 0;
    Plan1 plan{};
    size_t tw_off = 0, pos_off = 0;   // byte offsets into the constant buffer
};
struct Level2D { Axis a0, a1; };

template <typename T> class Plan2D final : public scat_plan2d {
public:
    explicit Plan2D(const scat_plan2d_desc& d) : d_(d) {
        if (d.J < 0 || d.L < 1) throw std::runtime_error("invalid J or L");
        if (d.max_order != 1 && d.max_order != 2) throw std::runtime_error("max_order must be 1 or 2");
        const int s = 1 << d.J;
        if (d.pre_pad) {
            P0_ = d.M; P1_ = d.N;
            if (P0_ % s || P1_ % s) throw std::runtime_error("pre-padded size must be a multiple of 2^J");
        } else {
            if (s > d.M || s > d.N) throw std::runtime_error("The smallest dimension should be larger than 2^J.");
            P0_ = ((d.M + s) / s + 1) * s;   // kymatio/scattering2d/utils.py:21-22
            P1_ = ((d.N + s) / s + 1) * s;
            top_ = (P0_ - d.M) / 2;          // scattering2d/frontend/base_frontend.py:27-28
            left_ = (P1_ - d.N) / 2;
        }
        m0_ = P0_ >> d.J; m1_ = P1_ >> d.J;
        o0_ = m0_ - 2; o1_ = m1_ - 2;
        if (o0_ < 1 || o1_ < 1) throw std::runtime_error("padded size too small for unpad");
        K_ = 1 + d.L * d.J + (d.max_order == 2 ? d.L * d.L * d.J * (d.J - 1) / 2 : 0);

        // levels 0..J, tables
        size_t off = 0;
        lev_.resize(d.J + 1);
        for (int j = 0; j <= d.J; ++j) {
            Axis* ax[2] = {&lev_[j].a0, &lev_[j].a1};
            const int nn[2] = {P0_ >> j, P1_ >> j};
            for (int a = 0; a < 2; ++a) {
                ax[a]->n = nn[a];
                ax[a]->plan = make_plan1(nn[a]);
                ax[a]->tw_off = off; off = align_up(off + (size_t)nn[a] * sizeof(cx<T>), 256);
                ax[a]->pos_off = off; off = align_up(off + (size_t)nn[a] * sizeof(int), 256);
            }
        }
        tables_bytes_ = off;
        host_tables_.assign(tables_bytes_, 0);
        for (int j = 0; j <= d.J; ++j) {
            const Axis* ax[2] = {&lev_[j].a0, &lev_[j].a1};
            for (int a = 0; a < 2; ++a) {
                auto tw = twiddle_table<T>(ax[a]->n);
                auto pos = scramble_table(ax[a]->plan);
                std::memcpy(host_tables_.data() + ax[a]->tw_off, tw.data(), (size_t)ax[a]->n * sizeof(cx<T>));
                std::memcpy(host_tables_.data() + ax[a]->pos_off, pos.data(), (size_t)ax[a]->n * sizeof(int));
            }
        }
        // scrambled filter bank
        phi_off_.resize(d.J);
        for (int j = 0; j < d.J; ++j) { phi_off_[j] = off; off = align_up(off + fsize(j) * sizeof(T), 256); }
        psi_off_.resize(d.J);
        n_psi_expected_ = 0;
        for (int j = 0; j < d.J; ++j) {
            const int nres = std::min(j + 1, std::max(d.J - 1, 1));   // filter_bank.py:40
            psi_off_[j].resize(nres);
            for (int r = 0; r < nres; ++r) {
                psi_off_[j][r] = off;
                off = align_up(off + (size_t)d.L * fsize(r) * sizeof(T), 256);
            }
            n_psi_expected_ += d.L * nres;
        }
        const_bytes_ = off;

        // slab configurations per level
        row_cfg_.resize(d.J + 1); col_cfg_.resize(d.J + 1);
        for (int j = 0; j <= d.J; ++j) {
            row_cfg_[j] = slab_cfg(lev_[j].a1.plan, lev_[j].a0.n, sizeof(cx<T>));
            col_cfg_[j] = slab_cfg(lev_[j].a0.plan, lev_[j].a1.n, sizeof(cx<T>));
        }
        lowW_ = m1_ | 1;
        low_smem_ = ((size_t)m0_ * lowW_ + m0_ + m1_) * sizeof(cx<T>);
        low_tile_ = low_smem_ <= 96 * 1024;

        // workspace per image, in cx<T> elements
        const size_t lvl0 = (size_t)P0_ * P1_;
        ws_u0_ = lvl0;
        ws_u1_ = d.J >= 1 ? (size_t)d.L * lvl0 : 0;
        ws_u2_ = (d.max_order == 2 && d.J >= 2) ? (size_t)d.L * d.L * lev_[1].a0.n * lev_[1].a1.n : 0;
        ws_low_ = 0;
        if (!low_tile_) ws_low_ = (size_t)std::max(1, (d.max_order == 2 && d.J >= 2) ? d.L * d.L : d.L) * m0_ * m1_;
        per_img_ = (ws_u0_ + ws_u1_ + ws_u2_ + ws_low_) * sizeof(cx<T>);

        enable_big_smem(k2d_pad_rowfft<T>);
        enable_big_smem(k2d_colpass<T, COL_FWD>);
        enable_big_smem(k2d_colpass<T, COL_INV>);
        enable_big_smem(k2d_colpass<T, COL_INV_MOD_FWD>);
        enable_big_smem(k2d_rowpass_prod<T>);
        enable_big_smem(k2d_rowpass<T, false>);
        enable_big_smem(k2d_rowpass<T, true>);
        enable_big_smem(k2d_lowpass<T>);
    }

    void info(int32_t* Mp, int32_t* Np, int32_t* oh, int32_t* ow, int32_t* K) const override {
        if (Mp) *Mp = P0_; if (Np) *Np = P1_; if (oh) *oh = o0_; if (ow) *ow = o1_; if (K) *K = K_;
    }
    size_t const_bytes() const override { return const_bytes_; }

    void bind(void* const_dev, const void* const* phi, int n_phi, const void* const* psi, int n_psi,
              cudaStream_t st) override {
        if (!const_dev) throw std::runtime_error("const buffer is null");
        if (n_phi != d_.J) throw std::runtime_error("expected J low-pass levels");
        if (n_psi != n_psi_expected_) throw std::runtime_error("unexpected number of band-pass levels");
        cbuf_ = static_cast<unsigned char*>(const_dev);
        SB_CUDA(cudaMemcpyAsync(cbuf_, host_tables_.data(), tables_bytes_, cudaMemcpyHostToDevice, st));
        auto scramble = [&](const void* src, size_t dst_off, int res) {
            const int n0 = lev_[res].a0.n, n1 = lev_[res].a1.n;
            dim3 grid(ceil_div(n1, 128), n0);
            launch("scramble", 2.0 * n0 * n1 * sizeof(T), st, [&] {
                k2d_scramble<T><<<grid, 128, 0, st>>>(static_cast<const T*>(src), reinterpret_cast<T*>(cbuf_ + dst_off),
                                                       pos(lev_[res].a0), pos(lev_[res].a1), n0, n1);
            });
        };
        for (int j = 0; j < d_.J; ++j) scramble(phi[j], phi_off_[j], j);
        int n = 0;
        for (int j = 0; j < d_.J; ++j)
            for (int th = 0; th < d_.L; ++th)
                for (size_t r = 0; r < psi_off_[j].size(); ++r)
                    scramble(psi[n++], psi_off_[j][r] + (size_t)th * fsize((int)r) * sizeof(T), (int)r);
        bound_ = true;
    }

    size_t workspace_bytes(int64_t batch) const override {
        const int64_t cap = std::max<int64_t>(1, (int64_t)(kWsCap / per_img_));
        return per_img_ * (size_t)std::max<int64_t>(1, std::min<int64_t>(batch, cap));
    }

    void forward(const void* x, void* out, void* ws, size_t ws_bytes, int64_t batch, cudaStream_t st) override {
        if (!bound_) throw std::runtime_error("filters are not bound to the plan");
        if (batch <= 0) return;
        const int64_t chunk = std::min<int64_t>(batch, (int64_t)(ws_bytes / per_img_));
        if (chunk < 1) throw std::runtime_error("workspace too small");
        const size_t in_img = d_.pre_pad ? (size_t)P0_ * P1_ : (size_t)d_.M * d_.N;
        const size_t out_img = (size_t)K_ * o0_ * o1_;
        for (int64_t b0 = 0; b0 < batch; b0 += chunk) {
            const int B = (int)std::min<int64_t>(chunk, batch - b0);
            forward_chunk(static_cast<const T*>(x) + b0 * in_img, static_cast<T*>(out) + b0 * out_img,
                          static_cast<cx<T>*>(ws), B, st);
        }
    }

private:
    static constexpr size_t kWsCap = (size_t)12 << 30;

    size_t fsize(int res) const { return (size_t)lev_[res].a0.n * lev_[res].a1.n; }
    const cx<T>* tw(const Axis& a) const { return reinterpret_cast<const cx<T>*>(cbuf_ + a.tw_off); }
    const int* pos(const Axis& a) const { return reinterpret_cast<const int*>(cbuf_ + a.pos_off); }
    const T* phi(int j) const { return reinterpret_cast<const T*>(cbuf_ + phi_off_[j]); }
    const T* psi(int j, int res) const { return reinterpret_cast<const T*>(cbuf_ + psi_off_[j][res]); }

    // out[G][n0][n1] = rows-inverse-DIT( periodise_k( parent * filt ) ), G = Bp * NF
    void row_prod(const cx<T>* parent, const T* filt, cx<T>* out, int parent_res, int out_res, int Bp, int NF,
                  T scale, cudaStream_t st) {
        RowProdArgs<T> a{};
        a.parent = parent; a.filt = filt; a.out = out;
        a.P0 = lev_[parent_res].a0.n; a.P1 = lev_[parent_res].a1.n;
        a.k = 1 << (out_res - parent_res);
        a.n0 = lev_[out_res].a0.n; a.n1 = lev_[out_res].a1.n;
        a.NP = 1; a.NF = NF; a.scale = scale;
        const SlabCfg& c = row_cfg_[out_res];
        a.lines = c.lines; a.LP = c.LP; a.plan = lev_[out_res].a1.plan; a.tw = tw(lev_[out_res].a1);
        dim3 grid((unsigned)(Bp * NF), ceil_div(a.n0, c.lines));
        const double G = (double)Bp * NF;
        const double bytes = G * a.P0 * a.P1 * sizeof(cx<T>) + (double)NF * a.P0 * a.P1 * sizeof(T) +
                             G * a.n0 * a.n1 * sizeof(cx<T>);
        launch("rowpass_prod:L" + std::to_string(parent_res) + ">L" + std::to_string(out_res), bytes, st,
               [&] { k2d_rowpass_prod<T><<<grid, c.block, c.smem, st>>>(a); });
    }
    template <int MODE> void col_pass(cx<T>* data, int res, int G, cudaStream_t st) {
        ColArgs<T> a{};
        a.in = data; a.out = data; a.n0 = lev_[res].a0.n; a.n1 = lev_[res].a1.n;
        const SlabCfg& c = col_cfg_[res];
        a.lines = c.lines; a.LP = c.LP; a.plan = lev_[res].a0.plan; a.tw = tw(lev_[res].a0);
        dim3 grid((unsigned)G, ceil_div(a.n1, c.lines));
        launch(std::string(MODE == COL_FWD ? "colpass_fwd" : MODE == COL_INV ? "colpass_inv" : "colpass_inv_mod_fwd") +
                   ":L" + std::to_string(res) + ":G" + std::to_string(G / std::max(1, last_B_)),
               2.0 * G * a.n0 * a.n1 * sizeof(cx<T>), st,
               [&] { k2d_colpass<T, MODE><<<grid, c.block, c.smem, st>>>(a); });
    }
    template <bool INV> void row_pass(cx<T>* data, int res, int G, cudaStream_t st) {
        RowArgs<T> a{};
        a.in = data; a.out = data; a.n0 = lev_[res].a0.n; a.n1 = lev_[res].a1.n;
        const SlabCfg& c = row_cfg_[res];
        a.lines = c.lines; a.LP = c.LP; a.plan = lev_[res].a1.plan; a.tw = tw(lev_[res].a1);
        dim3 grid((unsigned)G, ceil_div(a.n0, c.lines));
        launch(std::string(INV ? "rowpass_inv" : "rowpass_fwd") + ":L" + std::to_string(res) + ":G" +
                   std::to_string(G / std::max(1, last_B_)),
               2.0 * G * a.n0 * a.n1 * sizeof(cx<T>), st,
               [&] { k2d_rowpass<T, INV><<<grid, c.block, c.smem, st>>>(a); });
    }
    // S[b][ch] = unpad(Re ifft2(periodise(spec * phi[res]))) for G = B*PP spectra at resolution res
    void low_pass(const cx<T>* spec, int res, T* out, int B, int PP, int NF, int ch0, int chs, cx<T>* tmp,
                  cudaStream_t st) {
        const int k = 1 << (d_.J - res);
        const T scale = T(1) / (T(k) * T(k) * T(m0_) * T(m1_));
        if (low_tile_) {
            LowArgs<T> a{};
            a.in = spec; a.filt = phi(res); a.out = out;
            a.P0 = lev_[res].a0.n; a.P1 = lev_[res].a1.n; a.k = k; a.m0 = m0_; a.m1 = m1_; a.W = lowW_;
            a.PP = PP; a.NF = NF; a.ch0 = ch0; a.chs = chs; a.K = K_; a.scale = scale;
            a.plan0 = lev_[d_.J].a0.plan; a.plan1 = lev_[d_.J].a1.plan;
            a.tw0 = tw(lev_[d_.J].a0); a.tw1 = tw(lev_[d_.J].a1);
            const double G = (double)B * PP;
            launch("lowpass:L" + std::to_string(res) + ":G" + std::to_string(PP),
                   G * a.P0 * a.P1 * sizeof(cx<T>) + (double)a.P0 * a.P1 * sizeof(T) + G * o0_ * o1_ * sizeof(T), st,
                   [&] { k2d_lowpass<T><<<(unsigned)(B * PP), dim3(32, 8), low_smem_, st>>>(a); });
        } else {
            // streaming fallback for outputs too large for one CTA
            const int G = B * PP;
            RowProdArgs<T> a{};
            a.parent = spec; a.filt = phi(res); a.out = tmp;
            a.P0 = lev_[res].a0.n; a.P1 = lev_[res].a1.n; a.k = k; a.n0 = m0_; a.n1 = m1_;
            a.NP = 1; a.NF = 1; a.scale = scale;
            const SlabCfg& c = row_cfg_[d_.J];
            a.lines = c.lines; a.LP = c.LP; a.plan = lev_[d_.J].a1.plan; a.tw = tw(lev_[d_.J].a1);
            dim3 grid((unsigned)G, ceil_div(m0_, c.lines));
            launch("rowpass_prod(low):L" + std::to_string(res),
                   (double)G * (a.P0 * a.P1 + m0_ * m1_) * sizeof(cx<T>), st,
                   [&] { k2d_rowpass_prod<T><<<grid, c.block, c.smem, st>>>(a); });
            col_pass<COL_INV>(tmp, d_.J, G, st);
            CropArgs<T> ca{};
            ca.in = tmp; ca.out = out; ca.m0 = m0_; ca.m1 = m1_; ca.PP = PP; ca.NF = NF; ca.ch0 = ch0; ca.chs = chs;
            ca.K = K_;
            dim3 g2((unsigned)G, ceil_div(o0_ * o1_, 256));
            launch("crop_real", (double)G * (m0_ * m1_ * sizeof(cx<T>) + o0_ * o1_ * sizeof(T)), st,
                   [&] { k2d_crop_real<T><<<g2, 256, 0, st>>>(ca); });
        }
    }

    void forward_chunk(const T* x, T* out, cx<T>* ws, int B, cudaStream_t st) {
        const int J = d_.J, L = d_.L;
        last_B_ = B;
        cx<T>* U0 = ws;
        cx<T>* U1 = U0 + (size_t)B * ws_u0_;
        cx<T>* U2 = U1 + (size_t)B * ws_u1_;
        cx<T>* TL = U2 + (size_t)B * ws_u2_;
        // U0 = fft2(pad(x))   (core/scattering2d.py:14-16)
        {
            PadRowArgs<T> a{};
            a.x = x; a.out = U0; a.M = d_.pre_pad ? P0_ : d_.M; a.N = d_.pre_pad ? P1_ : d_.N;
            a.top = top_; a.left = left_; a.P0 = P0_; a.P1 = P1_;
            const SlabCfg& c = row_cfg_[0];
            a.lines = c.lines; a.LP = c.LP; a.plan = lev_[0].a1.plan; a.tw = tw(lev_[0].a1);
            dim3 grid((unsigned)B, ceil_div(P0_, c.lines));
            launch("pad_rowfft", (double)B * (a.M * a.N * sizeof(T) + (double)P0_ * P1_ * sizeof(cx<T>)), st,
                   [&] { k2d_pad_rowfft<T><<<grid, c.block, c.smem, st>>>(a); });
            col_pass<COL_FWD>(U0, 0, B, st);
        }
        // S0 (core/scattering2d.py:18-28); J == 0 has no filters: handled by the caller
        if (J == 0) throw std::runtime_error("J = 0 is not supported");
        low_pass(U0, 0, out, B, 1, 1, 0, 0, TL, st);
        int ch2 = 1 + L * J;   // first second-order channel
        for (int j1 = 0; j1 < J; ++j1) {
            // U1 = fft2(|ifft2(periodise(U0 * psi_j1))|)   (core:30-40)
            const T sc1 = T(1) / (T(1 << j1) * T(1 << j1) * T(lev_[j1].a0.n) * T(lev_[j1].a1.n));
            row_prod(U0, psi(j1, 0), U1, 0, j1, B, L, sc1, st);
            col_pass<COL_INV_MOD_FWD>(U1, j1, B * L, st);
            row_pass<false>(U1, j1, B * L, st);
            low_pass(U1, j1, out, B, L, L, 1 + j1 * L, 0, TL, st);   // core:42-51
            if (d_.max_order < 2) continue;
            const int nchild = (J - 1 - j1) * L;
            for (int j2 = j1 + 1; j2 < J; ++j2) {
                // U2 = fft2(|ifft2(periodise(U1 * psi_j2[level j1]))|)   (core:55-68)
                const int kk = 1 << (j2 - j1);
                const T sc2 = T(1) / (T(kk) * T(kk) * T(lev_[j2].a0.n) * T(lev_[j2].a1.n));
                row_prod(U1, psi(j2, j1), U2, j1, j2, B * L, L, sc2, st);
                col_pass<COL_INV_MOD_FWD>(U2, j2, B * L * L, st);
                row_pass<false>(U2, j2, B * L * L, st);
                low_pass(U2, j2, out, B, L * L, L, ch2 + (j2 - j1 - 1) * L, nchild, TL, st);   // core:70-83
            }
            ch2 += L * nchild;
        }
    }

    scat_plan2d_desc d_;
    int P0_ = 0, P1_ = 0, top_ = 0, left_ = 0, m0_ = 0, m1_ = 0, o0_ = 0, o1_ = 0, K_ = 0;
    std::vector<Level2D> lev_;
    std::vector<SlabCfg> row_cfg_, col_cfg_;
    size_t tables_bytes_ = 0, const_bytes_ = 0;
    std::vector<unsigned char> host_tables_;
    std::vector<size_t> phi_off_;
    std::vector<std::vector<size_t>> psi_off_;
    int n_psi_expected_ = 0;
    int lowW_ = 0; size_t low_smem_ = 0; bool low_tile_ = true;
    size_t ws_u0_ = 0, ws_u1_ = 0, ws_u2_ = 0, ws_low_ = 0, per_img_ = 0;
    unsigned char* cbuf_ = nullptr;
    bool bound_ = false;
    int last_B_ = 1;
};

}  // namespace sb
