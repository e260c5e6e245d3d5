// stream_inst.cu - instances of the streaming slab kernels: compile-time specialised for the
// full-resolution line lengths of the BASELINE.json configurations (272 = 256^2 J=3 padded,
// 256 = 224^2 J=4 padded, 240 = 224^2 J=3 padded) plus the generic runtime-size versions.
#include "kernels2d.cuh"
#include "common.cuh"

namespace sb {

#define SB_STREAM_SIZES(X) X(272) X(256) X(240)

template <typename T, int NS> static StreamKernels<T> make_table() {
    StreamKernels<T> k;
    k.pad_rowfft = k2d_pad_rowfft<T, NS>;
    k.col_fwd = k2d_colpass<T, COL_FWD, NS>;
    k.col_inv = k2d_colpass<T, COL_INV, NS>;
    k.col_imf = k2d_colpass<T, COL_INV_MOD_FWD, NS>;
    k.row_prod = k2d_rowpass_prod<T, NS>;
    k.row_fwd = k2d_rowpass<T, false, NS>;
    k.row_inv = k2d_rowpass<T, true, NS>;
    if constexpr (NS > 0 && NS % 2 == 0) { k.col_imrf = k2d_colpass_imrf<T, NS>; k.row_fwdh = k2d_rowpass_fwdh<T, NS>; }
    else { k.col_imrf = nullptr; k.row_fwdh = nullptr; }
    k.is_static = NS > 0;
    return k;
}

template <typename T> StreamKernels<T> stream_kernels_lookup(int n, bool allow_static) {
    if (allow_static) {
#define SB_CASE(N) if (n == N) return make_table<T, N>();
        SB_STREAM_SIZES(SB_CASE)
#undef SB_CASE
    }
    return make_table<T, 0>();
}

template <typename T, int NS> static void enable_table() {
    StreamKernels<T> k = make_table<T, NS>();
    enable_big_smem(k.pad_rowfft); enable_big_smem(k.col_fwd); enable_big_smem(k.col_inv);
    enable_big_smem(k.col_imf); enable_big_smem(k.row_prod); enable_big_smem(k.row_fwd);
    enable_big_smem(k.row_inv);
    if (k.col_imrf) { enable_big_smem(k.col_imrf); enable_big_smem(k.row_fwdh); }
}
template <typename T> void stream_kernels_enable_smem() {
#define SB_EN(N) enable_table<T, N>();
    SB_STREAM_SIZES(SB_EN)
#undef SB_EN
    enable_table<T, 0>();
}

template StreamKernels<float> stream_kernels_lookup<float>(int, bool);
template StreamKernels<double> stream_kernels_lookup<double>(int, bool);
template void stream_kernels_enable_smem<float>();
template void stream_kernels_enable_smem<double>();

}  // namespace sb
