// scat1d_inst.cu - instances of the fused 1-D kernels (kernels1d.cuh) for every power-of-two split
// N = NA * NB with NA in 1..512, NB in 16..512 (N = 16 .. 2^18), and low-pass lengths M = 8..1024.
#include "kernels1d.cuh"
#include "common.cuh"

namespace sb {

#define SB_NA_SIZES(X) X(1) X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512)
#define SB_NB_SIZES(X) X(16) X(32) X(64) X(128) X(256) X(512)
#define SB_M_SIZES(X) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024)

template <typename T> Kern1d<T> kern1d_cols(int NA) {
    Kern1d<T> k{nullptr, nullptr};
#define SB_CASE(N) if (NA == N) { k.col_prod = k1d_col_prod<T, N>; k.col_fwd = k1d_col_fwd<T, N>; }
    SB_NA_SIZES(SB_CASE)
#undef SB_CASE
    return k;
}
template <typename T> KernRow1d<T> kern1d_rows(int NB) {
    KernRow1d<T> k{nullptr, nullptr, nullptr};
#define SB_CASE(N) if (NB == N) { k.parent = k1d_row_mod<T, N, false>; k.leaf = k1d_row_mod<T, N, true>; k.real = k1d_row_real<T, N>; }
    SB_NB_SIZES(SB_CASE)
#undef SB_CASE
    return k;
}
template <typename T> void (*kern1d_finish(int M))(Finish1<T>) {
#define SB_CASE(N) if (M == N) return k1d_finish<T, N>;
    SB_M_SIZES(SB_CASE)
#undef SB_CASE
    return nullptr;
}
#define SB_TILE1_SIZES(X) X(1, 16) X(2, 16) X(4, 16) X(8, 16) X(16, 16) X(16, 32) X(32, 32) X(32, 64) X(64, 64) X(64, 128)
template <typename T> void (*kern1d_tile(int NA, int NB))(Tile1<T>) {
#define SB_CASE(A, B) if (NA == A && NB == B) return k1d_tile<T, A, B>;
    SB_TILE1_SIZES(SB_CASE)
#undef SB_CASE
    return nullptr;
}
void kern1d_enable_smem() {
#define SB_EN(A, B) enable_big_smem(k1d_tile<float, A, B>);
    SB_TILE1_SIZES(SB_EN)
#undef SB_EN
#define SB_EN(N) enable_big_smem(k1d_col_prod<float, N>); enable_big_smem(k1d_col_fwd<float, N>);
    SB_NA_SIZES(SB_EN)
#undef SB_EN
#define SB_EN(N) enable_big_smem(k1d_row_mod<float, N, false>); enable_big_smem(k1d_row_mod<float, N, true>); \
                 enable_big_smem(k1d_row_real<float, N>);
    SB_NB_SIZES(SB_EN)
#undef SB_EN
#define SB_EN(N) enable_big_smem(k1d_finish<float, N>);
    SB_M_SIZES(SB_EN)
#undef SB_EN
}

template Kern1d<float> kern1d_cols<float>(int);
template KernRow1d<float> kern1d_rows<float>(int);
template void (*kern1d_finish<float>(int))(Finish1<float>);
template void (*kern1d_tile<float>(int, int))(Tile1<float>);

}  // namespace sb
