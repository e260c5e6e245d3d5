// tma_inst.cu - instances of the TMA-fed streaming row passes (kernels2d_tma.cuh).
#include "kernels2d_tma.cuh"
#include "common.cuh"

namespace sb {

#define SB_TMA_SIZES(X) X(272)

RowProdTmaKernel rowprod_tma_lookup(int n) {
#define SB_CASE(N) if (n == N) return k2d_rowprod_tma<N>;
    SB_TMA_SIZES(SB_CASE)
#undef SB_CASE
    return nullptr;
}
RowFwdhTmaKernel rowfwdh_tma_lookup(int n) {
#define SB_CASE(N) if (n == N) return k2d_rowfwdh_tma<N>;
    SB_TMA_SIZES(SB_CASE)
#undef SB_CASE
    return nullptr;
}
size_t tma_row_smem(int n) {
#define SB_CASE(N) if (n == N) return tma_row_smem_bytes<N>();
    SB_TMA_SIZES(SB_CASE)
#undef SB_CASE
    return 0;
}
void tma_kernels_enable_smem() {
#define SB_EN(N) enable_big_smem(k2d_rowprod_tma<N>); enable_big_smem(k2d_rowfwdh_tma<N>);
    SB_TMA_SIZES(SB_EN)
#undef SB_EN
}

}  // namespace sb
