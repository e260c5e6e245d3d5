// scat3d_inst.cu - instances of the fused 3-D kernels (kernels3d.cuh): M-axis lengths 8..256, square planes 8..128.
#include "kernels3d.cuh"
#include "common.cuh"

namespace sb {

#define SB_M3_SIZES(X) X(8) X(16) X(32) X(64) X(96) X(128) X(192) X(256)
// planes (N, O): N a power of two (the |.|^2 accumulation hooks the last, untwiddled power-of-two pass), O % 16 == 0
#define SB_P3_SIZES(X) X(16, 16) X(32, 32) X(64, 64) X(128, 128) X(128, 96) X(64, 96) X(64, 48)

template <typename T> void (*kern3d_col_prod(int M))(ColProd3<T>) {
#define SB_CASE(N) if (M == N) return k3d_col_prod<T, N>;
    SB_M3_SIZES(SB_CASE)
#undef SB_CASE
    return nullptr;
}
template <typename T> void (*kern3d_col_fwd(int M))(ColFwd3<T>) {
#define SB_CASE(N) if (M == N) return k3d_col_fwd<T, N>;
    SB_M3_SIZES(SB_CASE)
#undef SB_CASE
    return nullptr;
}
template <typename T> void (*kern3d_plane(int N, int O))(Plane3<T>) {
#define SB_CASE(A, B) if (N == A && O == B) return k3d_plane<T, A, B / 2>;
    SB_P3_SIZES(SB_CASE)
#undef SB_CASE
    return nullptr;
}
template <typename T> void (*kern3d_plane_real(int N, int O))(PlaneReal3<T>) {
#define SB_CASE(A, B) if (N == A && O == B) return k3d_plane_real<T, A, B / 2>;
    SB_P3_SIZES(SB_CASE)
#undef SB_CASE
    return nullptr;
}
void kern3d_enable_smem() {
#define SB_EN(N) enable_big_smem(k3d_col_prod<float, N>); enable_big_smem(k3d_col_fwd<float, N>);
    SB_M3_SIZES(SB_EN)
#undef SB_EN
#define SB_EN(A, B) enable_big_smem(k3d_plane<float, A, B / 2>); enable_big_smem(k3d_plane_real<float, A, B / 2>);
    SB_P3_SIZES(SB_EN)
#undef SB_EN
}

template void (*kern3d_col_prod<float>(int))(ColProd3<float>);
template void (*kern3d_plane<float>(int, int))(Plane3<float>);
template void (*kern3d_col_fwd<float>(int))(ColFwd3<float>);
template void (*kern3d_plane_real<float>(int, int))(PlaneReal3<float>);

}  // namespace sb
