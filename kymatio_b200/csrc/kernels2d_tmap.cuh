// kernels2d_tmap.cuh - the product row pass for POWER-OF-TWO line lengths (256: BASELINE configs[4], 224 x 224 padded,
// J = 4), fed by TMA TENSOR copies with the hardware 128-byte swizzle.
//
// Why a second TMA variant: the 1-D bulk copies of kernels2d_tma.cuh deliver dense row-major rows.  For 272 = 16 x 17 that
// layout is bank-conflict free; for 256 = 16 x 16 the second radix pass reads, per thread, 16 CONTIGUOUS elements - a
// thread stride of exactly 128 bytes, i.e. every lane on the same banks.  Padding is impossible (a bulk copy needs dense,
// 16-byte aligned rows), so the slab is described to the TMA unit as a 2-D tensor instead:
//     global  [rows][512 floats] (a row of 256 complex values)        box {32 floats = 128 B, 16 rows}
// and copied box by box with CU_TENSOR_MAP_SWIZZLE_128B: box k (the 16 complex values 16k .. 16k+15 of each of the 16 rows)
// lands as 16 rows x 128 B in which the 16-byte chunk index is XORed with (row & 7).  With that layout
//   * pass 0 (radix 16, stride 16 elements): thread (row, e) touches the SAME offset of the 16 boxes - a half-warp covers
//     one permuted 128-byte row: conflict free;
//   * pass 1 (radix 16, contiguous): thread (box, row) with the ROW fastest across lanes reads its 8 chunks with LDS.128 at
//     chunk (c ^ (row & 7)): the 8 lanes of a quarter-warp hit 8 different chunk positions: conflict free.
// The store goes back through the same tensor map, which undoes the swizzle: HBM sees the usual dense scrambled-order
// rows, so the column pass and every other consumer are unchanged.  SASS: UTMALDG / UTMASTG + SYNCS.
#pragma once
#include "kernels2d.cuh"
#include "tma.cuh"

namespace sb {

constexpr int kTmapRows = 16;
constexpr int kTmapComputeThreads = 256;                        // 16 rows x 16 butterflies per pass
constexpr int kTmapThreads = kTmapComputeThreads + 32;          // + the producer warp
constexpr uint32_t kTmapBoxBytes = 16 * 128;                    // 16 rows x 16 complex
constexpr uint32_t kTmapSlabBytes = 16 * kTmapBoxBytes;         // 16 boxes = 16 rows x 256 complex

constexpr size_t tmap_row_smem_bytes() { return 1024 + 2 * (size_t)kTmapSlabBytes + 256 * sizeof(cx<float>) + 4 * sizeof(uint64_t); }

// host side (tmap_inst.cu): encode the two tensor maps and launch; false when the driver entry point is unavailable
bool rowprod_tmap256_launch(const RowProdArgs<float>& a, int Bp, int m, int npairs, cudaStream_t st);
// Hermitian forward row pass (in place) of G paths; grid = persistent CTAs
bool rowfwdh_tmap256_launch(const RowArgs<float>& a, int G, int grid, cudaStream_t st);
// column pass (inverse, modulus, real forward) of G 272 x 272 or 256 x 256 paths, in place, staged by dense tensor copies
bool colpass_imrf_tmap_launch(const ColArgs<float>& a, int G, int ctas_per_sm, int num_sms, cudaStream_t st);
void tmap_kernels_enable_smem();

}  // namespace sb
