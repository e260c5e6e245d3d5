/* scat_b200.h - C ABI of libscat_b200.so, the sm_100a wavelet-scattering engine.
 *
 * Boundary replaced: the per-primitive backend protocol of the kymatio torch frontends
 *   kymatio/frontend/base_frontend.py:32-50          (backend binding)
 *   kymatio/scattering2d/core/scattering2d.py:3-9    (rfft, ifft, irfft, cdgmm,
 *                                                     subsample_fourier, modulus, stack)
 *   kymatio/scattering2d/frontend/torch_frontend.py:72-111 (caller of the core)
 * and, for the fused path, the whole core call
 *   kymatio/scattering2d/frontend/torch_frontend.py:98-99  scattering2d(...)
 *
 * Conventions: every function returns 0 on success, non-zero on error
 * (scat_last_error() gives the message, thread-local).  All pointers named *_dev are
 * device pointers owned by the caller (torch tensors); the library never allocates
 * device memory, never synchronises the device and launches only on the given stream.
 * dtype: 0 = float32, 1 = float64 (the reference's gradcheck runs in float64:
 * tests/scattering2d/test_torch_scattering2d.py:238-248).
 */
#ifndef SCAT_B200_H
#define SCAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scat_plan2d scat_plan2d;

typedef struct scat_plan2d_desc {
    int32_t M, N;          /* un-padded spatial size (ScatteringBase2D.shape)            */
    int32_t J, L;          /* scales / angles        (base_frontend.py:8-17)              */
    int32_t max_order;     /* 1 or 2                 (core/scattering2d.py:53-54)          */
    int32_t pre_pad;       /* input already padded   (torch_frontend.py:16-18)             */
    int32_t dtype;         /* 0 = f32, 1 = f64                                           */
    int32_t reserved;
} scat_plan2d_desc;

/* library / device ---------------------------------------------------------------- */
int  scat_version(void);
const char* scat_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
uint64_t scat_launch_count(void);

/* per-launch timing with CUDA events on the launching stream (bench.py's live roofline):
 * enable, run, then report -> "label\tcount\ttotal_ms\ttotal_algorithmic_bytes\n" per kernel label.
 * scat_timing_report synchronises on the recorded events; returns the bytes needed. */
void scat_timing_enable(int on);
size_t scat_timing_report(char* buf, size_t buflen);
/* profiling build only (make prof -> libscat_b200_prof.so): cycles spent per phase of the 2-D tile kernels,
 * out[kind * 8 + phase]; returns the number of slots written, 0 in the production library.  Synchronises the device. */
int scat_phase_prof_read(unsigned long long* out, int max_n, int reset);

/* 2-D plan ----------------------------------------------------------------------- */
int  scat_plan2d_create(const scat_plan2d_desc* desc, scat_plan2d** out_plan);
void scat_plan2d_destroy(scat_plan2d* plan);

/* padded size, output size, number of channels K = 1 + LJ + L^2 J(J-1)/2 */
int  scat_plan2d_info(const scat_plan2d* plan, int32_t* Mp, int32_t* Np, int32_t* out_h, int32_t* out_w,
                      int32_t* K);

/* bytes of plan constants (twiddles, scramble tables, scrambled filter bank) the caller
 * must provide as one device buffer */
size_t scat_plan2d_const_bytes(const scat_plan2d* plan);

/* Bind the filter bank.  phi_dev[J]: low-pass at resolutions 0..J-1, natural order,
 * (Mp/2^r, Np/2^r) real; psi_dev[n_psi]: band-pass levels flattened in the frontend's
 * registration order (torch_frontend.py:28-43): for n in 0..J*L-1, for each level.
 * Filters are copied (scrambled) into const_dev; the originals are not referenced later.
 * Call again whenever the module buffers change (.to(), .double(), load_state_dict). */
int  scat_plan2d_bind(scat_plan2d* plan, void* const_dev, const void* const* phi_dev, int32_t n_phi,
                      const void* const* psi_dev, int32_t n_psi, void* stream);

/* workspace bytes for a forward over `batch` images (the plan chunks large batches) */
size_t scat_plan2d_workspace_bytes(const scat_plan2d* plan, int64_t batch);

/* x_dev: (batch, M, N) real [or (batch, Mp, Np) when pre_pad]; out_dev: (batch, K, out_h, out_w) */
int  scat_plan2d_forward(scat_plan2d* plan, const void* x_dev, void* out_dev, void* workspace_dev,
                         size_t workspace_bytes, int64_t batch, void* stream);
/* The same forward with every coefficient plane ALSO stored at the same offset of n_peers (<= 7) more buffers: the
 * output tensors of the peer GPUs (symmetric memory mapped over NVLink).  Batch-sharded ranks that each pass the others'
 * buffers end up with the full (total_batch, K, oh, ow) tensor without a separate all-gather pass; the caller places a
 * cross-rank barrier after the call (kymatio_b200/parallel.py: PeerGatherScattering).  peer_out_dev is a HOST array of
 * device pointers, each already offset to this rank's block.  n_peers == -1: peer_out_dev[0] is the MULTICAST (NVLS)
 * address of the block; every plane is then stored ONCE with multimem.st and replicated by the NVSwitch into all ranks'
 * buffers, the caller's own included (out_dev is not written). */
int  scat_plan2d_forward_peers(scat_plan2d* plan, const void* x_dev, void* out_dev, void* const* peer_out_dev,
                               int32_t n_peers, void* ws_dev, size_t ws_bytes, int64_t batch, void* stream);

/* The same forward that KEEPS the first-order spectra for a later backward pass (what an autograd engine saves between
 * kymatio/scattering2d/core/scattering2d.py:43-45 and its backward): saved_u1_dev is a HOST array of J device pointers;
 * entry j1 (NULL: not kept) receives U1 of scale j1 for the whole batch, (batch*L, Mp/2^j1, Np/2^j1) complex, natural order -
 * the `u1_dev` operand of scat_plan2d_order2_backward and the tensor scat_plan2d_order1_forward would recompute. */
int  scat_plan2d_forward_save(scat_plan2d* plan, const void* x_dev, void* out_dev, void* const* saved_u1_dev, void* ws_dev,
                              size_t ws_bytes, int64_t batch, void* stream);

/* First-order block of scale j1 as a stand-alone differentiable operator on a caller-provided U0 = fft2(pad(x))
 * (kymatio/scattering2d/core/scattering2d.py:30-51; gradients: SURVEY Appendix B, kymatio/backend/torch_backend.py:64-96).
 * mode 1 (the field fits one CTA): forward returns S1 (batch, L, oh, ow) and, when u1_dev != NULL, U1 = fft2(|.|)
 *   (batch*L, n0, n1) complex; backward takes gs1 and (optionally) gu1.
 * mode 2 (full resolution, streaming chain): forward returns U1 only (s1_dev = NULL; the caller low-passes U1);
 *   backward takes gu1.  mode 0: not available (caller uses the per-primitive ops).
 * backward ACCUMULATES into gu0_dev (batch, Mp, Np) complex; ws_dev: scat_plan2d_order1_workspace_bytes. */
int32_t scat_plan2d_order1_mode(const scat_plan2d* plan, int32_t j1);
size_t scat_plan2d_order1_workspace_bytes(const scat_plan2d* plan, int32_t j1, int64_t batch);
int  scat_plan2d_order1_forward(scat_plan2d* plan, int32_t j1, const void* u0_dev, void* s1_dev, void* u1_dev, int64_t batch,
                                void* stream);
int  scat_plan2d_order1_backward(scat_plan2d* plan, int32_t j1, const void* u0_dev, const void* gs1_dev, const void* gu1_dev,
                                 void* gu0_dev, void* ws_dev, size_t ws_bytes, int64_t batch, void* stream);

/* The second-order block of first-order scale j1 as a stand-alone differentiable operator (used by the
 * autograd path): u1_dev = (batch*L, n0_j1, n1_j1) complex natural-order spectra of the first-order moduli
 * (what `rfft(modulus(...))` returns at core/scattering2d.py:38-40); out = (batch, C2, out_h, out_w) with
 * C2 = scat_plan2d_order2_channels(j1) channels ordered (theta1, j2, theta2) (core/scattering2d.py:55-83).
 * backward overwrites gu1_dev (same shape as u1_dev) with the gradient; it accumulates with atomics, so
 * results may differ in the last bits from run to run.  _channels returns 0 when the block is unavailable. */
int32_t scat_plan2d_order2_channels(const scat_plan2d* plan, int32_t j1);
int  scat_plan2d_order2_forward(scat_plan2d* plan, int32_t j1, const void* u1_dev, void* out_dev, int64_t batch,
                                void* stream);
int  scat_plan2d_order2_backward(scat_plan2d* plan, int32_t j1, const void* u1_dev, const void* gout_dev,
                                 void* gu1_dev, int64_t batch, void* stream);

/* eager primitives ----------------------------------------------------------------
 * One entry point per backend primitive of kymatio/scattering2d/core/scattering2d.py:3-9, on
 * contiguous device tensors in the torch backend's layout (real: trailing axis 1, complex: trailing
 * axis 2).  Not used by the fused plan; they complete the backend protocol for the `torch_b200`
 * backend object. */

/* natural-order complex 2-D FFT on (G, n0, n1, 2): tables -> const_dev (caller-owned), then exec.
 * inverse: 0 = forward, 1 = inverse normalised by 1/(n0 n1), 2 = inverse without normalisation (adjoint of the forward).
 * replaces torch.fft.fft2 / ifft2 at kymatio/scattering2d/backend/torch_backend.py:10-12,134-155 */
size_t scat_fft2d_const_bytes(int32_t n0, int32_t n1, int32_t dtype);
int  scat_fft2d_init(void* const_dev, int32_t n0, int32_t n1, int32_t dtype, void* stream);
int  scat_fft2d_exec(const void* const_dev, const void* in_dev, void* out_dev, int64_t G, int32_t n0, int32_t n1,
                     int32_t inverse, int32_t dtype, void* stream);
/* reflect padding, (B, M, N) real -> (B, M+top+bottom, N+left+right) real  (torch_backend.py:36-86) */
int  scat_pad2d(const void* x_dev, void* out_dev, int64_t B, int32_t M, int32_t N, int32_t top, int32_t bottom,
                int32_t left, int32_t right, int32_t dtype, void* stream);
/* out[b][i] = a[b][i] * b[i], a complex (batch, n), b real (n) or complex (n)  (backend/torch_backend.py:148-219) */
int  scat_cdgmm(const void* a_dev, const void* b_dev, void* out_dev, int64_t batch, int64_t n, int32_t b_is_complex,
                int32_t dtype, void* stream);
/* Fourier-domain periodisation (G, n0, n1) -> (G, n0/k, n1/k)  (scattering2d/backend/torch_backend.py:93-129) */
int  scat_subsample_fourier2d(const void* in_dev, void* out_dev, int64_t G, int32_t n0, int32_t n1, int32_t k,
                              int32_t dtype, void* stream);
/* |z| on n complex values  (backend/torch_backend.py:138-141) */
int  scat_modulus(const void* in_dev, void* out_dev, int64_t n, int32_t dtype, void* stream);
/* real <-> complex views used by rfft / irfft  (scattering2d/backend/torch_backend.py:134-148) */
int  scat_complex_from_real(const void* in_dev, void* out_dev, int64_t n, int32_t dtype, void* stream);
int  scat_real_part(const void* in_dev, void* out_dev, int64_t n, int32_t dtype, void* stream);

/* 1-D primitives (kymatio/scattering1d/backend/torch_backend.py:19-141) --------------------------------
 * natural-order complex FFT of any length N = Na*Nb on (G, N, 2) via the four-step algorithm (column pass,
 * twiddle, row pass, transposed store); tmp_dev is a caller-owned scratch of the input's size. */
size_t scat_fft1d_const_bytes(int32_t N, int32_t dtype);
int  scat_fft1d_init(void* const_dev, int32_t N, int32_t dtype, void* stream);
int  scat_fft1d_exec(const void* const_dev, const void* in_dev, void* tmp_dev, void* out_dev, int64_t G, int32_t N,
                     int32_t inverse, int32_t dtype, void* stream);
/* reflect padding along time, (G, N) -> (G, N + pad_left + pad_right)  (torch_backend.py:51-82) */
int  scat_pad1d(const void* x_dev, void* out_dev, int64_t G, int32_t N, int32_t pad_left, int32_t pad_right,
                int32_t dtype, void* stream);
/* Fourier periodisation (G, N) -> (G, N/k)  (torch_backend.py:19-48; CUDA: torch_skcuda_backend.py:133-164) */
int  scat_subsample_fourier1d(const void* in_dev, void* out_dev, int64_t G, int32_t N, int32_t k, int32_t dtype,
                              void* stream);

/* fused 1-D path (kymatio/scattering1d/core/scattering1d.py:40-107), float32 ------------------------------
 * The cascade (which paths, filters, channels) is driven by the host engine (kymatio_b200/engine1d.py); the
 * library exposes one entry point per fused kernel.  A path transform of length N = Na*Nb (powers of two,
 * 16 <= N <= 2^18) is three passes: col_prod -> row_mod -> col_fwd (parents) or col_prod -> row_mod(leaf)
 * (paths whose spectrum only feeds the low-pass), then finish.  `algo_bytes` is the caller's algorithmic byte
 * count of the launch (only recorded by scat_timing_*).  tables_dev / fin_tables_dev are caller-owned. */
int    scat1d_split(int32_t N, int32_t* Na, int32_t* Nb);
size_t scat1d_tables_bytes(int32_t N);
int    scat1d_tables_init(void* tables_dev, int32_t N, void* stream);
size_t scat1d_fin_tables_bytes(int32_t M);
int    scat1d_fin_tables_init(void* tables_dev, int32_t M, void* stream);
/* Y[g] = twiddled column-inverse of periodise_{Npar/N}(parent(g) * filt[g % NI]) / (N k); path g = b*NI + i reads the
 * natural-order parent spectrum at parent_dev + b*ps_b + i*ps_i (complex elements); filt_ptrs_dev: device array of NI
 * pointers to real filters of length Npar (cdgmm + subsample_fourier + first half of ifft,
 * core/scattering1d.py:61-63,92-94); supp_dev: NI (start, len) int32 pairs, circular support of each filter */
int    scat1d_col_prod(const void* tables_dev, const void* parent_dev, int64_t ps_b, int64_t ps_i, const void* filt_ptrs_dev,
                       const void* supp_dev, void* y_dev, int64_t G, int32_t NI, int32_t Npar, int32_t N, double algo_bytes,
                       void* stream);
/* second half of ifft, modulus, first half of rfft (core/scattering1d.py:63-69,94-100), in place on y_dev;
 * part_dev != NULL: leaf mode, writes the Fc lowest bins of the spectrum as ceil(Na/16) partial sums per path,
 * part_dev[(g*ceil(Na/16) + c)*Fc + f] */
int    scat1d_row_mod(const void* tables_dev, void* y_dev, int64_t G, int32_t N, void* part_dev, int32_t Fc, double algo_bytes,
                      void* stream);
/* second half of rfft: natural-order spectrum (G, N) out */
int    scat1d_col_fwd(const void* tables_dev, const void* z_dev, void* out_dev, int64_t G, int32_t N, double algo_bytes,
                      void* stream);
/* U_hat = rfft(x) of the padded real signals x_dev (G, N) (core/scattering1d.py:41; the real -> complex copy of
 * scattering1d/backend/torch_backend.py:109-113 stays in shared memory); z_dev: (G, N) complex scratch, may alias out_dev */
int    scat1d_rfft(const void* tables_dev, const void* x_dev, void* z_dev, void* out_dev, int64_t G, int32_t N, void* stream);
/* whole path in ONE launch for short transforms (N <= scat1d_tile_max()): product + periodise, inverse, modulus, forward
 * all in one CTA's shared memory; writes the natural-order spectrum to spec_dev (G, N) when non-NULL (parents) and/or
 * the Fc lowest bins to part_dev (G, Fc) when non-NULL (leaves: one "partial" per path for scat1d_finish) */
int    scat1d_tile_max(void);
int    scat1d_tile(const void* tables_dev, const void* parent_dev, int64_t ps_b, int64_t ps_i, const void* filt_ptrs_dev,
                   const void* supp_dev, void* spec_dev, void* part_dev, int32_t Fc, int64_t G, int32_t NI, int32_t Npar,
                   int32_t N, double algo_bytes, void* stream);
/* cdgmm(phi) -> subsample_fourier(N/M) -> irfft -> unpad[i0:i0+W] (core/scattering1d.py:72-77,101-105 and
 * frontend/base_frontend.py:137-139) for every path of a batch chunk in ONE launch.  Line (= path) `line` belongs to
 * the last segment with line0 <= line; with gl = line - line0 = b*NI + i its spectrum is
 * X[f] = sum_{q<nparts} base[which][src_off + gl*ss_g + q*ss_part + f], f < Fc (the negative bins follow from the
 * Hermitian symmetry), base = {u0_dev, u1_dev, part_dev}; writes out[b*os_b + chan[i]*W + n]. */
typedef struct scat1d_finseg {
    int64_t src_off, ss_g, ss_part;   /* complex elements */
    const void* phi_dev;              /* real low-pass on the length-N grid */
    const void* chan_dev;             /* int32[NI] */
    int32_t which, nparts, N, Fc, NI, line0;
} scat1d_finseg;
/* T = 0 (no averaging: kymatio/scattering1d/core/scattering1d.py:75-76,104-105 yield the modulus fields themselves):
 * the same path kernels, which also store |u| in NATURAL time order at mod_dev[g*N + t] (float); is_leaf: nothing else to
 * do (no spectrum for children); scat1d_tile_t0 with spec_dev != NULL also writes the natural-order spectrum. */
int scat1d_row_mod_t0(const void* tables_dev, void* y_dev, int64_t G, int32_t N, void* mod_dev, int32_t is_leaf,
                      double algo_bytes, void* stream);
int scat1d_tile_t0(const void* tables_dev, const void* parent_dev, int64_t ps_b, int64_t ps_i, const void* filt_ptrs_dev,
                   const void* supp_dev, void* spec_dev, void* mod_dev, int64_t G, int32_t NI, int32_t Npar, int32_t N,
                   double algo_bytes, void* stream);
/* average='global' tail (kymatio/scattering1d/frontend/base_frontend.py:137-138): out[b*os_b + chan] = bin 0 of the
 * path's spectrum = the sum over time of its modulus field; same segment table as scat1d_finish. */
int scat1d_finish_global(const void* u0_dev, const void* u1_dev, const void* part_dev, const void* segs_dev, int32_t nseg,
                         int64_t total_lines, void* out_dev, int64_t os_b, void* stream);
size_t scat1d_finseg_bytes(void);
int    scat1d_finish(const void* fin_tables_dev, const void* u0_dev, const void* u1_dev, const void* part_dev,
                     const void* segs_dev, int32_t nseg, int64_t total_lines, int32_t M, void* out_dev, int64_t os_b,
                     int32_t i0, int32_t W, double algo_bytes, void* stream);

/* 3-D primitives (kymatio/scattering3d/backend/torch_backend.py:73-151) --------------------------------
 * natural-order complex 3-D FFT on (G, M, N, O, 2) - replaces torch.fft.fftn / ifftn (torch_backend.py:39-40) */
size_t scat_fft3d_const_bytes(int32_t M, int32_t N, int32_t O, int32_t dtype);
int  scat_fft3d_init(void* const_dev, int32_t M, int32_t N, int32_t O, int32_t dtype, void* stream);
int  scat_fft3d_exec(const void* const_dev, const void* in_dev, void* out_dev, int64_t G, int32_t M, int32_t N, int32_t O,
                     int32_t inverse, int32_t dtype, void* stream);
/* out = sqrt(prev^2 + |x|^2), prev may be NULL  (modulus_rotation, torch_backend.py:102-124) */
int  scat_modulus_rotation(const void* x_dev, const void* prev_dev, void* out_dev, int64_t n, int32_t dtype, void* stream);
/* out[b][p] = sum_i x[b][i]^powers[p], float64 accumulators (compute_integrals, torch_backend.py:127-151) */
int  scat_compute_integrals(const void* x_dev, void* out_f64_dev, int64_t B, int64_t n, const void* powers_f32_dev,
                            int32_t P, int32_t dtype, void* stream);

/* fused 3-D path (kymatio/scattering3d/core/scattering3d.py:24-73), float32, power-of-two volumes ---------
 * One band (l, j) = cdgmm3d + ifft + modulus_rotation over its nm = 2l+1 filters + compute_integrals (+ rfft for
 * parents) runs as: col_prod (filter product + inverse transform along M, all m in one launch) -> plane (2-D inverse
 * of every (N, O) plane in shared memory, sum_m |.|^2 in registers, sqrt, voxel sums of U^q added to
 * integ[b*istride + ioff + p] (float64 atomics), and for parents the 2-D forward of (U, 0)) -> col_fwd (forward
 * along M).  u_dev / out_dev: (B, M, N, O) complex natural-order spectra; filt_dev: (nm, M, N, O) complex;
 * y_dev: (B*nm, M, N, O) complex scratch; spec_dev: (B, M, N, O) complex scratch or NULL for a leaf band. */
int    scat3d_supported(int32_t M, int32_t N, int32_t O);
size_t scat3d_tables_bytes(int32_t M, int32_t N, int32_t O);
int    scat3d_tables_init(void* tables_dev, int32_t M, int32_t N, int32_t O, void* stream);
/* U0_hat = rfft(x) of real volumes x_dev (B, M, N, O) -> out_dev (B, M, N, O) complex natural-order spectrum
 * (core/scattering3d.py:24; replaces the zero-imaginary copy + fftn of scattering3d/backend/torch_backend.py:81-87) */
int    scat3d_rfft(const void* tables_dev, const void* x_dev, void* out_dev, int64_t B, int32_t M, int32_t N, int32_t O,
                   void* stream);
int    scat3d_col_prod(const void* tables_dev, const void* u_dev, const void* filt_dev, void* y_dev, int64_t B, int32_t nm,
                       int32_t M, int32_t N, int32_t O, void* stream);
int    scat3d_plane(const void* tables_dev, const void* y_dev, void* spec_dev, void* integ_f64_dev, int64_t istride,
                    int32_t ioff, const void* powers_f32_dev, int32_t P, int64_t B, int32_t nm, int32_t M, int32_t N,
                    int32_t O, void* stream);
int    scat3d_col_fwd(const void* tables_dev, const void* z_dev, void* out_dev, int64_t B, int32_t M, int32_t N, int32_t O,
                      void* stream);

/* adjoints for the autograd graph ------------------------------------------------
 * filter multiply with the filters broadcast over the batch: out[b][f][i] = a[b][i] * w[f][i] (w real);
 * adjoint = 1 computes ga[b][i] = sum_f a[b][f][i] * w[f][i]  (backward of cdgmm, backend/torch_backend.py:205-206) */
int  scat_cdgmm_bcast(const void* a_dev, const void* w_dev, void* out_dev, int64_t nb, int32_t nf, int64_t n,
                      int32_t adjoint, int32_t dtype, void* stream);
/* adjoint of subsample_fourier: replicate / k^2, (G, n0/k, n1/k) -> (G, n0, n1) */
int  scat_subsample_fourier2d_bwd(const void* gout_dev, void* gin_dev, int64_t G, int32_t n0, int32_t n1, int32_t k,
                                  int32_t dtype, void* stream);
/* ModulusStable.backward (backend/torch_backend.py:64-96): gx = x g / |x|, 0 where |x| = 0 */
int  scat_modulus_bwd(const void* x_dev, const void* g_dev, void* gx_dev, int64_t n, int32_t dtype, void* stream);
/* adjoint of the reflect padding: fold-add (gx is zeroed first) */
int  scat_pad2d_bwd(const void* gout_dev, void* gx_dev, int64_t B, int32_t M, int32_t N, int32_t top, int32_t bottom,
                    int32_t left, int32_t right, int32_t dtype, void* stream);

/* adjoints of the 1-D / 3-D eager primitives (gradients through backend='torch_b200' Scattering1D / HarmonicScattering3D):
 * periodisation -> replicate / k; sqrt(prev^2 + |x|^2) -> (x, prev) g / out (gprev_dev, prev_dev may be NULL);
 * integrals -> sum_p g[b][p] q_p x^(q_p - 1).  scat_cdgmm with b_is_complex = 2 multiplies by conj(b) (adjoint of cdgmm3d). */
int  scat_subsample_fourier1d_bwd(const void* gout_dev, void* gin_dev, int64_t G, int32_t N, int32_t k, int32_t dtype,
                                  void* stream);
int  scat_modulus_rotation_bwd(const void* x_dev, const void* prev_dev, const void* out_dev, const void* g_dev, void* gx_dev,
                               void* gprev_dev, int64_t n, int32_t dtype, void* stream);
int  scat_compute_integrals_bwd(const void* x_dev, const void* g_dev, void* gx_dev, int64_t B, int64_t n,
                                const void* powers_f32_dev, int32_t P, int32_t dtype, void* stream);

/* filter synthesis on the device (constructor path) -----------------------------------------
 * 2-D Morlet / Gabor bank, replaces the numpy loops of kymatio/scattering2d/filter_bank.py:5-53 (bank), :94-175 (wavelets):
 *   params_dev: n_filters x 8 doubles {c00, cross, c11, fu, fv, 1/norm, zero_mean flag, 0}: the filter is
 *     exp(-(c00 u^2 + cross u v + c11 v^2) + i (fu u + fv v)) / norm summed over the 5x5 neighbouring periods, minus
 *     beta * (the same with fu = fv = 0) when the flag is set, beta chosen so that the filter sums to zero;
 *   carrier_dev: (n_filters, M, N) complex128 out (the spatial filters); envelope_dev: (n_filters, M, N) float64 scratch;
 *   sums_dev: n_filters x 3 float64 scratch.
 * scat_filters2d_fold: band-limit + alias fold of the real part of one (M, N) complex128 spectrum to resolution res,
 *   (M/2^res, N/2^res) float32 out  - periodize_filter_fft, filter_bank.py:56-91. */
int scat_filters2d_spatial(const void* params_dev, int32_t n_filters, int32_t M, int32_t N, void* carrier_dev,
                           void* envelope_dev, void* sums_dev, void* stream);
int scat_filters2d_fold(const void* spec_dev, void* out_dev, int32_t M, int32_t N, int32_t res, void* stream);
/* 3-D solid harmonic wavelets of order l at n_scales widths, Fourier domain, closed form
 * (kymatio/scattering3d/filter_bank.py:100-166): out_dev (n_scales, 2l+1, M, N, O) complex64; sigmas_dev: n_scales float64;
 * norm: the real normalisation c_l (2 pi)^(3/2) (the factor (-i)^l is applied by the kernel).
 * scat_filters3d_gaussian: out_dev (n_scales, M, N, O) complex64 (filter_bank.py:39-97). */
int scat_filters3d_solid_harmonic(void* out_dev, const void* sigmas_dev, int32_t n_scales, int32_t l, double norm, int32_t M,
                                  int32_t N, int32_t O, void* stream);
int scat_filters3d_gaussian(void* out_dev, const void* sigmas_dev, int32_t n_scales, int32_t M, int32_t N, int32_t O,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCAT_B200_H */
