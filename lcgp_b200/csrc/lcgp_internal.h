// Internal (C++) declarations shared between the translation units of liblcgp_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace lcgp {

struct FactorView;
struct GemmSrcs;

// potrf.cu
// Optional look-ahead resources of one Cholesky call: a (high-priority) stream for the serial panel chain and
// two events (disable-timing) it may record freely.  panel == nullptr -> everything runs on `stream`.
struct Lookahead {
    cudaStream_t panel = nullptr;
    cudaEvent_t ev_panel = nullptr, ev_bulk = nullptr;
};
// sync != nullptr (potrf_pll_sync_ints(nb, batch) ints of device scratch) selects the persistent left-looking kernel
// of potrf_pll.cu (one launch per call) unless LCGP_POTRF=panels; otherwise the launch-per-block-column path below.
cudaError_t potrf_batched(const FactorView& v, double* DLw, double* DUw, int batch, double* logdet_part,
                          int* info, int panel_width, cudaStream_t stream, const Lookahead& la = Lookahead(),
                          int* sync = nullptr, bool fused_inverse = false);
// LCGP_FUSE_TRTRI (default 1): with the persistent kernel the triangular inverse is computed by the same launch
bool potrf_fuse_trtri();
bool potrf_use_pll();                     // the persistent kernel is enabled at all (LCGP_POTRF != panels, TMA engine)
bool potrf_use_pll(int nb, int batch);    // ... and selected for this batch of matrices
// potrf_pll.cu
size_t potrf_pll_sync_ints(int nb, int batch);
// fused_inverse: the same launch also fills the strictly-upper blocks with U = L^-T (what trtri_batched computes)
cudaError_t potrf_pll(const FactorView& v, double* DLw, double* DUw, int batch, double* logdet_part, int* info,
                      int* sync, cudaStream_t stream, bool fused_inverse = false);
size_t trtri_scratch_blocks(int nb);
void factor_srcs(const FactorView& v, GemmSrcs& s, int rows[]);
cudaError_t trtri_batched(const FactorView& v, double* scratch, size_t tstride, int batch, cudaStream_t stream);

// Per-latent kernel hyper-parameters, device arrays of length q_loc (ell: q_loc x d).
struct KernelParams {
    const double* ell;
    const double* s0;
    const double* lnug;
    const double* D;  // diag_D of the local latents
    int q_per;        // latents per emulator: latent k belongs to emulator k / q_per, whose X / sr / YR / ... follow
                      // those of emulator 0 at the natural strides (one emulator: q_per = q_loc)
};

// matern.cu
cudaError_t launch_matern_rect(const double* x1, int n1, const double* x2, int n2, int d, const double* ell,
                               const double* s0, const double* lnug, int same, const double* colscale,
                               double* out, int ld_out, int rows_out, int cols_out, int batch,
                               size_t out_stride, cudaStream_t stream);
cudaError_t launch_build_A(const double* X, const double* sr, int n, int d, int np, KernelParams kp,
                           double* F, size_t fstride, int batch, cudaStream_t stream);

// solve_grad.cu
struct SolveArgs {
    int n, d, p, np, nb, q_loc;
    const double* X;     // n x d
    const double* sr;    // n
    const double* B;     // q_loc x np   (b_k, zero padded)
    KernelParams kp;
    double* alpha;       // q_loc x np   out
    double* mk;          // q_loc x np   out
    double* atil;        // q_loc x np   out  (A^{-1}(b / sqrt r))
    double* gemv_part;   // q_loc x nb x np scratch
    double* quad;        // q_loc out: b^T m
};
cudaError_t solve_alpha(const FactorView& v, const SolveArgs& a, cudaStream_t stream);
// g_kern: kernel-gradient part of emulator 0's out block; emulator e's follows at e * out_stride doubles
cudaError_t contract_grad(const FactorView& v, const SolveArgs& a, double* tile_part /* q_loc x ntiles x (d+2) */,
                          double* g_kern, size_t out_stride, cudaEvent_t ev_before, cudaEvent_t ev_after,
                          cudaStream_t stream);

cudaError_t grad_phi(const FactorView& v, int n, int p, int q_loc, double scale, const double* sr, const double* mk,
                     const double* lsig, const double* t, const double* phi, const double* D, const double* Z,
                     double* part /* 2 * q_loc * nb scratch */, double* g_phi /* p x q_loc */, cudaStream_t stream);

// predict.cu
cudaError_t predict_latents(const FactorView& v, int n, int d, const double* X, const double* sr, KernelParams kp,
                            const double* atil, const double* x0s, int n0, int same, double* scratch,
                            int q_loc, double* ghat, double* gvar, cudaStream_t stream);

cudaError_t launch_predict_outputs(const double* Psi, const double* ghat, const double* gvar, const double* noise_var,
                                   const double* scale, const double* shift, int p, int q, int n0, double* ypred,
                                   double* ypredvar, double* yconfvar, cudaStream_t stream);

// fullcov.cu
cudaError_t launch_fullcov(const double* psi, const double* gvar, const double* sig2, const double* sv, int q, int p,
                           int n0, double* out, cudaStream_t stream);

// prep.cu
cudaError_t prep_segment_mean(const double* y, const int* order, const int* off, int p, int N, int n, double* ybar,
                              cudaStream_t st);
cudaError_t prep_row_select(const double* Y, const double* center, int p, int m, int k, double* out, cudaStream_t st);
cudaError_t prep_standardize(const double* Y, const double* c, const double* s, const double* r, int p, int n,
                             double* Ys, double* YR, double* w, cudaStream_t st);

}  // namespace lcgp
