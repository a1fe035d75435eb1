/* lcgp_b200 -- C-ABI of the B200 (sm_100a) implementation of LCGP's emulator-fitting hot path.
 *
 * The reference (mosesyhc/LCGP, pure Python on TensorFlow) has no FFI: the seam is cut where its
 * Python loops over the latent components call dense TF linear algebra.  Each entry point below
 * names the reference code it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every array argument is a DEVICE pointer to row-major contiguous IEEE fp64 unless its name
 *     ends in _host; the caller (torch) owns every buffer including the workspace; the library
 *     allocates no device memory and keeps no state between calls (exceptions: a process-wide pool of
 *     side streams / events, and the explicit lcgp_plan objects)
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the call returns
 *     without synchronising (the *_host variants synchronise the stream before returning)
 *   - return value: 0 = ok; < 0 = invalid argument (LCGP_E_*); >= 1000 = 1000 + cudaError_t.
 *     Numerical failure (non-positive pivot) is reported per latent through `info`
 *     (LAPACK style: 1-based index of the first bad pivot, 0 = ok), never by the return value
 *   - dense factors are stored padded: np = LCGP_NB * ceil(n / LCGP_NB)
 */
#ifndef LCGP_B200_H
#define LCGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCGP_NB 128
#define LCGP_MAX_D 64
#define LCGP_N_STAGE_EVENTS 7

#define LCGP_E_ARG (-1)        /* null pointer / non-positive size */
#define LCGP_E_DIM (-2)        /* d > LCGP_MAX_D or np not a multiple of LCGP_NB */
#define LCGP_E_WORKSPACE (-3)  /* workspace too small */

#define LCGP_FLAG_GRAD 1            /* lcgp_nll_grad flags bit0 */
#define LCGP_FLAG_NO_LOOKAHEAD 256  /* bit8: keep each group's whole Cholesky on one stream */

/* Constant data of one rank's share of an emulator (all device pointers).
 * rep mode (lcgp.py:554-630):  X = x_unique_s, sr = sqrt(r), YR = ybar_s * r (or ybar * r),
 *   w_j = sum_i r_i ybar_ji^2, t = ybar_std (or ones), scale = 1/n, sum_log_r = sum_i log r_i
 * full mode (lcgp.py:635-666): X = x (standardised), sr = ones, YR = y, w_j = sum_i y_ji^2, t = ones,
 *   scale = 1, sum_log_r = 0 */
typedef struct lcgp_problem {
    int32_t n;       /* unique training inputs */
    int32_t d;       /* input dimension */
    int32_t p;       /* output dimension */
    int32_t q_loc;   /* latent components handled by this rank */
    int32_t include_host_terms; /* 1: add the O(p) data-fit / noise / -p/2 sum log r terms (one rank only) */
    int32_t n_emu;      /* 0 or 1: one emulator.  E > 1: the call evaluates E INDEPENDENT emulators of identical
                           (n, d, p) and q_loc / E latents each (BASELINE config 5: batched fits / restarts); every
                           array below then carries a leading emulator dimension (X: E x n x d, sr: E x n, YR: E x p x n,
                           w, t: E x p, phi: E x p x q_per, D: E x q_per), the parameters are lLmb (q_loc x d), lLmb0,
                           lnugGPs (q_loc) and lsigma2_p (E x p), and `out` holds E consecutive blocks of
                           lcgp_out_len(p, d, q_loc / E) doubles, one objective and gradient per emulator */
    double scale;       /* objective multiplier: 1/n (rep) or 1 (full) */
    double sum_log_r;
    const double* X;    /* n x d */
    const double* sr;   /* n */
    const double* YR;   /* p x n */
    const double* w;    /* p */
    const double* t;    /* p */
    const double* phi;  /* p x q_loc  (this rank's columns of the basis) */
    const double* D;    /* q_loc      (diag_D of this rank's latents) */
    const double* emu_consts; /* n_emu > 1: E x 2 device array (scale, sum_log_r) per emulator; otherwise NULL */
} lcgp_problem;

/* Length (in doubles) of the `out` vector of lcgp_nll_grad:
 *   [0]                      objective contribution of this rank (scaled)
 *   [1 .. p]                 d/d lsigma2 (p-vector, before the diag_error_structure segment-sum)
 *   [1+p .. ]                d/d lLmb (q_loc x d), d/d lLmb0 (q_loc), d/d lnugGPs (q_loc)   (constrained values)
 *   then                     logdet A_k (q_loc), b_k^T S_k b_k (q_loc)                       (diagnostics) */
size_t lcgp_out_len(int32_t p, int32_t d, int32_t q_loc);

/* Bytes of workspace needed by lcgp_nll_grad / lcgp_predict for a problem of this size. */
size_t lcgp_workspace_bytes(int32_t n, int32_t d, int32_t p, int32_t q_loc);
/* Same for a batch of n_emu emulators with q_per latents each (lcgp_problem.n_emu > 1). */
size_t lcgp_workspace_bytes_batched(int32_t n, int32_t d, int32_t p, int32_t q_per, int32_t n_emu);
/* Bytes of scratch needed by lcgp_predict for n0 test points. */
size_t lcgp_predict_scratch_bytes(int32_t n, int32_t q_loc, int32_t n0);

/* Objective and analytic gradient of this rank's latents.
 * Replaces the body of neglpost_rep (lcgp.py:554-630) / neglpost (lcgp.py:635-666) and the TF
 * autodiff driven from fit() (lcgp.py:537-540).  Parameters are the CONSTRAINED values:
 * lLmb (q_loc x d length-scales), lLmb0 (q_loc variances), lnugGPs (q_loc), lsigma2s expanded to
 * the p-vector of get_param (lcgp.py:515-532).
 * flags: bit0 = also compute the gradient (otherwise only out[0] and the diagnostics are valid);
 *        bits 4-7 = number of internal stream groups the latents are factored on (0 = default min(4, q_loc);
 *        1 = everything on `stream` (this also disables the look-ahead below), e.g. when the caller runs many
 *        small emulators on its own streams);
 *        bit8 (LCGP_FLAG_NO_LOOKAHEAD) = do not run the Cholesky panel chain on the library's internal
 *        high-priority stream (default: it overlaps the bulk of the previous trailing update).
 *        The internal streams are a process-wide pool per device: concurrent callers with the default group count
 *        serialise on it, so multi-threaded callers should pass 1 in bits 4-7 (lcgp_plan_* does); a call made while
 *        `stream` is being captured is detected and kept on `stream` alone.
 * After the call the workspace holds L_k, L_k^{-T}, alpha_k (= CinvMs, lcgp.py:781) and m_k (= mks,
 * lcgp.py:779) for lcgp_predict / lcgp_get_aux -- i.e. it also replaces
 * _compute_aux_predictive_quantities_rep (lcgp.py:728-803) and compute_aux_predictive_quantities
 * (lcgp.py:685-726).
 * stage_events: NULL or LCGP_N_STAGE_EVENTS cudaEvent_t recorded at: [0] start, [1] after the
 * kernel-matrix build, [2] after Cholesky, [3] after the triangular inverse, [4] before and [5] after the
 * fused A^-1 / gradient-contraction GEMM launch (the single largest kernel), [6] end.  Passing events
 * also makes the call join its internal stream groups between [2] and [3]. */
int lcgp_nll_grad(const lcgp_problem* prob, const double* lLmb, const double* lLmb0, const double* lnugGPs,
                  const double* lsigma2_p, void* workspace, size_t workspace_bytes, double* out,
                  int32_t* info, int32_t flags, void* const* stage_events, void* stream);

/* Same, with the parameters and results in HOST memory (pinned or pageable): copies the
 * parameters to the device, runs lcgp_nll_grad, copies `out` and `info` back and synchronises
 * the stream.  This is the call the Python LCGP.loss()/fit() path makes once per evaluation. */
int lcgp_nll_grad_host(const lcgp_problem* prob, const double* lLmb_host, const double* lLmb0_host,
                       const double* lnugGPs_host, const double* lsigma2_p_host, void* workspace,
                       size_t workspace_bytes, double* out_host, int32_t* info_host, int32_t flags,
                       void* const* stage_events, void* stream);

/* Evaluation plan: lcgp_nll_grad_host with fixed buffers, captured once as a CUDA graph (parameter upload, every
 * kernel, result download) and replayed with one launch per evaluation.  Evaluations of small problems are bound
 * by the kernel-launch rate -- above all when several host threads fit independent emulators on one GPU (BASELINE
 * config 5) -- not by the GPU.  params_host = [lLmb (q_loc x d) | lLmb0 (q_loc) | lnugGPs (q_loc) | lsigma2_p (p)],
 * out_host (lcgp_out_len doubles) and info_host (q_loc) must stay valid (pinned host memory) for the life of the
 * plan; the caller rewrites params_host before each lcgp_plan_run.  lcgp_plan_run synchronises `stream`.  The plan
 * is the one object the library allocates for the caller (host memory + the graph); lcgp_plan_destroy frees it.
 * A plan runs on ONE private stream of its own (the stream-group bits of `flags` are forced to 1; a private
 * stream because a stream under capture must not be touched by another thread, and frameworks hand out pooled
 * streams), ordered after everything already queued on `stream`; the first lcgp_plan_run evaluates launch by
 * launch and then captures.  If the driver refuses the capture the plan keeps running the same kernels launch by
 * launch (lcgp_plan_is_graph tells which). */
typedef struct lcgp_plan lcgp_plan;
int lcgp_plan_create(const lcgp_problem* prob, void* workspace, size_t workspace_bytes, const double* params_host,
                     double* out_host, int32_t* info_host, int32_t flags, lcgp_plan** plan);
int lcgp_plan_run(lcgp_plan* plan, void* stream);
int lcgp_plan_is_graph(const lcgp_plan* plan);
void lcgp_plan_destroy(lcgp_plan* plan);

/* Latent predictive mean and variance at n0 standardised test inputs, from the factor left in the
 * workspace by the last lcgp_nll_grad with the same parameters.
 * Replaces the k-loops of predict_rep (lcgp.py:883-897) and predict_full (lcgp.py:827-835).
 * same_inputs = 1 reproduces the reference's nugget-on-equal-inputs behaviour (covmat.py:46-53). */
int lcgp_predict(const lcgp_problem* prob, const double* lLmb, const double* lLmb0, const double* lnugGPs,
                 void* workspace, size_t workspace_bytes, const double* x0s, int32_t n0, int32_t same_inputs,
                 void* scratch, size_t scratch_bytes, double* ghat /* q_loc x n0 */, double* gvar /* q_loc x n0 */,
                 void* stream);

/* Output maps of predict_rep (lcgp.py:915-926) / predict_full (lcgp.py:840-848) on the device: from the latent
 * moments ghat, gvar (q x n0, ALL q latents) to the p outputs,
 *   predmean = Psi ghat, confvar = Psi^2 gvar (element-wise squares), predvar = confvar + noise_var,
 *   ypred = predmean * scale + shift, yconfvar = confvar * scale^2, ypredvar = predvar * scale^2    (each p x n0).
 * Psi (p x q) = phi * sqrt(sigma^2_used) per row, noise_var (p) = sigma^2_used, scale / shift (p, may be NULL = 1 / 0)
 * = ybar_std / ybar_mean (rep) or ystd / ymean (full).  All device pointers. */
int lcgp_predict_outputs(const double* Psi, const double* ghat, const double* gvar, const double* noise_var,
                         const double* scale, const double* shift, int32_t p, int32_t q, int32_t n0, double* ypred,
                         double* ypredvar, double* yconfvar, void* stream);

/* Full predictive covariance of the p outputs at each of n0 test points (predict_full with
 * return_fullcov=True, lcgp.py:850-857):
 *   out[i][a][b] = ystd[a] ystd[b] ( sum_k psi[k][a] gvar[k][i] psi[k][b] + [a == b] sig2[a] )
 * psi (q x p) = phi^T * sqrt(exp(lsigma2)) as in lcgp.py:838, gvar (q x n0) from lcgp_predict, sig2 = exp(lsigma2)
 * (p), ystd (p); out is n0 x p x p row-major.  All device pointers. */
int lcgp_predict_fullcov(const double* psi, const double* gvar, const double* sig2, const double* ystd, int32_t q,
                         int32_t p, int32_t n0, double* out, void* stream);

/* ---- one-off preprocessing on the device (constructor pipeline, lcgp.py:312-324, 358-395) ----------------
 * Replicate means: ybar[j][i] = mean of y[j][order[t]], t in [offsets[i], offsets[i+1]) -- `order` is the stable
 * argsort of the group ids of _group_unique_rows_np (lcgp.py:349-356), offsets its n+1 segment boundaries.
 * Replaces the Python loop of _compute_ybar_np (lcgp.py:358-367).  y: p x N, ybar: p x n. */
int lcgp_prep_segment_mean(const double* y, const int32_t* order, const int32_t* offsets, int32_t p, int32_t N,
                           int32_t n, double* ybar, void* stream);
/* out[j] = k-th smallest (0-based) element of row j of Y (p x m), or of |Y[j][:] - center[j]| when center != NULL:
 * tfp.stats.percentile(., 50, interpolation='nearest') of lcgp.py:317-318 / 388-389 with k = round((m-1)/2),
 * exact (radix select, no interpolation). */
int lcgp_prep_row_select(const double* Y, const double* center, int32_t p, int32_t m, int32_t k, double* out,
                         void* stream);
/* Ys = (Y - center_j) / spread_j (lcgp.py:320, 433); YR = Ys * r_i and w_j = sum_i r_i Ys[j][i]^2, the constant
 * arrays lcgp_problem takes.  r == NULL means r = 1; Ys, YR, w may each be NULL when not wanted. */
int lcgp_prep_standardize(const double* Y, const double* center, const double* spread, const double* r, int32_t p,
                          int32_t n, double* Ys, double* YR, double* w, void* stream);

/* Gradient of this rank's part of the objective with respect to its columns of the latent basis phi (p x q_loc),
 * diag_D treated as the function d_k = sum_j phi_jk^2 of it.  The reference keeps phi constant (lcgp.py:164), so this
 * has no reference counterpart; it exists for callers that train the basis (north-star "shared parameters").
 * Valid right after lcgp_nll_grad with the gradient bit at the same parameters (reads the factor, m_k and Z from the
 * workspace).  lsigma2_p, g_phi: device pointers. */
int lcgp_grad_phi(const lcgp_problem* prob, const double* lsigma2_p, void* workspace, size_t workspace_bytes,
                  double* g_phi, void* stream);

/* Sharded evaluation (one rank per GPU, latents k = rank mod world; the reference's q-loop, lcgp.py:605-624, has
 * no cross-k dependence): rewrites this rank's `out` / `info` of lcgp_nll_grad as the flat vector every rank
 * all-reduces once per evaluation,
 *   flat = [objective | d/d lsigma2 (p) | d/d lLmb (q x d) | d/d lLmb0 (q) | d/d lnugGPs (q) | failed latents]
 * in GLOBAL latent order (1 + p + q d + 2 q + 1 doubles), zero in the rows of latents other ranks own;
 * loc_of[k] (q int32, device) = local index of latent k on this rank, -1 if not owned.  The last slot counts this
 * rank's latents whose Cholesky failed, so that after the all-reduce every rank raises together. */
int lcgp_pack_sharded(const double* out, const int32_t* info, const int32_t* loc_of, int32_t p, int32_t d, int32_t q,
                      int32_t q_loc, double* flat, void* stream);

/* Copies alpha (CinvMs) and m (mks), each q_loc x n, out of the workspace. */
int lcgp_get_aux(const lcgp_problem* prob, void* workspace, size_t workspace_bytes, double* CinvMs, double* mks,
                 void* stream);

/* A^{-1} of latent k (n x n, dense, symmetric) rebuilt from the factor in the workspace; used to
 * materialise the reference's Tks / Ths on request (lcgp.py:783-788, 709-715). */
int lcgp_get_Ainv(const lcgp_problem* prob, void* workspace, size_t workspace_bytes, int32_t k, double* Ainv,
                  void* stream);

/* Matern32(x1, x2, llmb, llmb0, lnug) of covmat.py:5-55 (diag_only=False branch): out is n1 x n2.
 * same_inputs: the caller's evaluation of `x1.shape == x2.shape and all(x1 == x2)` (covmat.py:46-49). */
int lcgp_kernel_matrix(const double* x1, int32_t n1, const double* x2, int32_t n2, int32_t d, const double* llmb,
                       const double* llmb0, const double* lnug, int32_t same_inputs, double* out, void* stream);

/* ---- stage entry points (unit parity tests and the stage rooflines of bench.py) ---- */

/* A_k = I + d_k (C_k o sqrt r sqrt r^T) (lcgp.py:616) for `batch` latents into padded np x np
 * buffers F (lower triangle + diagonal; identity in the pad). */
int lcgp_build_A(const double* X, const double* sr, int32_t n, int32_t d, const double* lLmb, const double* lLmb0,
                 const double* lnugGPs, const double* D, int32_t batch, double* F, int32_t np, void* stream);

/* In-place batched Cholesky of the lower triangles of `batch` padded np x np matrices
 * (tf.linalg.cholesky, lcgp.py:617/775).  DL/DU receive the inverses of the NB x NB diagonal
 * blocks and their transposes (batch x np/NB x NB x NB each); logdet_part (batch x np/NB, may be
 * NULL) receives sum log L_ii per diagonal block.
 * scratch: lcgp_potrf_scratch_bytes(np, batch) bytes of device memory (tile-dependency flags of the persistent
 * left-looking kernel, ONE launch per call); NULL selects the launch-per-block-column path (as does the
 * environment variable LCGP_POTRF=panels).  info[b] < 0 reports an internal dependency wait that timed out. */
size_t lcgp_potrf_scratch_bytes(int32_t np, int32_t batch);
int lcgp_potrf_batched(double* F, int32_t np, int32_t batch, double* DL, double* DU, double* logdet_part,
                       int32_t* info, void* scratch, size_t scratch_bytes, void* stream);

/* Cholesky AND triangular inverse in ONE persistent launch (what lcgp_nll_grad runs): as lcgp_potrf_batched with
 * scratch, and on return the strictly-upper NB-blocks of F also hold L^{-T} (what lcgp_trtri_batched computes).  The
 * inverse tiles are queued one block column behind the factorisation and fill the SMs its dependency chain leaves
 * idle.  (tf.linalg.cholesky + the identity solves / tf.linalg.inv of lcgp.py:775-787.) */
int lcgp_potrf_trtri_batched(double* F, int32_t np, int32_t batch, double* DL, double* DU, double* logdet_part,
                             int32_t* info, void* scratch, size_t scratch_bytes, void* stream);

/* Blocked triangular inverse: fills the strictly-upper NB-blocks of F with L^{-T}.
 * scratch: lcgp_trtri_scratch_bytes(np, batch). */
size_t lcgp_trtri_scratch_bytes(int32_t np, int32_t batch);
int lcgp_trtri_batched(double* F, int32_t np, int32_t batch, const double* DL, const double* DU, void* scratch,
                       size_t scratch_bytes, void* stream);

/* Library / build identification. */
const char* lcgp_version(void);

/* Number of CUDA kernels this library has launched in the calling process so far (all entry points, all
 * threads).  bench.py reports the difference across its timed region as "gpu_launches". */
unsigned long long lcgp_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LCGP_B200_H */
