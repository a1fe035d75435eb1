// Batched predictive mean / variance of the latent GPs from the single factor
// (reference: lcgp.py:883-897 predict_rep, :827-835 predict_full; identities SURVEY.md A.6):
//   c0    = Matern32(x0, Xtrain)                       (n0 x n, nugget only under the equality quirk)
//   ghat  = c0 alpha
//   gvar  = s0 - d_k || L^{-1} (sqrt r o c0^T) ||_col^2 = s0 - d_k sum_i ( (c0s U)[t][i] )^2,  c0s = c0 o sqrt r
// The cross-covariance tile matrix c0s is built once per latent into scratch; the product c0s * U
// runs on the DMMA GEMM (N-major B operand, triangular K range) and its epilogue reduces the
// squared tile to row sums, so the n0 x n product is never stored.
#include "gemm_dmma.cuh"
#include "lcgp_internal.h"

namespace lcgp {

struct PredictParams {
    FactorView v;
    const double* c0s;   // [batch][n0p][np]
    size_t cstride;
    double* part;        // [batch][nb][n0p]  row sums of squares per column block
    int n0p;
};

struct PredictJob {
    static constexpr bool kBNMajor = true;
    // Skipped MMA steps: the last K block of operand B is the upper-triangular DU_Ib (column c needs k <= c); for the last
    // column block the columns >= n are padding and so are the last K steps (c0s is zero there).
    static constexpr bool kSkips = true;
    typedef PredictParams Params;
    int kb0, kb1, Tb, Ib, tail_skip;
    __device__ bool init(const Params& p) {
        Tb = blockIdx.x / p.v.nb;
        Ib = blockIdx.x % p.v.nb;
        kb0 = 0;
        kb1 = Ib + 1;
        tail_skip = Ib == p.v.nb - 1 ? KSTEPS - (last_block_rows(p.v) + BK - 1) / BK : 0;
        return true;
    }
    __device__ int head_steps(const Params& p) const { return (Ib == p.v.nb - 1 && last_block_rows(p.v) <= NB - BK) ? kAllSteps : 0; }
    __device__ int tail_steps(const Params&) const { return KSTEPS - 1 - tail_skip > 0 ? KSTEPS - 1 - tail_skip : 0; }
    __device__ StepMask mask(const Params& p, int it) const {
        const int itl = it - Ib * KSTEPS;                    // >= 0 inside the last K block
        return StepMask{NB, itl > 0 ? BK * itl : 0, Ib == p.v.nb - 1 ? last_block_rows(p.v) : NB};
    }
    __device__ TileRef a_ref(const Params&, int kb) const { return TileRef{SRC_C, Tb * NB, kb * NB}; }
    __device__ TileRef b_ref(const Params&, int kb) const {   // N-major: rows = k, cols = i
        return kb == Ib ? TileRef{SRC_DU, Ib * NB, 0} : TileRef{SRC_F, kb * NB, Ib * NB};
    }
    template <class Coord>
    __device__ void epilogue(const Params& p, double (&acc)[8][4][2], double* smem, const Coord& wc) const {
        double* rs = smem;  // [NB][4]
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
            double s = 0.0;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) s += acc[mi][ni][0] * acc[mi][ni][0] + acc[mi][ni][1] * acc[mi][ni][1];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (wc.t == 0) rs[wc.row(mi) * 4 + wc.wn] = s;
        }
        __syncthreads();
        if (threadIdx.x < NB) {
            const int r = threadIdx.x;
            const double s = (rs[r * 4 + 0] + rs[r * 4 + 1]) + (rs[r * 4 + 2] + rs[r * 4 + 3]);
            p.part[((size_t)blockIdx.y * p.v.nb + Ib) * p.n0p + (size_t)Tb * NB + r] = s;
        }
    }
};

__global__ void __launch_bounds__(256)
predict_finish_kernel(int np, int nb, int n0, int n0p, const double* __restrict__ c0s, size_t cstride,
                      const double* __restrict__ atil, const double* __restrict__ part,
                      const double* __restrict__ s0v, const double* __restrict__ Dk,
                      double* __restrict__ ghat, double* __restrict__ gvar) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + warp;
    const int k = blockIdx.y;
    if (t >= n0) return;
    const double* row = c0s + (size_t)k * cstride + (size_t)t * np;
    const double* at = atil + (size_t)k * np;
    double s = 0.0;
    for (int i = lane; i < np; i += 32) s += row[i] * at[i];   // c0s[t][i] * atil[i] = c0[t][i] * alpha[i]
    s = warp_sum(s);
    double q = 0.0;
    for (int b = lane; b < nb; b += 32) q += part[((size_t)k * nb + b) * n0p + t];
    q = warp_sum(q);
    if (lane == 0) {
        ghat[(size_t)k * n0 + t] = s;
        gvar[(size_t)k * n0 + t] = s0v[k] - Dk[k] * q;
    }
}

// ---- output maps (lcgp.py:915-926 rep, :840-848 full): from the latent moments to the p outputs, on the device ----
//   predmean = Psi ghat,  confvar = Psi^2 gvar,  predvar = confvar + noise_var
//   ypred = predmean * scale + shift,  yconfvar = confvar * scale^2,  ypredvar = predvar * scale^2
// One thread per test point and 8 outputs; Psi rows of the block in shared memory; ghat / gvar columns stream
// through L2 (q x n0 each, re-read by every block of outputs).
constexpr int PO_J = 8;
__global__ void __launch_bounds__(128)
predict_outputs_kernel(int p, int q, int n0, const double* __restrict__ Psi, const double* __restrict__ ghat,
                       const double* __restrict__ gvar, const double* __restrict__ noise_var,
                       const double* __restrict__ scale, const double* __restrict__ shift,
                       double* __restrict__ ypred, double* __restrict__ ypredvar, double* __restrict__ yconfvar) {
    extern __shared__ double ps[];      // [PO_J][q]
    const int j0 = blockIdx.y * PO_J;
    for (int idx = threadIdx.x; idx < PO_J * q; idx += blockDim.x) {
        const int jj = idx / q, k = idx % q;
        ps[idx] = (j0 + jj < p) ? Psi[(size_t)(j0 + jj) * q + k] : 0.0;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0) return;
    double m[PO_J], c[PO_J];
#pragma unroll
    for (int jj = 0; jj < PO_J; ++jj) m[jj] = c[jj] = 0.0;
    for (int k = 0; k < q; ++k) {
        const double gh = ghat[(size_t)k * n0 + i], gv = gvar[(size_t)k * n0 + i];
#pragma unroll
        for (int jj = 0; jj < PO_J; ++jj) {
            const double w = ps[jj * q + k];
            m[jj] = fma(w, gh, m[jj]);
            c[jj] = fma(w * w, gv, c[jj]);
        }
    }
#pragma unroll
    for (int jj = 0; jj < PO_J; ++jj) {
        const int j = j0 + jj;
        if (j >= p) break;
        const double sc = scale ? scale[j] : 1.0, sh = shift ? shift[j] : 0.0;
        const size_t o = (size_t)j * n0 + i;
        ypred[o] = m[jj] * sc + sh;
        yconfvar[o] = c[jj] * (sc * sc);
        ypredvar[o] = (c[jj] + noise_var[j]) * (sc * sc);
    }
}

cudaError_t launch_predict_outputs(const double* Psi, const double* ghat, const double* gvar, const double* noise_var,
                                   const double* scale, const double* shift, int p, int q, int n0, double* ypred,
                                   double* ypredvar, double* yconfvar, cudaStream_t stream) {
    const size_t smem = sizeof(double) * PO_J * q;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(predict_outputs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    note_launch();
    predict_outputs_kernel<<<dim3((n0 + 127) / 128, (p + PO_J - 1) / PO_J), 128, smem, stream>>>(
        p, q, n0, Psi, ghat, gvar, noise_var, scale, shift, ypred, ypredvar, yconfvar);
    return cudaGetLastError();
}

cudaError_t predict_latents(const FactorView& v, int n, int d, const double* X, const double* sr, KernelParams kp,
                            const double* atil, const double* x0s, int n0, int same, double* scratch,
                            int q_loc, double* ghat, double* gvar, cudaStream_t stream) {
    const int n0p = round_up(n0, NB);
    double* c0s = scratch;
    const size_t cstride = (size_t)n0p * v.np;
    double* part = scratch + (size_t)q_loc * cstride;
    cudaError_t e = launch_matern_rect(x0s, n0, X, n, d, kp.ell, kp.s0, kp.lnug, same, sr, c0s, v.np, n0p, v.np,
                                       q_loc, cstride, stream);
    if (e != cudaSuccess) return e;
    PredictParams p{v, c0s, cstride, part, n0p};
    GemmCtx ctx;
    {
        GemmSrcs srcs;
        int rows[NSRC];
        factor_srcs(v, srcs, rows);
        srcs.base[SRC_C] = c0s; srcs.ld[SRC_C] = v.np; srcs.bstride[SRC_C] = cstride; rows[SRC_C] = n0p;
        if ((e = gemm_make_ctx(ctx, srcs, rows, q_loc)) != cudaSuccess) return e;
    }
    e = gemm_launch<PredictJob>(ctx, p, dim3((n0p / NB) * v.nb, q_loc, 1), stream);
    if (e != cudaSuccess) return e;
    note_launch(); predict_finish_kernel<<<dim3((n0 + 7) / 8, q_loc), 256, 0, stream>>>(v.np, v.nb, n0, n0p, c0s, cstride, atil, part,
                                                                          kp.s0, kp.D, ghat, gvar);
    return cudaGetLastError();
}

}  // namespace lcgp
