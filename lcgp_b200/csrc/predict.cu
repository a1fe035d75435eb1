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
    typedef PredictParams Params;
    int kb0, kb1, Tb, Ib;
    __device__ bool init(const Params& p) {
        Tb = blockIdx.x / p.v.nb;
        Ib = blockIdx.x % p.v.nb;
        kb0 = 0;
        kb1 = Ib + 1;
        return true;
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
