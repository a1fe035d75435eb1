// Product Matern-3/2 kernel-matrix construction (reference: src/lcgp/covmat.py:5-55).
//
//   C0[i][j] = prod_m (1 + S_m) * exp(-sum_m S_m),   S_m = |x1[i][m]/ell_m - x2[j][m]/ell_m|
//   C        = s0 * ((1 - nu) C0 + nu * I [only when x1 and x2 are the same point set]),  nu = lnug/(1+lnug)
//
// build_A writes A_k = I + d_k (C_k o sqrt(r) sqrt(r)^T) (lcgp.py:616) straight into the padded
// factor buffer -- C_k itself never exists in HBM -- lower-triangular 64x64 tiles only, X tiles
// staged (pre-divided by the length-scales) in shared memory, 32 B of contiguous output per
// thread.  matern_rect is the general rectangular form used for the public Matern32() operator
// and for the prediction cross-covariances.
#include "common.cuh"
#include "lcgp_internal.h"

namespace lcgp {

constexpr int MT = 64;  // tile edge

__global__ void __launch_bounds__(256)
build_A_kernel(const double* __restrict__ X, const double* __restrict__ sr, int n, int d, int np,
               KernelParams kp, double* __restrict__ F, size_t fstride) {
    extern __shared__ __align__(16) double sm[];
    double* xi = sm;           // [d][64]
    double* xj = sm + d * MT;  // [d][64]
    const int k = blockIdx.y;
    {   // this latent's emulator
        const int e = k / kp.q_per;
        X += (size_t)e * n * d;
        sr += (size_t)e * n;
    }
    int TI, TJ;
    tri_decode(blockIdx.x, TI, TJ);
    const int i0 = TI * MT, j0 = TJ * MT;
    const double* ell = kp.ell + (size_t)k * d;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < MT * d; idx += 256) {
        const int m = idx / MT, r = idx % MT;
        const int gi = i0 + r, gj = j0 + r;
        const double l = ell[m];
        xi[idx] = gi < n ? X[(size_t)gi * d + m] / l : 0.0;
        xj[idx] = gj < n ? X[(size_t)gj * d + m] / l : 0.0;
    }
    __syncthreads();

    const int ty = tid >> 4, tx = tid & 15;
    double P[4][4], V[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { P[a][b] = 1.0; V[a][b] = 0.0; }
    for (int m = 0; m < d; ++m) {
        double av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = xi[m * MT + ty + 16 * a];
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = xj[m * MT + 4 * tx + b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double S = fabs(av[a] - bv[b]);
                P[a][b] *= (1.0 + S);
                V[a][b] -= S;
            }
    }
    const double s0 = kp.s0[k], lnug = kp.lnug[k], dk = kp.D[k];
    const double nu = lnug / (1.0 + lnug);
    double srj[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) { const int gj = j0 + 4 * tx + b; srj[b] = gj < n ? sr[gj] : 0.0; }
    double* Fk = F + (size_t)k * fstride;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int gi = i0 + ty + 16 * a;
        const double sri = gi < n ? sr[gi] : 0.0;
        double out[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int gj = j0 + 4 * tx + b;
            double v;
            if (gi < n && gj < n) {
                const double c0 = P[a][b] * exp(V[a][b]);
                double c = (1.0 - nu) * c0;
                if (gi == gj) c += nu;
                c *= s0;
                v = dk * ((c * srj[b]) * sri);
                if (gi == gj) v += 1.0;
            } else {
                v = (gi == gj) ? 1.0 : 0.0;
            }
            out[b] = v;
        }
        double2* dst = reinterpret_cast<double2*>(Fk + (size_t)gi * np + j0 + 4 * tx);
        dst[0] = make_double2(out[0], out[1]);
        dst[1] = make_double2(out[2], out[3]);
    }
}

cudaError_t launch_build_A(const double* X, const double* sr, int n, int d, int np, KernelParams kp,
                           double* F, size_t fstride, int batch, cudaStream_t stream) {
    const int nt = np / MT;
    const size_t smem = sizeof(double) * 2 * d * MT;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(build_A_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    note_launch(); build_A_kernel<<<dim3(nt * (nt + 1) / 2, batch), 256, smem, stream>>>(X, sr, n, d, np, kp, F, fstride);
    return cudaGetLastError();
}

// out[b][t][i] = colscale[i] * s0_b * ((1-nu_b) C0(x1_t, x2_i) + nu_b [same && t == i])  for t < n1, i < n2;
// zero in the pad (t in [n1, rows_out), i in [n2, cols_out)).
__global__ void __launch_bounds__(256)
matern_rect_kernel(const double* __restrict__ x1, int n1, const double* __restrict__ x2, int n2, int d,
                   const double* __restrict__ ellv, const double* __restrict__ s0v,
                   const double* __restrict__ lnugv, int same, const double* __restrict__ colscale,
                   double* __restrict__ out, int ld_out, int rows_out, int cols_out, size_t out_stride) {
    extern __shared__ __align__(16) double sm[];
    double* xi = sm;
    double* xj = sm + d * MT;
    const int k = blockIdx.z;
    const int i0 = blockIdx.y * MT, j0 = blockIdx.x * MT;
    const double* ell = ellv + (size_t)k * d;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < MT * d; idx += 256) {
        const int m = idx / MT, r = idx % MT;
        const int gi = i0 + r, gj = j0 + r;
        const double l = ell[m];
        xi[idx] = gi < n1 ? x1[(size_t)gi * d + m] / l : 0.0;
        xj[idx] = gj < n2 ? x2[(size_t)gj * d + m] / l : 0.0;
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    double P[4][4], V[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { P[a][b] = 1.0; V[a][b] = 0.0; }
    for (int m = 0; m < d; ++m) {
        double av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = xi[m * MT + ty + 16 * a];
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = xj[m * MT + tx + 16 * b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double S = fabs(av[a] - bv[b]);
                P[a][b] *= (1.0 + S);
                V[a][b] -= S;
            }
    }
    const double s0 = s0v[k], lnug = lnugv[k];
    const double nu = lnug / (1.0 + lnug);
    double* o = out + (size_t)k * out_stride;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int gi = i0 + ty + 16 * a;
        if (gi >= rows_out) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int gj = j0 + tx + 16 * b;
            if (gj >= cols_out) continue;
            double v = 0.0;
            if (gi < n1 && gj < n2) {
                double c = (1.0 - nu) * (P[a][b] * exp(V[a][b]));
                if (same && gi == gj) c += nu;
                v = s0 * c;
                if (colscale) v *= colscale[gj];
            }
            o[(size_t)gi * ld_out + gj] = v;
        }
    }
}

cudaError_t launch_matern_rect(const double* x1, int n1, const double* x2, int n2, int d, const double* ell,
                               const double* s0, const double* lnug, int same, const double* colscale,
                               double* out, int ld_out, int rows_out, int cols_out, int batch,
                               size_t out_stride, cudaStream_t stream) {
    const size_t smem = sizeof(double) * 2 * d * MT;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(matern_rect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((cols_out + MT - 1) / MT, (rows_out + MT - 1) / MT, batch);
    note_launch(); matern_rect_kernel<<<grid, 256, smem, stream>>>(x1, n1, x2, n2, d, ell, s0, lnug, same, colscale, out,
                                                    ld_out, rows_out, cols_out, out_stride);
    return cudaGetLastError();
}

}  // namespace lcgp
