// Batched blocked FP64 Cholesky (right-looking, NB = 128) and blocked triangular inverse.
//
//   per block column jb:   diag kernel  : L_jj = chol(A_jj), DL = inv(L_jj), DU = DL^T, logdet part
//                          TrsmJob      : P = A[jb+1:, jb] * DL^T                 (DMMA GEMM)
//                          SyrkJob      : A[I][J] -= P_I P_J^T, jb < J <= I       (DMMA GEMM)
//   TRTRI: bottom-up binary merges of already-inverted diagonal ranges, two DMMA GEMMs per level.
#include "gemm_dmma.cuh"
#include "lcgp_internal.h"

namespace lcgp {

// ------------------------------------------------------------------------------------------
// Diagonal-block kernel: one CTA per matrix.  The block lives in shared memory as S[128][129]:
// the lower triangle is the running Schur complement / L, the strictly-upper triangle holds the
// transpose of the running inverse Z (Z[i][j] at S[j][i]); zd[] is the diagonal of Z.  Column
// step c scales column c by 1/L_cc (which finalises both column c of L and row c of Z) and then
// applies one rank-1 update that serves the Cholesky trailing block and the forward elimination
// of the identity at the same time.
// ------------------------------------------------------------------------------------------
constexpr int DIAG_THREADS = 256;
constexpr int DPITCH = NB + 1;
constexpr size_t DIAG_SMEM = sizeof(double) * (NB * DPITCH + 3 * NB + 8);

__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_kernel(FactorView v, double* DLw, double* DUw, int jb, double* logdet_part /* [batch][nb] */,
                  int* info /* [batch] */) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;
    double* cv = sm + NB * DPITCH;   // scaled column c (cv[c] = 1/L_cc)
    double* zd = cv + NB;            // diagonal of the inverse
    double* ldg = zd + NB;           // diagonal of L
    double* red = ldg + NB;
    const int tid = threadIdx.x;
    const int bz = blockIdx.x;
    double* blk = v.F + (size_t)bz * v.fstride + (size_t)jb * NB * v.np + (size_t)jb * NB;

    for (int idx = tid; idx < NB * NB; idx += DIAG_THREADS) {
        const int r = idx >> 7, c = idx & (NB - 1);
        S[r * DPITCH + c] = (c <= r) ? blk[(size_t)r * v.np + c] : 0.0;
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    for (int c = 0; c < NB; ++c) {
        // S[c][c] is final after the barrier that closed step c-1 and is never written during
        // step c (the diagonal of L is kept in ldg[] until write-back), so no barrier is needed
        // between this read and the column scaling below.
        const double piv = S[c * DPITCH + c];
        if (!(piv > 0.0) && tid == 0) atomicCAS(&info[bz], 0, jb * NB + c + 1);
        const double dd = sqrt(piv);
        const double invd = 1.0 / dd;
        if (tid < NB) {
            const int r = tid;
            if (r == c) {
                ldg[c] = dd;
                cv[c] = invd;
                zd[c] = invd;
            } else {
                const double val = S[r * DPITCH + c] * invd;
                S[r * DPITCH + c] = val;
                cv[r] = val;
            }
        }
        __syncthreads();
        // rank-1 update: for i > c, rho <= i :  T(i,rho) -= cv[i] * cv[rho]
        //   T(i,rho) = S[i][rho] if rho > c (Schur complement),  S[rho][i] if rho <= c (inverse^T)
        for (int i = c + 1 + warp; i < NB; i += DIAG_THREADS / 32) {
            const double li = cv[i];
            for (int rho = lane; rho <= i; rho += 32) {
                const int off = (rho > c) ? (i * DPITCH + rho) : (rho * DPITCH + i);
                S[off] -= li * cv[rho];
            }
        }
        __syncthreads();
    }

    // write back L (lower triangle only), DL = Z, DU = Z^T
    double* dl = DLw + (size_t)bz * v.dstride + (size_t)jb * NB * NB;
    double* du = DUw + (size_t)bz * v.dstride + (size_t)jb * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += DIAG_THREADS) {
        const int r = idx >> 7, c = idx & (NB - 1);
        if (c <= r) blk[(size_t)r * v.np + c] = (c == r) ? ldg[r] : S[r * DPITCH + c];
        dl[idx] = (c < r) ? S[c * DPITCH + r] : (c == r ? zd[r] : 0.0);
        du[idx] = (c > r) ? S[r * DPITCH + c] : (c == r ? zd[r] : 0.0);
    }
    // log det part: sum_c log L_cc = -sum_c log zd[c]
    double lg = (tid < NB) ? -log(zd[tid]) : 0.0;
    lg = block_sum(lg, red);
    if (tid == 0 && logdet_part) logdet_part[(size_t)bz * v.nb + jb] = lg;
}

static cudaError_t diag_configure() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM);
    if (e == cudaSuccess) done = true;
    return e;
}

cudaError_t potrf_batched(const FactorView& v, double* DLw, double* DUw, int batch, double* logdet_part,
                          int* info, cudaStream_t stream) {
    cudaError_t e = diag_configure();
    if (e != cudaSuccess) return e;
    for (int jb = 0; jb < v.nb; ++jb) {
        potrf_diag_kernel<<<batch, DIAG_THREADS, DIAG_SMEM, stream>>>(v, DLw, DUw, jb, logdet_part, info);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const int T = v.nb - jb - 1;
        if (T == 0) break;
        TrsmJob::Params tp{v, jb};
        e = gemm_launch<TrsmJob>(tp, dim3(T, batch, 1), stream);
        if (e != cudaSuccess) return e;
        SyrkJob::Params sp{v, jb};
        e = gemm_launch<SyrkJob>(sp, dim3(T * (T + 1) / 2, batch, 1), stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Scratch blocks (NB x NB each) needed per matrix by trtri_batched.
size_t trtri_scratch_blocks(int nb) {
    size_t mx = 0;
    for (int s = 1; s < nb; s *= 2) {
        const size_t merges = (size_t)(nb + 2 * s - 1) / (2 * s);
        mx = merges * s * s > mx ? merges * s * s : mx;
    }
    return mx;
}

cudaError_t trtri_batched(const FactorView& v, double* scratch, size_t tstride, int batch, cudaStream_t stream) {
    for (int s = 1; s < v.nb; s *= 2) {
        const int merges = (v.nb + 2 * s - 1) / (2 * s);
        TrtriParams p{v, scratch, tstride, s};
        dim3 grid(s * s, batch, merges);
        cudaError_t e = gemm_launch<TrtriG1Job>(p, grid, stream);
        if (e != cudaSuccess) return e;
        e = gemm_launch<TrtriG2Job>(p, grid, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace lcgp
