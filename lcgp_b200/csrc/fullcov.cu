// Full predictive covariance of the outputs at each test point (lcgp.py:850-857, predict_full with
// return_fullcov=True):
//     out[i][a][b] = sv_a sv_b ( sum_k psi[k][a] gvar[k][i] psi[k][b]  +  [a == b] sig2_a )
// i.e. a batch of n0 rank-q updates of a diagonal, each a p x p matrix (32 MB per test point at p = 2000).
// The output write is the algorithmic traffic (8 n0 p^2 bytes); with q = 32 the 2 q flop per element are also
// close to the FP64 tensor rate per SM, so the products run on DMMA (m8n8k4) from shared-memory tiles of
// psi^T that stay resident while the CTA walks over its test points.
#include <atomic>

#include "common.cuh"
#include "lcgp_internal.h"

namespace lcgp {

constexpr int FC_KC = 64;            // latents per shared-memory chunk
constexpr int FC_PITCH = FC_KC + 4;  // 68 = 4 mod 16: conflict-free LDS.64 fragment loads
constexpr int FC_IPC = 16;           // most test points one CTA walks over
constexpr size_t FC_SMEM = sizeof(double) * (2 * NB * FC_PITCH + FC_IPC * FC_KC);

// As[a][k] = sv[a0 + a] psi[k0 + k][a0 + a]  (zero outside q x p); same for Bs with b0
__device__ __forceinline__ void fc_load_tile(double* dst, const double* __restrict__ psi, const double* __restrict__ sv,
                                             int q, int p, int r0, int k0) {
    for (int idx = threadIdx.x; idx < NB * FC_KC; idx += GEMM_THREADS) {
        const int a = idx & (NB - 1), k = idx >> 7;   // consecutive threads -> consecutive a: coalesced reads
        const int ga = r0 + a, gk = k0 + k;
        dst[a * FC_PITCH + k] = (ga < p && gk < q) ? psi[(size_t)gk * p + ga] * sv[ga] : 0.0;
    }
}

// gvs[ii][k] = sqrt(gvar[k0 + k][i0 + ii]) for the CTA's test points.  Both operands carry sqrt(gvar)
// (CH = sqrt(gvar) psi, lcgp.py:851-853): out[a][b] and out[b][a] are then sums of the same products in the
// same order, i.e. the result is exactly symmetric.
__device__ __forceinline__ void fc_load_gv(double* gvs, const double* __restrict__ gvar, int q, int n0, int k0,
                                           int i0, int cnt) {
    for (int idx = threadIdx.x; idx < cnt * FC_KC; idx += GEMM_THREADS) {
        const int ii = idx / FC_KC, k = idx % FC_KC;
        gvs[idx] = (k0 + k < q) ? sqrt(gvar[(size_t)(k0 + k) * n0 + i0 + ii]) : 0.0;
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
fullcov_kernel(const double* __restrict__ psi, const double* __restrict__ gvar, const double* __restrict__ sig2,
               const double* __restrict__ sv, int q, int p, int n0, int ipc, double* __restrict__ out) {
    extern __shared__ __align__(16) double fsm[];
    double* As = fsm;
    double* Bs = fsm + NB * FC_PITCH;
    double* gvs = Bs + NB * FC_PITCH;      // [FC_IPC][FC_KC]
    const int a0 = blockIdx.y * NB, b0 = blockIdx.x * NB;
    const int i0 = blockIdx.z * ipc, i1 = min(n0, i0 + ipc);
    const int nchunks = (q + FC_KC - 1) / FC_KC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
    const bool vec = (p & 1) == 0;   // (i p + a) p + b is even for even b: 16-byte stores are aligned
    // diagonal term of this lane's 8 rows: sig2_a sv_a^2 (only tiles on the diagonal use it)
    double dg[8];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        const int a = a0 + wm * 64 + mi * 8 + g;
        dg[mi] = (a0 == b0 && a < p) ? sig2[a] * sv[a] * sv[a] : 0.0;
    }
    if (nchunks == 1) {   // q <= 64: tiles and the sqrt(gvar) of all the CTA's test points stay resident -- no
        fc_load_tile(As, psi, sv, q, p, a0, 0);   // barrier and no global load inside the loop over test points
        fc_load_tile(Bs, psi, sv, q, p, b0, 0);
        fc_load_gv(gvs, gvar, q, n0, 0, i0, i1 - i0);
        __syncthreads();
    }

    for (int i = i0; i < i1; ++i) {
        double acc[8][4][2];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        for (int c = 0; c < nchunks; ++c) {
            const int k0 = c * FC_KC;
            const double* gv_i = gvs + (i - i0) * FC_KC;
            if (nchunks > 1) {                             // q > 64: chunks are reloaded per test point
                __syncthreads();                           // previous users of the tiles / gvs are done
                fc_load_tile(As, psi, sv, q, p, a0, k0);
                fc_load_tile(Bs, psi, sv, q, p, b0, k0);
                fc_load_gv(gvs, gvar, q, n0, k0, i, 1);
                gv_i = gvs;
                __syncthreads();
            }
            const int kc = min(FC_KC, round_up(q - k0, 4));
            for (int kk = 0; kk < kc / 4; ++kk) {
                const double gv = gv_i[kk * 4 + t];
                double a[8], b[4];
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) a[mi] = As[(wm * 64 + mi * 8 + g) * FC_PITCH + kk * 4 + t] * gv;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[(wn * 32 + ni * 8 + g) * FC_PITCH + kk * 4 + t] * gv;
#pragma unroll
                for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
        }
        double* o = out + (size_t)i * p * p;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
            const int a = a0 + wm * 64 + mi * 8 + g;
            if (a >= p) continue;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int b = b0 + wn * 32 + ni * 8 + 2 * t;
                if (b >= p) continue;
                const double v0 = acc[mi][ni][0] + (a == b ? dg[mi] : 0.0);
                const double v1 = acc[mi][ni][1] + (a == b + 1 ? dg[mi] : 0.0);
                double* ptr = o + (size_t)a * p + b;
                if (vec) {
                    *reinterpret_cast<double2*>(ptr) = make_double2(v0, v1);   // p even and b even -> b + 1 < p
                } else {
                    ptr[0] = v0;
                    if (b + 1 < p) ptr[1] = v1;
                }
            }
        }
    }
}

cudaError_t launch_fullcov(const double* psi, const double* gvar, const double* sig2, const double* sv, int q, int p,
                           int n0, double* out, cudaStream_t stream) {
    static std::atomic<bool> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!configured[dev].load(std::memory_order_acquire)) {   // racing first calls both set the attribute: harmless
        cudaError_t e = cudaFuncSetAttribute(fullcov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FC_SMEM);
        if (e != cudaSuccess) return e;
        configured[dev].store(true, std::memory_order_release);
    }
    const int tiles = (p + NB - 1) / NB;
    // test points per CTA: as many as keep >= ~4 CTAs per SM in flight (the psi tiles are loaded once per CTA)
    long long ipc = (long long)n0 * tiles * tiles / (148 * 4);
    ipc = ipc < 1 ? 1 : (ipc > FC_IPC ? FC_IPC : ipc);
    const long long gz = (n0 + ipc - 1) / ipc;
    if (gz > 65535) return cudaErrorInvalidValue;   // n0 > 1M test points per call: the host side chunks long before
    note_launch(); fullcov_kernel<<<dim3(tiles, tiles, (unsigned)gz), GEMM_THREADS, FC_SMEM, stream>>>(psi, gvar, sig2, sv, q, p, n0, (int)ipc, out);
    return cudaGetLastError();
}

}  // namespace lcgp
