// Solves against the factor and the fused analytic-gradient contraction.
//
// With U = L^{-T} held in the upper triangle (see gemm_dmma.cuh) the per-latent quantities are
//   atil  = A^{-1} (b / sqrt r) = U (U^T v)                       (two memory-bound triangular mat-vecs)
//   alpha = sqrt(r) o atil      (= CinvM_k, lcgp.py:781)
//   m     = (b - alpha) / (d_k r)   (= C_k alpha = S_k b_k = mks, lcgp.py:779, from A atil = v)
//   quad  = b^T m
// and the gradient of T_k = 1/2 logdet A_k - 1/2 b^T S_k b wrt the kernel hyper-parameters is
//   dT/dtheta = 1/2 sum_ij G_ij dC_ij/dtheta,   G = d_k (sqrt r sqrt r^T) o A^{-1} - alpha alpha^T.
// ContractJob produces every 128x128 tile of A^{-1} = U U^T on the FP64 tensor pipe and consumes
// it in registers against dC/dtheta tiles recomputed from X in shared memory; neither A^{-1} nor
// any dC/dtheta matrix is written to HBM.
#include "gemm_dmma.cuh"
#include "lcgp_internal.h"

namespace lcgp {

// ---- y = U^T v :  y[k] = sum_{i <= k} U[i][k] v[i] -----------------------------------------------
__global__ void __launch_bounds__(NB)
gemv_ut_part_kernel(FactorView v, const double* __restrict__ B, const double* __restrict__ sr, int n, int q_per,
                    double* __restrict__ part /* [batch][nb][np] */) {
    const int Kb = blockIdx.x, Ib = blockIdx.y, bz = blockIdx.z;
    if (Ib > Kb) return;
    sr += (size_t)(bz / q_per) * n;
    __shared__ double vs[NB];
    const int c = threadIdx.x;
    {
        const int gi = Ib * NB + c;
        vs[c] = gi < n ? B[(size_t)bz * v.np + gi] / sr[gi] : 0.0;
    }
    __syncthreads();
    const double* tile;
    size_t ld;
    if (Ib == Kb) { tile = v.DU + (size_t)bz * v.dstride + (size_t)Ib * NB * NB; ld = NB; }
    else { tile = v.F + (size_t)bz * v.fstride + (size_t)Ib * NB * v.np + (size_t)Kb * NB; ld = v.np; }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 4
    for (int i = 0; i < NB; i += 4) {
        s0 += tile[(size_t)(i + 0) * ld + c] * vs[i + 0];
        s1 += tile[(size_t)(i + 1) * ld + c] * vs[i + 1];
        s2 += tile[(size_t)(i + 2) * ld + c] * vs[i + 2];
        s3 += tile[(size_t)(i + 3) * ld + c] * vs[i + 3];
    }
    part[((size_t)bz * v.nb + Ib) * v.np + (size_t)Kb * NB + c] = (s0 + s1) + (s2 + s3);
}

__global__ void gemv_ut_reduce_kernel(int np, int nb, const double* __restrict__ part, double* __restrict__ y) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int bz = blockIdx.y;
    if (k >= np) return;
    const int Kb = k / NB;
    double s = 0.0;
    for (int Ib = 0; Ib <= Kb; ++Ib) s += part[((size_t)bz * nb + Ib) * np + k];
    y[(size_t)bz * np + k] = s;
}

// ---- atil = U y (one warp per row), then alpha, m ----------------------------------------------
__global__ void __launch_bounds__(256)
gemv_u_kernel(FactorView v, const double* __restrict__ y, const double* __restrict__ B,
              const double* __restrict__ sr, const double* __restrict__ Dk, int n, int q_per,
              double* __restrict__ atil, double* __restrict__ alpha, double* __restrict__ mk) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    const int bz = blockIdx.y;
    if (i >= v.np) return;
    sr += (size_t)(bz / q_per) * n;
    const int I = i / NB, il = i % NB;
    const double* yk = y + (size_t)bz * v.np;
    const double* durow = v.DU + (size_t)bz * v.dstride + (size_t)I * NB * NB + (size_t)il * NB;
    double s = 0.0;
    for (int c = lane; c < NB; c += 32) s += durow[c] * yk[I * NB + c];
    const double* frow = v.F + (size_t)bz * v.fstride + (size_t)i * v.np;
    for (int k = (I + 1) * NB + lane; k < v.np; k += 32) s += frow[k] * yk[k];
    s = warp_sum(s);
    if (lane == 0) {
        const size_t o = (size_t)bz * v.np + i;
        if (i < n) {
            const double r = sr[i];
            const double al = r * s;
            atil[o] = s;
            alpha[o] = al;
            mk[o] = (B[o] - al) / (Dk[bz] * (r * r));
        } else {
            atil[o] = 0.0;
            alpha[o] = 0.0;
            mk[o] = 0.0;
        }
    }
}

__global__ void __launch_bounds__(256)
quad_kernel(int np, const double* __restrict__ B, const double* __restrict__ mk, double* __restrict__ quad) {
    __shared__ double red[8];
    const int bz = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < np; i += 256) s += B[(size_t)bz * np + i] * mk[(size_t)bz * np + i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) quad[bz] = s;
}

// gemv_part holds q_loc * nb * np partial sums followed by q_loc * np doubles for y = U^T v.
cudaError_t solve_alpha(const FactorView& v, const SolveArgs& a, cudaStream_t stream) {
    note_launch(); gemv_ut_part_kernel<<<dim3(v.nb, v.nb, a.q_loc), NB, 0, stream>>>(v, a.B, a.sr, a.n, a.kp.q_per, a.gemv_part);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    double* ybuf = a.gemv_part + (size_t)a.q_loc * v.nb * v.np;
    note_launch(); gemv_ut_reduce_kernel<<<dim3((v.np + 255) / 256, a.q_loc), 256, 0, stream>>>(v.np, v.nb, a.gemv_part, ybuf);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    note_launch(); gemv_u_kernel<<<dim3((v.np + 7) / 8, a.q_loc), 256, 0, stream>>>(v, ybuf, a.B, a.sr, a.kp.D, a.n, a.kp.q_per,
                                                                      a.atil, a.alpha, a.mk);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    note_launch(); quad_kernel<<<a.q_loc, 256, 0, stream>>>(v.np, a.B, a.mk, a.quad);
    return cudaGetLastError();
}

// 1 / x for x >= 1 (here x = 1 + S): hardware seed r0 (MUFU.RCP64H, relative error e = 1 - x r0 of ~2^-20) and ONE
// third-order step 1/x = r0 / (1 - e) ~ r0 (1 + e + e^2): error e^3 ~ 2^-60, three FMAs (two Newton steps take four).
// No range checks, no long dependent chain of the IEEE division; the quotient only weights gradient terms that are
// summed over n^2 pairs (tolerance 1e-8).
__device__ __forceinline__ double rcp_ge1(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    return fma(r, fma(e, e, e), r);
}

// ---- fused A^{-1} tile + gradient contraction ---------------------------------------------------
struct ContractParams {
    FactorView v;
    int n, d;
    const double* X;
    const double* sr;
    const double* alpha;  // [batch][np]
    KernelParams kp;
    double* tile_part;    // [batch][ntiles][d + 2] : (s0, lnug, ell_0..ell_{d-1}), unscaled
    int ntiles;
};

struct ContractJob {
    static constexpr bool kBNMajor = false;
    // Skipped MMA steps: the columns >= n of U in the last block are zero in every data row (tail_skip K steps), and the
    // diagonal block of operand A is the upper-triangular DU_I (row r is zero left of column r).  The K steps are ORDERED
    // so that the masked DU steps come LAST (step_ref): first the full blocks I + 1 .. nb - 2, then the valid steps of the
    // last block, then DU_I.  At the head of the run the row mask took 1.7 % of the tensor-pipe cycles off the kernel
    // and not a microsecond off its duration (the first steps of a tile run while the pipeline fills); at the tail the
    // ring is full and the sub-partition's other warp has the pipe to itself.
    static constexpr bool kSkips = true;
    static constexpr bool kCustomSteps = true;
    typedef ContractParams Params;
    int kb0, kb1, I, J, tail_skip, nfull, nlast, ndu;     // K steps of the full blocks, of the last block, of DU_I
    __device__ bool init(const Params& p) {
        if ((int)blockIdx.x >= p.ntiles) return false;
        tri_decode(blockIdx.x, I, J);
        kb0 = I;
        kb1 = p.v.nb;
        tail_skip = KSTEPS - (last_block_rows(p.v) + BK - 1) / BK;
        const int nblk = kb1 - kb0;
        nfull = (nblk >= 2 ? nblk - 2 : 0) * KSTEPS;
        nlast = nblk >= 2 ? KSTEPS - tail_skip : 0;
        ndu = nblk >= 2 ? KSTEPS : KSTEPS - tail_skip;        // I = nb - 1: DU_I is the last block
        return true;
    }
    __device__ void step_ref(int it, int& kb, int& ks) const {
        if (it < nfull) { kb = I + 1 + it / KSTEPS; ks = it % KSTEPS; }
        else if (it < nfull + nlast) { kb = kb1 - 1; ks = it - nfull; }
        else { kb = I; ks = it - nfull - nlast; }
    }
    __device__ int head_steps(const Params&) const { return 0; }
    __device__ int tail_steps(const Params&) const { return ndu; }
    __device__ StepMask mask(const Params&, int it) const {   // diagonal tiles: operand B is DU_I as well
        const int ks = it - nfull - nlast;
        const int lim = ks >= 0 ? BK * (ks + 1) : NB;
        return StepMask{lim, 0, I == J ? lim : NB};
    }
    __device__ TileRef a_ref(const Params&, int kb) const {
        return kb == I ? TileRef{SRC_DU, I * NB, 0} : TileRef{SRC_F, I * NB, kb * NB};
    }
    __device__ TileRef b_ref(const Params&, int kb) const {
        return kb == J ? TileRef{SRC_DU, J * NB, 0} : TileRef{SRC_F, J * NB, kb * NB};
    }
    template <class Coord>
    __device__ void epilogue(const Params& p, double (&acc)[8][4][2], double* smem, const Coord& wc) const {
        const int d = p.d, n = p.n, k = blockIdx.y, tid = threadIdx.x;
        double* xi = smem;                 // [d][NB]   x_i / ell
        double* xj = xi + d * NB;          // [d][NB]
        double* sri = xj + d * NB;         // sqrt r (0 in the pad)
        double* srj = sri + NB;
        double* ai = srj + NB;             // alpha
        double* aj = ai + NB;
        double* red = aj + NB;             // [(d + 2)][8]
        const double* ell = p.kp.ell + (size_t)k * d;
        const double* Xe = p.X + (size_t)(k / p.kp.q_per) * n * d;      // this latent's emulator
        const double* sre = p.sr + (size_t)(k / p.kp.q_per) * n;
        for (int idx = tid; idx < NB * d; idx += GEMM_THREADS) {
            const int m = idx / NB, r = idx % NB;
            const int gi = I * NB + r, gj = J * NB + r;
            const double l = ell[m];
            xi[idx] = gi < n ? Xe[(size_t)gi * d + m] / l : 0.0;
            xj[idx] = gj < n ? Xe[(size_t)gj * d + m] / l : 0.0;
        }
        if (tid < NB) {
            const int gi = I * NB + tid, gj = J * NB + tid;
            sri[tid] = gi < n ? sre[gi] : 0.0;
            srj[tid] = gj < n ? sre[gj] : 0.0;
            ai[tid] = p.alpha[(size_t)k * p.v.np + gi];
            aj[tid] = p.alpha[(size_t)k * p.v.np + gj];
        }
        __syncthreads();
        const double s0 = p.kp.s0[k], lnug = p.kp.lnug[k], dk = p.kp.D[k];
        const double nu = lnug / (1.0 + lnug);
        int cc[8];                          // this lane's 8 tile columns: (ni, e) -> cc[2 ni + e]
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) { cc[2 * ni] = wc.col(ni, 0); cc[2 * ni + 1] = wc.col(ni, 1); }
        // pass 1: C0 per element; acc <- G * C0.  With sumW = sum G C0 and sumD = sum over the diagonal of G (diagonal
        // tiles only), the two scalar derivatives are (1 - nu) sumW + nu sumD and sumD - sumW.
        double sumW = 0.0, sumD = 0.0;
        const bool dtile = (I == J);
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
            const int r = wc.row(mi);
            double P[8], V[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { P[e] = 1.0; V[e] = 0.0; }
            for (int m = 0; m < d; ++m) {
                const double a = xi[m * NB + r];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const double S = fabs(a - xj[m * NB + cc[e]]);
                    P[e] = fma(P[e], S, P[e]);      // P (1 + S)
                    V[e] -= S;
                }
            }
            const double dsr = dk * sri[r], al = ai[r];
            const int gi = I * NB + r;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = cc[2 * ni + e];
                    const double c0 = P[2 * ni + e] * exp(V[2 * ni + e]);
                    const double G = dsr * srj[c] * acc[mi][ni][e] - al * aj[c];
                    if (dtile && gi == J * NB + c) sumD += G;
                    const double w = G * c0;
                    sumW += w;
                    acc[mi][ni][e] = w;
                }
        }
        const int warp = tid >> 5, lane = tid & 31;
        sumW = warp_sum(sumW);
        sumD = warp_sum(sumD);
        const double acc_s0 = (1.0 - nu) * sumW + nu * sumD, acc_nug = sumD - sumW;
        if (lane == 0) { red[0 * 8 + warp] = acc_s0; red[1 * 8 + warp] = acc_nug; }
        // pass 2: sum_ij (G C0)_ij S_m^2 / (1 + S_m) per input dimension m
        for (int m = 0; m < d; ++m) {
            double a[8], b[8];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) a[mi] = xi[m * NB + wc.row(mi)];
#pragma unroll
            for (int e = 0; e < 8; ++e) b[e] = xj[m * NB + cc[e]];
            double sum = 0.0;
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const double S0 = fabs(a[mi] - b[2 * ni]), S1 = fabs(a[mi] - b[2 * ni + 1]);
                    sum = fma(acc[mi][ni][0] * S0, S0 * rcp_ge1(1.0 + S0), sum);
                    sum = fma(acc[mi][ni][1] * S1, S1 * rcp_ge1(1.0 + S1), sum);
                }
            sum = warp_sum(sum);
            if (lane == 0) red[(2 + m) * 8 + warp] = sum;
        }
        __syncthreads();
        if (tid < d + 2) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[tid * 8 + w];
            const double wt = (I == J) ? 0.5 : 1.0;  // 1/2 * (1 or 2 for the mirrored tile)
            double val;
            if (tid == 0) val = wt * s;
            else if (tid == 1) val = wt * s0 * s / ((1.0 + lnug) * (1.0 + lnug));
            else val = wt * s0 * (1.0 - nu) * s / ell[tid - 2];
            p.tile_part[((size_t)k * p.ntiles + blockIdx.x) * (d + 2) + tid] = val;
        }
    }
};

// g_kern: the kernel-gradient part of emulator 0's `out` block ([d/d lLmb (q_per x d) | d/d lLmb0 (q_per) | d/d lnugGPs
// (q_per)]); emulator e's block follows at e * out_stride
__global__ void contract_reduce_kernel(int ntiles, int d, int q_per, size_t out_stride, const double* __restrict__ tile_part,
                                       double* __restrict__ g_kern) {
    const int k = blockIdx.x, c = threadIdx.x;
    if (c >= d + 2) return;
    double s = 0.0;
    for (int t = 0; t < ntiles; ++t) s += tile_part[((size_t)k * ntiles + t) * (d + 2) + c];
    const int kk = k % q_per;
    double* g = g_kern + (size_t)(k / q_per) * out_stride;
    if (c == 0) g[(size_t)q_per * d + kk] = s;
    else if (c == 1) g[(size_t)q_per * d + q_per + kk] = s;
    else g[(size_t)kk * d + (c - 2)] = s;
}

cudaError_t contract_grad(const FactorView& v, const SolveArgs& a, double* tile_part, double* g_kern, size_t out_stride,
                          cudaEvent_t ev_before, cudaEvent_t ev_after, cudaStream_t stream) {
    const int ntiles = v.nb * (v.nb + 1) / 2;
    ContractParams p{v, a.n, a.d, a.X, a.sr, a.alpha, a.kp, tile_part, ntiles};
    GemmCtx ctx;
    {
        GemmSrcs srcs;
        int rows[NSRC];
        factor_srcs(v, srcs, rows);
        cudaError_t e0 = gemm_make_ctx(ctx, srcs, rows, a.q_loc);
        if (e0 != cudaSuccess) return e0;
    }
    if (ev_before) cudaEventRecord(ev_before, stream);
    cudaError_t e = gemm_launch<ContractJob>(ctx, p, dim3(ntiles, a.q_loc, 1), stream);
    if (ev_after) cudaEventRecord(ev_after, stream);
    if (e != cudaSuccess) return e;
    note_launch(); contract_reduce_kernel<<<a.q_loc, round_up(a.d + 2, 32), 0, stream>>>(ntiles, a.d, a.kp.q_per, out_stride, tile_part, g_kern);
    return cudaGetLastError();
}

// ---- gradient with respect to the latent basis (no counterpart in the reference, where phi is a
// constant): per latent k it needs tr A_k^{-1} = sum_{i<n} sum_{kk>=i} U[i][kk]^2 and sum_i r_i m_ki^2.
// One CTA per (row block, latent) streams its 128 rows of U once (HBM bound: 8 * np^2 / 2 bytes per latent).
__global__ void __launch_bounds__(256)
trace_part_kernel(FactorView v, int n, const double* __restrict__ sr, const double* __restrict__ mk,
                  double* __restrict__ part /* [q_loc][nb][2] */) {
    __shared__ double red[8];
    const int rb = blockIdx.x, k = blockIdx.y, tid = threadIdx.x;
    const double* F = v.F + (size_t)k * v.fstride;
    const double* DU = v.DU + (size_t)k * v.dstride + (size_t)rb * NB * NB;
    double acc = 0.0;
    const int rows = min(NB, n - rb * NB);           // rows i < n of this block (pad rows are the identity)
    for (int idx = tid; idx < rows * NB; idx += 256) {   // diagonal block: dense DU (zeros below the diagonal)
        const double u = DU[idx];
        acc += u * u;
    }
    for (int kb = rb + 1; kb < v.nb; ++kb)
        for (int idx = tid; idx < rows * NB; idx += 256) {
            const int r = idx >> 7, c = idx & (NB - 1);
            const double u = F[(size_t)(rb * NB + r) * v.np + (size_t)kb * NB + c];
            acc += u * u;
        }
    acc = block_sum(acc, red);
    double rm = 0.0;
    if (tid < rows) {
        const int i = rb * NB + tid;
        const double m = mk[(size_t)k * v.np + i], s = sr[i];
        rm = s * s * m * m;
    }
    rm = block_sum(rm, red);
    if (tid == 0) {
        part[((size_t)k * v.nb + rb) * 2] = acc;
        part[((size_t)k * v.nb + rb) * 2 + 1] = rm;
    }
}

// g_phi[j][k] = scale * ( -s_j Z[j][k] + 2 phi[j][k] dT_k/dd_k ),  dT_k/dd_k = (n - tr A_k^-1) / (2 d_k) + 1/2 sum_i r_i m_ki^2
__global__ void phi_grad_kernel(int n, int p, int q, int nb, double scale, const double* __restrict__ lsig,
                                const double* __restrict__ t, const double* __restrict__ phi, const double* __restrict__ D,
                                const double* __restrict__ Z, const double* __restrict__ part, double* __restrict__ g) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p * q) return;
    const int j = idx / q, k = idx % q;
    double tr = 0.0, rm = 0.0;
    for (int b = 0; b < nb; ++b) { tr += part[((size_t)k * nb + b) * 2]; rm += part[((size_t)k * nb + b) * 2 + 1]; }
    const double dTd = ((double)n - tr) / (2.0 * D[k]) + 0.5 * rm;
    const double sj = exp(-0.5 * lsig[j]) * t[j];
    g[idx] = scale * (-sj * Z[idx] + 2.0 * phi[idx] * dTd);
}

cudaError_t grad_phi(const FactorView& v, int n, int p, int q_loc, double scale, const double* sr, const double* mk,
                     const double* lsig, const double* t, const double* phi, const double* D, const double* Z,
                     double* part, double* g_phi, cudaStream_t stream) {
    note_launch(); trace_part_kernel<<<dim3(v.nb, q_loc), 256, 0, stream>>>(v, n, sr, mk, part);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    note_launch(); phi_grad_kernel<<<(p * q_loc + 255) / 256, 256, 0, stream>>>(n, p, q_loc, v.nb, scale, lsig, t, phi, D, Z, part, g_phi);
    return cudaGetLastError();
}

}  // namespace lcgp
