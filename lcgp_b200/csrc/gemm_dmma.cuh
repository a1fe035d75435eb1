// Batched 128x128-tile FP64 GEMM on the Blackwell FP64 tensor pipe (mma.sync m8n8k4 -> DMMA.8x8x4).
//
// One CTA (8 warps, warp tile 64x32, 64 accumulator doubles per thread) produces one 128x128
// tile  acc[m][n] = sum_k A[m][k] * B[n][k].  Operand A is always K-major (row m contiguous in
// k); operand B is K-major or N-major (B[k][n], row k contiguous in n).  Tiles are staged through
// a 3-deep cp.async ring (32-deep K stages) in shared memory with padded pitches chosen so that every DMMA fragment
// load (LDS.64) is bank-conflict free.  The K range is a run of NB-wide blocks [kb0, kb1); a Job
// supplies the tile base pointer per K block (so triangular operands can switch between the big
// factor buffer and the dense diagonal-block arrays) and an epilogue that consumes the register
// accumulators (store, read-modify-write, or a fused reduction).
#pragma once
#include "common.cuh"

namespace lcgp {

struct WarpCoord {
    int wm, wn, g, t;
    __device__ __forceinline__ WarpCoord() {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        wm = warp >> 2;  // 0..1  -> 64-row slab
        wn = warp & 3;   // 0..3  -> 32-col slab
        g = lane >> 2;   // fragment row / col group
        t = lane & 3;    // fragment k index / col pair
    }
    __device__ __forceinline__ int row(int mi) const { return wm * 64 + mi * 8 + g; }
    __device__ __forceinline__ int col(int ni) const { return wn * 32 + ni * 8 + 2 * t; }
};

template <class Job>
__device__ __forceinline__ void gemm_load_stage(const typename Job::Params& p, const Job& job, int it,
                                                unsigned As, unsigned Bs) {
    const int kb = job.kb0 + it / KSTEPS;
    const int ks = it % KSTEPS;
    const double* pa;
    const double* pb;
    int lda, ldb;
    job.a_src(p, kb, pa, lda);
    job.b_src(p, kb, pb, ldb);
    const int tid = threadIdx.x;
    constexpr int CPR = BK / 2;                       // 16-byte chunks per K-major tile row
    constexpr int NCH = NB * CPR / GEMM_THREADS;      // chunks per thread per operand
#pragma unroll
    for (int u = 0; u < NCH; ++u) {
        const int c = tid + u * GEMM_THREADS;
        const int row = c / CPR, kc = c % CPR;
        cp_async16(As + (row * LDS_K + kc * 2) * 8, pa + (size_t)row * lda + ks * BK + kc * 2);
    }
    if (!Job::kBNMajor) {
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
            const int c = tid + u * GEMM_THREADS;
            const int row = c / CPR, kc = c % CPR;
            cp_async16(Bs + (row * LDS_K + kc * 2) * 8, pb + (size_t)row * ldb + ks * BK + kc * 2);
        }
    } else {
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
            const int c = tid + u * GEMM_THREADS;
            const int krow = c >> 6, nc = c & 63;     // 64 chunks per 128-wide N-major row
            cp_async16(Bs + (krow * LDS_N + nc * 2) * 8, pb + (size_t)(ks * BK + krow) * ldb + nc * 2);
        }
    }
}

template <bool BNMAJOR, int KK0, int KK1>
__device__ __forceinline__ void gemm_compute_stage(const double* __restrict__ As, const double* __restrict__ Bs,
                                                   double (&acc)[8][4][2], const WarpCoord& wc) {
#pragma unroll
    for (int kk = KK0; kk < KK1; ++kk) {
        double a[8], b[4];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) a[mi] = As[(wc.wm * 64 + mi * 8 + wc.g) * LDS_K + kk * 4 + wc.t];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
            b[ni] = BNMAJOR ? Bs[(kk * 4 + wc.t) * LDS_N + wc.wn * 32 + ni * 8 + wc.g]
                            : Bs[(wc.wn * 32 + ni * 8 + wc.g) * LDS_K + kk * 4 + wc.t];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}

// Runs the pipelined K loop of `job` and leaves the tile in `acc`.  On return all cp.async
// groups have drained and all threads have passed a barrier, so `smem` may be reused.
// One barrier per BK-deep stage; the loads for stage it+STAGES-1 (which overwrite the buffer
// consumed in iteration it-1, free since the barrier) are issued after the first quarter of the
// stage's MMAs so that the tensor pipe already has work queued while the LSU issues them.
template <class Job>
__device__ __forceinline__ void gemm_mainloop(const typename Job::Params& p, const Job& job,
                                              double (&acc)[8][4][2], double* smem, const WarpCoord& wc) {
    constexpr int B_STAGE = Job::kBNMajor ? BN_STAGE : A_STAGE;
    constexpr int KK = BK / 4;
    double* As = smem;
    double* Bs = smem + STAGES * A_STAGE;
    const unsigned As_u = (unsigned)__cvta_generic_to_shared(As);
    const unsigned Bs_u = (unsigned)__cvta_generic_to_shared(Bs);
    const int niter = (job.kb1 - job.kb0) * KSTEPS;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < niter) gemm_load_stage<Job>(p, job, s, As_u + s * A_STAGE * 8, Bs_u + s * B_STAGE * 8);
        cp_async_commit();
    }
    int cur = 0, nst = STAGES - 1;
    for (int it = 0; it < niter; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        gemm_compute_stage<Job::kBNMajor, 0, KK / 4>(As + cur * A_STAGE, Bs + cur * B_STAGE, acc, wc);
        const int nxt = it + STAGES - 1;
        if (nxt < niter) gemm_load_stage<Job>(p, job, nxt, As_u + nst * A_STAGE * 8, Bs_u + nst * B_STAGE * 8);
        cp_async_commit();
        gemm_compute_stage<Job::kBNMajor, KK / 4, KK>(As + cur * A_STAGE, Bs + cur * B_STAGE, acc, wc);
        cur = (cur + 1 == STAGES) ? 0 : cur + 1;
        nst = (nst + 1 == STAGES) ? 0 : nst + 1;
    }
    cp_async_wait<0>();
    __syncthreads();
}

template <class Job>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_dmma_kernel(const typename Job::Params p) {
    extern __shared__ __align__(16) double smem[];
    Job job;
    if (!job.init(p)) return;
    WarpCoord wc;
    double acc[8][4][2];
    gemm_mainloop<Job>(p, job, acc, smem, wc);
    job.epilogue(p, acc, smem, wc);
}

template <class Job>
inline cudaError_t gemm_launch(const typename Job::Params& p, dim3 grid, cudaStream_t stream) {
    constexpr size_t smem = Job::kBNMajor ? GEMM_SMEM_NMAJOR : GEMM_SMEM_KMAJOR;
    static bool configured = false;  // per-process, per-Job; attribute is idempotent
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_dmma_kernel<Job>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    gemm_dmma_kernel<Job><<<grid, GEMM_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------
// Storage convention for one latent's factor buffer F (np x np, row-major, np % NB == 0):
//   strictly-lower NB-blocks and the lower triangle of diagonal blocks : L   (A = L L^T)
//   strictly-upper NB-blocks                                           : U = L^{-T} (U[i][k] = W[k][i], W = L^{-1})
//   DL[b], DU[b] (dense NB x NB, explicit zeros)                        : inverse of diagonal block b and its transpose
// -------------------------------------------------------------------------------------------
struct FactorView {
    double* F;          // batch base
    const double* DL;   // batch base, nb * NB * NB per matrix
    const double* DU;
    int np, nb;
    size_t fstride, dstride;
};

// ---- Cholesky updates:  C[I][J] -= sum_{kb in [kb0,kb1)} P_I,kb P_J,kb^T ------------------------
//   col <  0 : trailing update, all tiles jstart <= J <= I < nb (lower triangle of the trailing matrix)
//   col >= 0 : left-looking update of block column `col` inside a panel: tiles (I, col), I >= col
struct SyrkJob {
    static constexpr bool kBNMajor = false;
    struct Params { FactorView v; int kb0, kb1, jstart, col; };
    int kb0, kb1, I, J;
    double* base;
    __device__ bool init(const Params& p) {
        if (p.col >= 0) {
            I = p.col + blockIdx.x;
            J = p.col;
            if (I >= p.v.nb) return false;
        } else {
            const int T = p.v.nb - p.jstart;
            if ((int)blockIdx.x >= T * (T + 1) / 2) return false;
            int ti, tj;
            tri_decode(blockIdx.x, ti, tj);
            I = p.jstart + ti;
            J = p.jstart + tj;
        }
        kb0 = p.kb0;
        kb1 = p.kb1;
        base = p.v.F + (size_t)blockIdx.y * p.v.fstride;
        return true;
    }
    __device__ void a_src(const Params& p, int kb, const double*& ptr, int& ld) const {
        ptr = base + (size_t)I * NB * p.v.np + (size_t)kb * NB;
        ld = p.v.np;
    }
    __device__ void b_src(const Params& p, int kb, const double*& ptr, int& ld) const {
        ptr = base + (size_t)J * NB * p.v.np + (size_t)kb * NB;
        ld = p.v.np;
    }
    __device__ void epilogue(const Params& p, double (&acc)[8][4][2], double*, const WarpCoord& wc) const {
        double* C = base + (size_t)I * NB * p.v.np + (size_t)J * NB;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                double2* q = reinterpret_cast<double2*>(C + (size_t)wc.row(mi) * p.v.np + wc.col(ni));
                double2 c = *q;
                c.x -= acc[mi][ni][0];
                c.y -= acc[mi][ni][1];
                *q = c;
            }
    }
};

// ---- Cholesky panel solve:  P_I = A[I][jb] * inv(L_jb,jb)^T  (in place), I > jb --------------
struct TrsmJob {
    static constexpr bool kBNMajor = false;
    struct Params { FactorView v; int jb; };
    int kb0, kb1, I;
    double* base;
    const double* dl;
    __device__ bool init(const Params& p) {
        I = p.jb + 1 + blockIdx.x;
        if (I >= p.v.nb) return false;
        kb0 = p.jb;
        kb1 = p.jb + 1;
        base = p.v.F + (size_t)blockIdx.y * p.v.fstride;
        dl = p.v.DL + (size_t)blockIdx.y * p.v.dstride + (size_t)p.jb * NB * NB;
        return true;
    }
    __device__ void a_src(const Params& p, int kb, const double*& ptr, int& ld) const {
        ptr = base + (size_t)I * NB * p.v.np + (size_t)kb * NB;
        ld = p.v.np;
    }
    __device__ void b_src(const Params&, int, const double*& ptr, int& ld) const {
        ptr = dl;  // B[n=c][k] = Linv[c][k]
        ld = NB;
    }
    __device__ void epilogue(const Params& p, double (&acc)[8][4][2], double*, const WarpCoord& wc) const {
        double* C = base + (size_t)I * NB * p.v.np + (size_t)p.jb * NB;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
                *reinterpret_cast<double2*>(C + (size_t)wc.row(mi) * p.v.np + wc.col(ni)) =
                    make_double2(acc[mi][ni][0], acc[mi][ni][1]);
    }
};

// ---- TRTRI merge of block ranges [a,b) and [b,c) ---------------------------------------------
//   W[b:c, a:b] = -W[b:c,b:c] * L[b:c,a:b] * W[a:b,a:b], stored transposed in the upper triangle.
//   G1:  Tt[j][k'] = sum_k U[j][k] L[k'][k]        (j in [a,b), k' in [b,c), k in [j,b))
//   G2:  U[j][i]   = -sum_k' Tt[j][k'] U[k'][i]    (i in [b,c), k' in [b,i])
struct TrtriParams {
    FactorView v;
    double* T;        // batch base of the merge scratch
    size_t tstride;   // doubles per matrix
    int s;            // left range size in blocks at this level; merge m covers blocks [2sm, 2sm+2s)
};

struct TrtriG1Job {
    static constexpr bool kBNMajor = false;
    typedef TrtriParams Params;
    int kb0, kb1, J, Kp, a, b;
    double* base;
    const double* du;
    double* tt;
    __device__ bool init(const Params& p) {
        a = blockIdx.z * 2 * p.s;
        b = a + p.s;
        if (b >= p.v.nb) return false;
        const int c = min(a + 2 * p.s, p.v.nb);
        const int jl = blockIdx.x / p.s, kl = blockIdx.x % p.s;
        J = a + jl;
        Kp = b + kl;
        if (Kp >= c) return false;
        kb0 = J;
        kb1 = b;
        base = p.v.F + (size_t)blockIdx.y * p.v.fstride;
        du = p.v.DU + (size_t)blockIdx.y * p.v.dstride;
        const size_t ldt = (size_t)p.s * NB;
        tt = p.T + (size_t)blockIdx.y * p.tstride + (size_t)blockIdx.z * ldt * ldt + (size_t)jl * NB * ldt + (size_t)kl * NB;
        return true;
    }
    __device__ void a_src(const Params& p, int kb, const double*& ptr, int& ld) const {
        if (kb == J) { ptr = du + (size_t)J * NB * NB; ld = NB; }
        else { ptr = base + (size_t)J * NB * p.v.np + (size_t)kb * NB; ld = p.v.np; }
    }
    __device__ void b_src(const Params& p, int kb, const double*& ptr, int& ld) const {
        ptr = base + (size_t)Kp * NB * p.v.np + (size_t)kb * NB;
        ld = p.v.np;
    }
    __device__ void epilogue(const Params& p, double (&acc)[8][4][2], double*, const WarpCoord& wc) const {
        const size_t ldt = (size_t)p.s * NB;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
                *reinterpret_cast<double2*>(tt + (size_t)wc.row(mi) * ldt + wc.col(ni)) =
                    make_double2(acc[mi][ni][0], acc[mi][ni][1]);
    }
};

struct TrtriG2Job {
    static constexpr bool kBNMajor = true;
    typedef TrtriParams Params;
    int kb0, kb1, J, I, a, b;
    double* base;
    const double* du;
    const double* tt;  // row block J of the merge scratch
    __device__ bool init(const Params& p) {
        a = blockIdx.z * 2 * p.s;
        b = a + p.s;
        if (b >= p.v.nb) return false;
        const int c = min(a + 2 * p.s, p.v.nb);
        const int jl = blockIdx.x / p.s, il = blockIdx.x % p.s;
        J = a + jl;
        I = b + il;
        if (I >= c) return false;
        kb0 = b;
        kb1 = I + 1;
        base = p.v.F + (size_t)blockIdx.y * p.v.fstride;
        du = p.v.DU + (size_t)blockIdx.y * p.v.dstride;
        const size_t ldt = (size_t)p.s * NB;
        tt = p.T + (size_t)blockIdx.y * p.tstride + (size_t)blockIdx.z * ldt * ldt + (size_t)jl * NB * ldt;
        return true;
    }
    __device__ void a_src(const Params& p, int kb, const double*& ptr, int& ld) const {
        ptr = tt + (size_t)(kb - b) * NB;
        ld = p.s * NB;
    }
    __device__ void b_src(const Params& p, int kb, const double*& ptr, int& ld) const {
        if (kb == I) { ptr = du + (size_t)I * NB * NB; ld = NB; }
        else { ptr = base + (size_t)kb * NB * p.v.np + (size_t)I * NB; ld = p.v.np; }
    }
    __device__ void epilogue(const Params& p, double (&acc)[8][4][2], double*, const WarpCoord& wc) const {
        double* C = base + (size_t)J * NB * p.v.np + (size_t)I * NB;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
                *reinterpret_cast<double2*>(C + (size_t)wc.row(mi) * p.v.np + wc.col(ni)) =
                    make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
    }
};

}  // namespace lcgp
