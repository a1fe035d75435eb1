// Batched 128x128-tile FP64 GEMM on the Blackwell FP64 tensor pipe (mma.sync m8n8k4 -> DMMA.8x8x4).
//
// One CTA (8 warps, warp tile 64x32, 64 accumulator doubles per thread) produces one 128x128 tile
// acc[m][n] = sum_k A[m][k] * B[n][k].  Operand A is always K-major (row m contiguous in k); operand B
// is K-major or N-major (B[k][n], row k contiguous in n).  The K range is a run of NB-wide blocks
// [kb0, kb1); a Job names, per K block, the 128x128 source region of each operand (TileRef: which
// buffer, top-left row / column) -- so triangular operands can switch between the big factor buffer
// and the dense diagonal-block arrays and skip structurally-zero blocks -- and supplies an epilogue
// that consumes the register accumulators (store, read-modify-write, or a fused reduction).
//
// Two staging engines feed the same MMA code:
//   gemm_tma_kernel  : TMA tensor tiles (cp.async.bulk.tensor, SASS UTMALDG) with the 128-byte swizzle into a
//                      3-deep mbarrier full/empty ring; one elected lane issues 4 (K-major B) or 10
//                      (N-major B) box loads per 32-deep K stage; no CTA-wide barrier in the K loop.
//                      DMMA fragment loads stay bank-conflict free because fragment rows / columns are
//                      PERMUTED (see TmaCoord): the swizzle XORs the 16-byte chunk index with row % 8,
//                      so the four rows a half-warp touches must differ in bits 1-2, not bit 0.
//   gemm_dmma_kernel : cp.async (LDGSTS) into padded tiles, one barrier per stage (fallback engine).
#pragma once
#include <cuda.h>

#include <atomic>
#include <type_traits>

#include "common.cuh"

namespace lcgp {

// ---------------------------------------------------------------------------------------------
// Operand sources
// ---------------------------------------------------------------------------------------------
enum { SRC_F = 0, SRC_DU = 1, SRC_DL = 2, SRC_T = 3, SRC_C = 4, NSRC = 5 };

struct TileRef { int src, row, col; };   // top-left of a 128 x 128 region inside one batch element of `src`

struct GemmSrcs {                         // pointer view (batch element b starts at base + b * bstride)
    const double* base[NSRC];
    int ld[NSRC];
    size_t bstride[NSRC];
};

struct GemmMaps {                         // TMA view: 3-D maps (column, row, batch) of the same buffers
    CUtensorMap km[NSRC];                 // box 16 x 128 x 1  (K-major operand: 16 k-columns of 128 rows)
    CUtensorMap nm[2];                    // box 16 x 32 x 1   (N-major operand from SRC_F / SRC_DU)
    CUtensorMap km64;                     // box 16 x 64 x 1   (operand A of half-tile launches, from SRC_F)
    CUtensorMap km32;                     // box 16 x 32 x 1   (operand A of quarter-tile tasks of the persistent Cholesky)
    CUtensorMap du64, du32;               // the same two box shapes over SRC_DU (fused triangular inverse)
};

// Storage convention for one latent's factor buffer F (np x np, row-major, np % NB == 0):
//   strictly-lower NB-blocks and the lower triangle of diagonal blocks : L   (A = L L^T)
//   strictly-upper NB-blocks                                           : U = L^{-T} (U[i][k] = W[k][i], W = L^{-1})
//   DL[b], DU[b] (dense NB x NB, explicit zeros)                        : inverse of diagonal block b and its transpose
struct FactorView {
    double* F;          // batch base
    const double* DL;   // batch base, nb * NB * NB per matrix
    const double* DU;
    int np, nb;
    size_t fstride, dstride;
    int n;              // rows / columns that hold data (np when unknown): rows and columns >= n are the identity pad, and the
                        // kernels skip the MMA steps that would only multiply its structural zeros
};
// number of data rows in the last 128-block (1 .. NB)
__host__ __device__ inline int last_block_rows(const FactorView& v) {
    const int r = v.n - (v.nb - 1) * NB;
    return (r < 1 || r > NB) ? NB : r;
}

// ---------------------------------------------------------------------------------------------
// Accumulator <-> tile coordinates
// ---------------------------------------------------------------------------------------------
// cp.async engine: natural m8n8k4 layout.  lane = 4 g + t : a = A[g][t], b = B[t][g], d = D[g][2t, 2t+1]
struct WarpCoord {
    static constexpr bool kPairs = true;   // a lane's two accumulator columns are adjacent
    static constexpr int kMI = 8;          // 8-row accumulator groups per warp (warp tile 64 x 32)
    int wm, wn, g, t;
    __device__ __forceinline__ WarpCoord() {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        wm = warp >> 2;
        wn = warp & 3;
        g = lane >> 2;
        t = lane & 3;
    }
    __device__ __forceinline__ int row(int mi) const { return wm * 64 + mi * 8 + g; }
    __device__ __forceinline__ int col(int ni, int e) const { return wn * 32 + ni * 8 + 2 * t + e; }
};

// TMA engine: MMA row g of an 8-row group sits in tile row perm8(g) = 2 (g % 4) + g / 4, so that the
// rows of a half-warp (g = 0..3 or 4..7) are {0,2,4,6} / {1,3,5,7} and the A fragment loads are
// conflict free.  A K-major B operand keeps the natural n order instead (its 4 fragment loads per k-step
// are then 2-way conflicted, which costs 4 extra LDS wavefronts per 32 DMMAs) so that a lane's two
// accumulator columns stay adjacent for 16-byte epilogue accesses; an N-major B operand spreads the 8 n
// of an MMA over two 16-byte chunk groups of a 16-wide box: conflict free AND adjacent.
__device__ __forceinline__ int perm8(int g) { return ((g & 3) << 1) | (g >> 2); }

// MT = rows of the CTA tile: NB (warp tile 64 x 32) or NB / 2 (warp tile 32 x 32; blockIdx.z picks the upper or
// lower half of the 128-row tile the Job describes -- used for launches with fewer tiles than SMs, where
// halving the work per CTA halves the latency of a kernel that sits on a serial chain).
template <bool BNMAJOR, int MT = NB>
struct TmaCoord {
    static constexpr bool kPairs = true;
    static constexpr int kMI = MT / 16;
    int wm, wn, g, t, pg, moff;
    // Warps w and w + 4 share an SM sub-partition (one tensor pipe).  They take (wm, wn) = (0, s) and (1, 3 - s): every
    // sub-partition owns one warp of each row half, one of column chunks {0, 1} and one of {2, 3}, and chunks c and
    // 3 - c -- so that skipping structurally-zero MMA steps by row half, by column half or by triangular column chunk
    // (tma_compute_stage_skip) always takes the same share of work off every tensor pipe.
    __device__ __forceinline__ TmaCoord() {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        wm = warp >> 2;
        wn = wm ? 3 - (warp & 3) : (warp & 3);
        g = lane >> 2;
        t = lane & 3;
        pg = perm8(g);
        moff = (MT == NB) ? 0 : (int)blockIdx.z * MT;
    }
    __device__ __forceinline__ int row(int mi) const { return moff + wm * (MT / 2) + mi * 8 + pg; }
    __device__ __forceinline__ int col(int ni, int e) const {
        if (BNMAJOR) return wn * 32 + 16 * (ni >> 1) + 4 * (ni & 1) + ((t & 1) << 3) + (t & 2) + e;   // {0,8,2,10}[t]
        return wn * 32 + ni * 8 + 2 * t + e;
    }
};

// ---------------------------------------------------------------------------------------------
// cp.async engine
// ---------------------------------------------------------------------------------------------
template <class Job>
__device__ __forceinline__ void gemm_load_stage(const typename Job::Params& p, const Job& job, const GemmSrcs& s,
                                                int it, unsigned As, unsigned Bs) {
    const int kb = job.kb0 + it / KSTEPS;
    const int ks = it % KSTEPS;
    const TileRef ra = job.a_ref(p, kb), rb = job.b_ref(p, kb);
    const int lda = s.ld[ra.src], ldb = s.ld[rb.src];
    const double* pa = s.base[ra.src] + (size_t)blockIdx.y * s.bstride[ra.src] + (size_t)ra.row * lda + ra.col;
    const double* pb = s.base[rb.src] + (size_t)blockIdx.y * s.bstride[rb.src] + (size_t)rb.row * ldb + rb.col;
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

// One barrier per BK-deep stage; the loads for stage it+STAGES-1 (which overwrite the buffer consumed in
// iteration it-1, free since the barrier) are issued after the first quarter of the stage's MMAs.
template <class Job>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_dmma_kernel(const typename Job::Params p, const GemmSrcs srcs) {
    extern __shared__ __align__(16) double smem[];
    Job job;
    if (!job.init(p)) return;
    WarpCoord wc;
    double acc[8][4][2];
    constexpr int B_STAGE = Job::kBNMajor ? BN_STAGE : A_STAGE;
    constexpr int KK = BK / 4;
    double* As = smem;
    double* Bs = smem + STAGES * A_STAGE;
    const unsigned As_u = smem_u32(As), Bs_u = smem_u32(Bs);
    const int niter = (job.kb1 - job.kb0) * KSTEPS;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < niter) gemm_load_stage<Job>(p, job, srcs, s, As_u + s * A_STAGE * 8, Bs_u + s * B_STAGE * 8);
        cp_async_commit();
    }
    int cur = 0, nst = STAGES - 1;
    for (int it = 0; it < niter; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        gemm_compute_stage<Job::kBNMajor, 0, KK / 4>(As + cur * A_STAGE, Bs + cur * B_STAGE, acc, wc);
        const int nxt = it + STAGES - 1;
        if (nxt < niter) gemm_load_stage<Job>(p, job, srcs, nxt, As_u + nst * A_STAGE * 8, Bs_u + nst * B_STAGE * 8);
        cp_async_commit();
        gemm_compute_stage<Job::kBNMajor, KK / 4, KK>(As + cur * A_STAGE, Bs + cur * B_STAGE, acc, wc);
        cur = (cur + 1 == STAGES) ? 0 : cur + 1;
        nst = (nst + 1 == STAGES) ? 0 : nst + 1;
    }
    cp_async_wait<0>();
    __syncthreads();
    job.epilogue(p, acc, smem, wc);
}

// ---------------------------------------------------------------------------------------------
// TMA engine
// ---------------------------------------------------------------------------------------------
constexpr int TMA_BOX_K = 16;                                  // doubles per swizzled 128-byte row
constexpr int TMA_B_BYTES = NB * BK * 8;                       // 32 KB: operand B per stage
template <int MT> struct TmaShape {
    static constexpr int kABytes = MT * BK * 8;                // operand A per stage (32 KB, or 16 KB for half tiles)
    static constexpr int kStageBytes = kABytes + TMA_B_BYTES;
    static constexpr size_t kSmem = (size_t)STAGES * kStageBytes + 64 + 1024;   // + mbarriers + alignment slack
};
constexpr size_t GEMM_SMEM_TMA = TmaShape<NB>::kSmem;

__device__ __forceinline__ void tma_load_3d(unsigned smem_dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                 ::"r"(smem_dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<unsigned long long>(map)) : "memory");
}

// K-major operand tile in shared memory: 2 boxes (k 0..15 | 16..31) of [128 rows][128 B], chunk ^= row % 8.
// N-major operand tile: 8 boxes (16 n each) of [32 k rows][128 B].
template <bool BNMAJOR, int MT>
__device__ __forceinline__ void tma_compute_stage(const unsigned char* __restrict__ As, const unsigned char* __restrict__ Bs,
                                                  double (&acc)[MT / 16][4][2], const TmaCoord<BNMAJOR, MT>& wc) {
    constexpr int MI = MT / 16;
    const int arow = (wc.wm * (MT / 2) + wc.pg) * 128;   // + mi * 1024
    const int brow = (wc.wn * 32 + wc.g) * 128;       // K-major B (natural row order): + ni * 1024
    // N-major B: column c = nperm(ni & 1, g) inside box (wn * 2 + ni / 2)
    const int cn0 = (wc.g & 1) + ((wc.g & 2) << 2) + ((wc.g & 4) >> 1);
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
        double a[MI], b[4];
        const int akoff = (kk >> 2) * (MT * 128) + (((((kk & 3) << 1) | (wc.t >> 1)) ^ wc.pg) << 4) + ((wc.t & 1) << 3);
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) a[mi] = *reinterpret_cast<const double*>(As + arow + mi * 1024 + akoff);
        if (!BNMAJOR) {
            const int bkoff = (kk >> 2) * (NB * 128) + (((((kk & 3) << 1) | (wc.t >> 1)) ^ wc.g) << 4) + ((wc.t & 1) << 3);
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = *reinterpret_cast<const double*>(Bs + brow + ni * 1024 + bkoff);
        } else {
            const int k = kk * 4 + wc.t;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int c = cn0 + 4 * (ni & 1);
                b[ni] = *reinterpret_cast<const double*>(Bs + (wc.wn * 2 + (ni >> 1)) * (BK * 128) + k * 128 +
                                                         (((c >> 1) ^ (k & 7)) << 4) + ((c & 1) << 3));
            }
        }
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}

// Structural-zero skipping.  Of the 128-row tile the Job describes, only rows < rows and columns in [col_lo, col_hi)
// can be nonzero in this K step (a triangular diagonal block, the identity pad of the last block): a warp whose
// 32 columns lie outside does nothing, one whose row slab starts at slab0 works on its leading 8-row groups only.
// All conditions are warp uniform.  Masked steps are rare (the first or last K block of a tile, the tiles of the last
// block row / column), so this variant is COMPACT -- one k-step body in a rolled loop, ~1 KB of code against ~9 KB for
// the unrolled stage: a second unrolled copy next to the hot loop cost more in instruction fetch (ncu: no-instruction
// stalls x 5) than the skipped MMAs saved.
struct StepMask { int rows, col_lo, col_hi; };
template <bool BNMAJOR, int MT>
__device__ __forceinline__ void tma_compute_stage_skip(const unsigned char* __restrict__ As, const unsigned char* __restrict__ Bs,
                                                       double (&acc)[MT / 16][4][2], const TmaCoord<BNMAJOR, MT>& wc,
                                                       const StepMask& m, int slab0) {
    constexpr int MI = MT / 16;
    if (wc.wn * 32 + 32 <= m.col_lo || wc.wn * 32 >= m.col_hi) return;
    const int groups = (m.rows - slab0 + 7) >> 3;        // leading row groups of this warp's slab that can be nonzero
    if (groups <= 0) return;
    const int arow = (wc.wm * (MT / 2) + wc.pg) * 128;
    const int brow = (wc.wn * 32 + wc.g) * 128;
    const int cn0 = (wc.g & 1) + ((wc.g & 2) << 2) + ((wc.g & 4) >> 1);
#pragma unroll 1
    for (int kk = 0; kk < BK / 4; ++kk) {
        double a[MI], b[4];
        const int akoff = (kk >> 2) * (MT * 128) + (((((kk & 3) << 1) | (wc.t >> 1)) ^ wc.pg) << 4) + ((wc.t & 1) << 3);
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) a[mi] = *reinterpret_cast<const double*>(As + arow + mi * 1024 + akoff);
        if (!BNMAJOR) {
            const int bkoff = (kk >> 2) * (NB * 128) + (((((kk & 3) << 1) | (wc.t >> 1)) ^ wc.g) << 4) + ((wc.t & 1) << 3);
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = *reinterpret_cast<const double*>(Bs + brow + ni * 1024 + bkoff);
        } else {
            const int k = kk * 4 + wc.t;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int c = cn0 + 4 * (ni & 1);
                b[ni] = *reinterpret_cast<const double*>(Bs + (wc.wn * 2 + (ni >> 1)) * (BK * 128) + k * 128 +
                                                         (((c >> 1) ^ (k & 7)) << 4) + ((c & 1) << 3));
            }
        }
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
            if (mi < groups) {
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
    }
}

// A Job opts in with `static constexpr bool kSkips = true`, a member `tail_skip` (K steps at the end of its last K
// block that hold nothing but padding), `int head_steps(const Params&) const` / `int tail_steps(const Params&) const`
// (how many K steps at the start / end of the run may have a non-trivial mask; kAllSteps = the whole run) and
// `StepMask mask(const Params&, int it) const`.
#ifndef LCGP_LATE_FILL
#define LCGP_LATE_FILL 1   // 0: the producer always waits for the ring slot before its own MMAs (A/B builds)
#endif
#ifndef LCGP_SKIP_MODE
#define LCGP_SKIP_MODE 1   // 0: no structural-zero skipping (A/B builds)
#endif
constexpr int kAllSteps = 1 << 24;
template <class Job, class = void> struct JobSkips { static constexpr bool value = false; };
template <class Job> struct JobSkips<Job, decltype((void)Job::kSkips)> { static constexpr bool value = Job::kSkips; };
// A Job with `static constexpr bool kCustomSteps = true` orders its K steps itself: `void step_ref(int it, int& kb, int& ks)`
// names the K block and the BK-wide step inside it that pipeline iteration `it` loads (default: blocks kb0, kb0 + 1, ...).
template <class Job, class = void> struct JobSteps { static constexpr bool value = false; };
template <class Job> struct JobSteps<Job, decltype((void)Job::kCustomSteps)> { static constexpr bool value = Job::kCustomSteps; };

template <class Job, int MT = NB>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tma_kernel(const __grid_constant__ typename Job::Params p, const __grid_constant__ GemmMaps maps) {
    typedef TmaShape<MT> Shape;
    const int moff = (MT == NB) ? 0 : (int)blockIdx.z * MT;   // half tiles: rows [moff, moff + MT) of the Job's 128-row tile
    extern __shared__ unsigned char smem_raw[];
    Job job;
    if (!job.init(p)) return;   // uniform over the CTA
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms are 1 KB
    const unsigned sm_u = smem_u32(sm);
    const unsigned full0 = sm_u + STAGES * Shape::kStageBytes, empty0 = full0 + 8 * STAGES;
    if (threadIdx.x == 32) {   // hide the first descriptor fetches behind the barrier set-up
        const TileRef ra = job.a_ref(p, job.kb0), rb = job.b_ref(p, job.kb0);
        tma_prefetch_desc(MT == NB ? &maps.km[ra.src] : &maps.km64);
        tma_prefetch_desc(Job::kBNMajor ? &maps.nm[rb.src == SRC_F ? 0 : 1] : &maps.km[rb.src]);
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);                  // one arrive.expect_tx per fill
            mbar_init(empty0 + 8 * s, GEMM_THREADS / 32); // one arrive per MMA warp
        }
        mbar_fence_init();
    }
    __syncthreads();
    int niter = (job.kb1 - job.kb0) * KSTEPS;
    if constexpr (JobSkips<Job>::value) niter -= job.tail_skip;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // fill iteration `nxt` into ring slot `slot` (`use` = earlier fills of that slot); one lane issues it
    auto produce = [&](int nxt, int slot, int use) {
        if (lane == 0) {
            mbar_wait(empty0 + 8 * slot, (use & 1) ^ 1);   // passes at once for the first fill
            mbar_arrive_expect_tx(full0 + 8 * slot, Shape::kStageBytes);
            int kb = job.kb0 + nxt / KSTEPS, ks = nxt % KSTEPS;
            if constexpr (JobSteps<Job>::value) job.step_ref(nxt, kb, ks);
            const TileRef ra = job.a_ref(p, kb), rb = job.b_ref(p, kb);
            const unsigned dA = sm_u + slot * Shape::kStageBytes, dB = dA + Shape::kABytes;
            const CUtensorMap* ma = (MT == NB) ? &maps.km[ra.src] : &maps.km64;   // half tiles read operand A from F only
            const unsigned fb = full0 + 8 * slot;
            const int bz = blockIdx.y;
#pragma unroll
            for (int h = 0; h < BK / TMA_BOX_K; ++h)
                tma_load_3d(dA + h * (MT * 128), ma, ra.col + ks * BK + h * TMA_BOX_K, ra.row + moff, bz, fb);
            if (!Job::kBNMajor) {
#pragma unroll
                for (int h = 0; h < BK / TMA_BOX_K; ++h)
                    tma_load_3d(dB + h * (NB * 128), &maps.km[rb.src], rb.col + ks * BK + h * TMA_BOX_K, rb.row, bz, fb);
            } else {
                const CUtensorMap* mb = &maps.nm[rb.src == SRC_F ? 0 : 1];
#pragma unroll
                for (int j = 0; j < NB / TMA_BOX_K; ++j)
                    tma_load_3d(dB + j * (BK * 128), mb, rb.col + j * TMA_BOX_K, rb.row + ks * BK, bz, fb);
            }
        }
        __syncwarp();
    };
    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s)
            if (s < niter) produce(s, s, 0);
    }

    TmaCoord<Job::kBNMajor, MT> wc;
    double acc[MT / 16][4][2];
#pragma unroll
    for (int mi = 0; mi < MT / 16; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    int st = 0, nslot = STAGES - 1, nuse = 0;
    unsigned ph = 0;
    // one K step; MASKED steps go through the structural-zero dispatcher, the others run the plain stage code
    auto kstep = [&](int it, auto masked_c) {
        // The MMA warps take turns as producer.  The slot to refill is the one of step it - 1: if a slower warp still
        // reads it, the producer does its own MMAs first and fills afterwards (one stage of prefetch distance is still
        // several TMA latencies) instead of idling until the slowest warp has caught up.
        const int nxt = it + STAGES - 1;
        bool late = false;
        if (nxt < niter && warp == (it & 7)) {
#if LCGP_LATE_FILL
            unsigned ok = 0;
            if (lane == 0) ok = mbar_test(empty0 + 8 * nslot, (nuse & 1) ^ 1);
            ok = __shfl_sync(0xffffffffu, ok, 0);
            if (ok) produce(nxt, nslot, nuse);
            else late = true;
#else
            produce(nxt, nslot, nuse);
#endif
        }
        mbar_wait(full0 + 8 * st, ph);
        if constexpr (decltype(masked_c)::value)
            tma_compute_stage_skip<Job::kBNMajor, MT>(sm + st * Shape::kStageBytes, sm + st * Shape::kStageBytes + Shape::kABytes,
                                                      acc, wc, job.mask(p, it), moff + wc.wm * (MT / 2));
        else
            tma_compute_stage<Job::kBNMajor, MT>(sm + st * Shape::kStageBytes, sm + st * Shape::kStageBytes + Shape::kABytes, acc, wc);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * st);
        if (late) produce(nxt, nslot, nuse);
        if (++st == STAGES) { st = 0; ph ^= 1; }
        if (++nslot == STAGES) { nslot = 0; ++nuse; }
    };
    // The masked steps sit at the head and / or the tail of the K run; they get loops of their own so that the body
    // loop is exactly the plain pipeline (its instruction schedule is what the throughput of the whole kernel hangs on).
    int it = 0;
    if constexpr (JobSkips<Job>::value && LCGP_SKIP_MODE != 0) {
        const int head = min(job.head_steps(p), niter);
        const int body_end = max(head, niter - job.tail_steps(p));
        for (; it < head; ++it) kstep(it, std::true_type{});
        for (; it < body_end; ++it) kstep(it, std::false_type{});
        for (; it < niter; ++it) kstep(it, std::true_type{});
    } else {
        for (; it < niter; ++it) kstep(it, std::false_type{});
    }
    __syncthreads();   // every warp is done with the tiles: the epilogue may reuse shared memory
    job.epilogue(p, acc, reinterpret_cast<double*>(sm), wc);
}

// ---------------------------------------------------------------------------------------------
// Host side: sources, tensor maps, launch
// ---------------------------------------------------------------------------------------------
struct GemmCtx {
    GemmSrcs srcs;
    GemmMaps maps;
    bool tma;        // maps are valid and the TMA engine is selected
    int device;      // current device (the opt-in shared-memory attribute is per device)
};
constexpr int MAX_DEVICES = 64;

bool gemm_use_tma();   // engine selection (env LCGP_GEMM=tma|cpasync), defined in potrf.cu
int gemm_tma_min_kblocks();   // launches whose longest K run is shorter use the cp.async engine (env LCGP_TMA_MIN_KB)
// Fills ctx.srcs from the arguments and, when the TMA engine is selected, encodes the tensor maps.
// rows[] = rows per batch element of each source (0 = source unused).
cudaError_t gemm_make_ctx(GemmCtx& ctx, const GemmSrcs& srcs, const int rows[NSRC], int batch);

// max_kblocks: length (in 128-blocks) of the longest K run any tile of this launch has.
// MT = NB / 2 (TMA engine only, operand A from SRC_F, grid.z = 2): each 128-row tile is computed by two CTAs.
template <class Job, int MT = NB>
inline cudaError_t gemm_launch(const GemmCtx& ctx, const typename Job::Params& p, dim3 grid, cudaStream_t stream,
                               int max_kblocks = 1 << 20) {
    const int dev = (ctx.device >= 0 && ctx.device < MAX_DEVICES) ? ctx.device : 0;
    if (ctx.tma && (MT != NB || max_kblocks >= gemm_tma_min_kblocks())) {
        static std::atomic<bool> configured[MAX_DEVICES];
        if (!configured[dev].load(std::memory_order_acquire)) {   // racing first calls both set the attribute: harmless
            cudaError_t e = cudaFuncSetAttribute(gemm_tma_kernel<Job, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)TmaShape<MT>::kSmem);
            if (e != cudaSuccess) return e;
            configured[dev].store(true, std::memory_order_release);
        }
        note_launch(); gemm_tma_kernel<Job, MT><<<grid, GEMM_THREADS, TmaShape<MT>::kSmem, stream>>>(p, ctx.maps);
    } else {
        if (MT != NB) return cudaErrorNotSupported;   // half tiles exist for the TMA engine only
        constexpr size_t smem = Job::kBNMajor ? GEMM_SMEM_NMAJOR : GEMM_SMEM_KMAJOR;
        static std::atomic<bool> configured[MAX_DEVICES];
        if (!configured[dev].load(std::memory_order_acquire)) {   // racing first calls both set the attribute: harmless
            cudaError_t e = cudaFuncSetAttribute(gemm_dmma_kernel<Job>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[dev].store(true, std::memory_order_release);
        }
        note_launch(); gemm_dmma_kernel<Job><<<grid, GEMM_THREADS, smem, stream>>>(p, ctx.srcs);
    }
    return cudaGetLastError();
}

// Stores one accumulator pair of a tile to row-major C (ld in doubles) for either coordinate layout.
template <class Coord>
__device__ __forceinline__ void store_pair(double* C, size_t ld, const Coord& wc, int mi, int ni, double v0, double v1) {
    double* q = C + (size_t)wc.row(mi) * ld;
    if (Coord::kPairs) {
        *reinterpret_cast<double2*>(q + wc.col(ni, 0)) = make_double2(v0, v1);
    } else {
        q[wc.col(ni, 0)] = v0;
        q[wc.col(ni, 1)] = v1;
    }
}

// ---- Cholesky updates:  C[I][J] -= sum_{kb in [kb0,kb1)} P_I,kb P_J,kb^T ------------------------
//   col == -1 : trailing update, all tiles jstart <= J <= I < nb (lower triangle of the trailing matrix)
//   col == -2 : trailing update restricted to block columns jstart <= J < jend (the look-ahead part: the next
//               panel's columns); grid.x = (jend - jstart) * (nb - jstart), tiles above the diagonal exit
//   col >= 0  : left-looking update of block column `col` inside a panel: tiles (I, col), I >= col
struct SyrkJob {
    static constexpr bool kBNMajor = false;
    static constexpr bool kSkips = true;     // tiles of the last block row: the pad rows of P_I are zero
    static constexpr int tail_skip = 0;
    struct Params { FactorView v; int kb0, kb1, jstart, col, jend; };
    int kb0, kb1, I, J;
    double* base;
    __device__ int head_steps(const Params& p) const { return (I == p.v.nb - 1 && last_block_rows(p.v) <= NB - 8) ? kAllSteps : 0; }
    __device__ int tail_steps(const Params&) const { return 0; }
    __device__ StepMask mask(const Params& p, int) const { return StepMask{I == p.v.nb - 1 ? last_block_rows(p.v) : NB, 0, NB}; }
    __device__ bool init(const Params& p) {
        if (p.col >= 0) {
            I = p.col + blockIdx.x;
            J = p.col;
            if (I >= p.v.nb) return false;
        } else if (p.col == -2) {
            const int R = p.v.nb - p.jstart;
            J = p.jstart + (int)blockIdx.x / R;
            I = p.jstart + (int)blockIdx.x % R;
            if (J >= p.jend || I < J) return false;
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
    __device__ TileRef a_ref(const Params&, int kb) const { return TileRef{SRC_F, I * NB, kb * NB}; }
    __device__ TileRef b_ref(const Params&, int kb) const { return TileRef{SRC_F, J * NB, kb * NB}; }
    template <class Coord>
    __device__ void epilogue(const Params& p, double (&acc)[Coord::kMI][4][2], double*, const Coord& wc) const {
        double* C = base + (size_t)I * NB * p.v.np + (size_t)J * NB;
#pragma unroll
        for (int mi = 0; mi < Coord::kMI; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                double* q = C + (size_t)wc.row(mi) * p.v.np;
                if (Coord::kPairs) {
                    double2* q2 = reinterpret_cast<double2*>(q + wc.col(ni, 0));
                    double2 c = *q2;
                    c.x -= acc[mi][ni][0];
                    c.y -= acc[mi][ni][1];
                    *q2 = c;
                } else {
                    q[wc.col(ni, 0)] -= acc[mi][ni][0];
                    q[wc.col(ni, 1)] -= acc[mi][ni][1];
                }
            }
    }
};

// ---- Cholesky panel solve:  P_I = A[I][jb] * inv(L_jb,jb)^T  (in place), I > jb --------------
struct TrsmJob {
    static constexpr bool kBNMajor = false;
    static constexpr bool kSkips = true;     // inv(L_jb,jb) is lower triangular: output column c needs k <= c; last block row: pad rows
    static constexpr int tail_skip = 0;
    struct Params { FactorView v; int jb; };
    int kb0, kb1, I;
    double* base;
    __device__ int head_steps(const Params&) const { return kAllSteps; }
    __device__ int tail_steps(const Params&) const { return 0; }
    __device__ StepMask mask(const Params& p, int it) const {
        return StepMask{I == p.v.nb - 1 ? last_block_rows(p.v) : NB, BK * it, NB};
    }
    __device__ bool init(const Params& p) {
        I = p.jb + 1 + blockIdx.x;
        if (I >= p.v.nb) return false;
        kb0 = p.jb;
        kb1 = p.jb + 1;
        base = p.v.F + (size_t)blockIdx.y * p.v.fstride;
        return true;
    }
    __device__ TileRef a_ref(const Params&, int kb) const { return TileRef{SRC_F, I * NB, kb * NB}; }
    __device__ TileRef b_ref(const Params& p, int) const { return TileRef{SRC_DL, p.jb * NB, 0}; }   // B[n=c][k] = Linv[c][k]
    template <class Coord>
    __device__ void epilogue(const Params& p, double (&acc)[Coord::kMI][4][2], double*, const Coord& wc) const {
        double* C = base + (size_t)I * NB * p.v.np + (size_t)p.jb * NB;
#pragma unroll
        for (int mi = 0; mi < Coord::kMI; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) store_pair(C, p.v.np, wc, mi, ni, acc[mi][ni][0], acc[mi][ni][1]);
    }
};

// ---- TRTRI merge of block ranges [a,b) and [b,c) ---------------------------------------------
//   W[b:c, a:b] = -W[b:c,b:c] * L[b:c,a:b] * W[a:b,a:b], stored transposed in the upper triangle.
//   G1:  Tt[j][k'] = sum_k U[j][k] L[k'][k]        (j in [a,b), k' in [b,c), k in [j,b))
//   G2:  U[j][i]   = -sum_k' Tt[j][k'] U[k'][i]    (i in [b,c), k' in [b,i])
// Scratch of merge m (blockIdx.z) at level s: rows [m s NB, (m+1) s NB) of an (s NB)-wide matrix.
struct TrtriParams {
    FactorView v;
    double* T;        // batch base of the merge scratch
    size_t tstride;   // doubles per matrix
    int s;            // left range size in blocks at this level; merge m covers blocks [2sm, 2sm+2s)
};

struct TrtriG1Job {
    static constexpr bool kBNMajor = false;
    // first K block: A operand = upper-triangular DU_J; Kp = last block: pad rows of L are zero.  (Moving the DU block to
    // the tail of the K run, which pays for the contraction, made this launch 1.3 % slower -- measured, not kept.)
    static constexpr bool kSkips = true;
    static constexpr int tail_skip = 0;
    typedef TrtriParams Params;
    int kb0, kb1, J, Kp, a, b, jl, kl;
    __device__ int head_steps(const Params& p) const { return (Kp == p.v.nb - 1 && last_block_rows(p.v) <= NB - BK) ? kAllSteps : KSTEPS; }
    __device__ int tail_steps(const Params&) const { return 0; }
    __device__ StepMask mask(const Params& p, int it) const {
        return StepMask{it < KSTEPS ? BK * (it + 1) : NB, 0, Kp == p.v.nb - 1 ? last_block_rows(p.v) : NB};
    }
    __device__ bool init(const Params& p) {
        a = blockIdx.z * 2 * p.s;
        b = a + p.s;
        if (b >= p.v.nb) return false;
        const int c = min(a + 2 * p.s, p.v.nb);
        jl = blockIdx.x / p.s;
        kl = blockIdx.x % p.s;
        J = a + jl;
        Kp = b + kl;
        if (Kp >= c) return false;
        kb0 = J;
        kb1 = b;
        return true;
    }
    __device__ TileRef a_ref(const Params&, int kb) const {
        return kb == J ? TileRef{SRC_DU, J * NB, 0} : TileRef{SRC_F, J * NB, kb * NB};
    }
    __device__ TileRef b_ref(const Params&, int kb) const { return TileRef{SRC_F, Kp * NB, kb * NB}; }
    template <class Coord>
    __device__ void epilogue(const Params& p, double (&acc)[8][4][2], double*, const Coord& wc) const {
        const size_t ldt = (size_t)p.s * NB;
        double* tt = p.T + (size_t)blockIdx.y * p.tstride + ((size_t)blockIdx.z * ldt + (size_t)jl * NB) * ldt + (size_t)kl * NB;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) store_pair(tt, ldt, wc, mi, ni, acc[mi][ni][0], acc[mi][ni][1]);
    }
};

struct TrtriG2Job {
    static constexpr bool kBNMajor = true;
    static constexpr bool kSkips = true;     // last K block: B operand = upper-triangular DU_I (column c needs k' <= c); I = last block: pad columns
    static constexpr int tail_skip = 0;
    typedef TrtriParams Params;
    int kb0, kb1, J, I, a, b, jl;
    __device__ int head_steps(const Params& p) const { return (I == p.v.nb - 1 && last_block_rows(p.v) <= NB - BK) ? kAllSteps : 0; }
    __device__ int tail_steps(const Params&) const { return KSTEPS - 1; }
    __device__ StepMask mask(const Params& p, int it) const {
        const int itl = it - (kb1 - 1 - kb0) * KSTEPS;       // >= 0 inside the last K block
        return StepMask{NB, itl > 0 ? BK * itl : 0, I == p.v.nb - 1 ? last_block_rows(p.v) : NB};
    }
    __device__ bool init(const Params& p) {
        a = blockIdx.z * 2 * p.s;
        b = a + p.s;
        if (b >= p.v.nb) return false;
        const int c = min(a + 2 * p.s, p.v.nb);
        jl = blockIdx.x / p.s;
        const int il = blockIdx.x % p.s;
        J = a + jl;
        I = b + il;
        if (I >= c) return false;
        kb0 = b;
        kb1 = I + 1;
        return true;
    }
    __device__ TileRef a_ref(const Params& p, int kb) const {
        return TileRef{SRC_T, (int)blockIdx.z * p.s * NB + jl * NB, (kb - b) * NB};
    }
    __device__ TileRef b_ref(const Params&, int kb) const {   // N-major: rows = k', cols = i
        return kb == I ? TileRef{SRC_DU, I * NB, 0} : TileRef{SRC_F, kb * NB, I * NB};
    }
    template <class Coord>
    __device__ void epilogue(const Params& p, double (&acc)[8][4][2], double*, const Coord& wc) const {
        double* C = p.v.F + (size_t)blockIdx.y * p.v.fstride + (size_t)J * NB * p.v.np + (size_t)I * NB;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) store_pair(C, p.v.np, wc, mi, ni, -acc[mi][ni][0], -acc[mi][ni][1]);
    }
};

}  // namespace lcgp
