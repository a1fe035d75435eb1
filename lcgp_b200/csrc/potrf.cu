// Batched blocked FP64 Cholesky (right-looking, NB = 128) and blocked triangular inverse.
//
//   per block column jb:   diag kernel  : L_jj = chol(A_jj), DL = inv(L_jj), DU = DL^T, logdet part
//                          TrsmJob      : P = A[jb+1:, jb] * DL^T                 (DMMA GEMM)
//                          SyrkJob      : A[I][J] -= P_I P_J^T (panel-column and trailing updates, DMMA GEMM)
//   TRTRI: bottom-up binary merges of already-inverted diagonal ranges, two DMMA GEMMs per level.
#include <atomic>

#include "chol128.cuh"
#include "gemm_dmma.cuh"
#include "lcgp_internal.h"

#include <cstdlib>
#include <cstring>

namespace lcgp {

// ---- staging-engine selection and tensor-map encoding ---------------------------------------------
static bool gemm_tma_requested() {
    static const bool v = [] {
        const char* e = std::getenv("LCGP_GEMM");
        return !(e && std::strcmp(e, "cpasync") == 0);   // default: TMA engine
    }();
    return v;
}

static int env_clamped(const char* name, int dflt, int lo, int hi) {
    const char* e = std::getenv(name);
    if (!e || !*e) return dflt;
    const int v = std::atoi(e);
    return v < lo ? lo : (v > hi ? hi : v);
}
// LCGP_TAIL_N / LCGP_TAIL_W: panels are LCGP_TAIL_W block columns wide once at most LCGP_TAIL_N remain
static int potrf_tail_n() { static const int v = env_clamped("LCGP_TAIL_N", 24, 0, 1 << 20); return v; }
static int potrf_tail_w() { static const int v = env_clamped("LCGP_TAIL_W", 4, 1, 16); return v; }
constexpr int LCGP_MAX_PANELS = 4096;
// LCGP_HALF_TILES: panel kernels (column update, solve) with at most this many 128 x 128 tiles are launched as twice
// as many 64 x 128 tiles: they sit on the serial panel chain and leave SMs idle anyway, so halving the work per CTA
// halves their latency.  0 = never.
static int half_tile_limit() { static const int v = env_clamped("LCGP_HALF_TILES", 74, 0, 1 << 20); return v; }

// LCGP_POTRF = pll (one persistent kernel per factorisation, potrf_pll.cu) | panels (launch chain below) | auto
// (default): the persistent kernel except for large batches of large matrices (>= 24 matrices of >= 32 block columns),
// where the launch chain with its K = 1024 trailing updates + the merge-based inverse is still 0.5 % faster (config 4,
// factor + inverse at the end of round 2: 322.8 ms vs 324.6 ms at 32 latents per GPU; 164.4 vs 162.7 at 16, 83.1 vs 83.2
// at 8: every tile of the persistent kernel pays one extra K block for its solve against the dense
// inverse of the diagonal block) -- everywhere else the persistent kernel wins by up to 45 % (profiles/r2_potrf_*.txt).
static int potrf_mode() {   // 0 = auto, 1 = pll, 2 = panels
    static const int v = [] {
        const char* e = std::getenv("LCGP_POTRF");
        if (e && std::strcmp(e, "panels") == 0) return 2;
        if (e && std::strcmp(e, "pll") == 0) return 1;
        return 0;
    }();
    return v;
}
bool potrf_use_pll() { return potrf_mode() != 2 && gemm_use_tma(); }
bool potrf_fuse_trtri() {
    static const bool v = [] {
        const char* e = std::getenv("LCGP_FUSE_TRTRI");
        return !(e && std::strcmp(e, "0") == 0);
    }();
    return v;
}
bool potrf_use_pll(int nb, int batch) {
    if (!potrf_use_pll()) return false;
    return potrf_mode() == 1 || !(batch >= 24 && nb >= 32);
}
// LCGP_DIAG = v2 (default: blocked shared-memory kernel, chol128.cuh) | v1 (register-resident column-by-column kernel)
static bool diag_use_v2() {
    static const bool v = [] {
        const char* e = std::getenv("LCGP_DIAG");
        return !(e && std::strcmp(e, "v1") == 0);
    }();
    return v;
}

int gemm_tma_min_kblocks() {
    static const int v = [] {
        const char* e = std::getenv("LCGP_TMA_MIN_KB");
        return (e && *e) ? std::atoi(e) : 1;
    }();
    return v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn();
// TMA engine unless LCGP_GEMM=cpasync -- or the driver does not export cuTensorMapEncodeTiled, in which case every GEMM
// (and the Cholesky, through the launch chain) runs on the cp.async engine instead of failing
bool gemm_use_tma() { return gemm_tma_requested() && encode_fn() != nullptr; }

static EncodeTiledFn encode_fn() {
    static const EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// 3-D map (column, row, batch) over a row-major fp64 buffer; box = 16 columns x box_rows rows, 128-byte swizzle
static cudaError_t encode_map(CUtensorMap* m, const double* base, int cols, int rows, int batch, int ld,
                              size_t bstride, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cudaErrorNotSupported;
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)bstride * 8};
    const cuuint32_t box[3] = {(cuuint32_t)TMA_BOX_K, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t gemm_make_ctx(GemmCtx& ctx, const GemmSrcs& srcs, const int rows[NSRC], int batch) {
    ctx.srcs = srcs;
    ctx.tma = gemm_use_tma();
    ctx.device = 0;
    cudaGetDevice(&ctx.device);
    if (!ctx.tma) return cudaSuccess;
    std::memset(&ctx.maps, 0, sizeof(ctx.maps));
    for (int i = 0; i < NSRC; ++i) {
        if (!srcs.base[i] || rows[i] <= 0) continue;
        cudaError_t e = encode_map(&ctx.maps.km[i], srcs.base[i], srcs.ld[i], rows[i], batch, srcs.ld[i], srcs.bstride[i], NB);
        if (e != cudaSuccess) return e;
        if (i == SRC_F || i == SRC_DU) {
            e = encode_map(&ctx.maps.nm[i == SRC_F ? 0 : 1], srcs.base[i], srcs.ld[i], rows[i], batch, srcs.ld[i],
                           srcs.bstride[i], BK);
            if (e != cudaSuccess) return e;
        }
        if (i == SRC_DU) {
            e = encode_map(&ctx.maps.du64, srcs.base[i], srcs.ld[i], rows[i], batch, srcs.ld[i], srcs.bstride[i], NB / 2);
            if (e != cudaSuccess) return e;
            e = encode_map(&ctx.maps.du32, srcs.base[i], srcs.ld[i], rows[i], batch, srcs.ld[i], srcs.bstride[i], NB / 4);
            if (e != cudaSuccess) return e;
        }
        if (i == SRC_F) {   // half-tile launches: operand A boxes of 64 rows
            e = encode_map(&ctx.maps.km64, srcs.base[i], srcs.ld[i], rows[i], batch, srcs.ld[i], srcs.bstride[i], NB / 2);
            if (e != cudaSuccess) return e;
            e = encode_map(&ctx.maps.km32, srcs.base[i], srcs.ld[i], rows[i], batch, srcs.ld[i], srcs.bstride[i], NB / 4);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

// sources every factor-based GEMM has: F, DU, DL
void factor_srcs(const FactorView& v, GemmSrcs& s, int rows[NSRC]) {
    std::memset(&s, 0, sizeof(s));
    for (int i = 0; i < NSRC; ++i) rows[i] = 0;
    s.base[SRC_F] = v.F;   s.ld[SRC_F] = v.np;  s.bstride[SRC_F] = v.fstride;   rows[SRC_F] = v.np;
    s.base[SRC_DU] = v.DU; s.ld[SRC_DU] = NB;   s.bstride[SRC_DU] = v.dstride;  rows[SRC_DU] = v.nb * NB;
    s.base[SRC_DL] = v.DL; s.ld[SRC_DL] = NB;   s.bstride[SRC_DL] = v.dstride;  rows[SRC_DL] = v.nb * NB;
}

// ------------------------------------------------------------------------------------------
// Diagonal-block kernel: one CTA (16 x 16 threads) per matrix; the 128 x 128 block is REGISTER
// resident, thread (ty, tx) owning S[ty + 16 i][tx + 16 j], i, j < 8.  The lower triangle is the
// running Schur complement, the strictly-upper triangle the transpose of the running inverse Z of
// the factor (Z[i][j] at S[j][i]).  Column step c applies, with u = column c (u[c] := 1),
//     S[a][b] -= u[a] u[b] / piv     for  b > c  and  (a >= b  or  a <= c)
// which is at once the Cholesky trailing update and the forward elimination of the identity.
// Column c is dead afterwards, so its scaling by 1/sqrt(piv_c) is deferred to the write-back.
//
// Synchronisation is split-phase (mbarriers, no CTA-wide rendezvous in the loop): the 16 owners of
// column c+1 update that column FIRST, publish it to shared memory and arrive on ready[(c+1)&1];
// everybody then finishes the rest of step c while the publication propagates, and waits on the
// mbarrier only at the top of step c+1.  The two-deep u buffer needs no second barrier: every warp
// holds two owners of every column (tx = lane % 16), so ready[c] completing means every warp has
// published column c -- in its step c-1, after the loads of that step returned (the published values
// depend on them) -- and buffer (c+1)&1, last read in step c-1, is free when column c+1 is written in
// step c.  The critical path per column is  wait -> LDS -> reciprocal -> 8 FMAs -> STS -> arrive.
// ------------------------------------------------------------------------------------------
constexpr int DIAG_THREADS = 256;
constexpr int DPITCH = NB + 1;
// The published column is stored permuted and padded: row t + 16 i at t * UP + i.  A thread's 8 values (its
// rows ty + 16 i, its columns tx + 16 j) are then contiguous -- 4 LDS.128 instead of 8 LDS.64 for each of
// u[a], u[b], 4 STS.128 to publish -- and the 80-byte stride keeps the quarter-warp phases conflict free.
constexpr int UP = 10;
constexpr int UBUF = 16 * UP;
constexpr size_t DIAG_SMEM = sizeof(double) * (NB * DPITCH + 2 * UBUF + NB + NB + 16);

// Owner thread (rows ty + 16 i) publishes val[0..7]; dg >= 0: val[dg] is the pivot (diagonal entry := 1).
__device__ __forceinline__ void diag_publish(const double (&val)[8], int dg, double* un, double* pivs, int cn) {
    double2* u2 = reinterpret_cast<double2*>(un);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        double v0 = val[i], v1 = val[i + 1];
        if (i == dg) { pivs[cn] = v0; v0 = 1.0; }
        if (i + 1 == dg) { pivs[cn] = v1; v1 = 1.0; }
        u2[i >> 1] = make_double2(v0, v1);
    }
}

__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_kernel(FactorView v, double* DLw, double* DUw, int jb, double* logdet_part /* [batch][nb] */,
                  int* info /* [batch] */) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                     // write-back staging only
    double* ucol = sm + NB * DPITCH;    // [2][UBUF], 16-byte aligned (NB * DPITCH is even)
    double* pivs = ucol + 2 * UBUF;     // pivots of all columns
    double* invd = pivs + NB;           // 1 / L_cc
    double* red = invd + NB;            // 8 doubles of reduction scratch, then 2 mbarriers
    const unsigned ready0 = smem_u32(red + 8);      // ready[2]
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int bz = blockIdx.x;
    double* blk = v.F + (size_t)bz * v.fstride + (size_t)jb * NB * v.np + (size_t)jb * NB;

    if (tid == 0) {
        mbar_init(ready0, 16); mbar_init(ready0 + 8, 16);
    }
    double reg[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int a = ty + 16 * i, b = tx + 16 * j;
            reg[i][j] = (a >= b) ? blk[(size_t)a * v.np + b] : 0.0;
        }
    const bool p_low = ty >= tx;
    __syncthreads();   // barriers initialised
    if (tx == 0) {     // publish column 0
        double val[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) val[i] = reg[i][0];
        diag_publish(val, ty == 0 ? 0 : -1, ucol + ty * UP, pivs, 0);
        mbar_arrive(ready0);
    }

#pragma unroll
    for (int jc = 0; jc < 8; ++jc) {
        for (int cp = 0; cp < 8; ++cp) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {      // unrolled by two: the buffer index c & 1 == h is static
                const int cc = 2 * cp + h;
                const int c = jc * 16 + cc;
                mbar_wait(ready0 + 8 * h, cp & 1);             // tenant c >> 1 = jc * 8 + cp of buffer h
                const double ipiv = fast_rcp(pivs[c]);
                double ua[8], ub[8];
                {
                    const double2* pa = reinterpret_cast<const double2*>(ucol + h * UBUF + ty * UP);
                    const double2* pb = reinterpret_cast<const double2*>(ucol + h * UBUF + tx * UP);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double2 qa = pa[k], qb = pb[k];
                        ua[2 * k] = qa.x * ipiv; ua[2 * k + 1] = qa.y * ipiv;
                        ub[2 * k] = qb.x; ub[2 * k + 1] = qb.y;
                    }
                }
                const bool p_ale = ty <= cc;
                // columns b <= c of group jc are finished: a zero multiplier leaves them untouched (x - u * 0 == x)
                // instead of a predicate (and a select pair) per update
                const double ub_c = (tx > cc) ? ub[jc] : 0.0;
                // owners of column c+1: tx == (cc+1) & 15, two lanes of every warp
                const bool own_next = (c + 1 < NB) && (tx == ((cc + 1) & 15));
                // ---- phase 1: the column group(s) that can contain column c+1: j == jc and j == jc+1
#pragma unroll
                for (int j = jc; j < 8 && j <= jc + 1; ++j)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const bool lower = (i > j) || (i == j && p_low);
                        const bool a_le = (i < jc) || (i == jc && p_ale);
                        if (lower || a_le) reg[i][j] -= ua[i] * (j == jc ? ub_c : ub[j]);
                    }
                if (own_next) {
                    const int cn = c + 1;
                    double* un = ucol + (1 - h) * UBUF + ty * UP;
                    double val[8];
                    if (cc < 15) {      // column c+1 is in register column jc; its diagonal row is ty == cc+1, i == jc
#pragma unroll
                        for (int i = 0; i < 8; ++i) val[i] = reg[i][jc];
                        diag_publish(val, ty == cc + 1 ? jc : -1, un, pivs, cn);
                    } else if (jc < 7) {   // first column of the next group: diagonal row ty == 0, i == jc+1
#pragma unroll
                        for (int i = 0; i < 8; ++i) val[i] = reg[i][(jc + 1) & 7];
                        diag_publish(val, ty == 0 ? jc + 1 : -1, un, pivs, cn);
                    }
                    mbar_arrive(ready0 + 8 * (1 - h));
                }
                // ---- phase 2: all remaining column groups
#pragma unroll
                for (int j = jc + 2; j < 8; ++j)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const bool lower = (i > j) || (i == j && p_low);
                        const bool a_le = (i < jc) || (i == jc && p_ale);
                        if (lower || a_le) reg[i][j] -= ua[i] * ub[j];
                    }
            }
        }
    }
    __syncthreads();
    if (tid < NB) invd[tid] = 1.0 / sqrt(pivs[tid]);
    __syncthreads();
    // first bad pivot (if any), LAPACK style
    if (tid == 0) {
        for (int c = 0; c < NB; ++c)
            if (!(pivs[c] > 0.0)) { atomicCAS(&info[bz], 0, jb * NB + c + 1); break; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int a = ty + 16 * i, b = tx + 16 * j;
            S[a * DPITCH + b] = (a == b) ? sqrt(pivs[a]) : reg[i][j] * invd[b];
        }
    __syncthreads();

    // write back L (lower triangle only), DL = Z, DU = Z^T
    double* dl = DLw + (size_t)bz * v.dstride + (size_t)jb * NB * NB;
    double* du = DUw + (size_t)bz * v.dstride + (size_t)jb * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += DIAG_THREADS) {
        const int r = idx >> 7, c = idx & (NB - 1);
        if (c <= r) blk[(size_t)r * v.np + c] = S[r * DPITCH + c];
        dl[idx] = (c < r) ? S[c * DPITCH + r] : (c == r ? invd[r] : 0.0);
        du[idx] = (c > r) ? S[r * DPITCH + c] : (c == r ? invd[r] : 0.0);
    }
    // log det part: sum_c log L_cc = 1/2 sum_c log piv_c
    double lg = (tid < NB) ? 0.5 * log(pivs[tid]) : 0.0;
    lg = block_sum(lg, red);
    if (tid == 0 && logdet_part) logdet_part[(size_t)bz * v.nb + jb] = lg;
}

// Diagonal-block kernel, second generation: one CTA per matrix, block resident in shared memory as 32 x 32 sub-blocks;
// in-warp factorisation of the sub-blocks, substitutions with one lane per right-hand side, DMMA for the trailing
// updates and for the inverse (chol128.cuh).
__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag2_kernel(FactorView v, double* DLw, double* DUw, int jb, double* logdet_part, int* info) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, bz = blockIdx.x;
    double* blk = v.F + (size_t)bz * v.fstride + (size_t)jb * NB * v.np + (size_t)jb * NB;
    c128::load_block(sm, blk, v.np);
    __syncthreads();
    c128::factor_invert(sm, [&] { c128::store_L(sm, blk, v.np); });
    c128::store_inverse(sm, DLw + (size_t)bz * v.dstride + (size_t)jb * NB * NB, DUw + (size_t)bz * v.dstride + (size_t)jb * NB * NB);
    const double* pivs = sm + c128::OFF_PIV;
    if (tid == 0) {
        for (int c = 0; c < NB; ++c)
            if (!(pivs[c] > 0.0)) { atomicCAS(&info[bz], 0, jb * NB + c + 1); break; }
    }
    double lg = (tid < NB) ? 0.5 * log(pivs[tid]) : 0.0;
    lg = block_sum(lg, sm + c128::OFF_RED);
    if (tid == 0 && logdet_part) logdet_part[(size_t)bz * v.nb + jb] = lg;
}

static cudaError_t diag_configure() {
    static std::atomic<bool> done[MAX_DEVICES];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= MAX_DEVICES) dev = 0;
    if (done[dev].load(std::memory_order_acquire)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(potrf_diag2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c128::SMEM_BYTES);
    if (e == cudaSuccess) done[dev].store(true, std::memory_order_release);
    return e;
}

// Two-level blocking: panels of `pw` block columns are factored left-looking (block column j is
// first updated with the panel's earlier columns, K = 128 (j - j0), then its diagonal block is
// factored and the rows below solved); the matrix to the right of the panel then gets ONE trailing
// update with K = 128 pw, which halves / quarters the C read-modify-write traffic and the number of
// pipeline fills per flop compared with a rank-128 update per block column.
//
// Look-ahead (la.panel != nullptr): the serial panel chain (column update -> ~50 us diagonal kernel -> panel
// solve, ~150 us per block column) runs on the high-priority stream la.panel; the trailing update of panel k
// is split into the NEXT panel's block columns (on la.panel, so that panel k+1 can start at once) and the
// bulk (on `stream`), and panel k+1 is factored while the bulk of update k is still running:
//     panel:  factor k | wait bulk k-1 | update_k(cols of panel k+1) | factor k+1 | ...
//     bulk :            wait factor k  | update_k(cols right of panel k+1) | ...
// Events are re-recorded every panel (a wait refers to the record that precedes it in host order).
cudaError_t potrf_batched(const FactorView& v, double* DLw, double* DUw, int batch, double* logdet_part,
                          int* info, int pw, cudaStream_t stream, const Lookahead& la, int* sync, bool fused_inverse) {
    if (sync && potrf_use_pll()) return potrf_pll(v, DLw, DUw, batch, logdet_part, info, sync, stream, fused_inverse);   // callers apply the size rule
    if (fused_inverse) return cudaErrorNotSupported;
    cudaError_t e = diag_configure();
    if (e != cudaSuccess) return e;
    if (pw < 1) pw = batch >= 8 ? 16 : 8;   // auto: wide panels pay once the batch alone fills the SMs
    GemmCtx ctx;
    {
        GemmSrcs srcs;
        int rows[NSRC];
        factor_srcs(v, srcs, rows);
        if ((e = gemm_make_ctx(ctx, srcs, rows, batch)) != cudaSuccess) return e;
    }
    // panel boundaries: width pw, narrowed to tail_w once at most tail_n block columns remain (there the
    // trailing update is too small to hide the panel chain, whose column updates shorten with the panel)
    int bounds[LCGP_MAX_PANELS + 2];
    int npanels = 0;
    {
        const int tail_n = potrf_tail_n(), tail_w = potrf_tail_w();
        int j = 0;
        bounds[0] = 0;
        while (j < v.nb) {
            int w = (v.nb - j <= tail_n && tail_w < pw) ? tail_w : pw;
            if (npanels + 1 >= LCGP_MAX_PANELS) w = v.nb - j;   // cannot happen for nb <= 8 * LCGP_MAX_PANELS
            j = (j + w < v.nb) ? j + w : v.nb;
            bounds[++npanels] = j;
        }
    }
    const bool look = la.panel != nullptr && npanels > 1;
    cudaStream_t ps = look ? la.panel : stream;
    if (look) {   // the panel stream starts after everything already queued on `stream`
        if ((e = cudaEventRecord(la.ev_bulk, stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(ps, la.ev_bulk, 0)) != cudaSuccess) return e;
    }
    bool bulk_pending = false;   // la.ev_bulk holds a bulk update the panel stream has not yet waited for
    for (int k = 0; k < npanels; ++k) {
        const int j0 = bounds[k], j1 = bounds[k + 1];
        for (int j = j0; j < j1; ++j) {
            if (j > j0) {
                SyrkJob::Params cp{v, j0, j, 0, j, 0};
                if (ctx.tma && (v.nb - j) * batch <= half_tile_limit())
                    e = gemm_launch<SyrkJob, NB / 2>(ctx, cp, dim3(v.nb - j, batch, 2), ps, j - j0);
                else
                    e = gemm_launch<SyrkJob>(ctx, cp, dim3(v.nb - j, batch, 1), ps, j - j0);
                if (e != cudaSuccess) return e;
            }
            note_launch();
            if (diag_use_v2()) potrf_diag2_kernel<<<batch, DIAG_THREADS, c128::SMEM_BYTES, ps>>>(v, DLw, DUw, j, logdet_part, info);
            else potrf_diag_kernel<<<batch, DIAG_THREADS, DIAG_SMEM, ps>>>(v, DLw, DUw, j, logdet_part, info);
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            const int T = v.nb - j - 1;
            if (T > 0) {
                TrsmJob::Params tp{v, j};
                if (ctx.tma && T * batch <= half_tile_limit())
                    e = gemm_launch<TrsmJob, NB / 2>(ctx, tp, dim3(T, batch, 2), ps, 1);
                else
                    e = gemm_launch<TrsmJob>(ctx, tp, dim3(T, batch, 1), ps, 1);
                if (e != cudaSuccess) return e;
            }
        }
        const int T = v.nb - j1;
        if (T <= 0) break;
        // trailing updates with few tiles (the tail of every factorisation) are on the chain too: half tiles
        auto syrk = [&](const SyrkJob::Params& sp, int tiles, cudaStream_t st) {
            if (ctx.tma && tiles * batch <= half_tile_limit())
                return gemm_launch<SyrkJob, NB / 2>(ctx, sp, dim3(tiles, batch, 2), st, j1 - j0);
            return gemm_launch<SyrkJob>(ctx, sp, dim3(tiles, batch, 1), st, j1 - j0);
        };
        if (!look) {
            SyrkJob::Params sp{v, j0, j1, j1, -1, 0};
            e = syrk(sp, T * (T + 1) / 2, stream);
            if (e != cudaSuccess) return e;
            continue;
        }
        const int j2 = bounds[k + 2];                                                  // end of the next panel
        if ((e = cudaEventRecord(la.ev_panel, ps)) != cudaSuccess) return e;            // panel k is final
        if (bulk_pending) {                                                            // bulk k-1 wrote these columns
            if ((e = cudaStreamWaitEvent(ps, la.ev_bulk, 0)) != cudaSuccess) return e;
            bulk_pending = false;
        }
        SyrkJob::Params lp{v, j0, j1, j1, -2, j2};
        e = syrk(lp, (j2 - j1) * T, ps);
        if (e != cudaSuccess) return e;
        const int T2 = v.nb - j2;
        if (T2 > 0) {
            if ((e = cudaStreamWaitEvent(stream, la.ev_panel, 0)) != cudaSuccess) return e;
            SyrkJob::Params sp{v, j0, j1, j2, -1, 0};
            e = syrk(sp, T2 * (T2 + 1) / 2, stream);
            if (e != cudaSuccess) return e;
            if ((e = cudaEventRecord(la.ev_bulk, stream)) != cudaSuccess) return e;
            bulk_pending = true;
        }
    }
    if (look) {   // join: `stream` continues after the last panel
        if ((e = cudaEventRecord(la.ev_panel, ps)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(stream, la.ev_panel, 0)) != cudaSuccess) return e;
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
    GemmSrcs srcs;
    int rows[NSRC];
    factor_srcs(v, srcs, rows);
    for (int s = 1; s < v.nb; s *= 2) {
        const int merges = (v.nb + 2 * s - 1) / (2 * s);
        srcs.base[SRC_T] = scratch; srcs.ld[SRC_T] = s * NB; srcs.bstride[SRC_T] = tstride; rows[SRC_T] = merges * s * NB;
        GemmCtx ctx;
        cudaError_t e = gemm_make_ctx(ctx, srcs, rows, batch);
        if (e != cudaSuccess) return e;
        TrtriParams p{v, scratch, tstride, s};
        dim3 grid(s * s, batch, merges);
        e = gemm_launch<TrtriG1Job>(ctx, p, grid, stream, s);
        if (e != cudaSuccess) return e;
        e = gemm_launch<TrtriG2Job>(ctx, p, grid, stream, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace lcgp
