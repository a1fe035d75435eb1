// Cholesky factor AND inverse of one 128 x 128 diagonal block by ONE CTA (256 threads), shared-memory resident.
//
// The block is held as 32 x 32 sub-blocks (pitch 36 doubles: DMMA fragment loads of K-major and N-major operands
// are bank-conflict free for any pitch = 4 mod 16).  Blocked right-looking factorisation over 4 sub-block columns:
//   panel  the tall panel [diagonal sub-block ; rows below ; I_32] is eliminated column by column by all 256 threads
//          (panel32 below): the rows below come out as X = A L_ss^-T and the identity rows as the inverse of the
//          diagonal sub-block, so there is no separate triangular solve / inverse pass
//   P3     trailing update S -= X X^T on the FP64 tensor pipe (DMMA m8n8k4 from shared memory)
// and then the inverse of the whole block by two levels of block merges W21 = -W22 (L21 W11), again on DMMA.
// Everything is CTA-local (__syncthreads between phases).  Measured on B200 (tools/ubench/diag_phases.cu, cycles at
// 1.965 GHz): load 7 k, panels 4 x 10.5 k, P3 9 k, inverse merges 10 k, stores (L, DL, DU) 16 k: ~43 us per block, of
// which the stores of L and DU can follow the hand-off of DL (was 58 us for the register-resident column-by-column
// kernel).  What was tried on the way: one warp per 32 x 32 sub-block with 64-bit shuffles (12.2 k cycles per
// sub-block: a 64-bit SHFL costs ~13 issue cycles) or with a shared-memory column broadcast (6.7 k) plus per-lane
// forward substitutions (2.7 k): one warp can issue a DFMA only every ~6 cycles, so the rank-1 FMAs must be spread
// over all warps; what remains is the dependent chain LDS -> reciprocal (47) -> multiply -> FMA -> STS -> barrier.
#pragma once
#include "common.cuh"

#ifndef C128_STAMP              // profiling hook of tools/ubench/diag_phases.cu: records a time stamp per phase
#define C128_STAMP(id)
#endif

namespace lcgp {
namespace c128 {

constexpr int SB = 32;                 // sub-block edge
constexpr int BP = 36;                 // sub-block pitch (doubles)
constexpr int BLK = SB * BP;           // doubles per sub-block
constexpr int NBLK = 10;               // lower-triangular 4 x 4 arrangement
constexpr int LTP = 34;                // pitch of the transposed diagonal sub-block used by the substitutions
// layout (doubles): S blocks | W blocks | pivs[128] | invd[128] | red[16]
constexpr int OFF_S = 0;
constexpr int OFF_W = NBLK * BLK;
constexpr int OFF_PIV = 2 * NBLK * BLK;
constexpr int OFF_INVD = OFF_PIV + NB;
constexpr int OFF_RED = OFF_INVD + NB;
constexpr int OFF_COL = OFF_RED + 16;           // column broadcast buffers of panel32 (2 x 160)
constexpr int SMEM_DOUBLES = OFF_COL + 2 * 160;
constexpr size_t SMEM_BYTES = sizeof(double) * SMEM_DOUBLES;   // 189,056 B

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// ---- Panel factorisation of sub-block column S by the whole CTA (256 threads) ------------------------------------
// The tall panel  [ S(s,s) ; S(s+1..3, s) ; I_32 ]  (NR = 160 - 32 s rows x 32 columns) is eliminated column by column
// (right-looking, unscaled "LDL^T in flight" form: a_ij -= u_i u_j / d, the square roots are applied at the end).
// Rows below the diagonal sub-block come out as X = A L_ss^-T, the identity rows as L_ss^-T = W_ss^T -- so the
// panel solve and the inverse of the diagonal sub-block cost no separate pass.
// Thread (rg = tid / TXP, tx = tid % TXP) owns ROWS rows rho = rg + (256 / TXP) m and 32 / TXP columns j = tx + TXP n,
// with (ROWS, TXP) = (5, 8), (2, 4), (3, 8), (1, 4) for s = 0..3: every thread owns exactly NR * 32 / 256 elements, so
// no FP64 issue slot is spent on padding (a warp can issue a DFMA only every ~6 cycles: the issue slots, not the
// dependent chain, bound this phase; the first version, 3 rows x 8 columns for every s, took the same 10.5 k cycles
// for 64 rows as for 160).  Column c is published through shared memory (double buffered), ONE CTA barrier per
// column; the owners of column c + 1 update and publish it before the rest of the rank-1 update.
// Dependent chain per column:  LDS -> reciprocal (47 cycles) -> multiply -> FMA -> STS -> barrier.
constexpr int CBP = 160;                // doubles per column buffer
__device__ __forceinline__ void bar_cta() { __syncthreads(); }

template <int S>
__device__ __forceinline__ void panel32(double* __restrict__ sm) {
    constexpr int NBELOW = (3 - S) * SB;            // rows below the diagonal sub-block
    constexpr int NR = 2 * SB + NBELOW;             // + diagonal sub-block + identity rows
    constexpr int TXP = (S & 1) ? 4 : 8;            // column phases
    constexpr int LOGT = (S & 1) ? 2 : 3;
    constexpr int RG = 256 / TXP;                   // row groups
    constexpr int ROWS = NR / RG;                   // 5, 2, 3, 1
    constexpr int COLS = SB / TXP;                  // 4, 8, 4, 8
    static_assert(ROWS * RG == NR, "rows must divide evenly");
    double* Sb = sm + OFF_S;
    double* W = sm + OFF_W;
    double* pivs = sm + OFF_PIV + S * SB;
    double* invd = sm + OFF_INVD + S * SB;
    double* colbuf = sm + OFF_COL;
    const int tid = threadIdx.x, rg = tid / TXP, tx = tid % TXP;
    double a[ROWS][COLS];
    double* rowp[ROWS];                             // shared-memory row of a data row (null: identity row)
    int rho[ROWS];
#pragma unroll
    for (int m = 0; m < ROWS; ++m) {
        rho[m] = rg + RG * m;
        rowp[m] = nullptr;
        if (rho[m] < SB) rowp[m] = Sb + tri(S, S) * BLK + rho[m] * BP;
        else if (rho[m] < SB + NBELOW) rowp[m] = Sb + tri(S + 1 + ((rho[m] - SB) >> 5), S) * BLK + ((rho[m] - SB) & 31) * BP;
#pragma unroll
        for (int n = 0; n < COLS; ++n) {
            const int j = tx + TXP * n;
            a[m][n] = rowp[m] ? rowp[m][j] : ((rho[m] - SB - NBELOW == j) ? 1.0 : 0.0);
        }
    }
    if (tx == 0) {                                  // publish column 0
#pragma unroll
        for (int m = 0; m < ROWS; ++m) colbuf[rho[m]] = a[m][0];
    }
    bar_cta();
    // software pipelined by one column: the loads and the reciprocal of column c + 1 are issued before the bulk of
    // column c's rank-1 update
    double d = colbuf[0];
    double tt[ROWS], u[COLS];
#pragma unroll
    for (int m = 0; m < ROWS; ++m) tt[m] = colbuf[rho[m]];
#pragma unroll
    for (int n = 0; n < COLS; ++n) u[n] = colbuf[tx + TXP * n];
    double rd = fast_rcp(d);
#pragma unroll
    for (int c = 0; c < SB; ++c) {
        double* nb_ = colbuf + ((c + 1) & 1) * CBP;
        const int n0 = c >> LOGT;                   // column groups n < n0 are finished; n == n0 is active iff tx > c % TXP
        const int n1 = (c + 1) >> LOGT;             // group of column c + 1
        double t[ROWS];
#pragma unroll
        for (int m = 0; m < ROWS; ++m) t[m] = tt[m] * rd;
        if (tid == 0) pivs[c] = d;
        if (c + 1 < SB) {                           // column c + 1 first; its owners publish it at once
            if ((n1 > n0) || (tx > (c & (TXP - 1)))) {
#pragma unroll
                for (int m = 0; m < ROWS; ++m) a[m][n1] = fma(-t[m], u[n1], a[m][n1]);
            }
            if (tx == ((c + 1) & (TXP - 1))) {
#pragma unroll
                for (int m = 0; m < ROWS; ++m) nb_[rho[m]] = a[m][n1];
            }
        }
        bar_cta();
        double ucur[COLS];
#pragma unroll
        for (int n = 0; n < COLS; ++n) ucur[n] = u[n];
        if (c + 1 < SB) {                           // next column: loads and reciprocal, ahead of this column's bulk
            d = nb_[c + 1];
#pragma unroll
            for (int m = 0; m < ROWS; ++m) tt[m] = nb_[rho[m]];
#pragma unroll
            for (int n = n1; n < COLS; ++n) u[n] = nb_[tx + TXP * n];
            rd = fast_rcp(d);
        }
#pragma unroll
        for (int n = n0; n < COLS; ++n) {           // the rest of the rank-1 update (registers only)
            if (n == n1 && c + 1 < SB) continue;
            if ((n > n0) || (tx > (c & (TXP - 1)))) {
#pragma unroll
                for (int m = 0; m < ROWS; ++m) a[m][n] = fma(-t[m], ucur[n], a[m][n]);
            }
        }
    }
    // 1 / L_cc = rsqrt(d_c), once per column
    if (tid < SB) invd[tid] = rsqrt(pivs[tid]);
    bar_cta();
    double* Wd = W + tri(S, S) * BLK;
#pragma unroll
    for (int m = 0; m < ROWS; ++m) {
#pragma unroll
        for (int n = 0; n < COLS; ++n) {
            const int j = tx + TXP * n;
            const double v = a[m][n] * invd[j];
            if (rho[m] < SB) rowp[m][j] = (j <= rho[m]) ? v : 0.0;              // L_ss (zeros above the diagonal)
            else if (rowp[m]) rowp[m][j] = v;                                    // X = A L_ss^-T
            else {                                                               // identity row i: W_ss[j][i]
                const int i = rho[m] - SB - NBELOW;
                Wd[j * BP + i] = (j >= i) ? v : 0.0;
            }
        }
    }
}

// ---- The same panel factorisation in COMPACT form: one instantiation (3 rows x 8 columns per thread, rows beyond the
// panel predicated) for every s.  It wastes FP64 issue slots (10.5 k cycles per step whatever s) but its code is a
// quarter of the size of the four panel32<S> instantiations.  That matters inside the persistent Cholesky kernel when
// the factor buffers do not fit in L2: the diagonal-block code runs once in a while per SM, is evicted from L2 by the
// operand streams in between, and is then re-fetched line by line (measured, same kernel otherwise: 2048 x 10
// matrices 1.60 ms compact vs 1.72 ms specialised; 1024 x 1 0.53 vs 0.50 ms).

// The tall panel  [ S(s,s) ; S(s+1..3, s) ; I_32 ]  (160 - 32 s rows x 32 columns) is eliminated column by column
// (right-looking, unscaled "LDL^T in flight" form: a_ij -= u_i u_j / d, the square roots are applied at the end).
// Rows below the diagonal sub-block come out as X = A L_ss^-T, the identity rows as L_ss^-T = W_ss^T -- so the
// panel solve and the inverse of the diagonal sub-block cost no separate pass.
// Thread (rg = tid / 4, tx = tid % 4) owns rows rho = rg + 64 m (m < 3) and columns j = tx + 4 n (n < 8): the columns
// still active at step c are spread evenly over the threads (one warp can issue a DFMA only every ~6 cycles, so the
// rank-1 updates must be spread over all warps).  Column c is published through shared memory (double buffered),
// ONE CTA barrier per column; the owners of column c + 1 update and publish it before their other updates.
// Dependent chain per column:  LDS -> reciprocal -> multiply -> FMA -> STS -> barrier (~130 cycles); measured ~300 per
// column including the rank-1 FMAs (a warp issues one DFMA per ~6 cycles; 21 per thread and column at most).

// The 32 column steps for a thread that owns mr <= MR rows (rho[0..mr)) x 8 columns (tx + 4 n).  ONE instantiation for
// the whole CTA: per-warp instantiations by row count (fewer wasted FMAs) were measured and gave wrong results --
// barriers reached from different code copies did not order the column buffers -- for a 15 % gain at best.
template <int MR>
__device__ __forceinline__ void panel32c_cols(double (&a)[3][8], const int (&rho)[3], int tx, double* __restrict__ colbuf,
                                             double* __restrict__ pivs, int mr) {
#pragma unroll
    for (int c = 0; c < SB; ++c) {
        const double* cb = colbuf + (c & 1) * CBP;
        double* nb_ = colbuf + ((c + 1) & 1) * CBP;
        const int n0 = c >> 2;                      // column groups n < n0 are finished; n == n0 is active iff tx > (c & 3)
        const int n1 = (c + 1) >> 2;                // group of column c + 1
        const double d = cb[c];
        double t[MR], u[8];
#pragma unroll
        for (int m = 0; m < MR; ++m) t[m] = (m < mr) ? cb[rho[m]] : 0.0;
#pragma unroll
        for (int n = n0; n < 8; ++n) u[n] = cb[tx + 4 * n];      // all reads of this buffer happen before the barrier
        const double rd = fast_rcp(d);
#pragma unroll
        for (int m = 0; m < MR; ++m) t[m] *= rd;
        if (threadIdx.x == 0) pivs[c] = d;
        if (c + 1 < SB) {                           // column c + 1 first; its owners publish it at once
            if ((n1 > n0) || (tx > (c & 3))) {
#pragma unroll
                for (int m = 0; m < MR; ++m) a[m][n1] = fma(-t[m], u[n1], a[m][n1]);
            }
            if (tx == ((c + 1) & 3)) {
#pragma unroll
                for (int m = 0; m < MR; ++m)
                    if (m < mr) nb_[rho[m]] = a[m][n1];
            }
        }
        bar_cta();
        // the rest of the rank-1 update (registers only): overlaps the next column's loads and reciprocal
#pragma unroll
        for (int n = n0; n < 8; ++n) {
            if (n == n1 && c + 1 < SB) continue;
            if ((n > n0) || (tx > (c & 3))) {
#pragma unroll
                for (int m = 0; m < MR; ++m) a[m][n] = fma(-t[m], u[n], a[m][n]);
            }
        }
    }
}

__device__ __forceinline__ void panel32_compact(double* __restrict__ sm, int s) {
    double* S = sm + OFF_S;
    double* W = sm + OFF_W;
    double* pivs = sm + OFF_PIV + s * SB;
    double* invd = sm + OFF_INVD + s * SB;
    double* colbuf = sm + OFF_COL;
    const int tid = threadIdx.x, rg = tid >> 2, tx = tid & 3;
    const int nbelow = (3 - s) * SB;                // rows below the diagonal sub-block
    const int NR = 2 * SB + nbelow;                 // + diagonal sub-block + identity rows: 160, 128, 96, 64
    const int mr = (NR + 63 - rg) >> 6;             // rows rho = rg + 64 m < NR owned by this thread: warp-uniform, 1..3
    double a[3][8];
    double* rowp[3];                                // shared-memory row of a data row (null: identity row / unused)
    int rho[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        rho[m] = rg + 64 * m;
        const bool valid = m < mr;
        rowp[m] = nullptr;
        if (valid && rho[m] < SB) rowp[m] = S + tri(s, s) * BLK + rho[m] * BP;
        else if (valid && rho[m] < SB + nbelow) rowp[m] = S + tri(s + 1 + ((rho[m] - SB) >> 5), s) * BLK + ((rho[m] - SB) & 31) * BP;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int j = tx + 4 * n;
            a[m][n] = rowp[m] ? rowp[m][j] : ((valid && rho[m] - SB - nbelow == j) ? 1.0 : 0.0);
        }
    }
    if (tx == 0) {                                  // publish column 0
#pragma unroll
        for (int m = 0; m < 3; ++m)
            if (m < mr) colbuf[rho[m]] = a[m][0];
    }
    bar_cta();
    panel32c_cols<3>(a, rho, tx, colbuf, pivs, mr);
    // 1 / L_cc = rsqrt(d_c), once per column
    if (tid < SB) invd[tid] = rsqrt(pivs[tid]);
    bar_cta();
    double* Wd = W + tri(s, s) * BLK;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        if (m >= mr) continue;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int j = tx + 4 * n;
            const double v = a[m][n] * invd[j];
            if (rho[m] < SB) rowp[m][j] = (j <= rho[m]) ? v : 0.0;              // L_ss (zeros above the diagonal)
            else if (rowp[m]) rowp[m][j] = v;                                    // X = A L_ss^-T
            else {                                                               // identity row i: W_ss[j][i]
                const int i = rho[m] - SB - nbelow;
                Wd[j * BP + i] = (j >= i) ? v : 0.0;
            }
        }
    }
}

// ---- DMMA unit: acc (16 x 32) += A (16 rows x 32 k, K-major, pitch BP) * B
//   BN == false : B given as 32 rows (n) x 32 k, K-major       (acc += A B^T)
//   BN == true  : B given as 32 rows (k) x 32 n, N-major       (acc += A B)
// lane = 4 g + t :  a = A[g][t], b = B[k = t][n = g], d = D[g][2t, 2t+1]
template <bool BN>
__device__ __forceinline__ void mma_unit(const double* __restrict__ A, const double* __restrict__ B,
                                         double (&acc)[2][4][2], int g, int t) {
#pragma unroll
    for (int kk = 0; kk < SB / 4; ++kk) {
        double a[2], b[4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) a[mi] = A[(mi * 8 + g) * BP + kk * 4 + t];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
            b[ni] = BN ? B[(kk * 4 + t) * BP + ni * 8 + g] : B[(ni * 8 + g) * BP + kk * 4 + t];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}

__device__ __forceinline__ void zero_acc(double (&acc)[2][4][2]) {
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
}

// C (16 rows x 32, pitch BP) = sgn * acc   or   C += sgn * acc
template <bool ACCUM>
__device__ __forceinline__ void store_unit(double* __restrict__ C, const double (&acc)[2][4][2], double sgn, int g, int t) {
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            double2* q = reinterpret_cast<double2*>(C + (mi * 8 + g) * BP + ni * 8 + 2 * t);
            double2 v = ACCUM ? *q : make_double2(0.0, 0.0);
            v.x = fma(sgn, acc[mi][ni][0], v.x);
            v.y = fma(sgn, acc[mi][ni][1], v.y);
            *q = v;
        }
}

// Factor + invert the block held in sm (layout above; S blocks loaded, upper triangle of diagonal sub-blocks zero).
// On return: S blocks (i > j) and the lower triangles of S(i,i) hold L -- but the diagonal S blocks are CLOBBERED by
// the inverse phase, so `write_L` is invoked (by all threads, after a barrier) between the two phases to save L;
// W blocks hold inv(L) (lower triangular, zeros above the diagonal of the diagonal sub-blocks); pivs / invd filled.
// compact: use panel32_compact (small code) instead of the per-step specialisations
template <class WriteL>
__device__ __forceinline__ void factor_invert(double* __restrict__ sm, WriteL write_L, bool compact = false) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    double* S = sm + OFF_S;
    double* W = sm + OFF_W;

#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        C128_STAMP(4 * s + 0);
        // ---- P1 + P2: panel factorisation (diagonal sub-block, rows below, inverse of the sub-block)
        if (compact) panel32_compact(sm, s);
        else switch (s) {                                            // (ROWS, TXP) differ per step: see panel32
            case 0: panel32<0>(sm); break;
            case 1: panel32<1>(sm); break;
            case 2: panel32<2>(sm); break;
            default: panel32<3>(sm); break;
        }
        __syncthreads();
        C128_STAMP(4 * s + 1);
        C128_STAMP(4 * s + 2);
        // ---- P3: trailing update S(bi,bj) -= X_bi X_bj^T, units of 16 rows x 32 columns
        {
            const int nt = 3 - s;                                   // trailing sub-block rows
            const int nunits = nt * (nt + 1);                        // 2 per lower-triangular sub-block
            for (int u = warp; u < nunits; u += 8) {
                const int blk = u >> 1, half = u & 1;
                int bi = 0, bj = blk;                                // decode blk -> (bi >= bj) in the nt x nt triangle
                while (bj > bi) { bj -= bi + 1; ++bi; }
                const int I = s + 1 + bi, J = s + 1 + bj;
                double acc[2][4][2];
                zero_acc(acc);
                mma_unit<false>(S + tri(I, s) * BLK + half * 16 * BP, S + tri(J, s) * BLK, acc, g, t);
                store_unit<true>(S + tri(I, J) * BLK + half * 16 * BP, acc, -1.0, g, t);
            }
        }
        __syncthreads();
    }
    // ---- L is complete: let the caller save it before the diagonal S blocks become scratch
    C128_STAMP(16);
    write_L();
    __syncthreads();
    C128_STAMP(17);
    // ---- inverse, level 1: W(1,0) = -W(1,1) (L(1,0) W(0,0)),  W(3,2) = -W(3,3) (L(3,2) W(2,2))
    {
        double acc[2][4][2];
        if (warp < 4) {
            const int m = warp >> 1, half = warp & 1;               // merge m: blocks (2m+1, 2m)
            zero_acc(acc);
            mma_unit<true>(S + tri(2 * m + 1, 2 * m) * BLK + half * 16 * BP, W + tri(2 * m, 2 * m) * BLK, acc, g, t);
            store_unit<false>(S + tri(2 * m, 2 * m) * BLK + half * 16 * BP, acc, 1.0, g, t);    // T in S(2m,2m)
        }
        __syncthreads();
        if (warp < 4) {
            const int m = warp >> 1, half = warp & 1;
            zero_acc(acc);
            mma_unit<true>(W + tri(2 * m + 1, 2 * m + 1) * BLK + half * 16 * BP, S + tri(2 * m, 2 * m) * BLK, acc, g, t);
            store_unit<false>(W + tri(2 * m + 1, 2 * m) * BLK + half * 16 * BP, acc, -1.0, g, t);
        }
        __syncthreads();
        C128_STAMP(18);
        // ---- level 2: T2 = L[2:4][0:2] W[0:2][0:2] (into the diagonal S blocks), W[2:4][0:2] = -W[2:4][2:4] T2
        // unit = warp: output block (a, b) = (2 + (warp >> 2), (warp >> 1) & 1), half = warp & 1;  T2(a,b) lives in S(2(a-2)+b, same)
        const int a2 = 2 + (warp >> 2), b2 = (warp >> 1) & 1, half = warp & 1;
        zero_acc(acc);
        if (b2 == 0) {
            mma_unit<true>(S + tri(a2, 0) * BLK + half * 16 * BP, W + tri(0, 0) * BLK, acc, g, t);
            mma_unit<true>(S + tri(a2, 1) * BLK + half * 16 * BP, W + tri(1, 0) * BLK, acc, g, t);
        } else {
            mma_unit<true>(S + tri(a2, 1) * BLK + half * 16 * BP, W + tri(1, 1) * BLK, acc, g, t);
        }
        {
            const int tb = 2 * (a2 - 2) + b2;
            store_unit<false>(S + tri(tb, tb) * BLK + half * 16 * BP, acc, 1.0, g, t);
        }
        __syncthreads();
        zero_acc(acc);
        mma_unit<true>(W + tri(a2, 2) * BLK + half * 16 * BP, S + tri(b2, b2) * BLK, acc, g, t);              // W(a,2) T2(2,b)
        if (a2 == 3) mma_unit<true>(W + tri(3, 3) * BLK + half * 16 * BP, S + tri(2 + b2, 2 + b2) * BLK, acc, g, t);  // W(3,3) T2(3,b)
        store_unit<false>(W + tri(a2, b2) * BLK + half * 16 * BP, acc, -1.0, g, t);
        __syncthreads();
        C128_STAMP(19);
    }
}

// ---- global <-> shared ----------------------------------------------------------------------------------
// Loads rows [r0, r0 + NR) of the lower triangle of the 128 x 128 block at blk (row-major, ld) into the S sub-blocks
// (zeros above the diagonal of the diagonal sub-blocks).  L2 loads (the block may have been written by another SM of
// the same kernel); up to 16 loads per thread are in flight before the first is consumed.  NR = 128, 64 or 32.
template <int NR>
__device__ __forceinline__ void load_rows(double* __restrict__ sm, const double* __restrict__ blk, size_t ld, int r0) {
    double* S = sm + OFF_S;
    constexpr int PAIRS = NR * (NB / 2);
    constexpr int U = PAIRS / 256 < 16 ? PAIRS / 256 : 16;      // loads in flight per thread
    static_assert(PAIRS % (256 * U) == 0, "load_rows assumes 256 threads");
#pragma unroll 1
    for (int base = 0; base < PAIRS; base += 256 * U) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = base + u * 256 + threadIdx.x;
            const int r = r0 + (idx >> 6), c = (idx & 63) * 2;
            v[u] = (c <= r) ? __ldcg(reinterpret_cast<const double2*>(blk + (size_t)r * ld + c)) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = base + u * 256 + threadIdx.x;
            const int r = r0 + (idx >> 6), c = (idx & 63) * 2;
            if ((c >> 5) <= (r >> 5)) {             // sub-block on or below the diagonal (zeros above the diagonal inside it)
                double* q = S + tri(r >> 5, c >> 5) * BLK + (r & 31) * BP + (c & 31);
                *reinterpret_cast<double2*>(q) = make_double2(v[u].x, (c + 1 <= r) ? v[u].y : 0.0);
            }
        }
    }
}
__device__ __forceinline__ void load_block(double* __restrict__ sm, const double* __restrict__ blk, size_t ld) {
    load_rows<NB>(sm, blk, ld, 0);
}

// One accumulator pair (row r, columns c, c + 1 of the block; c even) -> S sub-blocks, as load_rows would place it
__device__ __forceinline__ void put_pair(double* __restrict__ sm, int r, int c, double v0, double v1) {
    if ((c >> 5) <= (r >> 5)) {
        double* q = sm + OFF_S + tri(r >> 5, c >> 5) * BLK + (r & 31) * BP + (c & 31);
        *reinterpret_cast<double2*>(q) = make_double2((c <= r) ? v0 : 0.0, (c + 1 <= r) ? v1 : 0.0);
    }
}

// L (lower triangle incl. diagonal) -> global block; DIAG selects the diagonal sub-blocks (which the inverse phase
// overwrites) or the sub-blocks below them (which it leaves alone)
template <bool DIAG>
__device__ __forceinline__ void store_L_part(const double* __restrict__ sm, double* __restrict__ blk, size_t ld) {
    const double* S = sm + OFF_S;
    constexpr int PAIRS = NB * (NB / 2);
#pragma unroll 1
    for (int base = 0; base < PAIRS; base += 256 * 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * 256 + threadIdx.x;
            const int r = idx >> 6, c = (idx & 63) * 2;
            if (c <= r && ((r >> 5) == (c >> 5)) == DIAG) {
                const double2 v = *reinterpret_cast<const double2*>(S + tri(r >> 5, c >> 5) * BLK + (r & 31) * BP + (c & 31));
                if (c + 1 <= r) *reinterpret_cast<double2*>(blk + (size_t)r * ld + c) = v;
                else blk[(size_t)r * ld + c] = v.x;
            }
        }
    }
}
__device__ __forceinline__ void store_L(const double* __restrict__ sm, double* __restrict__ blk, size_t ld) {
    store_L_part<true>(sm, blk, ld);
    store_L_part<false>(sm, blk, ld);
}

// DL = W (dense, zeros above the diagonal), NB x NB row-major
__device__ __forceinline__ void store_DL(const double* __restrict__ sm, double* __restrict__ dl) {
    const double* W = sm + OFF_W;
    constexpr int PAIRS = NB * (NB / 2);
#pragma unroll 1
    for (int base = 0; base < PAIRS; base += 256 * 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * 256 + threadIdx.x;
            const int r = idx >> 6, c = (idx & 63) * 2;
            double2 v = make_double2(0.0, 0.0);
            if (c <= r) {
                v = *reinterpret_cast<const double2*>(W + tri(r >> 5, c >> 5) * BLK + (r & 31) * BP + (c & 31));
                if (c + 1 > r) v.y = 0.0;
            }
            *reinterpret_cast<double2*>(dl + (size_t)r * NB + c) = v;
        }
    }
}

// DU = W^T.  Transposed read: a warp covers 4 rows x 8 columns of DU per step (64-byte global segments, conflict-free LDS)
__device__ __forceinline__ void store_DU(const double* __restrict__ sm, double* __restrict__ du) {
    const double* W = sm + OFF_W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cl = lane >> 2, rl = lane & 3;
#pragma unroll 1
    for (int p0 = warp; p0 < (NB / 4) * (NB / 8); p0 += 8 * 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int patch = p0 + u * 8;
            const int r = (patch >> 4) * 4 + rl, c = (patch & 15) * 8 + cl;
            du[r * NB + c] = (c >= r) ? W[tri(c >> 5, r >> 5) * BLK + (c & 31) * BP + (r & 31)] : 0.0;
        }
    }
}

__device__ __forceinline__ void store_inverse(const double* __restrict__ sm, double* __restrict__ dl, double* __restrict__ du) {
    store_DL(sm, dl);
    store_DU(sm, du);
}

}  // namespace c128
}  // namespace lcgp
