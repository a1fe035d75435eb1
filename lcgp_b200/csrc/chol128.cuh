// Cholesky factor AND inverse of one 128 x 128 diagonal block by ONE CTA (256 threads), shared-memory resident.
//
// The block is held as 32 x 32 sub-blocks (pitch 36 doubles: DMMA fragment loads of K-major and N-major operands
// are bank-conflict free for any pitch = 4 mod 16).  Blocked right-looking factorisation over 4 sub-block columns:
//   P1  warp 0 factors the 32 x 32 diagonal sub-block in registers (lane = row): per column ONE shuffle round
//       (the unscaled column), a reciprocal and a fused update -- the division-free "LDL^T in flight" form keeps
//       the square root off the dependent chain (~100 cycles per column instead of ~900 in the column-by-column
//       kernel this replaces, where every column was a CTA-wide hand-off)
//   P2  forward substitutions against that sub-block, one LANE per right-hand side: warp 0 solves for the identity
//       (-> the inverse of the sub-block), warps 1.. solve for the rows of the panel below (X = A L^-T)
//   P3  trailing update S -= X X^T on the FP64 tensor pipe (DMMA m8n8k4 from shared memory)
// and then the inverse of the whole block by two levels of block merges W21 = -W22 (L21 W11), again on DMMA.
// Everything is CTA-local (__syncthreads between phases); ~12 us per block instead of 58.
#pragma once
#include "common.cuh"

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
constexpr int SMEM_DOUBLES = OFF_RED + 16;
constexpr size_t SMEM_BYTES = sizeof(double) * SMEM_DOUBLES;   // 186,496 B

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// ---- P1: in-warp Cholesky of a 32 x 32 block.  Lane i holds row i (a[j], j <= i valid; 0 above the diagonal).
// On return a[j] (j <= i) = L[i][j]; pivs[c] / invd[c] (shared) = pivot d_c and 1 / L_cc.
__device__ __forceinline__ void potrf32_warp(double (&a)[SB], int lane, double* __restrict__ pivs,
                                             double* __restrict__ invd) {
#pragma unroll
    for (int c = 0; c < SB; ++c) {
        const double d = __shfl_sync(0xffffffffu, a[c], c);
        const double rd = fast_rcp(d);
        double t = a[c] * rd;                       // u_i / d  (u = unscaled column c)
        t = (lane >= c) ? t : 0.0;                  // rows above the column stay exactly zero
#pragma unroll
        for (int j = c + 1; j < SB; ++j) {
            const double uj = __shfl_sync(0xffffffffu, a[c], j);
            a[j] = fma(-t, uj, a[j]);               // a_ij -= u_i u_j / d   (meaningful for i >= j)
        }
        const double rs = rsqrt(d);                 // off the dependent chain: nothing below waits for it
        a[c] = (lane == c) ? d * rs : a[c] * rs;
        if (lane == c) { pivs[c] = d; invd[c] = rs; }
    }
}

// ---- P2: forward substitution  L x = r  against the transposed block Lt[k * LTP + i] = L[i][k]; one right-hand
// side per lane, in registers.  invd[k] = 1 / L_kk.
__device__ __forceinline__ void fwdsub32(double (&r)[SB], const double* __restrict__ Lt,
                                         const double* __restrict__ invd) {
#pragma unroll
    for (int k = 0; k < SB; ++k) {
        const double w = r[k] * invd[k];
        r[k] = w;
        if (((k + 1) & 1) && k + 1 < SB) r[k + 1] = fma(-Lt[k * LTP + k + 1], w, r[k + 1]);   // odd first row: single
#pragma unroll
        for (int i = (k + 2) & ~1; i < SB; i += 2) {       // aligned pairs: one 16-byte broadcast load each
            const double2 l2 = *reinterpret_cast<const double2*>(Lt + k * LTP + i);
            r[i] = fma(-l2.x, w, r[i]);
            r[i + 1] = fma(-l2.y, w, r[i + 1]);
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
template <class WriteL>
__device__ __forceinline__ void factor_invert(double* __restrict__ sm, WriteL write_L) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    double* S = sm + OFF_S;
    double* W = sm + OFF_W;
    double* pivs = sm + OFF_PIV;
    double* invd = sm + OFF_INVD;
    double* Lt = W + tri(3, 0) * BLK;       // scratch until the level-2 merge writes W(3,0)

#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        double* Sd = S + tri(s, s) * BLK;
        // ---- P1: diagonal sub-block (warp 0)
        if (warp == 0) {
            double a[SB];
#pragma unroll
            for (int j = 0; j < SB; ++j) a[j] = (j <= lane) ? Sd[lane * BP + j] : 0.0;
            potrf32_warp(a, lane, pivs + s * SB, invd + s * SB);
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                Sd[lane * BP + j] = (j <= lane) ? a[j] : 0.0;      // L_ss, row-major
                Lt[j * LTP + lane] = a[j];                          // transposed copy (entries i >= k are read)
            }
        }
        __syncthreads();
        // ---- P2: substitutions, one right-hand side per lane
        const int mrows = (3 - s) * SB;                             // panel rows below
        if (warp == 0) {                                            // identity -> W(s,s) = inv(L_ss), column `lane`
            double r[SB];
#pragma unroll
            for (int i = 0; i < SB; ++i) r[i] = (i == lane) ? 1.0 : 0.0;
            fwdsub32(r, Lt, invd + s * SB);
            double* Wd = W + tri(s, s) * BLK;
#pragma unroll
            for (int i = 0; i < SB; ++i) Wd[i * BP + lane] = (i >= lane) ? r[i] : 0.0;
        } else if ((warp - 1) * SB < mrows) {                        // warp w: panel sub-block (s + w, s), row `lane`
            double* Sp = S + tri(s + warp, s) * BLK;
            double r[SB];
#pragma unroll
            for (int i = 0; i < SB; ++i) r[i] = Sp[lane * BP + i];
            fwdsub32(r, Lt, invd + s * SB);
#pragma unroll
            for (int i = 0; i < SB; ++i) Sp[lane * BP + i] = r[i];
        }
        __syncthreads();
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
    write_L();
    __syncthreads();
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
        // (Lt scratch lived in W(3,0): last read in step s = 3, two barriers ago)
        store_unit<false>(W + tri(a2, b2) * BLK + half * 16 * BP, acc, -1.0, g, t);
        __syncthreads();
    }
}

// ---- global <-> shared ----------------------------------------------------------------------------------
// Loads the lower triangle of the 128 x 128 block at blk (row-major, ld) into the S sub-blocks (zeros above the diagonal
// of the diagonal sub-blocks).  L2 loads (the block may have been written by another SM of the same kernel).
__device__ __forceinline__ void load_block(double* __restrict__ sm, const double* __restrict__ blk, size_t ld) {
    double* S = sm + OFF_S;
    for (int idx = threadIdx.x; idx < NB * (NB / 2); idx += blockDim.x) {
        const int r = idx >> 6, c = (idx & 63) * 2;
        if (c > r) continue;
        const double2 v = __ldcg(reinterpret_cast<const double2*>(blk + (size_t)r * ld + c));
        double* q = S + tri(r >> 5, c >> 5) * BLK + (r & 31) * BP + (c & 31);
        q[0] = v.x;
        q[1] = (c + 1 <= r) ? v.y : 0.0;
    }
    // zero the rest of the upper triangle of the diagonal sub-blocks (pairs skipped above)
    for (int idx = threadIdx.x; idx < 4 * SB * (SB / 2); idx += blockDim.x) {
        const int s = idx >> 9, r = (idx >> 4) & 31, c = (idx & 15) * 2;
        if (c > r) {
            double* q = S + tri(s, s) * BLK + r * BP + c;
            q[0] = 0.0;
            q[1] = 0.0;
        }
    }
}

// L (lower triangle incl. diagonal) -> global block
__device__ __forceinline__ void store_L(const double* __restrict__ sm, double* __restrict__ blk, size_t ld) {
    const double* S = sm + OFF_S;
    for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
        const int r = idx >> 7, c = idx & (NB - 1);
        if (c <= r) blk[(size_t)r * ld + c] = S[tri(r >> 5, c >> 5) * BLK + (r & 31) * BP + (c & 31)];
    }
}

// DL = W (dense, zeros above the diagonal), DU = W^T; both NB x NB row-major
__device__ __forceinline__ void store_inverse(const double* __restrict__ sm, double* __restrict__ dl, double* __restrict__ du) {
    const double* W = sm + OFF_W;
    for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
        const int r = idx >> 7, c = idx & (NB - 1);
        dl[idx] = (c <= r) ? W[tri(r >> 5, c >> 5) * BLK + (r & 31) * BP + (c & 31)] : 0.0;
    }
    // transposed read: a warp covers 4 rows x 8 columns of DU per step (64-byte global segments, conflict-free LDS)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int cl = lane >> 2, rl = lane & 3;
    for (int patch = warp; patch < (NB / 4) * (NB / 8); patch += nw) {
        const int r = (patch >> 4) * 4 + rl, c = (patch & 15) * 8 + cl;
        du[r * NB + c] = (c >= r) ? W[tri(c >> 5, r >> 5) * BLK + (c & 31) * BP + (r & 31)] : 0.0;
    }
}

}  // namespace c128
}  // namespace lcgp
