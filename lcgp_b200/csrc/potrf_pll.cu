// Batched FP64 Cholesky as ONE persistent, dependency-driven kernel (left-looking at tile level).
//
// Replaces the launch chain of potrf.cu (per 128-wide block column: column update -> one-CTA diagonal kernel ->
// panel solve, ~140 us of serial latency per block column, 800+ launches per evaluation at n = 8000) by a single
// launch whose CTAs draw tile tasks from a ticket counter:
//
//   task (b, i, j, h):  rows [h MT, (h+1) MT) of tile (i, j), i >= j, of matrix b            (MT = 128 or 64)
//     acc  = sum_{k<j} P(i,k) P(j,k)^T        TMA-staged DMMA pipeline; before the loads of K block k the producer
//                                             waits for rowdone[b][i][h] > k and rowdone[b][j][*] > k
//     C    = A(i,j) - acc                     (the tile is read once, written once: no per-panel read-modify-write)
//     i >  j :  wait diagdone[b][j];  P(i,j) = C inv(L_jj)^T (GEMM with the explicit inverse, C staged in shared
//               memory as the A operand);  store;  rowdone[b][i][h] = j + 1
//     i == j :  store C;  the CTA finishing the LAST half of the diagonal tile factors and inverts the block
//               (chol128.cuh: ~12 us, CTA local) and publishes L_jj, DL_jj, DU_jj;  diagdone[b][j] = 1
//
// Tickets are handed out in column-major order (j, then i with the diagonal tile first, then h, then b), so every
// dependency of a task has a SMALLER ticket.  Hence the CTA holding the smallest unfinished ticket can always
// finish: no co-residency assumption, no cooperative launch, safe next to other kernels on other streams.
// Look-ahead is implicit: tiles of column j+1 consume K blocks 0..j-1 while column j is being finished.
// Cross-CTA visibility: generic-proxy stores -> __threadfence -> fence.proxy.async -> release store of the flag;
// acquire load of the flag -> fence.proxy.async -> TMA loads (async proxy, through L2).
// Every wait is bounded (clock64): on expiry the matrix' info is set to a negative code and the kernel carries on
// (wrong numbers, reported) instead of hanging the device.
#include "chol128.cuh"
#include "gemm_dmma.cuh"
#include "lcgp_internal.h"

#include <cstdlib>

namespace lcgp {

struct PllParams {
    FactorView v;
    double* DLw;
    double* DUw;
    int batch;
    double* logdet_part;   // [batch][nb] or null
    int* info;             // [batch]
    int* hdr;              // [0] ticket counter, [1] spare
    int* rowdone;          // [batch][nb][4]
    int* diagdone;         // [batch][nb]: 1 = L_jj and DL_jj published, 2 = DU_jj as well
    int* diagcnt;          // [batch][nb]
    int* urow;             // [batch][nb][4]: finished inverse tiles U(j, j+1 ..) per row slice (fused inverse only)
    int fused;             // 1: the triangular inverse U = L^-T (strictly upper blocks) is computed by the same kernel
    int ntasks;
    int inv_first;         // fused ticket order inside a block column: diagonal tile, inverse tiles, then the other Cholesky tiles
    int compact_diag;      // 1: small-code diagonal-block panels (factor buffers larger than L2), see chol128.cuh
    long long timeout;     // cycles
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Spins until *p >= target; returns the observed value.  Bounded: after `timeout` cycles reports through info and
// gives up waiting (returns target so that callers continue).
__device__ __forceinline__ int wait_flag_ge(const int* p, int target, const PllParams& P, int b, int code) {
    int v = ld_acquire_gpu(p);
    if (v >= target) return v;
    const long long t0 = clock64();
    unsigned ns = 20;
    while (true) {
        __nanosleep(ns);
        if (ns < 200) ns += 20;
        v = ld_acquire_gpu(p);
        if (v >= target) return v;
        if (clock64() - t0 > P.timeout) {
            atomicCAS(&P.info[b], 0, -code);
            return target;
        }
    }
}

// The diagonal block: rows [h * mt, (h + 1) * mt) are already in shared memory (the caller put them there from its
// accumulators), the other row slices come back from L2.  Factor, invert, hand DL over (flag), then write the rest.
// NOT inlined: the panel code (chol128.cuh) is large and register hungry; inlined into the persistent kernel it
// degraded the register allocation of the hot K loop (64-row tasks ran 7-9 % slower).
__device__ __noinline__ void pll_diag_block(double* dsm, double* blk, int np, double* dl, double* du, int* done_flag,
                                            int* info, double* logdet_slot, int j, int h, int mt, bool compact) {
    const int tid = threadIdx.x;
    if (mt == NB / 2) c128::load_rows<NB / 2>(dsm, blk, np, (1 - h) * (NB / 2));
    else if (mt == NB / 4) {
#pragma unroll 1
        for (int hh = 0; hh < 4; ++hh)
            if (hh != h) c128::load_rows<NB / 4>(dsm, blk, np, hh * (NB / 4));
    }
    __syncthreads();
    // the diagonal sub-blocks of L are scratch of the inverse phase: they go to global before it, the rest of L
    // and DU after DL has been handed over
    c128::factor_invert(dsm, [&] { c128::store_L_part<true>(dsm, blk, np); }, compact);
    c128::store_DL(dsm, dl);
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(done_flag), "r"(1) : "memory");   // the solves of column j may start
    c128::store_L_part<false>(dsm, blk, np);
    c128::store_DU(dsm, du);
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(done_flag), "r"(2) : "memory");   // DU is there too
    const double* pivs = dsm + c128::OFF_PIV;
    if (tid == 0) {
        for (int c = 0; c < NB; ++c)
            if (!(pivs[c] > 0.0)) { atomicCAS(info, 0, j * NB + c + 1); break; }
    }
    double lg = (tid < NB) ? 0.5 * log(pivs[tid]) : 0.0;
    lg = block_sum(lg, dsm + c128::OFF_RED);
    if (tid == 0 && logdet_slot) *logdet_slot = lg;
}

constexpr int PLL_DATA_BYTES = STAGES * (NB * BK * 8 + TMA_B_BYTES);     // 192 KB: ring / TRSM staging / diagonal block
constexpr int PLL_NBAR = 2 * STAGES + 4;
constexpr size_t PLL_SMEM = (size_t)PLL_DATA_BYTES + 8 * PLL_NBAR + 64 + 1024;
static_assert(c128::SMEM_BYTES <= (size_t)PLL_DATA_BYTES, "diagonal-block routine must fit in the ring area");

template <int MT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
potrf_pll_kernel(const __grid_constant__ PllParams P, const __grid_constant__ GemmMaps maps) {
    constexpr int H = NB / MT;                       // row slices per tile (1, 2 or 4)
    constexpr int MI = MT / 16;
    constexpr int ABYTES = MT * BK * 8;              // operand A per K stage
    constexpr int STAGE_BYTES = ABYTES + TMA_B_BYTES;
    constexpr int NDL = (MT == NB) ? 2 : 4;          // DL chunk slots of the solve step
    extern __shared__ unsigned char smem_raw[];
    __shared__ int s_ticket, s_last;
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const unsigned sm_u = smem_u32(sm);
    const unsigned full0 = sm_u + PLL_DATA_BYTES, empty0 = full0 + 8 * STAGES, dlbar0 = empty0 + 8 * STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = P.v.nb, np = P.v.np;
    const CUtensorMap* const mapA = (MT == NB) ? &maps.km[SRC_F] : (MT == NB / 2 ? &maps.km64 : &maps.km32);      // F, MT-row boxes
    const CUtensorMap* const mapDU = (MT == NB) ? &maps.km[SRC_DU] : (MT == NB / 2 ? &maps.du64 : &maps.du32);     // DU, MT-row boxes
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, GEMM_THREADS / 32);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) mbar_init(dlbar0 + 8 * s, 1);
        mbar_fence_init();
        tma_prefetch_desc(mapA);
        tma_prefetch_desc(&maps.km[SRC_F]);
        tma_prefetch_desc(&maps.km[SRC_DL]);
    }
    __syncthreads();

    TmaCoord<false, MT> wc;                          // moff is overridden per task (blockIdx.z unused here)
    int cslot = 0, cuse = 0;                         // ring position shared by producer and consumers at task boundaries
    unsigned dlpar = 0;                              // parity bits of the DL barriers

    while (true) {
        __syncthreads();                             // previous task is done with shared memory and s_ticket
        if (tid == 0) s_ticket = atomicAdd(&P.hdr[0], 1);
        __syncthreads();
        const int ticket = s_ticket;
        if (ticket >= P.ntasks) break;
        fence_proxy_async();                         // generic accesses of the previous task -> before this task's TMA writes
        // ---- decode.  ticket = (slot * H + h) * batch + b; `slot` enumerates tiles.
        //   Cholesky only : column-major lower triangle, slot = colbase(j) + (i - j), colbase(j) = j nb - j (j-1) / 2
        //   fused inverse : column 0: nb Cholesky tiles; column c >= 1: the nb - c Cholesky tiles (i, c), i >= c, followed by
        //                   the c - 1 inverse tiles U(jt, c - 1), jt < c - 1 (one column late: the tiles of the next
        //                   Cholesky column, which are on the dependency chain, are never queued behind them for long;
        //                   P.inv_first puts them between the diagonal tile and the other Cholesky tiles of the column);
        //                   at the very end the nb - 1 inverse tiles of the last block column.  nb^2 slots.
        const int b = ticket % P.batch;
        const int x = ticket / P.batch;
        const int h = x % H;
        const int y = x / H;
        int kind = 0, i, j;                          // kind 0: Cholesky tile (i, j); kind 1: inverse tile U(j, i), j < i
        if (!P.fused) {
            j = (int)(((double)(2 * nb + 1) - sqrt((double)(2 * nb + 1) * (double)(2 * nb + 1) - 8.0 * (double)y)) * 0.5);
            j = max(0, min(j, nb - 1));
            while (j > 0 && j * nb - j * (j - 1) / 2 > y) --j;
            while (j + 1 < nb && (j + 1) * nb - (j + 1) * j / 2 <= y) ++j;
            i = j + (y - (j * nb - j * (j - 1) / 2));
        } else if (y < nb) {
            j = 0; i = y;
        } else {
            const int c = 1 + (y - nb) / (nb - 1), r = (y - nb) % (nb - 1);
            if (c < nb && P.inv_first) {
                // diagonal tile first, then the c - 1 inverse tiles (they do not need DL_c: CTAs that would otherwise wait for
                // the diagonal block have work), then the off-diagonal Cholesky tiles.  Every dependency still has a smaller ticket.
                if (r == 0) { j = c; i = c; }
                else if (r < c) { kind = 1; i = c - 1; j = r - 1; }
                else { j = c; i = r + 1; }
            } else if (c < nb && r < nb - c) { j = c; i = c + r; }
            else { kind = 1; i = c - 1; j = (c < nb) ? r - (nb - c) : r; }
        }
        const int moff = h * MT;
        double* Fb = P.v.F + (size_t)b * P.v.fstride;
        // operands: acc += A_k B_k^T over K blocks [k0, k1)
        //   Cholesky : A_k = L(i, k) rows of this slice, B_k = L(j, k); k in [0, j); solve with DL_j; result -> tile (i, j)
        //   inverse  : A_k = U(j, k) rows of this slice (U(j, j) = DU_j), B_k = L(i, k); k in [j, i); solve with DL_i,
        //              negated (U(j,i) = -[sum_k U(j,k) L(i,k)^T] inv(L_ii)) -> tile (j, i) in the upper triangle
        const int arow = kind ? j : i, brow = kind ? i : j;
        const int k0 = kind ? j : 0, k1 = kind ? i : j;
        const int dlb = kind ? i : j;
        // structural zeros (tma_compute_stage_skip): pad rows of the last block row (Cholesky tiles (nb-1, j)), pad columns
        // of the last block column (inverse tiles U(j, nb-1)), and the triangular DU_j that is the first K block of an
        // inverse tile
        auto padlim = [&] { return (i == nb - 1) ? last_block_rows(P.v) : NB; };   // i: row of a Cholesky tile, column of an inverse tile
        double* Ct = Fb + ((size_t)arow * NB + moff) * np + (size_t)(kind ? i : j) * NB;       // this task's rows of its output tile
        const int* fl_a = (kind ? P.urow : P.rowdone) + ((size_t)b * nb + arow) * 4 + h;     // progress of operand A's row slice
        const int* fl_b = P.rowdone + ((size_t)b * nb + brow) * 4;                            // ... of operand B's block row

        // ---- acc = -A(i,j) (Cholesky; the tile loads overlap the pipeline fill) or 0 (inverse)
        double acc[MI][4][2];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            const int rl = wc.wm * (MT / 2) + mi * 8 + wc.pg;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                double2 c = make_double2(0.0, 0.0);
                if (!kind) c = *reinterpret_cast<const double2*>(Ct + (size_t)rl * np + wc.wn * 32 + ni * 8 + 2 * wc.t);
                acc[mi][ni][0] = -c.x;
                acc[mi][ni][1] = -c.y;
            }
        }

        // ---- K loop
        const int niter = (k1 - k0) * KSTEPS;
        int ready_a = 0, ready_b = 0;                // K blocks k < ready_* are known to be complete (flags only grow)
        const int cslot0 = cslot, cuse0 = cuse;      // iteration `it` of this task lives in ring position cslot0 + it
        auto fill = [&](int nxt) {
            if (lane == 0) {
                const int pos = cslot0 + nxt;
                const int slot = pos % STAGES, use = cuse0 + pos / STAGES;
                mbar_wait(empty0 + 8 * slot, (use & 1) ^ 1);
                const int k = k0 + nxt / KSTEPS, ks = nxt % KSTEPS;
                const bool a_is_du = kind && k == arow;
                if (ks == 0 || ready_a <= k || ready_b <= k) {
                    if (a_is_du) {
                        if (ready_a <= k) { wait_flag_ge(&P.diagdone[(size_t)b * nb + arow], 2, P, b, 5000 + arow); ready_a = k + 1; }
                    } else if (ready_a <= k) {
                        // Cholesky: rowdone = finished tiles of the row slice = K blocks; inverse: urow = finished U tiles to
                        // the right of the diagonal block: U(j, j+1 .. j+v) -> K blocks k <= j + v
                        const int need = kind ? k - arow : k + 1;
                        const int v = wait_flag_ge(fl_a, need, P, b, 1000 + k);
                        ready_a = kind ? arow + v + 1 : v;
                    }
                    if (ready_b <= k) {
                        int v = wait_flag_ge(fl_b, k + 1, P, b, 2000 + k);
#pragma unroll
                        for (int hh = 1; hh < H; ++hh) v = min(v, wait_flag_ge(fl_b + hh, k + 1, P, b, 3000 + k));
                        ready_b = v;
                    }
                    fence_proxy_async();
                }
                const unsigned fb = full0 + 8 * slot;
                mbar_arrive_expect_tx(fb, STAGE_BYTES);
                const unsigned dA = sm_u + slot * STAGE_BYTES, dB = dA + ABYTES;
#pragma unroll
                for (int hh = 0; hh < BK / TMA_BOX_K; ++hh) {
                    if (a_is_du) tma_load_3d(dA + hh * (MT * 128), mapDU, ks * BK + hh * TMA_BOX_K, arow * NB + moff, b, fb);
                    else tma_load_3d(dA + hh * (MT * 128), mapA, k * NB + ks * BK + hh * TMA_BOX_K, arow * NB + moff, b, fb);
                }
#pragma unroll
                for (int hh = 0; hh < BK / TMA_BOX_K; ++hh)
                    tma_load_3d(dB + hh * (NB * 128), &maps.km[SRC_F], k * NB + ks * BK + hh * TMA_BOX_K, brow * NB, b, fb);
            }
            __syncwarp();
        };
        if (warp == 0) {
#pragma unroll
            for (int s = 0; s < STAGES - 1; ++s)
                if (s < niter) fill(s);
        }
        // one K step; MASKED steps go through the structural-zero variant.  They get a loop of their own so that the
        // body loop is exactly the plain pipeline (see gemm_tma_kernel).
        auto kstep = [&](int it, auto masked_c) {
            const int nxt = it + STAGES - 1;
            bool late = false;                       // see gemm_tma_kernel: a producer whose slot is still being read fills after its MMAs
            if (nxt < niter && warp == (it & 7)) {
#if LCGP_LATE_FILL
                const int pos = cslot0 + nxt;
                unsigned ok = 0;
                if (lane == 0) ok = mbar_test(empty0 + 8 * (pos % STAGES), ((cuse0 + pos / STAGES) & 1) ^ 1);
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (ok) fill(nxt);
                else late = true;
#else
                fill(nxt);
#endif
            }
            mbar_wait(full0 + 8 * cslot, cuse & 1);
            if constexpr (decltype(masked_c)::value)
                tma_compute_stage_skip<false, MT>(sm + cslot * STAGE_BYTES, sm + cslot * STAGE_BYTES + ABYTES, acc, wc,
                                                  StepMask{kind ? (it < KSTEPS ? BK * (it + 1) : NB) : padlim(), 0, kind ? padlim() : NB},
                                                  moff + wc.wm * (MT / 2));
            else
                tma_compute_stage<false, MT>(sm + cslot * STAGE_BYTES, sm + cslot * STAGE_BYTES + ABYTES, acc, wc);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * cslot);
            if (late) fill(nxt);
            if (++cslot == STAGES) { cslot = 0; ++cuse; }
        };
        {
            int it = 0;
            if (LCGP_SKIP_MODE != 0) {
                const int head = (padlim() <= (kind ? NB - BK : NB - 8)) ? niter : (kind ? min(KSTEPS, niter) : 0);
                for (; it < head; ++it) kstep(it, std::true_type{});
            }
            for (; it < niter; ++it) kstep(it, std::false_type{});
        }
        __syncthreads();                             // every warp is done with the ring: shared memory is free

        if (!kind && i == j) {
            // ---- diagonal tile: store C = A - sum; the CTA that completes the tile factors it
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                const int rl = wc.wm * (MT / 2) + mi * 8 + wc.pg;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    *reinterpret_cast<double2*>(Ct + (size_t)rl * np + wc.wn * 32 + ni * 8 + 2 * wc.t) =
                        make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) s_last = (atomicAdd(&P.diagcnt[(size_t)b * nb + j], 1) == H - 1);
            __syncthreads();
            if (s_last) {
                __threadfence();
                double* dsm = reinterpret_cast<double*>(sm);
                double* blk = Fb + (size_t)j * NB * np + (size_t)j * NB;
                // this CTA's rows of the updated block go to shared memory straight from the accumulators; the other
                // slices (64- / 32-row tasks) come back from L2
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    const int rl = wc.wm * (MT / 2) + mi * 8 + wc.pg;
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni)
                        c128::put_pair(dsm, moff + rl, wc.wn * 32 + ni * 8 + 2 * wc.t, -acc[mi][ni][0], -acc[mi][ni][1]);
                }
                pll_diag_block(dsm, blk, np, P.DLw + (size_t)b * P.v.dstride + (size_t)j * NB * NB,
                               P.DUw + (size_t)b * P.v.dstride + (size_t)j * NB * NB, &P.diagdone[(size_t)b * nb + j],
                               &P.info[b], P.logdet_part ? &P.logdet_part[(size_t)b * nb + j] : nullptr, j, h, MT,
                               P.compact_diag != 0);
            }
        } else {
            // ---- solve step: out = (-acc) inv(L_dd)^T, d = dlb.  -acc goes to shared memory in the layout of a K-major
            //      operand-A stage (4 K chunks of [2 boxes][MT rows][128 B], 16-byte chunk index ^= row % 8)
            {
                const int ks = wc.wn;                // this lane's 8 columns wn*32 + ni*8 + 2t lie in K chunk wn
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    const int rl = wc.wm * (MT / 2) + mi * 8 + wc.pg;
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) {
                        const int chunk = ((ni & 1) << 2) | wc.t;
                        unsigned char* q = sm + ks * ABYTES + (ni >> 1) * (MT * 128) + rl * 128 + ((chunk ^ wc.pg) << 4);
                        *reinterpret_cast<double2*>(q) = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
                        acc[mi][ni][0] = 0.0;
                        acc[mi][ni][1] = 0.0;
                    }
                }
            }
            const unsigned dl_base = sm_u + KSTEPS * ABYTES;
            auto load_dl = [&](int ks) {             // thread 0: K chunk ks of DL_d (128 n-rows x 32 k) into its slot
                const int slot = ks % NDL;
                const unsigned bar = dlbar0 + 8 * slot;
                mbar_arrive_expect_tx(bar, TMA_B_BYTES);
#pragma unroll
                for (int hh = 0; hh < BK / TMA_BOX_K; ++hh)
                    tma_load_3d(dl_base + slot * TMA_B_BYTES + hh * (NB * 128), &maps.km[SRC_DL], ks * BK + hh * TMA_BOX_K,
                                dlb * NB, b, bar);
            };
            if (tid == 0) {
                wait_flag_ge(&P.diagdone[(size_t)b * nb + dlb], 1, P, b, 4000 + dlb);
                fence_proxy_async();
#pragma unroll
                for (int ks = 0; ks < NDL; ++ks) load_dl(ks);
            }
            __syncthreads();                         // -acc is in shared memory
            // DL is lower triangular: output columns [32 c, 32 c + 32) only need K chunks ks <= c.  The two warps of every SM
            // sub-partition (w and w + 4 share one tensor pipe) hold column chunks c and 3 - c (TmaCoord): 5 K chunks per
            // sub-partition instead of 8 (the solve sits on the dependency chain of every block column, and is one extra
            // K block of every tile).
            const TmaCoord<false, MT>& ws = wc;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const int slot = ks % NDL;
                mbar_wait(dlbar0 + 8 * slot, (dlpar >> slot) & 1);
                dlpar ^= 1u << slot;
                if (ks <= ws.wn)
                    tma_compute_stage<false, MT>(sm + ks * ABYTES, sm + KSTEPS * ABYTES + slot * TMA_B_BYTES, acc, ws);
                if (NDL < KSTEPS && ks + NDL < KSTEPS) {
                    __syncthreads();                 // every warp has read the slot
                    if (tid == 0) load_dl(ks + NDL);
                }
            }
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                const int rl = ws.wm * (MT / 2) + mi * 8 + ws.pg;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    *reinterpret_cast<double2*>(Ct + (size_t)rl * np + ws.wn * 32 + ni * 8 + 2 * ws.t) =
                        make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            }
            __threadfence();
            fence_proxy_async();
            __syncthreads();
            // progress of this row slice: Cholesky tile (i, j) -> j + 1 finished tiles; inverse tile U(j, i) -> i - j
            if (tid == 0) st_release_gpu(const_cast<int*>(fl_a), kind ? i - j : j + 1);
        }
    }
}

// ---- host side --------------------------------------------------------------------------------------------
static int env_i(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return (e && *e) ? std::atoi(e) : dflt;
}
// LCGP_PLL_HALF / LCGP_PLL_QUARTER: 64-row / 32-row tasks when nb * batch is at most this (few tiles: the latency of the
// tiles on the dependency chain matters more than operand reuse)
// Defaults: factorisation alone (a pure dependency chain for small batches): 64-row tasks up to 160, 32-row tasks up to 64
// for nb <= 16.  With the fused inverse the launch is work bound much earlier (its tiles fill the idle SMs), and the
// wider tiles' better operand reuse wins: 64-row tasks up to 96, no 32-row tasks (measured: config 5 one emulator
// 0.998 ms with 64-row tasks vs 1.07 (32) / 1.15 (128); config 3 4.29 ms with 128-row tasks vs 4.56 (64) / 5.35 (32)).
static int pll_half_limit(bool fused) {
    static const int v = env_i("LCGP_PLL_HALF", -1);
    return v >= 0 ? v : (fused ? 96 : 160);
}
static int pll_quarter_limit(bool fused) {
    static const int v = env_i("LCGP_PLL_QUARTER", -1);
    return v >= 0 ? v : (fused ? 0 : 64);
}

size_t potrf_pll_sync_ints(int nb, int batch) { return 8 + (size_t)10 * batch * nb; }

cudaError_t potrf_pll(const FactorView& v, double* DLw, double* DUw, int batch, double* logdet_part, int* info,
                      int* sync, cudaStream_t stream, bool fused_inverse) {
    static int sms[MAX_DEVICES];
    static std::atomic<bool> configured[MAX_DEVICES];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= MAX_DEVICES) dev = 0;
    if (!configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(potrf_pll_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PLL_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(potrf_pll_kernel<NB / 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PLL_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(potrf_pll_kernel<NB / 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PLL_SMEM);
        if (e != cudaSuccess) return e;
        int n = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        sms[dev] = n > 0 ? n : 148;
        configured[dev].store(true, std::memory_order_release);
    }
    GemmCtx ctx;
    {
        GemmSrcs srcs;
        int rows[NSRC];
        factor_srcs(v, srcs, rows);
        ctx.tma = true;
        cudaError_t e = gemm_make_ctx(ctx, srcs, rows, batch);
        if (e != cudaSuccess) return e;
        if (!ctx.tma) return cudaErrorNotSupported;
    }
    const size_t nints = potrf_pll_sync_ints(v.nb, batch);
    cudaError_t e = cudaMemsetAsync(sync, 0, nints * sizeof(int), stream);
    if (e != cudaSuccess) return e;
    // quarter tiles only for small matrices (measured: 1024 x 1 500 -> 450 us, 1024 x 8 533 -> 508, 2048 x 1 1016 -> 914; but
    // 8064 x 1 7.0 -> 8.4 ms: with many block columns the bulk efficiency of the wider tiles wins)
    const bool fused = fused_inverse && v.nb > 1;
    const bool quarter = v.nb <= 16 && v.nb * batch <= pll_quarter_limit(fused);
    const bool half = !quarter && v.nb * batch <= pll_half_limit(fused);
    const int H = quarter ? 4 : (half ? 2 : 1);
    PllParams P;
    P.v = v; P.DLw = DLw; P.DUw = DUw; P.batch = batch; P.logdet_part = logdet_part; P.info = info;
    P.hdr = sync;
    P.rowdone = sync + 8;
    P.diagdone = P.rowdone + (size_t)4 * batch * v.nb;
    P.diagcnt = P.diagdone + (size_t)batch * v.nb;
    P.urow = P.diagcnt + (size_t)batch * v.nb;
    P.fused = fused ? 1 : 0;
    P.ntasks = (P.fused ? v.nb * v.nb : v.nb * (v.nb + 1) / 2) * H * batch;
    P.timeout = (long long)env_i("LCGP_PLL_TIMEOUT_MS", 4000) * 2000000LL;   // ~2 GHz
    {   // LCGP_PLL_INVFIRST=0: the older order (all Cholesky tiles of a block column, then the inverse tiles).  Measured with
        // the inverse tiles right behind the diagonal tile: 64 x 1024: 2.25 -> 2.14 ms, config 3 2.33 -> 2.26 ms, config 4 per-GPU
        // share 41.5 -> 40.8 ms, 8 x 1024 unchanged (with the chain tile (c + 1, c) kept right behind the diagonal tile: 2.17 / 2.29 / 40.9)
        const int f = env_i("LCGP_PLL_INVFIRST", -1);
        P.inv_first = f >= 0 ? f : 1;
    }
    {   // LCGP_PLL_COMPACT: 0 / 1 force; default: compact code once the factor buffers exceed ~half of the 126 MB L2
        const int f = env_i("LCGP_PLL_COMPACT", -1);
        P.compact_diag = f >= 0 ? f : ((size_t)batch * v.fstride * sizeof(double) > ((size_t)64 << 20) ? 1 : 0);
    }
    const int grid = P.ntasks < sms[dev] ? P.ntasks : sms[dev];
    note_launch();
    if (quarter) potrf_pll_kernel<NB / 4><<<grid, GEMM_THREADS, PLL_SMEM, stream>>>(P, ctx.maps);
    else if (half) potrf_pll_kernel<NB / 2><<<grid, GEMM_THREADS, PLL_SMEM, stream>>>(P, ctx.maps);
    else potrf_pll_kernel<NB><<<grid, GEMM_THREADS, PLL_SMEM, stream>>>(P, ctx.maps);
    return cudaGetLastError();
}

}  // namespace lcgp
