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
    int* diagdone;         // [batch][nb]
    int* diagcnt;          // [batch][nb]
    int ntasks;
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
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, GEMM_THREADS / 32);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) mbar_init(dlbar0 + 8 * s, 1);
        mbar_fence_init();
        tma_prefetch_desc(MT == NB ? &maps.km[SRC_F] : (MT == NB / 2 ? &maps.km64 : &maps.km32));
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
        // ---- decode: ticket = ((colbase(j) + (i - j)) * H + h) * batch + b,  colbase(j) = j nb - j (j-1) / 2
        const int b = ticket % P.batch;
        const int x = ticket / P.batch;
        const int h = x % H;
        const int y = x / H;
        int j = (int)(((double)(2 * nb + 1) - sqrt((double)(2 * nb + 1) * (double)(2 * nb + 1) - 8.0 * (double)y)) * 0.5);
        j = max(0, min(j, nb - 1));
        while (j > 0 && j * nb - j * (j - 1) / 2 > y) --j;
        while (j + 1 < nb && (j + 1) * nb - (j + 1) * j / 2 <= y) ++j;
        const int i = j + (y - (j * nb - j * (j - 1) / 2));
        const int moff = h * MT;
        double* Fb = P.v.F + (size_t)b * P.v.fstride;
        double* Ct = Fb + ((size_t)i * NB + moff) * np + (size_t)j * NB;       // this task's rows of tile (i, j)
        const int* rd_i = P.rowdone + ((size_t)b * nb + i) * 4 + h;
        const int* rd_j = P.rowdone + ((size_t)b * nb + j) * 4;

        // ---- acc = -A(i,j): the tile loads overlap the pipeline fill
        double acc[MI][4][2];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            const int rl = wc.wm * (MT / 2) + mi * 8 + wc.pg;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const double2 c = *reinterpret_cast<const double2*>(Ct + (size_t)rl * np + wc.wn * 32 + ni * 8 + 2 * wc.t);
                acc[mi][ni][0] = -c.x;
                acc[mi][ni][1] = -c.y;
            }
        }

        // ---- K loop over blocks k < j
        const int niter = j * KSTEPS;
        int seen_i = 0, seen_j = 0;                  // flags observed by this lane (they only grow)
        // position helpers: iteration `it` of this task lives in ring position (cslot0 + it)
        const int cslot0 = cslot, cuse0 = cuse;
        auto fill = [&](int nxt) {
            if (lane == 0) {
                const int pos = cslot0 + nxt;
                const int slot = pos % STAGES, use = cuse0 + pos / STAGES;
                mbar_wait(empty0 + 8 * slot, (use & 1) ^ 1);
                const int k = nxt / KSTEPS, ks = nxt % KSTEPS;
                if (ks == 0 || seen_i <= k || seen_j <= k) {
                    if (seen_i <= k) seen_i = wait_flag_ge(rd_i, k + 1, P, b, 1000 + j);
                    if (seen_j <= k) {
                        int v = wait_flag_ge(rd_j, k + 1, P, b, 2000 + j);
#pragma unroll
                        for (int hh = 1; hh < H; ++hh) v = min(v, wait_flag_ge(rd_j + hh, k + 1, P, b, 3000 + j));
                        seen_j = v;
                    }
                    fence_proxy_async();
                }
                const unsigned fb = full0 + 8 * slot;
                mbar_arrive_expect_tx(fb, STAGE_BYTES);
                const unsigned dA = sm_u + slot * STAGE_BYTES, dB = dA + ABYTES;
                const CUtensorMap* ma = (MT == NB) ? &maps.km[SRC_F] : (MT == NB / 2 ? &maps.km64 : &maps.km32);
#pragma unroll
                for (int hh = 0; hh < BK / TMA_BOX_K; ++hh)
                    tma_load_3d(dA + hh * (MT * 128), ma, k * NB + ks * BK + hh * TMA_BOX_K, i * NB + moff, b, fb);
#pragma unroll
                for (int hh = 0; hh < BK / TMA_BOX_K; ++hh)
                    tma_load_3d(dB + hh * (NB * 128), &maps.km[SRC_F], k * NB + ks * BK + hh * TMA_BOX_K, j * NB, b, fb);
            }
            __syncwarp();
        };
        if (warp == 0) {
#pragma unroll
            for (int s = 0; s < STAGES - 1; ++s)
                if (s < niter) fill(s);
        }
        for (int it = 0; it < niter; ++it) {
            const int nxt = it + STAGES - 1;
            if (nxt < niter && warp == (it & 7)) fill(nxt);
            mbar_wait(full0 + 8 * cslot, cuse & 1);
            tma_compute_stage<false, MT>(sm + cslot * STAGE_BYTES, sm + cslot * STAGE_BYTES + ABYTES, acc, wc);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * cslot);
            if (++cslot == STAGES) { cslot = 0; ++cuse; }
        }
        __syncthreads();                             // every warp is done with the ring: shared memory is free

        if (i == j) {
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
                // half (64-row tasks) comes back from L2
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
            // ---- off-diagonal tile: P = C inv(L_jj)^T.  C goes to shared memory in the layout of a K-major operand-A
            //      stage (4 K chunks of [2 boxes][MT rows][128 B], 16-byte chunk index ^= row % 8)
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
            auto load_dl = [&](int ks) {             // thread 0: K chunk ks of DL_jj (128 n-rows x 32 k) into its slot
                const int slot = ks % NDL;
                const unsigned bar = dlbar0 + 8 * slot;
                mbar_arrive_expect_tx(bar, TMA_B_BYTES);
#pragma unroll
                for (int hh = 0; hh < BK / TMA_BOX_K; ++hh)
                    tma_load_3d(dl_base + slot * TMA_B_BYTES + hh * (NB * 128), &maps.km[SRC_DL], ks * BK + hh * TMA_BOX_K,
                                j * NB, b, bar);
            };
            if (tid == 0) {
                wait_flag_ge(&P.diagdone[(size_t)b * nb + j], 1, P, b, 4000 + j);
                fence_proxy_async();
#pragma unroll
                for (int ks = 0; ks < NDL; ++ks) load_dl(ks);
            }
            __syncthreads();                         // C is in shared memory
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const int slot = ks % NDL;
                mbar_wait(dlbar0 + 8 * slot, (dlpar >> slot) & 1);
                dlpar ^= 1u << slot;
                tma_compute_stage<false, MT>(sm + ks * ABYTES, sm + KSTEPS * ABYTES + slot * TMA_B_BYTES, acc, wc);
                if (NDL < KSTEPS && ks + NDL < KSTEPS) {
                    __syncthreads();                 // every warp has read the slot
                    if (tid == 0) load_dl(ks + NDL);
                }
            }
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                const int rl = wc.wm * (MT / 2) + mi * 8 + wc.pg;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    *reinterpret_cast<double2*>(Ct + (size_t)rl * np + wc.wn * 32 + ni * 8 + 2 * wc.t) =
                        make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            }
            __threadfence();
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) st_release_gpu(const_cast<int*>(rd_i), j + 1);
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
static int pll_half_limit() { static const int v = env_i("LCGP_PLL_HALF", 160); return v; }
static int pll_quarter_limit() { static const int v = env_i("LCGP_PLL_QUARTER", 64); return v; }

size_t potrf_pll_sync_ints(int nb, int batch) { return 8 + (size_t)6 * batch * nb; }

cudaError_t potrf_pll(const FactorView& v, double* DLw, double* DUw, int batch, double* logdet_part, int* info,
                      int* sync, cudaStream_t stream) {
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
    const bool quarter = v.nb <= 16 && v.nb * batch <= pll_quarter_limit();
    const bool half = !quarter && v.nb * batch <= pll_half_limit();
    const int H = quarter ? 4 : (half ? 2 : 1);
    PllParams P;
    P.v = v; P.DLw = DLw; P.DUw = DUw; P.batch = batch; P.logdet_part = logdet_part; P.info = info;
    P.hdr = sync;
    P.rowdone = sync + 8;
    P.diagdone = P.rowdone + (size_t)4 * batch * v.nb;
    P.diagcnt = P.diagdone + (size_t)batch * v.nb;
    P.ntasks = v.nb * (v.nb + 1) / 2 * H * batch;
    P.timeout = (long long)env_i("LCGP_PLL_TIMEOUT_MS", 4000) * 2000000LL;   // ~2 GHz
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
