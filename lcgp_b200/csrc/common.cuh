// Shared device helpers for the lcgp_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lcgp {

// All dense factors are stored padded to a multiple of NB; the pad region of A_k is the
// identity so that L, L^{-1} and A^{-1} are the identity there and never mix with real rows.
constexpr int NB = 128;          // Cholesky / TRTRI block and GEMM CTA tile edge
constexpr int GEMM_THREADS = 256;
constexpr int BK = 32;           // K extent of one pipeline stage (doubles) = 256 B per tile row
constexpr int KSTEPS = NB / BK;  // pipeline iterations per NB-wide K block
constexpr int STAGES = 3;
constexpr int LDS_K = BK + 4;    // smem pitch of a K-major tile row (36 = 4 mod 16: conflict-free LDS.64 fragments)
constexpr int LDS_N = NB + 4;    // smem pitch of an N-major tile row (132)
constexpr int A_STAGE = NB * LDS_K;              // doubles per stage, K-major operand
constexpr int BN_STAGE = BK * LDS_N;             // doubles per stage, N-major operand
constexpr size_t GEMM_SMEM_KMAJOR = sizeof(double) * STAGES * (A_STAGE + A_STAGE);
constexpr size_t GEMM_SMEM_NMAJOR = sizeof(double) * STAGES * (A_STAGE + BN_STAGE);

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Process-wide count of kernel launches made by the library (read through lcgp_launch_count()); defined in api.cu.
void note_launch();

// ---- cp.async (LDGSTS) -------------------------------------------------------------------
// dst is a 32-bit shared-window address (computed once per kernel with __cvta_generic_to_shared, so
// the generic->shared conversion -- an S2R of the CTA id -- stays out of the pipelined loop)
__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- mbarrier (split-phase CTA synchronisation and TMA completion; SASS SYNCS.*) --------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// non-blocking probe of a barrier phase
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// 1 / x by hardware seed (MUFU.RCP64H) + two Newton steps: ~1 ulp, no range checks (x normal, > 0)
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// ---- FP64 tensor core: D(8x8) += A(8x4, row) * B(4x8, col)  -> SASS DMMA.8x8x4 ---------------
// lane = 4*g + t :  a = A[g][t],  b = B[k=t][n=g],  d0,d1 = D[g][2t], D[g][2t+1]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ---- reductions ----------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in thread 0.  `red` must hold blockDim.x/32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int i = 0; i < nw; ++i) s += red[i];
    }
    return s;
}

// Decode a linear index into a lower-triangular pair (i >= j), row-major: x = i(i+1)/2 + j.
__device__ __forceinline__ void tri_decode(int x, int& i, int& j) {
    int r = (int)((sqrt(8.0 * (double)x + 1.0) - 1.0) * 0.5);
    while (r * (r + 1) / 2 > x) --r;
    while ((r + 1) * (r + 2) / 2 <= x) ++r;
    i = r;
    j = x - r * (r + 1) / 2;
}

}  // namespace lcgp
