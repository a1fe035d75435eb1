// Dependent-chain latencies of the operations on the diagonal-block critical path (single warp, clock64).
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__device__ __forceinline__ double fast_rcp(double x) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); r = fma(r, fma(-x, r, 1.0), r); r = fma(r, fma(-x, r, 1.0), r); return r; }
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void lat(double* out, long long* cyc, double x0) {
    __shared__ double sm[1024];
    int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double x = x0 + lane * 1e-9, y = 1.0000001;
    long long t0, t1; int k = 0;
    // 0: dependent DFMA
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
    // 1: dependent DMUL
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = x * y;
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
    // 2: dependent fast_rcp (MUFU.RCP64H + 2 Newton)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = fast_rcp(x) + 0.5;
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
    // 3: dependent rsqrt
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = rsqrt(x) + 0.5;
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
    // 4: dependent shfl (64-bit)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
    // 5: dependent LDS.64 (pointer chase through value)
    {
        int idx = lane;
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < N; ++i) { double v = sm[idx]; idx = (idx + (int)v) & 1023; }
        t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
        x += idx;
    }
    // 6: dependent DMMA (accumulator chain)
    {
        double d0 = x, d1 = y;
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < N; ++i) dmma884(d0, d1, 1e-3, 1e-3);
        t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
        x += d0 + d1;
    }
    // 7: independent DFMA throughput (8 chains)
    {
        double a[8]; for (int j = 0; j < 8; ++j) a[j] = x + j;
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[j]) : "d"(y), "d"(1e-9));
        t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
        for (int j = 0; j < 8; ++j) x += a[j];
    }
    // 8: independent shfl throughput (8 values)
    {
        double a[8]; for (int j = 0; j < 8; ++j) a[j] = x + j;
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < N / 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __shfl_sync(0xffffffffu, a[j], (lane + j + i) & 31);
        t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
        for (int j = 0; j < 8; ++j) x += a[j];
    }
    // 9: independent DMMA throughput (8 accumulator pairs)
    {
        double d[8][2]; for (int j = 0; j < 8; ++j) { d[j][0] = x; d[j][1] = y; }
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < N / 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma884(d[j][0], d[j][1], 1e-3, 1e-3);
        t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
        for (int j = 0; j < 8; ++j) x += d[j][0] + d[j][1];
    }
    // 12: LDS.128 broadcast throughput (independent)
    {
        double s0 = 0, s1 = 0;
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < N; ++i) { double2 v; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(sm + 2 * (i & 255)))); s0 += v.x; s1 += v.y; }
        asm volatile("" :: "d"(s0), "d"(s1));
        t1 = clock64(); if (threadIdx.x == 0) cyc[12] = t1 - t0;
        x += s0 + s1;
    }
    // 13: DMMA throughput, 16 accumulator pairs, clock read depends on the results
    {
        double d[16][2]; for (int j = 0; j < 16; ++j) { d[j][0] = x; d[j][1] = y; }
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < 64; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) dmma884(d[j][0], d[j][1], 1e-3, 1e-3);
        double sacc = 0; for (int j = 0; j < 16; ++j) sacc += d[j][0] + d[j][1];
        if (sacc == 123.456) t0 = 0;          // consume before the clock
        t1 = clock64(); if (threadIdx.x == 0) cyc[13] = t1 - t0;
        x += sacc;
    }
    // 10: __syncthreads round trip (all warps)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) __syncthreads();
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
    // 11: dependent FP32 FFMA for reference
    {
        float f = (float)x, g = 1.0000001f;
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < N; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(g), "f"(1e-9f));
        t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
        x += f;
    }
    out[threadIdx.x] = x;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 32); cudaMemset(cyc, 0, 8 * 32);
    const char* names[] = {"dep DFMA", "dep DMUL", "dep fast_rcp(+add)", "dep rsqrt(+add)", "dep SHFL.64", "dep LDS.64", "dep DMMA.8x8x4",
                           "indep DFMA x8 (per op)", "indep SHFL.64 x8 (per op)", "indep DMMA x8 (per op)", "__syncthreads", "dep FFMA", "indep LDS.128 bcast (per op)", "indep DMMA x16, 1024 ops, results consumed"};
    int per[] = {N, N, N, N, N, N, N, N * 8, N, N, 64, N, N, 1024};
    for (int threads : {32, 128, 256}) {
        for (int rep = 0; rep < 2; ++rep) lat<<<1, threads>>>(out, cyc, 1.5);
        cudaDeviceSynchronize();
        long long h[14]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("== %d threads (1 CTA): cycles per op as seen by warp 0\n", threads);
        for (int k = 0; k < 14; ++k) printf("  %-28s %8.1f\n", names[k], (double)h[k] / per[k]);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
