// Do DMMA (FP64 tensor MMA) and DFMA (FP64 FMA) share an execution resource on this GPU?   (developer tool)
// One CTA of 8 warps per SM.  Warps 0-3 (one per SM sub-partition) run a chain-free DMMA stream, warps 4-7 a chain-free
// DFMA stream.  Timed: DMMA warps alone, DFMA warps alone, both together.  If the pipes were independent the combined
// run would take max(t_dmma, t_dfma); if they share the datapath it takes the sum.
//   nvcc -arch=sm_100a -O3 -o dmma_dfma dmma_dfma.cu && ./dmma_dfma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1) k(int mode, int iters, double* out, long long* cyc) {
    const int warp = threadIdx.x >> 5;
    double acc[16][2], f[16];
    for (int i = 0; i < 16; ++i) { acc[i][0] = acc[i][1] = 0.0; f[i] = threadIdx.x * 1e-3 + i; }
    const double a = 1.0 + threadIdx.x * 1e-6, b = 0.5;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < 4) {
        if (mode & 1)
            for (int it = 0; it < iters; ++it)
#pragma unroll
                for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    } else {
        if (mode & 2)
            for (int it = 0; it < iters * 8; ++it)     // 8 DFMA per DMMA slot: 16 x 8 = 128 DFMA per iteration
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0.0;
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1] + f[i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 256 * sms); cudaMalloc(&cyc, sizeof(long long) * 8 * sms);
    const int iters = 20000;
    const char* names[4] = {"", "DMMA warps alone", "DFMA warps alone", "both"};
    for (int mode = 1; mode <= 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            k<<<sms, 256>>>(mode, iters, out, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            if (rep == 1)
                printf("%-18s %8.3f ms   cycles warp0 (DMMA) %lld  warp4 (DFMA) %lld   [per SMSP: %d DMMA, %d DFMA]\n", names[mode], ms,
                       h[0], h[4], (mode & 1) ? iters * 16 : 0, (mode & 2) ? iters * 128 : 0);
        }
    }
    // rates: DMMA.8x8x4 = 256 FMA; DFMA warp instruction = 32 FMA
    return 0;
}
