// Phase timing of the blocked diagonal-block kernel (chol128.cuh) with clock64 stamps; one CTA.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
__device__ long long g_stamp[32];
#define C128_STAMP(id) do { if (threadIdx.x == 0) g_stamp[id] = clock64(); } while (0)
#include "../../lcgp_b200/csrc/chol128.cuh"
namespace lcgp { void note_launch() {} }
using namespace lcgp;
__global__ void __launch_bounds__(256, 1) k(double* blk, double* dl, double* du, int ld) {
    extern __shared__ __align__(16) double sm[];
    if (threadIdx.x == 0) g_stamp[30] = clock64();
    c128::load_block(sm, blk, ld);
    __syncthreads();
    c128::factor_invert(sm, [&] { c128::store_L(sm, blk, ld); });
    C128_STAMP(20);
    c128::store_inverse(sm, dl, du);
    __syncthreads();
    C128_STAMP(21);
}
int main() {
    const int n = 128;
    std::vector<double> A(n * n), M(n * n);
    srand(1);
    for (auto& v : M) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k2 = 0; k2 < n; ++k2) s += M[i * n + k2] * M[j * n + k2]; A[i * n + j] = s / n + (i == j ? 2.0 : 0.0); }
    double *d, *dl, *du; cudaMalloc(&d, n * n * 8); cudaMalloc(&dl, n * n * 8); cudaMalloc(&du, n * n * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c128::SMEM_BYTES);
    long long h[32];
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(d, A.data(), n * n * 8, cudaMemcpyHostToDevice);
        k<<<1, 256, c128::SMEM_BYTES>>>(d, dl, du, n);
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_stamp, sizeof(h));
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    printf("load             %7lld cycles\n", h[0] - h[30]);
    for (int s = 0; s < 4; ++s) {
        long long e = (s < 3) ? h[4 * s + 4] : h[16];
        printf("step %d: panel %6lld  P3 syrk %6lld\n", s, h[4 * s + 1] - h[4 * s], e - h[4 * s + 2]);
    }
    printf("write L          %7lld\nlevel 1          %7lld\nlevel 2          %7lld\nstore inverse    %7lld\ntotal            %7lld cycles = %.1f us at 1.965 GHz\n",
           h[17] - h[16], h[18] - h[17], h[19] - h[18], h[21] - h[20], h[21] - h[30], (h[21] - h[30]) / 1965.0);
    std::vector<double> L(n * n), DL(n * n);
    cudaMemcpy(L.data(), d, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(DL.data(), dl, n * n * 8, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k2 = 0; k2 <= j; ++k2) s += L[i * n + k2] * L[j * n + k2]; e1 = fmax(e1, fabs(s - A[i * n + j])); }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k2 = 0; k2 < n; ++k2) s += DL[i * n + k2] * (k2 >= j ? L[k2 * n + j] : 0.0); e2 = fmax(e2, fabs(s - (i == j))); }
    printf("|L L^T - A| %.2e   |DL L - I| %.2e\n", e1, e2);
    // error of L against a host Cholesky, per 32 x 32 sub-block
    std::vector<double> R(A);
    for (int j = 0; j < n; ++j) { for (int k2 = 0; k2 < j; ++k2) R[j * n + j] -= R[j * n + k2] * R[j * n + k2]; R[j * n + j] = sqrt(R[j * n + j]);
        for (int i = j + 1; i < n; ++i) { for (int k2 = 0; k2 < j; ++k2) R[i * n + j] -= R[i * n + k2] * R[j * n + k2]; R[i * n + j] /= R[j * n + j]; } }
    for (int bi = 0; bi < 4; ++bi) { for (int bj = 0; bj <= bi; ++bj) { double e = 0; for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) { int r = bi * 32 + i, c = bj * 32 + j; if (c <= r) e = fmax(e, fabs(L[r * n + c] - R[r * n + c])); } printf(" %9.2e", e); } printf("\n"); }
    return 0;
}
