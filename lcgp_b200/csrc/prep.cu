// One-off preprocessing on the device: the O(p N) parts of the reference's constructor pipeline
//   _compute_ybar_np               (lcgp.py:358-367)  replicate means
//   _compute_center_spread_tf      (lcgp.py:383-395)  nearest-rank median and median absolute deviation per output
//   init_standard_y                (lcgp.py:312-324)
//   standardisation + the constant arrays of the objective (YR = r * ybar_s, w_j = sum_i r_i ybar_s[j][i]^2)
// All HBM-bound streaming kernels; the replicate grouping (a lexicographic sort of N rows) and the SVD basis
// stay on the host (LAPACK), see DESIGN.md.
#include "common.cuh"
#include "lcgp_internal.h"

namespace lcgp {

// ybar[j][i] = (sum over t in [off[i], off[i+1]) of y[j][order[t]]) / count, accumulated in ascending t --
// `order` is the stable sort of the group ids, so this is the original column order, as numpy's mean over
// y[:, idx] for the short segments replication produces.
__global__ void __launch_bounds__(256)
segment_mean_kernel(const double* __restrict__ y, const int* __restrict__ order, const int* __restrict__ off,
                    int p, int N, int n, double* __restrict__ ybar) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int t0 = off[i], t1 = off[i + 1];
    for (int j = blockIdx.y; j < p; j += gridDim.y) {
        const double* row = y + (size_t)j * N;
        double s = 0.0;
        for (int t = t0; t < t1; ++t) s += row[order[t]];
        ybar[(size_t)j * n + i] = s / (double)(t1 - t0);
    }
}

// Order-preserving map double -> uint64 (negative values: all bits flipped, others: sign bit set).
__device__ __forceinline__ unsigned long long key_of(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// out[j] = k-th smallest (0-based) of f(Y[j][i]), f(v) = v or |v - center[j]|.  MSB-first radix select, 8 bits
// per pass, one CTA per row; the row (m * 8 bytes) is re-read from L2 in each of the 8 passes.  Exact: the
// result is an element of the row, bit for bit what a full sort would put at position k.
__global__ void __launch_bounds__(256)
row_select_kernel(const double* __restrict__ Y, const double* __restrict__ center, int m, int k, double* __restrict__ out) {
    __shared__ int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_k;
    const int j = blockIdx.x, tid = threadIdx.x;
    const double* row = Y + (size_t)j * m;
    const bool dev = center != nullptr;
    const double c = dev ? center[j] : 0.0;
    if (tid == 0) { s_prefix = 0ull; s_k = k; }
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int i = tid; i < m; i += 256) {
            const double v = dev ? fabs(row[i] - c) : row[i];
            const unsigned long long key = key_of(v);
            if ((key & mask) == prefix) atomicAdd(&hist[(int)((key >> shift) & 255ull)], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int kk = s_k, b = 0;
            while (b < 255 && kk >= hist[b]) { kk -= hist[b]; ++b; }
            s_k = kk;
            s_prefix = prefix | ((unsigned long long)b << shift);
        }
        __syncthreads();
    }
    if (tid == 0) out[j] = value_of(s_prefix);
}

// Ys = (Y - c_j) / s_j ; YR = Ys * r_i ; w_j = sum_i YR[j][i] Ys[j][i].  r == nullptr -> r_i = 1.
// Ys / YR may be nullptr (not wanted).  One CTA per row, fixed reduction tree (deterministic).
__global__ void __launch_bounds__(256)
standardize_kernel(const double* __restrict__ Y, const double* __restrict__ c, const double* __restrict__ s,
                   const double* __restrict__ r, int n, double* __restrict__ Ys, double* __restrict__ YR,
                   double* __restrict__ w) {
    __shared__ double red[8];
    const int j = blockIdx.x;
    const double cj = c[j], sj = s[j];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const size_t o = (size_t)j * n + i;
        const double ys = (Y[o] - cj) / sj;
        const double yr = r ? ys * r[i] : ys;
        if (Ys) Ys[o] = ys;
        if (YR) YR[o] = yr;
        acc += yr * ys;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0 && w) w[j] = acc;
}

cudaError_t prep_segment_mean(const double* y, const int* order, const int* off, int p, int N, int n, double* ybar,
                              cudaStream_t st) {
    note_launch(); segment_mean_kernel<<<dim3((n + 255) / 256, p < 65535 ? p : 65535), 256, 0, st>>>(y, order, off, p, N, n, ybar);
    return cudaGetLastError();
}

cudaError_t prep_row_select(const double* Y, const double* center, int p, int m, int k, double* out, cudaStream_t st) {
    note_launch(); row_select_kernel<<<p, 256, 0, st>>>(Y, center, m, k, out);
    return cudaGetLastError();
}

cudaError_t prep_standardize(const double* Y, const double* c, const double* s, const double* r, int p, int n,
                             double* Ys, double* YR, double* w, cudaStream_t st) {
    note_launch(); standardize_kernel<<<p, 256, 0, st>>>(Y, c, s, r, n, Ys, YR, w);
    return cudaGetLastError();
}

}  // namespace lcgp
