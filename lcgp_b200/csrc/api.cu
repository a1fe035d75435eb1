// extern "C" surface of liblcgp_b200.so (see include/lcgp_b200.h) and the O(p n q) glue kernels
// around the per-latent dense stages.
#include "../../include/lcgp_b200.h"
#include "gemm_dmma.cuh"
#include "lcgp_internal.h"

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <new>

namespace lcgp {

static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- tunables (environment, read once) --------------------------------------------------------
//   LCGP_PANEL_W  : block columns per Cholesky panel (K = 128 * W in the trailing update); default: 16 for
//                   groups of >= 8 latents, else 8
//   LCGP_STREAMS  : independent groups of latents factored on separate streams so that one group's
//                   serial panel work overlaps another group's trailing updates; default min(4, q_loc)
static int env_int(const char* name, int dflt, int lo, int hi) {
    const char* e = std::getenv(name);
    if (!e || !*e) return dflt;
    int v = std::atoi(e);
    return v < lo ? lo : (v > hi ? hi : v);
}
static int panel_width() { static const int v = env_int("LCGP_PANEL_W", 0, 0, 16); return v; }   // 0 = auto
constexpr int MAX_GROUPS = 4;
static int stream_groups(int q) {
    static const int v = env_int("LCGP_STREAMS", MAX_GROUPS, 1, MAX_GROUPS);
    return q < v ? q : v;
}
//   LCGP_LOOKAHEAD: 1 (default) = Cholesky panel chain on a high-priority stream, overlapped with the bulk of
//                   the previous trailing update; 0 = single stream per group
static bool lookahead_on() { static const int v = env_int("LCGP_LOOKAHEAD", 1, 0, 1); return v != 0; }

// Side streams + fork/join events, created on first use for the current device.  The library
// still allocates no device memory; these are the only objects that outlive a call.
struct SidePool {
    bool ready = false;
    cudaStream_t s[MAX_GROUPS];
    cudaStream_t hp[MAX_GROUPS];                     // high-priority panel streams (Cholesky look-ahead)
    cudaEvent_t fork, join[MAX_GROUPS], evp[MAX_GROUPS], evb[MAX_GROUPS];
    std::mutex mu;
    cudaError_t ensure() {
        if (ready) return cudaSuccess;
        cudaError_t e;
        int least = 0, greatest = 0;
        if ((e = cudaDeviceGetStreamPriorityRange(&least, &greatest)) != cudaSuccess) return e;
        for (int i = 0; i < MAX_GROUPS; ++i) {
            if ((e = cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaStreamCreateWithPriority(&hp[i], cudaStreamNonBlocking, greatest)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&evp[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&evb[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        if ((e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming)) != cudaSuccess) return e;
        ready = true;
        return cudaSuccess;
    }
};
// one pool per device (streams and events belong to the device that was current when they were created)
static SidePool& side_pool() {
    static SidePool pools[MAX_DEVICES];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) dev = 0;
    return pools[dev];
}

// Runs f(first_latent, count, stream, group) for G contiguous groups of the q latents, each on its own
// side stream, between a fork from and a join back into `main`.  The pool stays locked while work that
// records its events is being enqueued.
template <class F>
static cudaError_t run_grouped(cudaStream_t main, int q, int G, bool need_pool, F f) {
    if (G <= 1 && !need_pool) return f(0, q, main, 0);   // nothing shared is touched: concurrent callers do not serialise
    SidePool& P = side_pool();
    std::lock_guard<std::mutex> lock(P.mu);
    cudaError_t e = P.ensure();
    if (e != cudaSuccess) return e;
    if (G <= 1) return f(0, q, main, 0);
    if ((e = cudaEventRecord(P.fork, main)) != cudaSuccess) return e;
    int g0 = 0;
    for (int g = 0; g < G; ++g) {
        const int cnt = q / G + (g < q % G ? 1 : 0);
        if ((e = cudaStreamWaitEvent(P.s[g], P.fork, 0)) != cudaSuccess) return e;
        if ((e = f(g0, cnt, P.s[g], g)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(P.join[g], P.s[g])) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(main, P.join[g], 0)) != cudaSuccess) return e;
        g0 += cnt;
    }
    return cudaSuccess;
}

static FactorView sub_view(const FactorView& v, int g0) {
    FactorView s = v;
    s.F = v.F + (size_t)g0 * v.fstride;
    s.DL = v.DL + (size_t)g0 * v.dstride;
    s.DU = v.DU + (size_t)g0 * v.dstride;
    return s;
}

constexpr int JSPLIT = 8;  // split of the p-reduction in the B = V^T YR product

// ---- workspace layout (doubles; every segment 32-double aligned) -------------------------------
struct Workspace {
    int n, d, p, q, np, nb, ntiles;
    size_t fstride, dstride, tstride;
    double *F, *DL, *DU, *T, *B, *alpha, *mk, *atil, *gemv_part, *logdet_part, *quad, *tile_part, *V, *Z, *bpart;
    double *ell, *s0, *lnug, *lsig, *out;
    int32_t* info;
    int* sync;          // flags / ticket counters of the persistent Cholesky kernel, one region per stream group
    size_t total_doubles;
};

static inline size_t al(size_t x) { return (x + 31) / 32 * 32; }
// group g (latents [g0, g0 + cnt)) uses the region at sync_off(nb, g, g0) of potrf_pll_sync_ints(nb, cnt) ints
static inline size_t sync_off(int nb, int g, int g0) { return (size_t)8 * g + (size_t)10 * nb * g0; }
static inline size_t sync_ints(int nb, int q) { return (size_t)8 * MAX_GROUPS + (size_t)10 * nb * q; }

// q = all latents of the call (E emulators x q / E latents each)
static Workspace layout(int n, int d, int p, int q, int E, void* basep) {
    Workspace w;
    w.n = n; w.d = d; w.p = p; w.q = q;
    if (E < 1) E = 1;
    w.np = round_up(n, NB);
    w.nb = w.np / NB;
    w.ntiles = w.nb * (w.nb + 1) / 2;
    w.fstride = (size_t)w.np * w.np;
    w.dstride = (size_t)w.nb * NB * NB;
    w.tstride = trtri_scratch_blocks(w.nb) * NB * NB;
    double* base = (double*)basep;
    size_t o = 0;
    auto take = [&](size_t cnt) { double* ptr = base ? base + o : nullptr; o += al(cnt); return ptr; };
    w.F = take((size_t)q * w.fstride);
    w.DL = take((size_t)q * w.dstride);
    w.DU = take((size_t)q * w.dstride);
    w.T = take((size_t)q * w.tstride);
    w.B = take((size_t)q * w.np);
    w.alpha = take((size_t)q * w.np);
    w.mk = take((size_t)q * w.np);
    w.atil = take((size_t)q * w.np);
    w.gemv_part = take((size_t)q * (w.nb + 1) * w.np);
    w.logdet_part = take((size_t)q * w.nb);
    w.quad = take((size_t)q);
    w.tile_part = take((size_t)q * w.ntiles * (d + 2));
    w.V = take((size_t)p * q);
    w.Z = take((size_t)p * q);
    w.bpart = take((size_t)JSPLIT * q * w.np);
    w.ell = take((size_t)q * d);
    w.s0 = take((size_t)q);
    w.lnug = take((size_t)q);
    w.lsig = take((size_t)E * p);
    w.out = take((size_t)E * lcgp_out_len(p, d, q / E));
    w.info = (int32_t*)take((size_t)(q + 1) / 2 + 1);
    w.sync = (int*)take((sync_ints(w.nb, q) + 1) / 2);
    w.total_doubles = o;
    return w;
}

static FactorView view_of(const Workspace& w) {
    FactorView v;
    v.F = w.F; v.DL = w.DL; v.DU = w.DU; v.np = w.np; v.nb = w.nb; v.fstride = w.fstride; v.dstride = w.dstride; v.n = w.n;
    return v;
}

// ---- glue kernels -----------------------------------------------------------------------------
// V[j][k] = s_j phi[j][k],  s_j = exp(-lsig_j / 2) t_j        (sigma_inv_sqrt * phi[:, k], lcgp.py:608)
// (E emulators: V, phi are [E][p][q], lsig, t are [E][p]; idx / q is the flat (e, j) index)
__global__ void prep_v_kernel(int pE, int q, const double* __restrict__ lsig, const double* __restrict__ t,
                              const double* __restrict__ phi, double* __restrict__ V) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= pE * q) return;
    const int j = idx / q;
    V[idx] = exp(-0.5 * lsig[j]) * t[j] * phi[idx];
}

// bpart[js][k][i] = sum_{j in slice js} V[j][k] YR[j][i]      (b_k = r * ybar^T v_k, lcgp.py:609-610)
template <int KC>
__global__ void __launch_bounds__(128)
// q = latents per emulator, qtot = all latents; blockIdx.z = (emulator, chunk of KC latents of that emulator)
bmat_part_kernel(int n, int np, int p, int q, int qtot, const double* __restrict__ V, const double* __restrict__ YR,
                 double* __restrict__ bpart) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    const int js = blockIdx.y;
    const int chunks = (q + KC - 1) / KC;
    const int e = blockIdx.z / chunks;
    const int k0 = (blockIdx.z % chunks) * KC;
    V += (size_t)e * p * q;
    YR += (size_t)e * p * n;
    bpart += (size_t)e * q * np;
    const int jper = (p + JSPLIT - 1) / JSPLIT;
    const int j0 = js * jper, j1 = min(p, j0 + jper);
    double acc[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) acc[c] = 0.0;
    if (i < n) {
        for (int j = j0; j < j1; ++j) {
            const double yv = YR[(size_t)j * n + i];
#pragma unroll
            for (int c = 0; c < KC; ++c)
                if (k0 + c < q) acc[c] += __ldg(V + (size_t)j * q + k0 + c) * yv;
        }
    }
    if (i < np) {
#pragma unroll
        for (int c = 0; c < KC; ++c)
            if (k0 + c < q) bpart[((size_t)js * qtot + k0 + c) * np + i] = acc[c];
    }
}

__global__ void bmat_reduce_kernel(int np, int q, const double* __restrict__ bpart, double* __restrict__ B) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)q * np) return;
    double s = 0.0;
#pragma unroll
    for (int js = 0; js < JSPLIT; ++js) s += bpart[(size_t)js * q * np + idx];
    B[idx] = s;
}

// Z[j][k] = sum_i YR[j][i] m_k[i]     (the only term coupling latents to the noise parameters)
template <int KC>
__global__ void __launch_bounds__(256)
zmat_kernel(int n, int np, int p, int q, const double* __restrict__ YR, const double* __restrict__ mk,
            double* __restrict__ Z) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    const int chunks = (q + KC - 1) / KC;             // q = latents per emulator; blockIdx.y = (emulator, chunk)
    const int e = blockIdx.y / chunks;
    const int k0 = (blockIdx.y % chunks) * KC;
    YR += (size_t)e * p * n;
    mk += (size_t)e * q * np;
    Z += (size_t)e * p * q;
    if (j >= p) return;
    double acc[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) acc[c] = 0.0;
    // 4 columns per lane and trip: 36 independent loads in flight (one warp per output row: few warps per SM, so the
    // loop is latency bound; it took 5 % of a config-3 evaluation with one column per trip)
    const double* yr = YR + (size_t)j * n;
    int i = lane;
    for (; i + 96 < n; i += 128) {
        double yv[4], mv[4][KC];
#pragma unroll
        for (int u = 0; u < 4; ++u) yv[u] = yr[i + 32 * u];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < KC; ++c) mv[u][c] = (k0 + c < q) ? mk[(size_t)(k0 + c) * np + i + 32 * u] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)                  // same summation order as the one-column loop
#pragma unroll
            for (int c = 0; c < KC; ++c) acc[c] += yv[u] * mv[u][c];
    }
    for (; i < n; i += 32) {
        const double yv = yr[i];
#pragma unroll
        for (int c = 0; c < KC; ++c)
            if (k0 + c < q) acc[c] += yv * mk[(size_t)(k0 + c) * np + i];
    }
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        const double s = warp_sum(acc[c]);
        if (lane == 0 && k0 + c < q) Z[(size_t)j * q + k0 + c] = s;
    }
}

// Objective value, d/d lsigma2 and scaling of the kernel-parameter gradients.
__global__ void __launch_bounds__(256)
finalize_kernel(lcgp_problem P, int nb, int with_grad, const double* __restrict__ lsig,
                const double* __restrict__ logdet_part, const double* __restrict__ quad,
                const double* __restrict__ Z, double* __restrict__ out) {
    __shared__ double red[8];
    const int tid = threadIdx.x;
    const int E = P.n_emu > 1 ? P.n_emu : 1;
    const int p = P.p, q = P.q_loc / E, d = P.d;
    {   // one block per emulator: shift every per-emulator array (the emulators' constants follow each other)
        const int e = blockIdx.x;
        lsig += (size_t)e * p; logdet_part += (size_t)e * q * nb; quad += (size_t)e * q; Z += (size_t)e * p * q;
        out += (size_t)e * (1 + p + (size_t)q * d + 4 * (size_t)q);
        P.w += (size_t)e * p; P.t += (size_t)e * p; P.phi += (size_t)e * p * q;
        if (P.n_emu > 1) { P.scale = P.emu_consts[2 * e]; P.sum_log_r = P.emu_consts[2 * e + 1]; }
    }
    double* g_sig = out + 1;
    double* g_kern = out + 1 + p;                       // q*d + q + q values
    double* diag_logdet = g_kern + (size_t)q * d + 2 * q;
    double* diag_quad = diag_logdet + q;
    // latent part
    double lat = 0.0;
    for (int k = tid; k < q; k += 256) {
        double ld = 0.0;
        for (int b = 0; b < nb; ++b) ld += logdet_part[(size_t)k * nb + b];
        diag_logdet[k] = 2.0 * ld;
        diag_quad[k] = quad[k];
        lat += ld - 0.5 * quad[k];
    }
    lat = block_sum(lat, red);
    // host part: 1/2 sum_j s_j^2 w_j + n/2 sum_j (lsig_j - 2 log t_j) - p/2 sum log r   (lcgp.py:589-597, 663-664)
    double host = 0.0;
    for (int j = tid; j < p; j += 256) {
        const double tj = P.t[j];
        const double sj = exp(-0.5 * lsig[j]) * tj;
        const double fit = 0.5 * sj * sj * P.w[j];
        if (P.include_host_terms) host += fit + 0.5 * (double)P.n * (lsig[j] - 2.0 * log(tj));
        if (with_grad) {
            double zs = 0.0;
            for (int k = 0; k < q; ++k) zs += P.phi[(size_t)j * q + k] * Z[(size_t)j * q + k];
            double g = 0.5 * sj * zs;
            if (P.include_host_terms) g += -fit + 0.5 * (double)P.n;
            g_sig[j] = P.scale * g;
        }
    }
    host = block_sum(host, red);
    if (tid == 0) {
        double tot = lat + host;
        if (P.include_host_terms) tot -= 0.5 * (double)p * P.sum_log_r;
        out[0] = P.scale * tot;
    }
    if (with_grad)
        for (int c = tid; c < q * d + 2 * q; c += 256) g_kern[c] *= P.scale;
}

// Sharded evaluation: this rank's `out` vector (local latent order) -> the flat all-reduce vector
//   [objective | d/d lsigma2 (p) | d/d lLmb (q x d) | d/d lLmb0 (q) | d/d lnugGPs (q) | number of failed latents]
// in GLOBAL latent order, zero in the rows of latents other ranks own.  loc_of[k] = local index of latent k, -1 if
// not owned.  The last slot makes a Cholesky failure collective: after the all-reduce every rank sees it.
__global__ void pack_sharded_kernel(int p, int d, int q, int q_loc, const int* __restrict__ loc_of,
                                    const double* __restrict__ out, const int32_t* __restrict__ info,
                                    double* __restrict__ flat) {
    const int nflat = 1 + p + q * d + 2 * q;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e > nflat) return;
    if (e == nflat) {
        int bad = 0;
        for (int k = 0; k < q_loc; ++k) bad += info[k] != 0;
        flat[e] = (double)bad;
        return;
    }
    double v = 0.0;
    if (e < 1 + p) {
        v = out[e];
    } else {
        const double* gk = out + 1 + p;    // local layout: g_ell (q_loc x d), g_s0 (q_loc), g_nug (q_loc)
        const int c = e - 1 - p;
        if (c < q * d) {
            const int l = loc_of[c / d];
            if (l >= 0) v = gk[(size_t)l * d + c % d];
        } else if (c < q * d + q) {
            const int l = loc_of[c - q * d];
            if (l >= 0) v = gk[(size_t)q_loc * d + l];
        } else {
            const int l = loc_of[c - q * d - q];
            if (l >= 0) v = gk[(size_t)q_loc * d + q_loc + l];
        }
    }
    flat[e] = v;
}

static inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? 0 : 1000 + (int)e; }
#define LCGP_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return cuda_rc(e__); } while (0)

static int check_problem(const lcgp_problem* P) {
    if (!P || P->n <= 0 || P->d <= 0 || P->p <= 0 || P->q_loc <= 0) return LCGP_E_ARG;
    if (!P->X || !P->sr || !P->YR || !P->w || !P->t || !P->phi || !P->D) return LCGP_E_ARG;
    if (P->d > LCGP_MAX_D) return LCGP_E_DIM;
    if (P->n_emu < 0 || (P->n_emu > 1 && (P->q_loc % P->n_emu != 0 || !P->emu_consts))) return LCGP_E_ARG;
    return 0;
}
static inline int n_emu(const lcgp_problem* P) { return P->n_emu > 1 ? P->n_emu : 1; }

static int nll_grad_impl(const lcgp_problem* P, const double* ell, const double* s0, const double* lnug,
                         const double* lsig, const Workspace& w, double* out, int32_t* info, int flags,
                         void* const* ev, cudaStream_t st) {
    const int n = P->n, d = P->d, p = P->p, q = P->q_loc;
    const int E = n_emu(P), qe = q / E;             // emulators in this call, latents per emulator
    const int with_grad = flags & 1;
    FactorView v = view_of(w);
    KernelParams kp{ell, s0, lnug, P->D, qe};
    auto rec = [&](int i) { if (ev && ev[i]) cudaEventRecord((cudaEvent_t)ev[i], st); };
    rec(0);
    LCGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t) * q, st));
    // b_k
    note_launch(); prep_v_kernel<<<(E * p * qe + 255) / 256, 256, 0, st>>>(E * p, qe, lsig, P->t, P->phi, w.V);
    LCGP_CUDA(cudaGetLastError());
    note_launch(); bmat_part_kernel<8><<<dim3(w.np / 128, JSPLIT, E * ((qe + 7) / 8)), 128, 0, st>>>(n, w.np, p, qe, q, w.V, P->YR, w.bpart);
    LCGP_CUDA(cudaGetLastError());
    note_launch(); bmat_reduce_kernel<<<(unsigned)(((size_t)q * w.np + 255) / 256), 256, 0, st>>>(w.np, q, w.bpart, w.B);
    LCGP_CUDA(cudaGetLastError());
    // A_k
    LCGP_CUDA(launch_build_A(P->X, P->sr, n, d, w.np, kp, w.F, w.fstride, q, st));
    rec(1);
    int G = (flags >> 4) & 15;
    G = G == 0 ? stream_groups(q) : (G > MAX_GROUPS ? MAX_GROUPS : (G > q ? q : G));
    // The persistent Cholesky kernel keeps every SM it gets until its tickets run out, so concurrent groups mostly run
    // one after the other and only overlap at their tails; few, small matrices are better off in ONE launch (all
    // their tiles advance together) unless the caller / environment asked for a group count.
    {   // under stream capture by the CALLER the library's pooled side streams must not be pulled into the capture
        // (another thread waiting on them would invalidate it): everything stays on the caller's stream
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
            G = 1;
            flags |= LCGP_FLAG_NO_LOOKAHEAD;
        }
    }
    const bool pll = potrf_use_pll(w.nb, q);        // decided on this call's whole batch, not per stream group
    if (pll && ((flags >> 4) & 15) == 0 && !std::getenv("LCGP_STREAMS") && (q < 8 || w.nb <= 16)) G = 1;
    // Look-ahead uses the library's shared high-priority streams: skipped when the caller asked for "everything on
    // my stream" (bits 4-7 == 1: many small emulators driven from several host threads would falsely serialise on
    // them) and for small matrices, where no trailing update is big enough to hide a panel behind.
    const bool fuse = pll && potrf_fuse_trtri();    // triangular inverse inside the persistent Cholesky kernel
    const bool look = !pll && lookahead_on() && !(flags & LCGP_FLAG_NO_LOOKAHEAD) && ((flags >> 4) & 15) != 1 &&
                      w.nb > 16;
    auto potrf_group = [&](int g0, int cnt, cudaStream_t s, int g) {
        Lookahead la;
        if (look) {   // run_grouped holds the pool lock
            SidePool& sp = side_pool();
            la.panel = sp.hp[g]; la.ev_panel = sp.evp[g]; la.ev_bulk = sp.evb[g];
        }
        return potrf_batched(sub_view(v, g0), w.DL + (size_t)g0 * w.dstride, w.DU + (size_t)g0 * w.dstride, cnt,
                             w.logdet_part + (size_t)g0 * w.nb, info + g0, panel_width(), s, la,
                             pll ? w.sync + sync_off(w.nb, g, g0) : nullptr, fuse);
    };
    auto trtri_group = [&](int g0, int cnt, cudaStream_t s, int) {
        if (fuse) return cudaSuccess;               // done by the persistent Cholesky launch
        return trtri_batched(sub_view(v, g0), w.T + (size_t)g0 * w.tstride, w.tstride, cnt, s);
    };
    if (ev) {
        // stage timing requested: join the groups between Cholesky and triangular inverse so that the
        // two stages can be timed separately
        LCGP_CUDA(run_grouped(st, q, G, look, potrf_group));
        rec(2);
        LCGP_CUDA(run_grouped(st, q, G, false, trtri_group));
        rec(3);
    } else {
        // production path: each group flows from its Cholesky straight into its triangular inverse, so
        // one group's serial panel work and launch tails overlap another group's GEMMs
        LCGP_CUDA(run_grouped(st, q, G, look, [&](int g0, int cnt, cudaStream_t s, int g) {
            cudaError_t e = potrf_group(g0, cnt, s, g);
            return e != cudaSuccess ? e : trtri_group(g0, cnt, s, g);
        }));
    }
    SolveArgs a;
    a.n = n; a.d = d; a.p = p; a.np = w.np; a.nb = w.nb; a.q_loc = q;
    a.X = P->X; a.sr = P->sr; a.B = w.B; a.kp = kp;
    a.alpha = w.alpha; a.mk = w.mk; a.atil = w.atil; a.gemv_part = w.gemv_part; a.quad = w.quad;
    LCGP_CUDA(solve_alpha(v, a, st));
    double* g_kern = out + 1 + p;
    if (with_grad) {
        LCGP_CUDA(contract_grad(v, a, w.tile_part, g_kern, lcgp_out_len(p, d, qe),
                                ev ? (cudaEvent_t)ev[4] : nullptr, ev ? (cudaEvent_t)ev[5] : nullptr, st));
        note_launch(); zmat_kernel<8><<<dim3((p + 7) / 8, E * ((qe + 7) / 8)), 256, 0, st>>>(n, w.np, p, qe, P->YR, w.mk, w.Z);
        LCGP_CUDA(cudaGetLastError());
    } else {
        rec(4);
        rec(5);
    }
    note_launch(); finalize_kernel<<<E, 256, 0, st>>>(*P, w.nb, with_grad, lsig, w.logdet_part, w.quad, w.Z, out);
    LCGP_CUDA(cudaGetLastError());
    rec(6);
    return 0;
}

}  // namespace lcgp

using namespace lcgp;

extern "C" {

unsigned long long lcgp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* lcgp_version(void) { return "lcgp_b200 0.1 (sm_100a; DMMA m8n8k4 GEMM, NB=128)"; }

size_t lcgp_out_len(int32_t p, int32_t d, int32_t q_loc) {
    return (size_t)1 + p + (size_t)q_loc * d + 2 * (size_t)q_loc + 2 * (size_t)q_loc;
}

size_t lcgp_workspace_bytes(int32_t n, int32_t d, int32_t p, int32_t q_loc) {
    if (n <= 0 || d <= 0 || p <= 0 || q_loc <= 0) return 0;
    Workspace w = layout(n, d, p, q_loc, 1, nullptr);
    return w.total_doubles * sizeof(double);
}

size_t lcgp_workspace_bytes_batched(int32_t n, int32_t d, int32_t p, int32_t q_per, int32_t n_emu) {
    if (n <= 0 || d <= 0 || p <= 0 || q_per <= 0 || n_emu <= 0) return 0;
    Workspace w = layout(n, d, p, q_per * n_emu, n_emu, nullptr);
    return w.total_doubles * sizeof(double);
}

size_t lcgp_predict_scratch_bytes(int32_t n, int32_t q_loc, int32_t n0) {
    if (n <= 0 || q_loc <= 0 || n0 <= 0) return 0;
    const size_t np = round_up(n, NB), n0p = round_up(n0, NB), nb = np / NB;
    return sizeof(double) * (size_t)q_loc * (n0p * np + nb * n0p);
}

int lcgp_nll_grad(const lcgp_problem* P, const double* lLmb, const double* lLmb0, const double* lnugGPs,
                  const double* lsigma2_p, void* workspace, size_t workspace_bytes, double* out, int32_t* info,
                  int32_t flags, void* const* stage_events, void* stream) {
    int rc = check_problem(P);
    if (rc) return rc;
    if (!lLmb || !lLmb0 || !lnugGPs || !lsigma2_p || !workspace || !out || !info) return LCGP_E_ARG;
    Workspace w = layout(P->n, P->d, P->p, P->q_loc, n_emu(P), workspace);
    if (workspace_bytes < w.total_doubles * sizeof(double)) return LCGP_E_WORKSPACE;
    return nll_grad_impl(P, lLmb, lLmb0, lnugGPs, lsigma2_p, w, out, info, flags, stage_events, (cudaStream_t)stream);
}

int lcgp_nll_grad_host(const lcgp_problem* P, const double* lLmb_h, const double* lLmb0_h, const double* lnug_h,
                       const double* lsig_h, void* workspace, size_t workspace_bytes, double* out_h,
                       int32_t* info_h, int32_t flags, void* const* stage_events, void* stream) {
    int rc = check_problem(P);
    if (rc) return rc;
    if (!lLmb_h || !lLmb0_h || !lnug_h || !lsig_h || !workspace || !out_h || !info_h) return LCGP_E_ARG;
    Workspace w = layout(P->n, P->d, P->p, P->q_loc, n_emu(P), workspace);
    if (workspace_bytes < w.total_doubles * sizeof(double)) return LCGP_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int q = P->q_loc;
    LCGP_CUDA(cudaMemcpyAsync(w.ell, lLmb_h, sizeof(double) * q * P->d, cudaMemcpyHostToDevice, st));
    LCGP_CUDA(cudaMemcpyAsync(w.s0, lLmb0_h, sizeof(double) * q, cudaMemcpyHostToDevice, st));
    LCGP_CUDA(cudaMemcpyAsync(w.lnug, lnug_h, sizeof(double) * q, cudaMemcpyHostToDevice, st));
    LCGP_CUDA(cudaMemcpyAsync(w.lsig, lsig_h, sizeof(double) * n_emu(P) * P->p, cudaMemcpyHostToDevice, st));
    rc = nll_grad_impl(P, w.ell, w.s0, w.lnug, w.lsig, w, w.out, w.info, flags, stage_events, st);
    if (rc) return rc;
    LCGP_CUDA(cudaMemcpyAsync(out_h, w.out, sizeof(double) * n_emu(P) * lcgp_out_len(P->p, P->d, q / n_emu(P)),
                              cudaMemcpyDeviceToHost, st));
    LCGP_CUDA(cudaMemcpyAsync(info_h, w.info, sizeof(int32_t) * q, cudaMemcpyDeviceToHost, st));
    LCGP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"

// ---- evaluation plans: one objective(+gradient) evaluation captured as a CUDA graph -----------------------
// Small problems (n <= 2048: ~100 kernels of 20-60 us) are bound by the driver's launch rate, not by the GPU --
// above all when several host threads fit independent emulators on one device (BASELINE config 5) and serialise
// on the context lock.  A plan captures the parameter upload, every kernel of lcgp_nll_grad (including the fork /
// join over the internal stream groups) and the result download ONCE and replays it with a single launch.
struct lcgp_plan {
    lcgp_problem prob;
    void* ws;
    size_t ws_bytes;
    const double* par_h;   // [lLmb | lLmb0 | lnugGPs | lsigma2_p], pinned host memory owned by the caller
    double* out_h;
    int32_t* info_h;
    int32_t flags;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool eager = false;    // capture was refused once: run the plain call from now on
    cudaStream_t own = nullptr;   // the plan's private stream: captured / replayed on, never shared with another plan
    cudaEvent_t ev = nullptr;     // orders the plan stream after the caller's stream
};

static int plan_enqueue(const lcgp_plan* pl, cudaStream_t st) {
    const lcgp_problem* P = &pl->prob;
    const int q = P->q_loc, d = P->d, p = P->p;
    const int E = n_emu(P);
    Workspace w = layout(P->n, d, p, q, E, pl->ws);
    const double* h = pl->par_h;
    LCGP_CUDA(cudaMemcpyAsync(w.ell, h, sizeof(double) * q * d, cudaMemcpyHostToDevice, st));
    LCGP_CUDA(cudaMemcpyAsync(w.s0, h + (size_t)q * d, sizeof(double) * q, cudaMemcpyHostToDevice, st));
    LCGP_CUDA(cudaMemcpyAsync(w.lnug, h + (size_t)q * d + q, sizeof(double) * q, cudaMemcpyHostToDevice, st));
    LCGP_CUDA(cudaMemcpyAsync(w.lsig, h + (size_t)q * d + 2 * q, sizeof(double) * E * p, cudaMemcpyHostToDevice, st));
    int rc = nll_grad_impl(P, w.ell, w.s0, w.lnug, w.lsig, w, w.out, w.info, pl->flags, nullptr, st);
    if (rc) return rc;
    LCGP_CUDA(cudaMemcpyAsync(pl->out_h, w.out, sizeof(double) * E * lcgp_out_len(p, d, q / E), cudaMemcpyDeviceToHost, st));
    LCGP_CUDA(cudaMemcpyAsync(pl->info_h, w.info, sizeof(int32_t) * q, cudaMemcpyDeviceToHost, st));
    return 0;
}

extern "C" {

int lcgp_plan_create(const lcgp_problem* P, void* workspace, size_t workspace_bytes, const double* params_host,
                     double* out_host, int32_t* info_host, int32_t flags, lcgp_plan** plan) {
    int rc = check_problem(P);
    if (rc) return rc;
    if (!workspace || !params_host || !out_host || !info_host || !plan) return LCGP_E_ARG;
    Workspace w = layout(P->n, P->d, P->p, P->q_loc, n_emu(P), workspace);
    if (workspace_bytes < w.total_doubles * sizeof(double)) return LCGP_E_WORKSPACE;
    lcgp_plan* pl = new (std::nothrow) lcgp_plan();
    if (!pl) return LCGP_E_ARG;
    pl->prob = *P; pl->ws = workspace; pl->ws_bytes = workspace_bytes;
    pl->par_h = params_host; pl->out_h = out_host; pl->info_h = info_host;
    // a plan always runs on the caller's stream alone (stream-group bits forced to 1, which also switches the
    // look-ahead off): the library's shared side streams must not be pulled into one thread's capture
    pl->flags = (flags & ~0xF0) | (1 << 4);
    *plan = pl;
    return 0;
}

int lcgp_plan_run(lcgp_plan* pl, void* stream) {
    if (!pl) return LCGP_E_ARG;
    if (!pl->own) {
        LCGP_CUDA(cudaStreamCreateWithFlags(&pl->own, cudaStreamNonBlocking));
        LCGP_CUDA(cudaEventCreateWithFlags(&pl->ev, cudaEventDisableTiming));
    }
    cudaStream_t st = pl->own;
    // everything the caller queued on `stream` (constant uploads, earlier work on the workspace) comes first
    LCGP_CUDA(cudaEventRecord(pl->ev, (cudaStream_t)stream));
    LCGP_CUDA(cudaStreamWaitEvent(st, pl->ev, 0));
    if (!pl->exec && !pl->eager) {
        // first evaluation: launch by launch (this also sets the lazily-configured kernel attributes outside
        // the capture), then capture the same sequence for the evaluations to come
        int rc = plan_enqueue(pl, st);
        if (rc) return rc;
        LCGP_CUDA(cudaStreamSynchronize(st));
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = plan_enqueue(pl, st);
            cudaGraph_t g = nullptr;
            const cudaError_t ec = cudaStreamEndCapture(st, &g);
            if (rc == 0 && ec == cudaSuccess && g && cudaGraphInstantiate(&pl->exec, g, 0) == cudaSuccess) {
                pl->graph = g;
            } else {
                if (g) cudaGraphDestroy(g);
                pl->exec = nullptr;
                pl->eager = true;
            }
        } else {
            pl->eager = true;
        }
        (void)cudaGetLastError();   // a refused capture leaves a sticky-looking error behind; the plan stays usable
        return 0;                   // results of the launch-by-launch evaluation above are already on the host
    }
    if (pl->exec) {
        note_launch();   // the graph launch itself; its kernel nodes were counted when they were captured
        LCGP_CUDA(cudaGraphLaunch(pl->exec, st));
    } else {
        int rc = plan_enqueue(pl, st);
        if (rc) return rc;
    }
    LCGP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int lcgp_plan_is_graph(const lcgp_plan* pl) { return pl && pl->exec ? 1 : 0; }

void lcgp_plan_destroy(lcgp_plan* pl) {
    if (!pl) return;
    if (pl->exec) cudaGraphExecDestroy(pl->exec);
    if (pl->graph) cudaGraphDestroy(pl->graph);
    if (pl->own) { cudaStreamSynchronize(pl->own); cudaStreamDestroy(pl->own); }
    if (pl->ev) cudaEventDestroy(pl->ev);
    delete pl;
}

int lcgp_predict(const lcgp_problem* P, const double* lLmb, const double* lLmb0, const double* lnugGPs,
                 void* workspace, size_t workspace_bytes, const double* x0s, int32_t n0, int32_t same_inputs,
                 void* scratch, size_t scratch_bytes, double* ghat, double* gvar, void* stream) {
    int rc = check_problem(P);
    if (rc) return rc;
    if (!lLmb || !lLmb0 || !lnugGPs || !workspace || !x0s || n0 <= 0 || !scratch || !ghat || !gvar) return LCGP_E_ARG;
    Workspace w = layout(P->n, P->d, P->p, P->q_loc, n_emu(P), workspace);
    if (workspace_bytes < w.total_doubles * sizeof(double)) return LCGP_E_WORKSPACE;
    if (scratch_bytes < lcgp_predict_scratch_bytes(P->n, P->q_loc, n0)) return LCGP_E_WORKSPACE;
    if (n_emu(P) > 1) return LCGP_E_ARG;            // prediction works on one emulator (build it from the fitted parameters)
    KernelParams kp{lLmb, lLmb0, lnugGPs, P->D, P->q_loc};
    return cuda_rc(predict_latents(view_of(w), P->n, P->d, P->X, P->sr, kp, w.atil, x0s, n0, same_inputs,
                                   (double*)scratch, P->q_loc, ghat, gvar, (cudaStream_t)stream));
}

int lcgp_predict_outputs(const double* Psi, const double* ghat, const double* gvar, const double* noise_var,
                         const double* scale, const double* shift, int32_t p, int32_t q, int32_t n0, double* ypred,
                         double* ypredvar, double* yconfvar, void* stream) {
    if (!Psi || !ghat || !gvar || !noise_var || !ypred || !ypredvar || !yconfvar || p <= 0 || q <= 0 || n0 <= 0) return LCGP_E_ARG;
    if ((size_t)q * 8 * sizeof(double) > 200 * 1024) return LCGP_E_DIM;
    return cuda_rc(launch_predict_outputs(Psi, ghat, gvar, noise_var, scale, shift, p, q, n0, ypred, ypredvar, yconfvar,
                                          (cudaStream_t)stream));
}

int lcgp_predict_fullcov(const double* psi, const double* gvar, const double* sig2, const double* ystd, int32_t q,
                         int32_t p, int32_t n0, double* out, void* stream) {
    if (!psi || !gvar || !sig2 || !ystd || !out || q <= 0 || p <= 0 || n0 <= 0) return LCGP_E_ARG;
    return cuda_rc(launch_fullcov(psi, gvar, sig2, ystd, q, p, n0, out, (cudaStream_t)stream));
}

int lcgp_prep_segment_mean(const double* y, const int32_t* order, const int32_t* offsets, int32_t p, int32_t N,
                           int32_t n, double* ybar, void* stream) {
    if (!y || !order || !offsets || !ybar || p <= 0 || N <= 0 || n <= 0 || n > N) return LCGP_E_ARG;
    return cuda_rc(prep_segment_mean(y, order, offsets, p, N, n, ybar, (cudaStream_t)stream));
}

int lcgp_prep_row_select(const double* Y, const double* center, int32_t p, int32_t m, int32_t k, double* out,
                         void* stream) {
    if (!Y || !out || p <= 0 || m <= 0 || k < 0 || k >= m) return LCGP_E_ARG;
    return cuda_rc(prep_row_select(Y, center, p, m, k, out, (cudaStream_t)stream));
}

int lcgp_prep_standardize(const double* Y, const double* center, const double* spread, const double* r, int32_t p,
                          int32_t n, double* Ys, double* YR, double* w, void* stream) {
    if (!Y || !center || !spread || p <= 0 || n <= 0) return LCGP_E_ARG;
    return cuda_rc(prep_standardize(Y, center, spread, r, p, n, Ys, YR, w, (cudaStream_t)stream));
}

int lcgp_grad_phi(const lcgp_problem* P, const double* lsigma2_p, void* workspace, size_t workspace_bytes, double* g_phi,
                  void* stream) {
    int rc = check_problem(P);
    if (rc) return rc;
    if (!lsigma2_p || !workspace || !g_phi || n_emu(P) > 1) return LCGP_E_ARG;
    Workspace w = layout(P->n, P->d, P->p, P->q_loc, n_emu(P), workspace);
    if (workspace_bytes < w.total_doubles * sizeof(double)) return LCGP_E_WORKSPACE;
    // scratch: the GEMV partial-sum area, free once lcgp_nll_grad has finished
    return cuda_rc(grad_phi(view_of(w), P->n, P->p, P->q_loc, P->scale, P->sr, w.mk, lsigma2_p, P->t, P->phi, P->D, w.Z,
                            w.gemv_part, g_phi, (cudaStream_t)stream));
}

int lcgp_pack_sharded(const double* out, const int32_t* info, const int32_t* loc_of, int32_t p, int32_t d, int32_t q,
                      int32_t q_loc, double* flat, void* stream) {
    if (!out || !info || !loc_of || !flat || p <= 0 || d <= 0 || q <= 0 || q_loc <= 0 || q_loc > q) return LCGP_E_ARG;
    const int nflat = 1 + p + q * d + 2 * q + 1;
    note_launch(); pack_sharded_kernel<<<(nflat + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, d, q, q_loc, loc_of, out, info, flat);
    return cuda_rc(cudaGetLastError());
}

int lcgp_get_aux(const lcgp_problem* P, void* workspace, size_t workspace_bytes, double* CinvMs, double* mks,
                 void* stream) {
    int rc = check_problem(P);
    if (rc) return rc;
    if (!workspace || !CinvMs || !mks) return LCGP_E_ARG;
    Workspace w = layout(P->n, P->d, P->p, P->q_loc, n_emu(P), workspace);
    if (workspace_bytes < w.total_doubles * sizeof(double)) return LCGP_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t rowb = sizeof(double) * P->n;
    LCGP_CUDA(cudaMemcpy2DAsync(CinvMs, rowb, w.alpha, sizeof(double) * w.np, rowb, P->q_loc, cudaMemcpyDeviceToDevice, st));
    LCGP_CUDA(cudaMemcpy2DAsync(mks, rowb, w.mk, sizeof(double) * w.np, rowb, P->q_loc, cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // extern "C"

// ---- A^{-1} materialisation (diagnostic / Tks, Ths on request) -----------------------------------
namespace lcgp {
// Ainv[i][j] = sum_{k >= max(i,j)} U[i][k] U[j][k]; plain FP64 FMA kernel, 16x16 output tile per CTA.
__global__ void __launch_bounds__(256)
ainv_kernel(FactorView v, int kz, int n, double* __restrict__ Ainv) {
    __shared__ double ui[16][17], uj[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
    const double* F = v.F + (size_t)kz * v.fstride;
    const double* DU = v.DU + (size_t)kz * v.dstride;
    auto U = [&](int i, int k) -> double {
        if (k < i) return 0.0;
        const int Ib = i / NB, Kb = k / NB;
        if (Ib == Kb) return DU[(size_t)Ib * NB * NB + (size_t)(i % NB) * NB + (k % NB)];
        return F[(size_t)i * v.np + k];
    };
    double acc = 0.0;
    const int kstart = (max(i0, j0) / 16) * 16;
    for (int k0 = kstart; k0 < v.np; k0 += 16) {
        ui[ty][tx] = U(i0 + ty, k0 + tx);
        uj[ty][tx] = U(j0 + ty, k0 + tx);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) acc += ui[ty][kk] * uj[tx][kk];
        __syncthreads();
    }
    const int i = i0 + ty, j = j0 + tx;
    if (i < n && j < n) Ainv[(size_t)i * n + j] = acc;
}
}  // namespace lcgp

extern "C" {

int lcgp_get_Ainv(const lcgp_problem* P, void* workspace, size_t workspace_bytes, int32_t k, double* Ainv, void* stream) {
    int rc = check_problem(P);
    if (rc) return rc;
    if (!workspace || !Ainv || k < 0 || k >= P->q_loc) return LCGP_E_ARG;
    Workspace w = layout(P->n, P->d, P->p, P->q_loc, n_emu(P), workspace);
    if (workspace_bytes < w.total_doubles * sizeof(double)) return LCGP_E_WORKSPACE;
    const int g = (P->n + 15) / 16;
    note_launch(); ainv_kernel<<<dim3(g, g), 256, 0, (cudaStream_t)stream>>>(view_of(w), k, P->n, Ainv);
    return cuda_rc(cudaGetLastError());
}

int lcgp_kernel_matrix(const double* x1, int32_t n1, const double* x2, int32_t n2, int32_t d, const double* llmb,
                       const double* llmb0, const double* lnug, int32_t same_inputs, double* out, void* stream) {
    if (!x1 || !x2 || !llmb || !llmb0 || !lnug || !out || n1 <= 0 || n2 <= 0 || d <= 0) return LCGP_E_ARG;
    if (d > LCGP_MAX_D) return LCGP_E_DIM;
    return cuda_rc(launch_matern_rect(x1, n1, x2, n2, d, llmb, llmb0, lnug, same_inputs, nullptr, out, n2, n1, n2, 1, 0,
                                      (cudaStream_t)stream));
}

int lcgp_build_A(const double* X, const double* sr, int32_t n, int32_t d, const double* lLmb, const double* lLmb0,
                 const double* lnugGPs, const double* D, int32_t batch, double* F, int32_t np, void* stream) {
    if (!X || !sr || !lLmb || !lLmb0 || !lnugGPs || !D || !F || n <= 0 || d <= 0 || batch <= 0) return LCGP_E_ARG;
    if (d > LCGP_MAX_D || np % NB != 0 || np < n) return LCGP_E_DIM;
    KernelParams kp{lLmb, lLmb0, lnugGPs, D, batch};
    return cuda_rc(launch_build_A(X, sr, n, d, np, kp, F, (size_t)np * np, batch, (cudaStream_t)stream));
}

size_t lcgp_potrf_scratch_bytes(int32_t np, int32_t batch) {
    if (np <= 0 || batch <= 0 || np % NB != 0) return 0;
    return sizeof(int) * potrf_pll_sync_ints(np / NB, batch);
}

int lcgp_potrf_batched(double* F, int32_t np, int32_t batch, double* DL, double* DU, double* logdet_part,
                       int32_t* info, void* scratch, size_t scratch_bytes, void* stream) {
    if (!F || !DL || !DU || !info || np <= 0 || batch <= 0) return LCGP_E_ARG;
    if (np % NB != 0) return LCGP_E_DIM;
    if (scratch && scratch_bytes < lcgp_potrf_scratch_bytes(np, batch)) return LCGP_E_WORKSPACE;
    FactorView v;
    v.F = F; v.DL = DL; v.DU = DU; v.np = np; v.nb = np / NB; v.n = np;
    v.fstride = (size_t)np * np; v.dstride = (size_t)v.nb * NB * NB;
    cudaStream_t st = (cudaStream_t)stream;
    LCGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t) * batch, st));
    if (scratch && potrf_use_pll(v.nb, batch))
        return cuda_rc(potrf_batched(v, DL, DU, batch, logdet_part, info, panel_width(), st, Lookahead(), (int*)scratch));
    return cuda_rc(run_grouped(st, batch, 1, lookahead_on(), [&](int, int, cudaStream_t s, int g) {
        Lookahead la;
        if (lookahead_on()) {
            SidePool& sp = side_pool();
            la.panel = sp.hp[g]; la.ev_panel = sp.evp[g]; la.ev_bulk = sp.evb[g];
        }
        return potrf_batched(v, DL, DU, batch, logdet_part, info, panel_width(), s, la);
    }));
}

int lcgp_potrf_trtri_batched(double* F, int32_t np, int32_t batch, double* DL, double* DU, double* logdet_part,
                             int32_t* info, void* scratch, size_t scratch_bytes, void* stream) {
    if (!F || !DL || !DU || !info || !scratch || np <= 0 || batch <= 0) return LCGP_E_ARG;
    if (np % NB != 0) return LCGP_E_DIM;
    if (scratch_bytes < lcgp_potrf_scratch_bytes(np, batch)) return LCGP_E_WORKSPACE;
    if (!potrf_use_pll()) return LCGP_E_ARG;        // the fused form exists for the persistent kernel only
    FactorView v;
    v.F = F; v.DL = DL; v.DU = DU; v.np = np; v.nb = np / NB; v.n = np;
    v.fstride = (size_t)np * np; v.dstride = (size_t)v.nb * NB * NB;
    cudaStream_t st = (cudaStream_t)stream;
    LCGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t) * batch, st));
    return cuda_rc(potrf_pll(v, DL, DU, batch, logdet_part, info, (int*)scratch, st, true));
}

size_t lcgp_trtri_scratch_bytes(int32_t np, int32_t batch) {
    if (np <= 0 || batch <= 0 || np % NB != 0) return 0;
    return sizeof(double) * (size_t)batch * trtri_scratch_blocks(np / NB) * NB * NB;
}

int lcgp_trtri_batched(double* F, int32_t np, int32_t batch, const double* DL, const double* DU, void* scratch,
                       size_t scratch_bytes, void* stream) {
    if (!F || !DL || !DU || np <= 0 || batch <= 0) return LCGP_E_ARG;
    if (np % NB != 0) return LCGP_E_DIM;
    const size_t need = lcgp_trtri_scratch_bytes(np, batch);
    if (need > 0 && (!scratch || scratch_bytes < need)) return LCGP_E_WORKSPACE;
    FactorView v;
    v.F = F; v.DL = DL; v.DU = DU; v.np = np; v.nb = np / NB; v.n = np;
    v.fstride = (size_t)np * np; v.dstride = (size_t)v.nb * NB * NB;
    return cuda_rc(trtri_batched(v, (double*)scratch, trtri_scratch_blocks(v.nb) * NB * NB, batch, (cudaStream_t)stream));
}

}  // extern "C"
