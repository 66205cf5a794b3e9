// Standalone per-row affinity kernels (HBM-bound: one read of C[n,k], one write of P[n,k]).
// All bisection steps run on-chip in registers; the reference makes ~30-130 full passes
// over C.
#include "rowsearch.cuh"

namespace tdr {

constexpr int kRowsPerBlock = 8;  // one warp per row

template <int EPL>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
umap_affinity_kernel(const float* __restrict__ C, int64_t n, int k, int max_iter, float target,
                     float* __restrict__ P, float* __restrict__ rho, float* __restrict__ sigma) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
    if (row >= n) return;
    UmapRow<EPL> r;
    r.k = k;
    r.lane = lane;
    r.target = target;
    const float* crow = C + row * k;
#pragma unroll
    for (int e = 0; e < EPL; ++e) r.c[e] = r.valid(e) ? __ldg(crow + lane + 32 * e) : INFINITY;
    r.init();
    const float s = r.solve(max_iter);
    float* prow = P + row * k;
#pragma unroll
    for (int e = 0; e < EPL; ++e)
        if (r.valid(e)) prow[lane + 32 * e] = r.p(e, s);
    if (lane == 0) {
        rho[row] = r.rho;
        sigma[row] = s;
    }
}

template <int EPL>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
entropic_affinity_kernel(const float* __restrict__ C, int64_t n, int k, int max_iter, EntropicConsts K,
                         float* __restrict__ logP, float* __restrict__ eps_out,
                         float* __restrict__ log_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
    if (row >= n) return;
    EntropicRow<EPL> r;
    r.k = k;
    r.lane = lane;
    r.target = K.target;
    const float* crow = C + row * k;
#pragma unroll
    for (int e = 0; e < EPL; ++e) r.c[e] = r.valid(e) ? __ldg(crow + lane + 32 * e) : INFINITY;
    const float eps = r.solve(K, max_iter);
    float l[EPL];
    const float z = r.lse(eps, l);  // entropic.py:299-303
    float* prow = logP + row * k;
#pragma unroll
    for (int e = 0; e < EPL; ++e)
        if (r.valid(e)) prow[lane + 32 * e] = __fsub_rn(__fsub_rn(l[e], z), K.log_n_total);  // :305-310
    if (lane == 0) {
        eps_out[row] = eps;
        log_norm[row] = z;
    }
}

template <int EPL>
static int launch_umap(const float* C, int64_t n, int k, int max_iter, float* P, float* rho,
                       float* sigma, cudaStream_t st) {
    const int64_t blocks = (n + kRowsPerBlock - 1) / kRowsPerBlock;
    umap_affinity_kernel<EPL><<<(unsigned)blocks, kRowsPerBlock * 32, 0, st>>>(
        C, n, k, max_iter, log2f((float)k), P, rho, sigma);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

template <int EPL>
static int launch_entropic(const float* C, int64_t n, int k, int max_iter, const EntropicConsts& K,
                           float* logP, float* eps, float* log_norm, cudaStream_t st) {
    const int64_t blocks = (n + kRowsPerBlock - 1) / kRowsPerBlock;
    entropic_affinity_kernel<EPL><<<(unsigned)blocks, kRowsPerBlock * 32, 0, st>>>(
        C, n, k, max_iter, K, logP, eps, log_norm);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

// ---- dense rows (EntropicAffinity(sparsity=False), BASELINE config 3) -----------------------------
// One CTA per row of the dense N x M distance matrix; the row (400 KB at M = 100 k) does not fit on
// chip, so every bisection step is one streaming pass over it (HBM-bound: 4 M bytes per step per
// row).  Entropy in one pass:  with  m = max_j l_j = -C_min/eps,  S = sum e^{l_j - m},
// T = sum e^{l_j - m}(l_j - m):  H = log S - T/S + 1  (same quantity as entropic.py:274-277 /
// utils/utils.py:167-168, reassociated so that logsumexp and entropy share a pass).
constexpr int kDenseThreads = 512;

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    for (int w = 0; w < kDenseThreads / 32; ++w) t += red[w];
    return t;
}
__device__ __forceinline__ float block_min(float v, float* red) {
    v = warp_min(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = INFINITY;
    for (int w = 0; w < kDenseThreads / 32; ++w) t = fminf(t, red[w]);
    return t;
}

__global__ void __launch_bounds__(kDenseThreads)
entropic_dense_kernel(const float* C, int64_t n_rows, int64_t m, int max_iter, EntropicConsts K,
                      float* logP /* may alias C */, float* __restrict__ eps_out, float* __restrict__ log_norm) {
    __shared__ float red[kDenseThreads / 32];
    const int64_t row = blockIdx.x;
    const float* c = C + row * m;
    // pass 0: smallest, second smallest (with multiplicity) and largest entry (entropic.py:93-98)
    float mn = INFINITY, mn2 = INFINITY, mx = -INFINITY;
    for (int64_t j = threadIdx.x; j < m; j += kDenseThreads) {
        const float v = c[j];
        if (v < mn) {
            mn2 = mn;
            mn = v;
        } else if (v < mn2) {
            mn2 = v;
        }
        mx = fmaxf(mx, v);
    }
    const float d1 = block_min(mn, red);
    // second smallest overall: every thread contributes its own second smallest, or its smallest if that
    // is not the (unique) holder of d1; count holders of d1 to honour multiplicity
    const float cnt = block_sum(mn == d1 ? 1.0f : 0.0f, red);
    const float cand = (mn == d1 && cnt < 1.5f) ? mn2 : (mn == d1 ? fminf(mn2, d1) : mn);
    const float d2 = block_min(cnt > 1.5f ? d1 : cand, red);
    const float dN = -block_min(-mx, red);

    auto gap = [&](float eps) {
        const float mm = __fdiv_rn(-d1, eps);
        float s = 0.0f, t = 0.0f;
        for (int64_t j = threadIdx.x; j < m; j += kDenseThreads) {
            const float x = __fsub_rn(__fdiv_rn(-c[j], eps), mm);
            const float e = expf(x);
            s += e;
            t = fmaf(e, x, t);
        }
        s = block_sum(s, red);
        t = block_sum(t, red);
        return __fsub_rn(__fadd_rn(__fsub_rn(logf(s), __fdiv_rn(t, s)), 1.0f), K.target);
    };
    float b = 1.0f, e = 1.0f;
    if (K.use_bounds) {  // entropic.py:99-115
        const float span = __fsub_rn(dN, d1), step = __fsub_rn(d2, d1);
        const float t1 = __fdiv_rn(K.b_num, __fmul_rn(K.b_den, span));
        const float t2 = sqrtf(__fdiv_rn(K.b_lr, __fsub_rn(__fmul_rn(dN, dN), __fmul_rn(d1, d1))));
        float beta_lo = fmaxf(t1, t2);
        if (t1 != t1 || t2 != t2) beta_lo = NAN;
        b = __fadd_rn(__fdiv_rn(1.0f, __fdiv_rn(K.b_logp1, step)), 1e-6f);
        e = __fdiv_rn(1.0f, beta_lo);
    }
    const float eps = bracket_bisect(gap, b, e, max_iter);
    // final normalisation (entropic.py:299-310)
    const float mm = __fdiv_rn(-d1, eps);
    float s = 0.0f;
    for (int64_t j = threadIdx.x; j < m; j += kDenseThreads) s += expf(__fsub_rn(__fdiv_rn(-c[j], eps), mm));
    s = block_sum(s, red);
    const float z = __fadd_rn(logf(s), mm);
    if (logP) {
        float* out = logP + row * m;
        for (int64_t j = threadIdx.x; j < m; j += kDenseThreads)
            out[j] = __fsub_rn(__fsub_rn(__fdiv_rn(-c[j], eps), z), K.log_n_total);
    }
    if (threadIdx.x == 0) {
        eps_out[row] = eps;
        log_norm[row] = z;
    }
}

}  // namespace tdr

using namespace tdr;

extern "C" TDR_API int tdr_entropic_dense_f32(const float* C, int64_t n_rows, int64_t m, float target_entropy,
                                              float log_n_total, int use_bounds, float b_num, float b_den, float b_lr,
                                              float b_logp1, int max_iter, float* logP, float* eps, float* log_norm,
                                              tdr_stream_t stream) {
    TDR_CHECK_ARG(C && eps && log_norm, "tdr_entropic_dense_f32: null pointer");
    TDR_CHECK_ARG(n_rows >= 0 && m >= 2 && n_rows < 0x7fffffffLL, "tdr_entropic_dense_f32: bad shape");
    if (n_rows == 0) return TDR_OK;
    EntropicConsts K{target_entropy, log_n_total, use_bounds, b_num, b_den, b_lr, b_logp1};
    entropic_dense_kernel<<<(unsigned)n_rows, kDenseThreads, 0, (cudaStream_t)stream>>>(C, n_rows, m, max_iter, K, logP,
                                                                                      eps, log_norm);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

extern "C" TDR_API int tdr_umap_affinity_f32(const float* C, int64_t n, int k, int max_iter, float* P,
                                     float* rho, float* sigma, tdr_stream_t stream) {
    TDR_CHECK_ARG(C && P && rho && sigma, "tdr_umap_affinity_f32: null pointer");
    TDR_CHECK_ARG(n >= 0 && k >= 1 && k <= TDR_MAX_K, "tdr_umap_affinity_f32: k=%d outside [1,%d]", k, TDR_MAX_K);
    TDR_CHECK_ARG(n < ((int64_t)1 << 31) * kRowsPerBlock, "tdr_umap_affinity_f32: n too large");
    if (n == 0) return TDR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch ((k + 31) / 32) {
        case 1: return launch_umap<1>(C, n, k, max_iter, P, rho, sigma, st);
        case 2: return launch_umap<2>(C, n, k, max_iter, P, rho, sigma, st);
        case 3: return launch_umap<3>(C, n, k, max_iter, P, rho, sigma, st);
        case 4: return launch_umap<4>(C, n, k, max_iter, P, rho, sigma, st);
        default: return launch_umap<5>(C, n, k, max_iter, P, rho, sigma, st);
    }
}

extern "C" TDR_API int tdr_entropic_affinity_f32(const float* C, int64_t n, int k, float target_entropy,
                                         float log_n_total, int use_bounds, float b_num, float b_den,
                                         float b_lr, float b_logp1, int max_iter, float* logP,
                                         float* eps, float* log_norm, tdr_stream_t stream) {
    TDR_CHECK_ARG(C && logP && eps && log_norm, "tdr_entropic_affinity_f32: null pointer");
    TDR_CHECK_ARG(n >= 0 && k >= 2 && k <= TDR_MAX_K, "tdr_entropic_affinity_f32: k=%d outside [2,%d]", k, TDR_MAX_K);
    if (n == 0) return TDR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    EntropicConsts K{target_entropy, log_n_total, use_bounds, b_num, b_den, b_lr, b_logp1};
    switch ((k + 31) / 32) {
        case 1: return launch_entropic<1>(C, n, k, max_iter, K, logP, eps, log_norm, st);
        case 2: return launch_entropic<2>(C, n, k, max_iter, K, logP, eps, log_norm, st);
        case 3: return launch_entropic<3>(C, n, k, max_iter, K, logP, eps, log_norm, st);
        case 4: return launch_entropic<4>(C, n, k, max_iter, K, logP, eps, log_norm, st);
        default: return launch_entropic<5>(C, n, k, max_iter, K, logP, eps, log_norm, st);
    }
}
