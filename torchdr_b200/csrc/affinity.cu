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

}  // namespace tdr

using namespace tdr;

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
