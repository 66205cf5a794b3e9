// pairwise_distances_indexed (torchdr/distance/base.py:252-405), 2-D key_indices case:
// out[i, s] = dist(X[query[i]], Y[key[i, s]]) in the exact-difference form of base.py:384-385
// (sum over features of (x - y)^2, sqrt for "euclidean").  One warp per query row, lanes over keys;
// negative key indices wrap like torch indexing (the -1 padding of the symmetrised affinity reads
// the last row, umap.py:236-264).  The optimisation kernels fuse this gather; the entry point exists
// for the seam itself.
#include "common.cuh"

namespace tdr {

template <typename KeyT>
__global__ void __launch_bounds__(256)
indexed_dist_kernel(const float* __restrict__ X, const int64_t* __restrict__ query, int64_t nq, const float* __restrict__ Y,
                    int64_t ny, int d, const KeyT* __restrict__ key, int k, int metric, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= nq) return;
    const float* x = X + (query ? query[i] : i) * d;
    for (int s = lane; s < k; s += 32) {
        int64_t j = (int64_t)key[i * k + s];
        if (j < 0) j += ny;
        const float* y = Y + j * d;
        float acc = 0.0f;
        for (int c = 0; c < d; ++c) {
            const float t = __fsub_rn(__ldg(x + c), __ldg(y + c));
            acc = __fadd_rn(acc, __fmul_rn(t, t));
        }
        out[i * k + s] = metric == TDR_METRIC_EUCLIDEAN ? sqrtf(acc) : acc;
    }
}

}  // namespace tdr

using namespace tdr;

extern "C" TDR_API int tdr_indexed_dist_f32(const float* X, const int64_t* query_idx, int64_t nq, const float* Y,
                                            int64_t ny, int d, const void* key_idx, int key_is_int64, int k, int metric,
                                            float* out, tdr_stream_t stream) {
    TDR_CHECK_ARG(X && Y && key_idx && out, "tdr_indexed_dist_f32: null pointer");
    TDR_CHECK_ARG(nq >= 0 && ny >= 1 && d >= 1 && k >= 1, "tdr_indexed_dist_f32: bad shape");
    TDR_CHECK_ARG(metric == TDR_METRIC_SQEUCLIDEAN || metric == TDR_METRIC_EUCLIDEAN,
                  "[TorchDR] ERROR : metric id %d is not supported.", metric);
    if (nq == 0) return TDR_OK;
    const unsigned blocks = (unsigned)((nq + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (key_is_int64)
        indexed_dist_kernel<int64_t><<<blocks, 256, 0, st>>>(X, query_idx, nq, Y, ny, d, (const int64_t*)key_idx, k, metric, out);
    else
        indexed_dist_kernel<int32_t><<<blocks, 256, 0, st>>>(X, query_idx, nq, Y, ny, d, (const int32_t*)key_idx, k, metric, out);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}
