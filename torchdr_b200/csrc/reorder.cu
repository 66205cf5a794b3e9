// Nearest-centre assignment of the Voronoi-tree row ordering (torchdr_b200/reorder.py): every open row is handed to
// the nearest of the <= 16 centres of its own tree node.  No counterpart in the reference (it hands unordered rows to
// FAISS); the ordering only creates the index locality the pruned kNN sweep and the step kernel's gathers exploit,
// so its arithmetic does not enter any result.
// One warp per row: lanes take the row's features (coalesced, the row stays in registers), the node's centres are
// streamed once per row (8 KB per node, L1/L2-resident: the host hands the rows grouped by node), a warp reduction
// per centre gives |c|^2 - 2 x.c, ties go to the lower centre.
#include "common.cuh"

namespace tdr {

constexpr int kAssignWarps = 8;
constexpr int kAssignMaxPerLane = 16;  // d <= 512

__global__ void __launch_bounds__(kAssignWarps * 32)
tree_assign_kernel(const float* __restrict__ X, int d, const int64_t* __restrict__ rows, const int64_t* __restrict__ node,
                   int64_t m, const float* __restrict__ centres, const float* __restrict__ cnorm, int B,
                   int64_t* __restrict__ child) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * kAssignWarps + (threadIdx.x >> 5);
    if (i >= m) return;
    const float* x = X + rows[i] * d;
    const int64_t nd = node[i];
    float xv[kAssignMaxPerLane];
#pragma unroll
    for (int u = 0; u < kAssignMaxPerLane; ++u) {
        const int j = lane + 32 * u;
        xv[u] = j < d ? __ldg(x + j) : 0.0f;
    }
    float best = INFINITY;
    int best_c = 0;
    for (int c = 0; c < B; ++c) {
        const float cn = __ldg(cnorm + nd * B + c);
        if (cn == INFINITY) continue;  // warp-uniform: not a centre of this node
        const float* cv = centres + (nd * B + c) * d;
        float dot = 0.0f;
#pragma unroll
        for (int u = 0; u < kAssignMaxPerLane; ++u) {
            const int j = lane + 32 * u;
            if (j < d) dot = fmaf(xv[u], __ldg(cv + j), dot);
        }
        dot = warp_sum(dot);
        const float d2 = fmaf(-2.0f, dot, cn);
        if (d2 < best) {
            best = d2;
            best_c = c;
        }
    }
    if (lane == 0) child[i] = best_c;
}

// Per (node, child) sums and counts of the member rows (the Lloyd step of the tree).  The host hands the entries grouped
// by node, so a tile of 256 consecutive entries touches at most kAccNodes nodes (nodes being split hold > 128 rows): the
// CTA accumulates them in shared memory (lanes over features: no conflicts inside a warp) and flushes one global atomic
// per non-zero accumulator — torch.index_add_ would issue one global atomic per row and feature, all of them on the same
// 16 x d addresses at the root of the tree.
constexpr int kAccTile = 256;
constexpr int kAccNodes = 3;

__global__ void __launch_bounds__(256)
tree_accumulate_kernel(const float* __restrict__ X, int d, const int64_t* __restrict__ rows, const int64_t* __restrict__ node,
                       const int64_t* __restrict__ child, int64_t m, int B, float* __restrict__ sums,
                       float* __restrict__ cnt) {
    extern __shared__ float acc[];  // [kAccNodes * B][d] sums, then [kAccNodes * B] counts
    const int n_slots = kAccNodes * B;
    float* cnt_s = acc + (size_t)n_slots * d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * kAccTile;
    const int64_t t1 = min(t0 + kAccTile, m);
    for (int i = threadIdx.x; i < n_slots * d + n_slots; i += blockDim.x) acc[i] = 0.0f;
    __syncthreads();
    const int64_t node0 = node[t0];
    for (int64_t i = t0 + warp; i < t1; i += 8) {
        const int64_t rel = node[i] - node0;
        const int64_t c = child[i];
        const float* x = X + rows[i] * d;
        if (rel < kAccNodes) {
            float* a = acc + (size_t)(rel * B + c) * d;
            for (int j = lane; j < d; j += 32) atomicAdd(a + j, __ldg(x + j));
            if (lane == 0) atomicAdd(cnt_s + rel * B + c, 1.0f);
        } else {  // more nodes in the tile than planned for (tiny leaf size): straight to global memory
            float* a = sums + (size_t)(node[i] * B + c) * d;
            for (int j = lane; j < d; j += 32) atomicAdd(a + j, __ldg(x + j));
            if (lane == 0) atomicAdd(cnt + node[i] * B + c, 1.0f);
        }
    }
    __syncthreads();
    float* out = sums + (size_t)node0 * B * d;
    for (int i = threadIdx.x; i < n_slots * d; i += blockDim.x)
        if (acc[i] != 0.0f) atomicAdd(out + i, acc[i]);
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x)
        if (cnt_s[i] != 0.0f) atomicAdd(cnt + node0 * B + i, cnt_s[i]);
}

}  // namespace tdr

using namespace tdr;

extern "C" TDR_API int tdr_tree_accumulate_f32(const float* X, int d, const int64_t* rows, const int64_t* node,
                                               const int64_t* child, int64_t m, int64_t n_nodes, int B, float* sums,
                                               float* cnt, tdr_stream_t stream) {
    TDR_CHECK_ARG(X && rows && node && child && sums && cnt, "tdr_tree_accumulate_f32: null pointer");
    TDR_CHECK_ARG(d >= 1 && d <= 512 && B >= 1 && B <= 16 && m >= 0 && n_nodes >= 1, "tdr_tree_accumulate_f32: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    // the accumulators of the last node(s) of a tile may lie beyond n_nodes * B: the caller allocates kAccNodes - 1 spare nodes
    TDR_CUDA(cudaMemsetAsync(sums, 0, (size_t)(n_nodes + kAccNodes - 1) * B * d * sizeof(float), st));
    TDR_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(n_nodes + kAccNodes - 1) * B * sizeof(float), st));
    if (m == 0) return TDR_OK;
    const size_t smem = (size_t)kAccNodes * B * (d + 1) * sizeof(float);
    TDR_CUDA(cudaFuncSetAttribute(tree_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tree_accumulate_kernel<<<(unsigned)((m + kAccTile - 1) / kAccTile), 256, smem, st>>>(X, d, rows, node, child, m, B, sums, cnt);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

extern "C" TDR_API int tdr_tree_assign_f32(const float* X, int d, const int64_t* rows, const int64_t* node, int64_t m,
                                           const float* centres, const float* cnorm, int B, int64_t* child_out,
                                           tdr_stream_t stream) {
    TDR_CHECK_ARG(X && rows && node && centres && cnorm && child_out, "tdr_tree_assign_f32: null pointer");
    TDR_CHECK_ARG(d >= 1 && d <= 32 * kAssignMaxPerLane && B >= 1 && B <= 64 && m >= 0, "tdr_tree_assign_f32: bad shape");
    if (m == 0) return TDR_OK;
    tree_assign_kernel<<<(unsigned)((m + kAssignWarps - 1) / kAssignWarps), kAssignWarps * 32, 0, (cudaStream_t)stream>>>(
        X, d, rows, node, m, centres, cnorm, B, child_out);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}
