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

}  // namespace tdr

using namespace tdr;

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
