// Closed-form gradients of the two autograd-mode methods and the momentum-SGD update.
//
// LargeVis  torchdr/neighbor_embedding/largevis.py:181-201
//   L = lam * sum_is -P_is log Q_is  +  rep * (1/N) sum_in -log(1 - Q_in),
//   Q = q/(q+1), q = 1/(1+D)  (so Q = 1/(2+D));  dL/dD_is = lam P_is Q_is,
//   dL/dD_in = -(rep/N) Q_in^2 / (1 - Q_in).  D = |z_i - z_j|^2 (distance/base.py:384-385),
//   so each edge adds 2 c (z_i - z_j) to row i and subtracts it from row j — the scatter that
//   autograd's index_put(accumulate) performs (affinity_matcher.py:418-425).
// t-SNE     torchdr/neighbor_embedding/tsne.py:162-180
//   attraction on the directed kNN edges: dL/dD = lam P/(1+D);
//   repulsion = log sum_{all i,j} 1/(1+C_ij) with C in the expanded form of
//   distance/torch.py:89-91, diagonal included:  grad_i = -(4/S) sum_j w_ij^2 (z_i - z_j).
// InfoTSNE  torchdr/neighbor_embedding/infotsne.py:179-197
//   attraction as t-SNE; repulsion = (1/N) sum_i log sum_{s in Neg(i)} 1/(1+D_is):
//   dL/dD_is = -(rep/N) w_is^2 / S_i,  w = 1/(1+D), S_i = sum_s w_is  (softmax of log Q times dlogQ/dD).
// SNE       torchdr/neighbor_embedding/sne.py:162-179
//   attraction sum P_ij D_ij (dL/dD = lam P); repulsion = (1/N) sum_i log sum_j exp(-C_ij), C the expanded
//   form on the embedding, diagonal included:  grad_i = -(2 rep/N) sum_j e_ij (1/S_i + 1/S_j)(z_i - z_j).
// SGD       torch.optim.SGD with momentum (NE base.py:331-343).
#include <algorithm>

#include "common.cuh"

namespace tdr {

constexpr int kAgWarps = 8;

__device__ __forceinline__ int64_t draw_negative(const Philox& rng, int64_t n_iter, int64_t gi, int s,
                                                 int64_t n_total) {
    const uint4 u = rng((uint32_t)n_iter, (uint32_t)(n_iter >> 32) ^ (uint32_t)(gi >> 32), (uint32_t)gi,
                        (uint32_t)(s >> 2));
    const uint32_t w = (s & 3) == 0 ? u.x : (s & 3) == 1 ? u.y : (s & 3) == 2 ? u.z : u.w;
    int64_t j = (int64_t)(((uint64_t)w * (uint64_t)(n_total - 1)) >> 32);
    return j + ((j >= gi) ? 1 : 0);  // NE base.py:636
}

__global__ void __launch_bounds__(kAgWarps * 32)
largevis_grad_kernel(const float2* __restrict__ Z, int64_t n_total, int64_t row0, int64_t n_local,
                     const float* __restrict__ P, const int32_t* __restrict__ idx, int k,
                     const int64_t* __restrict__ neg, int n_neg, uint64_t seed, int64_t n_iter, float lam,
                     float rep_over_n, float* __restrict__ grad) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kAgWarps + (threadIdx.x >> 5);
    if (r >= n_local) return;
    const int64_t gi = row0 + r;
    const float2 zi = __ldg(Z + gi);
    float gx = 0.0f, gy = 0.0f;
    for (int s = lane; s < k; s += 32) {
        const int64_t j = __ldg(idx + r * k + s);
        const float2 zj = __ldg(Z + j);
        const float dx = __fsub_rn(zi.x, zj.x), dy = __fsub_rn(zi.y, zj.y);
        const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const float q = __fdiv_rn(1.0f, __fadd_rn(1.0f, D));   // largevis.py:199
        const float Q = __fdiv_rn(q, __fadd_rn(q, 1.0f));      // largevis.py:200
        const float c = 2.0f * lam * __ldg(P + r * k + s) * Q;
        const float cx = c * dx, cy = c * dy;
        gx += cx;
        gy += cy;
        atomicAdd(reinterpret_cast<float2*>(grad) + j, make_float2(-cx, -cy));  // one 8-byte red.global.add.v2.f32
    }
    const Philox rng(seed);
    for (int s = lane; s < n_neg; s += 32) {
        const int64_t j = neg ? __ldg(neg + r * n_neg + s) : draw_negative(rng, n_iter, gi, s, n_total);
        const float2 zj = __ldg(Z + j);
        const float dx = __fsub_rn(zi.x, zj.x), dy = __fsub_rn(zi.y, zj.y);
        const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const float q = __fdiv_rn(1.0f, __fadd_rn(1.0f, D));   // largevis.py:188
        const float Q = __fdiv_rn(q, __fadd_rn(q, 1.0f));      // largevis.py:189
        const float c = -2.0f * rep_over_n * __fdiv_rn(Q * Q, __fsub_rn(1.0f, Q));
        const float cx = c * dx, cy = c * dy;
        gx += cx;
        gy += cy;
        atomicAdd(reinterpret_cast<float2*>(grad) + j, make_float2(-cx, -cy));  // one 8-byte red.global.add.v2.f32
    }
    gx = warp_sum(gx);
    gy = warp_sum(gy);
    if (lane == 0) {
        atomicAdd(reinterpret_cast<float2*>(grad) + gi, make_float2(gx, gy));
    }
}

// ---- LargeVis, row-local form (tdr_largevis_step_f32) ------------------------------------------------------------
// The scatter of the autograd graph is turned into gathers, so that a rank produces the COMPLETE gradient of its
// own rows and the N x 2 all-reduce of affinity_matcher.py:418-425 disappears:
//   * attraction: row i receives 2 lam P_ij Q_ij (z_i - z_j) from its own edge (i, j) and 2 lam P_ji Q_ji (z_i - z_j)
//     from every edge (j, i) that points at it; Q is symmetric in (i, j), so with S = P + P^T (the union graph,
//     tdr_symmetrize_csr_f32 mode SUM) both are one term  2 lam S_ij Q_ij (z_i - z_j)  of a row-local sum;
//   * repulsion: row i pulls from its own negatives; the equal and opposite force on a negative j is the only true
//     scatter left.  The negatives are a counter-based stream (Philox, counter = iteration / row / slot), so the
//     owner of j does not need to be told: largevis_push_kernel re-generates the negatives of ALL rows — 2 Philox
//     blocks per row — and keeps the pairs whose target is one of its own rows (local fp32 atomics, ~n_neg per row).
// largevis_pull_update_kernel then adds the row-local sums, applies torch.optim.SGD with momentum to the rank's
// rows (NE base.py:331-343) and writes them to Z_out — and to every NVLink peer's Z_out — Jacobi like the UMAP step.
__device__ __forceinline__ float rcp_newton(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

__global__ void __launch_bounds__(256)
largevis_push_kernel(const float2* __restrict__ Z, int64_t n_total, int64_t row0, int64_t n_local, int n_neg,
                     uint64_t seed, int64_t n_iter, float rep_over_n, float* __restrict__ grad_local) {
    const Philox rng(seed);
    for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < n_total; gi += (int64_t)gridDim.x * blockDim.x) {
        float2 zi;
        bool have = false;
        for (int s0 = 0; s0 < n_neg; s0 += 4) {
            const uint4 u = rng((uint32_t)n_iter, (uint32_t)(n_iter >> 32) ^ (uint32_t)(gi >> 32), (uint32_t)gi,
                                (uint32_t)(s0 >> 2));
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (s0 + q >= n_neg) break;
                int64_t j = (int64_t)(((uint64_t)w[q] * (uint64_t)(n_total - 1)) >> 32);
                j += (j >= gi) ? 1 : 0;  // NE base.py:636
                if (j < row0 || j >= row0 + n_local) continue;
                if (!have) {
                    zi = __ldg(Z + gi);
                    have = true;
                }
                const float2 zj = __ldg(Z + j);
                const float dx = __fsub_rn(zi.x, zj.x), dy = __fsub_rn(zi.y, zj.y);
                const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                // Q^2 / (1 - Q) = 1 / ((2 + D)(1 + D))   (largevis.py:188-190), as in the pull kernel
                const float c = -2.0f * rep_over_n * rcp_newton(__fmul_rn(__fadd_rn(2.0f, D), __fadd_rn(1.0f, D)));
                atomicAdd(reinterpret_cast<float2*>(grad_local) + (j - row0), make_float2(-c * dx, -c * dy));
            }
        }
    }
}

struct LvPeers {
    float2* out[8];
    int n;
};

__global__ void __launch_bounds__(kAgWarps * 32)
largevis_pull_update_kernel(const float2* __restrict__ Zin, float2* __restrict__ Zout, int64_t n_total, int64_t row0,
                            int64_t n_local, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                            const float* __restrict__ val, int n_neg, uint64_t seed, int64_t n_iter, float lam,
                            float rep_over_n, const float2* __restrict__ push, float2* __restrict__ mom, float neg_lr,
                            float mu, int first, double* __restrict__ gnorm_sq, int* __restrict__ nan_flag,
                            const LvPeers peers) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kAgWarps + (threadIdx.x >> 5);
    if (r >= n_local) return;
    const int64_t gi = row0 + r;
    const float2 zi = __ldg(Zin + gi);
    float gx = 0.0f, gy = 0.0f;
    const int64_t e0 = __ldg(rowptr + r), e1 = __ldg(rowptr + r + 1);
    // A row's work is a chain  (col, val) -> z_j gather -> arithmetic  of dependent round trips: the loads of kLvU
    // lane-strided chunks are issued together, then all their gathers, so that a warp keeps kLvU gathers per lane in
    // flight (a row of the union graph holds ~100-200 entries = 4-6 chunks).
    // Q = q/(q+1) with q = 1/(1+D) (largevis.py:199-200) is 1/(2+D); one MUFU.RCP + Newton step (<= 1 ulp) replaces the
    // two IEEE divisions — the gradient stays within the 1e-5 the tests hold it to against the reference's autograd.
    constexpr int kLvU = 4;
    for (int64_t base = e0; base < e1; base += 32 * kLvU) {
        int cj[kLvU];
        float sv[kLvU];
        float2 zj[kLvU];
#pragma unroll
        for (int u = 0; u < kLvU; ++u) {
            const int64_t e = base + u * 32 + lane;
            const bool ok = e < e1;
            cj[u] = ok ? __ldg(col + e) : (int)gi;
            sv[u] = ok ? __ldg(val + e) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < kLvU; ++u) zj[u] = __ldg(Zin + cj[u]);
#pragma unroll
        for (int u = 0; u < kLvU; ++u) {
            const float dx = __fsub_rn(zi.x, zj[u].x), dy = __fsub_rn(zi.y, zj[u].y);
            const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            const float Q = rcp_newton(__fadd_rn(2.0f, D));
            const float c = 2.0f * lam * sv[u] * Q;
            gx = fmaf(c, dx, gx);
            gy = fmaf(c, dy, gy);
        }
    }
    const Philox rng(seed);
    for (int s = lane; s < n_neg; s += 32) {
        const int64_t j = draw_negative(rng, n_iter, gi, s, n_total);
        const float2 zj = __ldg(Zin + j);
        const float dx = __fsub_rn(zi.x, zj.x), dy = __fsub_rn(zi.y, zj.y);
        const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        // Q^2 / (1 - Q) = 1 / ((2 + D)(1 + D))   (largevis.py:188-190)
        const float c = -2.0f * rep_over_n * rcp_newton(__fmul_rn(__fadd_rn(2.0f, D), __fadd_rn(1.0f, D)));
        gx = fmaf(c, dx, gx);
        gy = fmaf(c, dy, gy);
    }
    gx = warp_sum(gx);
    gy = warp_sum(gy);
    if (lane == 0) {
        const float2 ps = push[r];
        const float g0 = gx + ps.x, g1 = gy + ps.y;
        // torch.optim.sgd: buf = g (first step) | buf.mul_(mu).add_(g) ; param.add_(buf, alpha=-lr)
        float2 b;
        if (first) {
            b = make_float2(g0, g1);
        } else {
            const float2 old = mom[r];
            b = make_float2(__fadd_rn(__fmul_rn(old.x, mu), g0), __fadd_rn(__fmul_rn(old.y, mu), g1));
        }
        mom[r] = b;
        const float2 zo = make_float2(fmaf(neg_lr, b.x, zi.x), fmaf(neg_lr, b.y, zi.y));
        Zout[gi] = zo;
        for (int q = 0; q < peers.n; ++q) peers.out[q][gi] = zo;
        if (gnorm_sq) atomicAdd(gnorm_sq, (double)g0 * g0 + (double)g1 * g1);
        if (nan_flag && (zo.x != zo.x || zo.y != zo.y)) atomicExch(nan_flag, 1);
    }
}

__global__ void __launch_bounds__(kAgWarps * 32)
tsne_attract_kernel(const float2* __restrict__ Z, int64_t row0, int64_t n_local, const float* __restrict__ P,
                    const int32_t* __restrict__ idx, int k, float lam, float* __restrict__ grad) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kAgWarps + (threadIdx.x >> 5);
    if (r >= n_local) return;
    const int64_t gi = row0 + r;
    const float2 zi = __ldg(Z + gi);
    float gx = 0.0f, gy = 0.0f;
    for (int s = lane; s < k; s += 32) {
        const int64_t j = __ldg(idx + r * k + s);
        const float2 zj = __ldg(Z + j);
        const float dx = __fsub_rn(zi.x, zj.x), dy = __fsub_rn(zi.y, zj.y);
        const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const float c = 2.0f * lam * __fdiv_rn(__ldg(P + r * k + s), __fadd_rn(1.0f, D));  // tsne.py:169
        const float cx = c * dx, cy = c * dy;
        gx += cx;
        gy += cy;
        atomicAdd(reinterpret_cast<float2*>(grad) + j, make_float2(-cx, -cy));  // one 8-byte red.global.add.v2.f32
    }
    gx = warp_sum(gx);
    gy = warp_sum(gy);
    if (lane == 0) {
        atomicAdd(reinterpret_cast<float2*>(grad) + gi, make_float2(gx, gy));
    }
}

// Dense repulsion: thread = one local row i, the j range streamed through shared memory in
// tiles of TJ points (x, y, |z|^2).  blockIdx.y splits the j range so small N still fills the chip.
constexpr int TSNE_TI = 128, TSNE_TJ = 512;

__global__ void __launch_bounds__(TSNE_TI)
tsne_repulse_kernel(const float2* __restrict__ Z, int64_t n_total, int64_t row0, int64_t n_local,
                    int64_t j_per_split, float* __restrict__ U /*[n_local,2]*/, double* __restrict__ S) {
    __shared__ float4 tile[TSNE_TJ];
    const int64_t r = (int64_t)blockIdx.x * TSNE_TI + threadIdx.x;
    const bool live = r < n_local;
    const float2 zi = live ? __ldg(Z + row0 + r) : make_float2(0.f, 0.f);
    const float ni = fmaf(zi.x, zi.x, zi.y * zi.y);
    float ux = 0.0f, uy = 0.0f, s = 0.0f;
    const int64_t j_begin = (int64_t)blockIdx.y * j_per_split;
    const int64_t j_end = min(n_total, j_begin + j_per_split);
    for (int64_t j0 = j_begin; j0 < j_end; j0 += TSNE_TJ) {
        const int cnt = (int)min((int64_t)TSNE_TJ, j_end - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += TSNE_TI) {
            const float2 zj = __ldg(Z + j0 + t);
            tile[t] = make_float4(zj.x, zj.y, fmaf(zj.x, zj.x, zj.y * zj.y), 0.f);
        }
        __syncthreads();
#pragma unroll 8
        for (int t = 0; t < cnt; ++t) {
            const float4 zj = tile[t];
            // distance/torch.py:89-91 expanded form on the embedding (diagonal included)
            const float C = __fsub_rn(__fadd_rn(ni, zj.z), 2.0f * fmaf(zi.x, zj.x, zi.y * zj.y));
            const float w = __frcp_rn(__fadd_rn(1.0f, C));  // exp(-log(1+C)), tsne.py:176-177
            const float w2 = w * w;
            s += w;
            ux = fmaf(w2, zi.x - zj.x, ux);
            uy = fmaf(w2, zi.y - zj.y, uy);
        }
    }
    if (live) {
        atomicAdd(U + 2 * r, ux);
        atomicAdd(U + 2 * r + 1, uy);
    }
    // block reduction of the normaliser in fp64
    double sd = live ? (double)s : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
    __shared__ double wsum[TSNE_TI / 32];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = sd;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < TSNE_TI / 32; ++w) t += wsum[w];
        atomicAdd(S, t);
    }
}

// mode 0: c = 2 lam P/(1+D) (t-SNE, InfoTSNE);  mode 1: c = 2 lam P (SNE, sne.py:170)
template <int MODE>
__global__ void __launch_bounds__(kAgWarps * 32)
edge_attract_kernel(const float2* __restrict__ Z, int64_t row0, int64_t n_local, const float* __restrict__ P,
                    const int32_t* __restrict__ idx, int k, float lam, float* __restrict__ grad) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kAgWarps + (threadIdx.x >> 5);
    if (r >= n_local) return;
    const int64_t gi = row0 + r;
    const float2 zi = __ldg(Z + gi);
    float gx = 0.0f, gy = 0.0f;
    for (int s = lane; s < k; s += 32) {
        const int64_t j = __ldg(idx + r * k + s);
        const float2 zj = __ldg(Z + j);
        const float dx = __fsub_rn(zi.x, zj.x), dy = __fsub_rn(zi.y, zj.y);
        float c = 2.0f * lam * __ldg(P + r * k + s);
        if (MODE == 0) c = __fdiv_rn(c, __fadd_rn(1.0f, __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
        const float cx = c * dx, cy = c * dy;
        gx += cx;
        gy += cy;
        atomicAdd(reinterpret_cast<float2*>(grad) + j, make_float2(-cx, -cy));
    }
    gx = warp_sum(gx);
    gy = warp_sum(gy);
    if (lane == 0) atomicAdd(reinterpret_cast<float2*>(grad) + gi, make_float2(gx, gy));
}

// InfoTSNE repulsion: one warp per row; each lane keeps up to kInfoCache of its negatives (index, w, dx, dy) in
// registers between the normaliser pass and the gradient pass (n_neg <= 32 * kInfoCache = 320 covers the
// default 300); slots beyond that are recomputed.
constexpr int kInfoCache = 10;

__global__ void __launch_bounds__(kAgWarps * 32)
infotsne_repulse_kernel(const float2* __restrict__ Z, int64_t n_total, int64_t row0, int64_t n_local,
                        const int64_t* __restrict__ neg, int n_neg, uint64_t seed, int64_t n_iter,
                        float rep_over_n, float* __restrict__ grad) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kAgWarps + (threadIdx.x >> 5);
    if (r >= n_local) return;
    const int64_t gi = row0 + r;
    const float2 zi = __ldg(Z + gi);
    const Philox rng(seed);
    int64_t cj[kInfoCache];
    float cw[kInfoCache], cdx[kInfoCache], cdy[kInfoCache];
    float S = 0.0f;
#pragma unroll
    for (int u = 0; u < kInfoCache; ++u) {
        const int s = lane + 32 * u;
        cw[u] = 0.0f; cdx[u] = 0.0f; cdy[u] = 0.0f; cj[u] = gi;
        if (s < n_neg) {
            cj[u] = neg ? __ldg(neg + r * n_neg + s) : draw_negative(rng, n_iter, gi, s, n_total);
            const float2 zj = __ldg(Z + cj[u]);
            cdx[u] = __fsub_rn(zi.x, zj.x);
            cdy[u] = __fsub_rn(zi.y, zj.y);
            cw[u] = __fdiv_rn(1.0f, __fadd_rn(1.0f, __fadd_rn(__fmul_rn(cdx[u], cdx[u]), __fmul_rn(cdy[u], cdy[u]))));
            S += cw[u];
        }
    }
    for (int s = lane + 32 * kInfoCache; s < n_neg; s += 32) {
        const int64_t j = neg ? __ldg(neg + r * n_neg + s) : draw_negative(rng, n_iter, gi, s, n_total);
        const float2 zj = __ldg(Z + j);
        const float dx = __fsub_rn(zi.x, zj.x), dy = __fsub_rn(zi.y, zj.y);
        S += __fdiv_rn(1.0f, __fadd_rn(1.0f, __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
    }
    S = warp_sum(S);
    const float scale = -2.0f * rep_over_n / S;  // infotsne.py:197: logsumexp over the row, / n_samples_in_
    float gx = 0.0f, gy = 0.0f;
#pragma unroll
    for (int u = 0; u < kInfoCache; ++u) {
        if (lane + 32 * u < n_neg) {
            const float c = scale * cw[u] * cw[u];
            const float cx = c * cdx[u], cy = c * cdy[u];
            gx += cx;
            gy += cy;
            atomicAdd(reinterpret_cast<float2*>(grad) + cj[u], make_float2(-cx, -cy));
        }
    }
    for (int s = lane + 32 * kInfoCache; s < n_neg; s += 32) {
        const int64_t j = neg ? __ldg(neg + r * n_neg + s) : draw_negative(rng, n_iter, gi, s, n_total);
        const float2 zj = __ldg(Z + j);
        const float dx = __fsub_rn(zi.x, zj.x), dy = __fsub_rn(zi.y, zj.y);
        const float w = __fdiv_rn(1.0f, __fadd_rn(1.0f, __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
        const float c = scale * w * w;
        const float cx = c * dx, cy = c * dy;
        gx += cx;
        gy += cy;
        atomicAdd(reinterpret_cast<float2*>(grad) + j, make_float2(-cx, -cy));
    }
    gx = warp_sum(gx);
    gy = warp_sum(gy);
    if (lane == 0) atomicAdd(reinterpret_cast<float2*>(grad) + gi, make_float2(gx, gy));
}

// SNE dense repulsion, same tiling as tsne_repulse_kernel.  PHASE 0: row normalisers S_i = sum_j exp(-C_ij)
// (atomically accumulated over the j splits);  PHASE 1: u_i = sum_j e_ij (1/S_i + 1/S_j)(z_i - z_j).
template <int PHASE>
__global__ void __launch_bounds__(TSNE_TI)
sne_repulse_kernel(const float2* __restrict__ Z, int64_t n_total, int64_t row0, int64_t n_local,
                   int64_t j_per_split, float* __restrict__ S /*[n_total]*/, float scale,
                   float* __restrict__ grad) {
    __shared__ float4 tile[TSNE_TJ];
    const int64_t r = (int64_t)blockIdx.x * TSNE_TI + threadIdx.x;
    const bool live = r < n_local;
    const float2 zi = live ? __ldg(Z + row0 + r) : make_float2(0.f, 0.f);
    const float ni = fmaf(zi.x, zi.x, zi.y * zi.y);
    const float inv_si = (PHASE == 1 && live) ? __frcp_rn(S[row0 + r]) : 0.0f;
    float ux = 0.0f, uy = 0.0f, s = 0.0f;
    const int64_t j_begin = (int64_t)blockIdx.y * j_per_split;
    const int64_t j_end = min(n_total, j_begin + j_per_split);
    for (int64_t j0 = j_begin; j0 < j_end; j0 += TSNE_TJ) {
        const int cnt = (int)min((int64_t)TSNE_TJ, j_end - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += TSNE_TI) {
            const float2 zj = __ldg(Z + j0 + t);
            tile[t] = make_float4(zj.x, zj.y, fmaf(zj.x, zj.x, zj.y * zj.y), PHASE == 1 ? __frcp_rn(S[j0 + t]) : 0.f);
        }
        __syncthreads();
#pragma unroll 8
        for (int t = 0; t < cnt; ++t) {
            const float4 zj = tile[t];
            const float C = __fsub_rn(__fadd_rn(ni, zj.z), 2.0f * fmaf(zi.x, zj.x, zi.y * zj.y));  // distance/torch.py:89-91
            const float e = __expf(-C);
            if (PHASE == 0) {
                s += e;
            } else {
                const float w = e * (inv_si + zj.w);
                ux = fmaf(w, zi.x - zj.x, ux);
                uy = fmaf(w, zi.y - zj.y, uy);
            }
        }
    }
    if (!live) return;
    if (PHASE == 0) {
        atomicAdd(S + row0 + r, s);
    } else {
        atomicAdd(grad + 2 * (row0 + r), scale * ux);
        atomicAdd(grad + 2 * (row0 + r) + 1, scale * uy);
    }
}

__global__ void __launch_bounds__(256)
tsne_finish_kernel(const float* __restrict__ U, const double* __restrict__ S, int64_t row0, int64_t n_local,
                   float repulsion, float* __restrict__ grad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n_local) return;
    const float scale = (float)(-4.0 * (double)repulsion / *S);  // NE base.py:237-241: repulsion_strength * repulsive
    atomicAdd(grad + 2 * row0 + i, scale * U[i]);
}

__global__ void __launch_bounds__(256)
sgd_momentum_kernel(float* __restrict__ Z, float* __restrict__ buf, const float* __restrict__ grad, int64_t n,
                    float neg_lr, float mu, int first, double* __restrict__ gnorm_sq, int* __restrict__ nan_flag) {
    double gn = 0.0;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = grad[i];
        // torch.optim.sgd: buf = g (first step) | buf.mul_(mu).add_(g) ; param.add_(buf, alpha=-lr)
        const float b = first ? g : __fadd_rn(__fmul_rn(buf[i], mu), g);
        buf[i] = b;
        const float z = fmaf(neg_lr, b, Z[i]);
        Z[i] = z;
        gn += (double)g * g;
        bad |= (z != z);
    }
    if (gnorm_sq) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gn += __shfl_xor_sync(0xffffffffu, gn, o);
        if ((threadIdx.x & 31) == 0 && gn != 0.0) atomicAdd(gnorm_sq, gn);
    }
    if (nan_flag && bad) atomicExch(nan_flag, 1);
}

}  // namespace tdr

using namespace tdr;

extern "C" TDR_API int tdr_largevis_grad_f32(const float* Z, int64_t n_total, int64_t row0, int64_t n_local, const float* P,
                                     const int32_t* idx, int k, const int64_t* neg, int n_neg, uint64_t seed,
                                     int64_t n_iter, float lam, float repulsion, float* grad, tdr_stream_t stream) {
    TDR_CHECK_ARG(Z && P && idx && grad, "tdr_largevis_grad_f32: null pointer");
    TDR_CHECK_ARG(n_total >= 2 && row0 >= 0 && n_local >= 0 && row0 + n_local <= n_total && k >= 1 && n_neg >= 0,
                  "tdr_largevis_grad_f32: bad shape");
    if (n_local == 0) return TDR_OK;
    const unsigned blocks = (unsigned)((n_local + kAgWarps - 1) / kAgWarps);
    largevis_grad_kernel<<<blocks, kAgWarps * 32, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(Z), n_total, row0, n_local, P, idx, k, neg, n_neg, seed, n_iter, lam,
        repulsion / (float)n_total, grad);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

extern "C" TDR_API int tdr_largevis_step_f32(const float* Z_in, float* Z_out, int64_t n_total, int64_t row0, int64_t n_local,
                                             const int64_t* rowptr, const int32_t* col, const float* val, int n_neg,
                                             uint64_t seed, int64_t n_iter, float lam, float repulsion,
                                             float* grad_scratch, float* mom, float lr, float momentum, int first,
                                             double* gnorm_sq, int* nan_flag, const uint64_t* peer_out_ptrs /*host*/,
                                             int n_peers, tdr_stream_t stream) {
    TDR_CHECK_ARG(Z_in && Z_out && Z_in != Z_out && rowptr && col && val && grad_scratch && mom,
                  "tdr_largevis_step_f32: null pointer / aliased buffers");
    TDR_CHECK_ARG(n_total >= 2 && row0 >= 0 && n_local >= 0 && row0 + n_local <= n_total && n_neg >= 0,
                  "tdr_largevis_step_f32: bad shape");
    TDR_CHECK_ARG(n_peers >= 0 && n_peers <= 8 && (n_peers == 0 || peer_out_ptrs), "tdr_largevis_step_f32: at most 8 peers");
    if (n_local == 0) return TDR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const float rep_over_n = repulsion / (float)n_total;
    TDR_CUDA(cudaMemsetAsync(grad_scratch, 0, (size_t)n_local * 8, st));
    if (n_neg > 0) {
        const unsigned pgrid = (unsigned)std::min<int64_t>((int64_t)kNumSMs * 16, (n_total + 255) / 256);
        largevis_push_kernel<<<pgrid, 256, 0, st>>>(reinterpret_cast<const float2*>(Z_in), n_total, row0, n_local, n_neg,
                                                    seed, n_iter, rep_over_n, grad_scratch);
    }
    LvPeers peers{};
    peers.n = n_peers;
    for (int q = 0; q < n_peers; ++q) {
        peers.out[q] = reinterpret_cast<float2*>(peer_out_ptrs[q]);
        TDR_CHECK_ARG(peers.out[q] && (const float*)peers.out[q] != Z_in, "tdr_largevis_step_f32: bad peer buffer");
    }
    const unsigned blocks = (unsigned)((n_local + kAgWarps - 1) / kAgWarps);
    largevis_pull_update_kernel<<<blocks, kAgWarps * 32, 0, st>>>(
        reinterpret_cast<const float2*>(Z_in), reinterpret_cast<float2*>(Z_out), n_total, row0, n_local, rowptr, col, val,
        n_neg, seed, n_iter, lam, rep_over_n, reinterpret_cast<const float2*>(grad_scratch),
        reinterpret_cast<float2*>(mom), -lr, momentum, first, gnorm_sq, nan_flag, peers);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

extern "C" TDR_API size_t tdr_tsne_workspace_bytes(int64_t n_local) { return 256 + (size_t)n_local * 8; }

extern "C" TDR_API int tdr_tsne_grad_f32(const float* Z, int64_t n_total, int64_t row0, int64_t n_local, const float* P,
                                 const int32_t* idx, int k, float lam, float repulsion, int phase, float* grad,
                                 void* ws, size_t ws_bytes, tdr_stream_t stream) {
    TDR_CHECK_ARG(Z && grad && ws, "tdr_tsne_grad_f32: null pointer");
    TDR_CHECK_ARG(n_total >= 2 && row0 >= 0 && n_local >= 0 && row0 + n_local <= n_total, "tdr_tsne_grad_f32: bad shape");
    TDR_CHECK_ARG(ws_bytes >= tdr_tsne_workspace_bytes(n_local) && (uintptr_t)ws % 16 == 0,
                  "tdr_tsne_grad_f32: workspace too small or misaligned");
    cudaStream_t st = (cudaStream_t)stream;
    double* S = reinterpret_cast<double*>(ws);
    float* U = reinterpret_cast<float*>((char*)ws + 256);
    const float2* Z2 = reinterpret_cast<const float2*>(Z);
    if (phase == 0) {
        // partial normaliser + unnormalised repulsion of the local rows, and the sparse attraction
        TDR_CUDA(cudaMemsetAsync(ws, 0, tdr_tsne_workspace_bytes(n_local), st));
        if (n_local > 0) {
            const unsigned bx = (unsigned)((n_local + TSNE_TI - 1) / TSNE_TI);
            unsigned by = 1;
            while ((int64_t)bx * by < 2 * kNumSMs && (int64_t)by * TSNE_TJ * 2 <= n_total) by *= 2;
            const int64_t per = (n_total + by - 1) / by;
            const int64_t per_al = (per + TSNE_TJ - 1) / TSNE_TJ * TSNE_TJ;
            tsne_repulse_kernel<<<dim3(bx, by), TSNE_TI, 0, st>>>(Z2, n_total, row0, n_local, per_al, U, S);
            if (P && idx && k > 0) {
                const unsigned blocks = (unsigned)((n_local + kAgWarps - 1) / kAgWarps);
                tsne_attract_kernel<<<blocks, kAgWarps * 32, 0, st>>>(Z2, row0, n_local, P, idx, k, lam, grad);
            }
        }
    } else {
        // S now holds the global normaliser (all-reduced by the host when distributed)
        if (n_local > 0)
            tsne_finish_kernel<<<(unsigned)((2 * n_local + 255) / 256), 256, 0, st>>>(U, S, row0, n_local, repulsion, grad);
    }
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

extern "C" TDR_API int tdr_infotsne_grad_f32(const float* Z, int64_t n_total, int64_t row0, int64_t n_local, const float* P,
                                     const int32_t* idx, int k, const int64_t* neg, int n_neg, uint64_t seed,
                                     int64_t n_iter, float lam, float repulsion, float* grad, tdr_stream_t stream) {
    TDR_CHECK_ARG(Z && P && idx && grad, "tdr_infotsne_grad_f32: null pointer");
    TDR_CHECK_ARG(n_total >= 2 && row0 >= 0 && n_local >= 0 && row0 + n_local <= n_total && k >= 1 && n_neg >= 1,
                  "tdr_infotsne_grad_f32: bad shape");
    if (n_local == 0) return TDR_OK;
    const unsigned blocks = (unsigned)((n_local + kAgWarps - 1) / kAgWarps);
    const float2* Z2 = reinterpret_cast<const float2*>(Z);
    edge_attract_kernel<0><<<blocks, kAgWarps * 32, 0, (cudaStream_t)stream>>>(Z2, row0, n_local, P, idx, k, lam, grad);
    infotsne_repulse_kernel<<<blocks, kAgWarps * 32, 0, (cudaStream_t)stream>>>(
        Z2, n_total, row0, n_local, neg, n_neg, seed, n_iter, repulsion / (float)n_total, grad);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

extern "C" TDR_API int tdr_sne_grad_f32(const float* Z, int64_t n_total, int64_t row0, int64_t n_local, const float* P,
                                const int32_t* idx, int k, float lam, float repulsion, int phase, float* grad,
                                float* row_sums, tdr_stream_t stream) {
    TDR_CHECK_ARG(Z && grad && row_sums, "tdr_sne_grad_f32: null pointer");
    TDR_CHECK_ARG(n_total >= 2 && row0 >= 0 && n_local >= 0 && row0 + n_local <= n_total, "tdr_sne_grad_f32: bad shape");
    TDR_CHECK_ARG(phase == 0 || phase == 1, "tdr_sne_grad_f32: phase must be 0 or 1");
    cudaStream_t st = (cudaStream_t)stream;
    const float2* Z2 = reinterpret_cast<const float2*>(Z);
    if (phase == 0) TDR_CUDA(cudaMemsetAsync(row_sums + row0, 0, (size_t)n_local * sizeof(float), st));
    if (n_local == 0) return TDR_OK;
    const unsigned bx = (unsigned)((n_local + TSNE_TI - 1) / TSNE_TI);
    unsigned by = 1;
    while ((int64_t)bx * by < 2 * kNumSMs && (int64_t)by * TSNE_TJ * 2 <= n_total) by *= 2;
    const int64_t per = (n_total + by - 1) / by;
    const int64_t per_al = (per + TSNE_TJ - 1) / TSNE_TJ * TSNE_TJ;
    if (phase == 0) {
        sne_repulse_kernel<0><<<dim3(bx, by), TSNE_TI, 0, st>>>(Z2, n_total, row0, n_local, per_al, row_sums, 0.0f, grad);
        if (P && idx && k > 0) {
            const unsigned blocks = (unsigned)((n_local + kAgWarps - 1) / kAgWarps);
            edge_attract_kernel<1><<<blocks, kAgWarps * 32, 0, st>>>(Z2, row0, n_local, P, idx, k, lam, grad);
        }
    } else {
        // row_sums now holds every row's normaliser (all-gathered by the host when distributed)
        sne_repulse_kernel<1><<<dim3(bx, by), TSNE_TI, 0, st>>>(Z2, n_total, row0, n_local, per_al, row_sums,
                                                               -2.0f * repulsion / (float)n_total, grad);
    }
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

extern "C" TDR_API int tdr_sgd_momentum_f32(float* Z, float* buf, const float* grad, int64_t n_elems, float lr,
                                    float momentum, int first, double* gnorm_sq, int* nan_flag,
                                    tdr_stream_t stream) {
    TDR_CHECK_ARG(Z && buf && grad && n_elems >= 0, "tdr_sgd_momentum_f32: bad arguments");
    if (n_elems == 0) return TDR_OK;
    const unsigned grid = (unsigned)min((int64_t)kNumSMs * 8, (n_elems + 255) / 256);
    sgd_momentum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Z, buf, grad, n_elems, -lr, momentum, first,
                                                               gnorm_sq, nan_flag);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}
