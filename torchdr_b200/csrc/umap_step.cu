// (iii) UMAP optimisation step: sampled attraction + repulsion + clamp + SGD in ONE kernel.
//
// Replaces, per iteration, the ~40 ATen launches of torchdr/neighbor_embedding/umap.py:236-292
// + neighbor_embedding/base.py:617-649 (negative sampling) + affinity_matcher.py:427 (SGD):
//   * graph as CSR (live edges only) instead of the 75 %-padded ELL;
//   * epoch_of_next_sample read-modify-written in place, only for due edges;
//   * negatives drawn in-kernel (Philox4x32-10) in throughput mode, or read from the
//     caller's table (the reference's neg_indices_) in parity mode;
//   * Jacobi update: every gradient is computed from Z_in, results go to Z_out.
// One warp per row: lanes stride over the row's CSR segment, then over the 5*active
// negatives; 8-byte gathers of z_j hit L2 (Z is 8 MB at 1 M points, 80 MB at 10 M).
//
// Arithmetic mirrors the reference op by op (each torch op is one rounding): explicit
// __f*_rn intrinsics keep the compiler from contracting them into FMAs.
#include <stdlib.h>

#include "common.cuh"

namespace tdr {

struct UmapStepParams {
    const float2* Zin;
    float2* Zout;
    int64_t n_total, row0, n_local;
    const int64_t* rowptr;
    const int32_t* col;
    const float* eps;
    float* eons;
    const int64_t* neg;
    int n_neg, rate;
    uint64_t seed;
    int64_t n_iter;
    float a, b, bm1, two_ab, neg_two_b, lam, rep, lr;
    float2* grad_out;
    double* gnorm_sq;
    int* nan_flag;
    unsigned long long* stats;  // [0] += sampled edges, [1] += negatives used (roofline accounting)
    // fused exchange (multi-GPU): the updated row is also stored into the Z_out buffer of every peer through
    // NVLink peer mappings, so no separate all-gather of the embedding is needed after the step
    float2* peer_out[8];
    int n_peers;
};

__device__ __forceinline__ void store_row(const UmapStepParams& p, int64_t gi, float2 zo) {
    p.Zout[gi] = zo;
    if (p.n_peers == 0) return;  // uniform: single-GPU launches skip the peer loop
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (q < p.n_peers) p.peer_out[q][gi] = zo;
}

// Per-block reduction of the optional diagnostics (gradient norm, NaN flag, sampled-edge counters) in
// shared memory, then one global atomic per block: thousands of same-address global atomics per launch
// would otherwise serialise in L2.
__device__ __forceinline__ void block_flush(bool leader, double gn, bool saw_nan, unsigned long long n_act,
                                            unsigned long long n_neg, const UmapStepParams& p) {
    if (p.nan_flag && leader && saw_nan) atomicExch(p.nan_flag, 1);  // rare: no reduction needed
    if (!p.gnorm_sq && !p.stats) return;  // uniform: the common per-iteration case has no block barrier
    __shared__ double s_gn;
    __shared__ unsigned long long s_cnt[2];
    __shared__ int s_nan;
    if (threadIdx.x == 0) {
        s_gn = 0.0;
        s_cnt[0] = s_cnt[1] = 0ull;
        s_nan = 0;
    }
    __syncthreads();
    if (leader) {
        if (gn != 0.0) atomicAdd(&s_gn, gn);
        if (n_act) atomicAdd(&s_cnt[0], n_act);
        if (n_neg) atomicAdd(&s_cnt[1], n_neg);
        if (saw_nan) s_nan = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (p.gnorm_sq && s_gn != 0.0) atomicAdd(p.gnorm_sq, s_gn);
        if (p.nan_flag && s_nan) atomicExch(p.nan_flag, 1);
        if (p.stats) {
            if (s_cnt[0]) atomicAdd(p.stats, s_cnt[0]);
            if (s_cnt[1]) atomicAdd(p.stats + 1, s_cnt[1]);
        }
    }
}

template <bool PRECISE>
__device__ __forceinline__ float pow_b(float x, float y) {
    // torch pow(tensor, scalar) evaluates powf in fp32 (Sleef, 1 ulp); parity mode goes through
    // fp64 so the result is the correctly rounded fp32 value in all but ~1e-9 of cases.
    if (PRECISE) return (float)pow((double)x, (double)y);
    return powf(x, y);
}

constexpr int kStepWarps = 8;

template <bool PRECISE>
__global__ void __launch_bounds__(kStepWarps * 32) umap_step_kernel(const UmapStepParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * kStepWarps + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * kStepWarps;
    const float due_before = (float)(p.n_iter + 1);  // umap.py:251 (long promoted to fp32)
    double gn_local = 0.0;
    bool saw_nan = false;
    unsigned long long n_act = 0, n_neg_used = 0;

    for (int64_t r = warp_global; r < p.n_local; r += n_warps) {
        const int64_t gi = p.row0 + r;
        const float2 zi = __ldg(p.Zin + gi);
        // ---- attraction (umap.py:236-264) over the row's live edges
        const int64_t e0 = p.rowptr[r], e1 = p.rowptr[r + 1];
        float gx = 0.0f, gy = 0.0f;
        int active = 0;
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const float nxt = p.eons[e];
            if (nxt <= due_before) {
                p.eons[e] = __fadd_rn(nxt, __ldg(p.eps + e));  // umap.py:253-255
                ++active;
                const float2 zj = __ldg(p.Zin + __ldg(p.col + e));
                const float dx = __fsub_rn(zi.x, zj.x), dy = __fsub_rn(zi.y, zj.y);
                const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));  // distance/base.py:384-385
                if (D > 0.0f) {  // umap.py:243,247
                    const float den = __fadd_rn(1.0f, __fmul_rn(p.a, pow_b<PRECISE>(D, p.b)));
                    const float coef = __fdiv_rn(__fmul_rn(pow_b<PRECISE>(D, p.bm1), p.two_ab), den);
                    gx = fmaf(dx, coef, gx);
                    gy = fmaf(dy, coef, gy);
                }
            }
        }
        gx = warp_sum(gx);
        gy = warp_sum(gy);
        active = warp_sum_int(active);
        gx = fminf(fmaxf(gx, -4.0f), 4.0f);  // umap.py:263
        gy = fminf(fmaxf(gy, -4.0f), 4.0f);

        // ---- repulsion (umap.py:266-292) on the first rate*active negatives
        int quota = active * p.rate;
        if (quota > p.n_neg) quota = p.n_neg;
        n_act += active;
        n_neg_used += quota;
        float rx = 0.0f, ry = 0.0f;
        const Philox rng(p.seed);
        for (int s = lane; s < quota; s += 32) {
            int64_t j;
            if (p.neg) {
                j = __ldg(p.neg + r * p.n_neg + s);
            } else {
                // one Philox block per 4 consecutive slots: counter = (n_iter, row, slot/4)
                const uint4 u = rng((uint32_t)p.n_iter, (uint32_t)(p.n_iter >> 32) ^ (uint32_t)(gi >> 32),
                                    (uint32_t)gi, (uint32_t)(s >> 2));
                const uint32_t w = (s & 3) == 0 ? u.x : (s & 3) == 1 ? u.y : (s & 3) == 2 ? u.z : u.w;
                j = (int64_t)(((uint64_t)w * (uint64_t)(p.n_total - 1)) >> 32);  // uniform on [0, N-2]
                j += (j >= gi) ? 1 : 0;                                         // NE base.py:636
            }
            const float2 zj = __ldg(p.Zin + j);
            const float dx = __fsub_rn(zi.x, zj.x), dy = __fsub_rn(zi.y, zj.y);
            const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            const float den = __fadd_rn(1.0f, __fmul_rn(p.a, pow_b<PRECISE>(D, p.b)));   // umap.py:273
            const float coef = __fmul_rn(__frcp_rn(__fmul_rn(__fadd_rn(D, 1e-3f), den)), p.neg_two_b);  // :274-276
            rx = fmaf(dx, coef, rx);
            ry = fmaf(dy, coef, ry);
        }
        rx = warp_sum(rx);
        ry = warp_sum(ry);
        rx = fminf(fmaxf(rx, -4.0f), 4.0f);  // umap.py:291
        ry = fminf(fmaxf(ry, -4.0f), 4.0f);

        if (lane == 0) {
            // NE base.py:237-241: lam * attractive + repulsion_strength * repulsive
            const float g0 = __fadd_rn(__fmul_rn(p.lam, gx), __fmul_rn(p.rep, rx));
            const float g1 = __fadd_rn(__fmul_rn(p.lam, gy), __fmul_rn(p.rep, ry));
            float2 zo;  // torch.optim.SGD: param.add_(grad, alpha=-lr)
            zo.x = fmaf(-p.lr, g0, zi.x);
            zo.y = fmaf(-p.lr, g1, zi.y);
            store_row(p, gi, zo);
            if (p.grad_out) p.grad_out[r] = make_float2(g0, g1);
            gn_local += (double)g0 * g0 + (double)g1 * g1;
            saw_nan |= (zo.x != zo.x) || (zo.y != zo.y);
        }
    }
    block_flush(lane == 0, gn_local, saw_nan, n_act, n_neg_used, p);
}

// ---------------------------------------------------------------------------------------------
// Throughput kernel (fp32 powf): the step is bound by memory latency, not bandwidth — a row's work
// is a chain  rowptr -> (eons, col, eps) -> z_j gathers  of dependent round trips — so the kernel is
// organised for memory-level parallelism: G = 8 lanes per row (4 rows per warp), the three edge
// arrays are loaded together and unconditionally for up to 4 lane-strided chunks, the gathers of all
// due edges are issued back to back, and each lane fetches 4 negatives per Philox call before any of
// them is consumed.  Same arithmetic as umap_step_kernel<false>; only the order of the fp32 partial
// sums differs.
constexpr int G = 8;           // lanes per row
constexpr int U = 4;           // edge chunks / negatives in flight per lane
constexpr int kV2Threads = 256;

__device__ __forceinline__ float group_sum(float v, unsigned mask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
__device__ __forceinline__ int group_sum_int(int v, unsigned mask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

__global__ void __launch_bounds__(kV2Threads, 4) umap_step_kernel_v2(const UmapStepParams p) {
    const int lane = threadIdx.x & 31;
    const int l = lane & (G - 1);
    const unsigned gmask = ((1u << G) - 1u) << (lane & ~(G - 1));
    const int64_t group_global = ((int64_t)blockIdx.x * kV2Threads + threadIdx.x) / G;
    const int64_t n_groups = (int64_t)gridDim.x * kV2Threads / G;
    const float due_before = (float)(p.n_iter + 1);
    const Philox rng(p.seed);
    double gn_local = 0.0;
    bool saw_nan = false;
    unsigned long long n_act = 0, n_neg_used = 0;

    for (int64_t r = group_global; r < p.n_local; r += n_groups) {
        const int64_t gi = p.row0 + r;
        const float2 zi = __ldg(p.Zin + gi);
        const int64_t e0 = __ldg(p.rowptr + r), e1 = __ldg(p.rowptr + r + 1);
        float gx = 0.0f, gy = 0.0f;
        int active = 0;
        for (int64_t base = e0; base < e1; base += G * U) {
            float nxt[U], ep[U];
            int cj[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t e = base + u * G + l;
                const bool ok = e < e1;
                nxt[u] = ok ? p.eons[e] : INFINITY;
                cj[u] = ok ? __ldg(p.col + e) : 0;
                ep[u] = ok ? __ldg(p.eps + e) : 0.0f;
            }
            float2 zj[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (nxt[u] <= due_before) zj[u] = __ldg(p.Zin + cj[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (nxt[u] <= due_before) {
                    p.eons[base + u * G + l] = __fadd_rn(nxt[u], ep[u]);  // umap.py:253-255
                    ++active;
                    const float dx = __fsub_rn(zi.x, zj[u].x), dy = __fsub_rn(zi.y, zj[u].y);
                    const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                    if (D > 0.0f) {  // umap.py:243-247
                        const float den = __fadd_rn(1.0f, __fmul_rn(p.a, powf(D, p.b)));
                        const float coef = __fdiv_rn(__fmul_rn(powf(D, p.bm1), p.two_ab), den);
                        gx = fmaf(dx, coef, gx);
                        gy = fmaf(dy, coef, gy);
                    }
                }
            }
        }
        gx = group_sum(gx, gmask);
        gy = group_sum(gy, gmask);
        active = group_sum_int(active, gmask);
        gx = fminf(fmaxf(gx, -4.0f), 4.0f);  // umap.py:263
        gy = fminf(fmaxf(gy, -4.0f), 4.0f);

        int quota = active * p.rate;  // umap.py:279-284
        if (quota > p.n_neg) quota = p.n_neg;
        float rx = 0.0f, ry = 0.0f;
        for (int s0 = 0; s0 < quota; s0 += G * U) {
            // lane l owns slots s0 + 4 l .. s0 + 4 l + 3 (one Philox block)
            const int sb = s0 + U * l;
            int64_t j[U];
            if (p.neg) {
#pragma unroll
                for (int u = 0; u < U; ++u) j[u] = (sb + u < quota) ? __ldg(p.neg + r * p.n_neg + sb + u) : gi;
            } else {
                const uint4 w = rng((uint32_t)p.n_iter, (uint32_t)(p.n_iter >> 32) ^ (uint32_t)(gi >> 32), (uint32_t)gi,
                                    (uint32_t)(sb >> 2));
                const uint32_t wv[U] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    int64_t t = (int64_t)(((uint64_t)wv[u] * (uint64_t)(p.n_total - 1)) >> 32);
                    j[u] = t + ((t >= gi) ? 1 : 0);  // NE base.py:636
                }
            }
            float2 zn[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (sb + u < quota) zn[u] = __ldg(p.Zin + j[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (sb + u < quota) {
                    const float dx = __fsub_rn(zi.x, zn[u].x), dy = __fsub_rn(zi.y, zn[u].y);
                    const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                    const float den = __fadd_rn(1.0f, __fmul_rn(p.a, powf(D, p.b)));  // umap.py:273
                    const float coef = __fmul_rn(__frcp_rn(__fmul_rn(__fadd_rn(D, 1e-3f), den)), p.neg_two_b);
                    rx = fmaf(dx, coef, rx);
                    ry = fmaf(dy, coef, ry);
                }
            }
        }
        rx = group_sum(rx, gmask);
        ry = group_sum(ry, gmask);
        rx = fminf(fmaxf(rx, -4.0f), 4.0f);  // umap.py:291
        ry = fminf(fmaxf(ry, -4.0f), 4.0f);
        if (l == 0) {
            const float g0 = __fadd_rn(__fmul_rn(p.lam, gx), __fmul_rn(p.rep, rx));
            const float g1 = __fadd_rn(__fmul_rn(p.lam, gy), __fmul_rn(p.rep, ry));
            float2 zo;
            zo.x = fmaf(-p.lr, g0, zi.x);
            zo.y = fmaf(-p.lr, g1, zi.y);
            store_row(p, gi, zo);
            if (p.grad_out) p.grad_out[r] = make_float2(g0, g1);
            gn_local += (double)g0 * g0 + (double)g1 * g1;
            saw_nan |= (zo.x != zo.x) || (zo.y != zo.y);
            n_act += active;
            n_neg_used += quota;
        }
    }
    block_flush(l == 0, gn_local, saw_nan, n_act, n_neg_used, p);
}

}  // namespace tdr
#include "umap_step_math.cuh"
#include "umap_step_fast3.cuh"
#include "umap_step_fast4.cuh"
namespace tdr {

template <int OCC, int CAP, bool NEG_CG, bool L2H = false, bool PF = false, bool CHEAP = false>
static cudaError_t launch_fast4(const UmapStepParams& p, unsigned blocks, cudaStream_t st) {
    constexpr size_t smem = sizeof(Warp4Smem<CAP>) * kWarps4;
    static const cudaError_t attr = cudaFuncSetAttribute(umap_step_kernel_fast4<OCC, CAP, NEG_CG, L2H, PF, CHEAP>,
                                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (attr != cudaSuccess) return attr;
    if (PF && blocks > (unsigned)(kNumSMs * OCC)) blocks = (unsigned)(kNumSMs * OCC);  // one resident wave
    umap_step_kernel_fast4<OCC, CAP, NEG_CG, L2H, PF, CHEAP><<<blocks, kFastThreads, smem, st>>>(p);
    return cudaSuccess;
}

// precise: 0 = throughput kernel (umap_step_fast4.cuh), 1 = parity kernel (fp64 pow, one warp per row),
//          2 = umap_step_kernel_v2 (libdevice powf; kept as the measured baseline of profiles/r1_step_kernel.md)
static int launch_step(const UmapStepParams& p, int precise, cudaStream_t st) {
    if (precise == 0) {
        const int64_t cap = (int64_t)kNumSMs * 4 * 8;  // grid-stride beyond ~8 waves of resident CTAs
        static const int variant = [] {
            const char* e = getenv("TDR_STEP_FAST");
            return e ? atoi(e) : 4;  // 4 = umap_step_fast4.cuh (lane per row), 3 = umap_step_fast3.cuh (8 rows per warp)
        }();
        static const int cfg = [] {
            const char* e = getenv("TDR_STEP_CFG");
            return e ? atoi(e) : 0;
        }();
        if (variant == 3) {
            int64_t b3 = (p.n_local + kWarps3 * kRows3 - 1) / (kWarps3 * kRows3);
            if (b3 > cap) b3 = cap;
            umap_step_kernel_fast3<4><<<(unsigned)b3, kFastThreads, 0, st>>>(p);
            TDR_LAUNCH_CHECK();
            return TDR_OK;
        }
        int64_t b4 = (p.n_local + kWarps4 * 32 - 1) / (kWarps4 * 32);
        if (b4 > cap) b4 = cap;
        const unsigned g4 = (unsigned)b4;
        cudaError_t err;
        // (CTAs per SM, list entries per warp, negatives through L2 only); ms per iteration at 1 M x 128, k = 15:
        //   (4, 256, cg) 0.225   (5, 256, cg) 0.225   (3, 384, cg) 0.237   (3, 384, ldg) 0.242   (4, 384, ldg) 0.361
        switch (cfg) {
            case 1: err = launch_fast4<3, 384, true>(p, g4, st); break;
            case 2: err = launch_fast4<4, 320, true>(p, g4, st); break;
            case 3: err = launch_fast4<4, 256, false>(p, g4, st); break;
            case 4: err = launch_fast4<4, 288, true>(p, g4, st); break;
            case 5: err = launch_fast4<4, 256, true, true>(p, g4, st); break;  // L2 eviction hints
            case 6: err = launch_fast4<4, 256, true, false, true>(p, g4, st); break;  // persistent + stream prefetch
            case 7: err = launch_fast4<4, 256, true, false, false, true>(p, g4, st); break;  // Philox-7 + cheap pow
            case 8: err = launch_fast4<4, 256, true, false, true, true>(p, g4, st); break;   // 6 + 7
            default: err = launch_fast4<4, 256, true>(p, g4, st); break;
        }
        TDR_CUDA(err);
        TDR_LAUNCH_CHECK();
        return TDR_OK;
    }
    if (precise == 1) {
        int64_t blocks = (p.n_local + kStepWarps - 1) / kStepWarps;
        const int64_t cap = (int64_t)kNumSMs * 32;
        if (blocks > cap) blocks = cap;
        umap_step_kernel<true><<<(unsigned)blocks, kStepWarps * 32, 0, st>>>(p);
    } else {
        const int rows_per_block = kV2Threads / G;
        int64_t blocks = (p.n_local + rows_per_block - 1) / rows_per_block;
        const int64_t cap = (int64_t)kNumSMs * 8 * 4;  // grid-stride beyond 4 waves of 8 resident CTAs per SM
        if (blocks > cap) blocks = cap;
        umap_step_kernel_v2<<<(unsigned)blocks, kV2Threads, 0, st>>>(p);
    }
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

static void fill_consts(UmapStepParams& p, float a, float b, double a64, double b64) {
    // python-side scalars are doubles cast to fp32 at the op (umap.py:244-246, 273-276)
    p.a = a;
    p.b = b;
    p.bm1 = (float)(b64 - 1.0);
    p.two_ab = (float)(2.0 * a64 * b64);
    p.neg_two_b = (float)(-2.0 * b64);
}

}  // namespace tdr

using namespace tdr;

static int umap_step_impl(const float* Z_in, float* Z_out, int64_t n_total, int64_t row0, int64_t n_local,
                          const int64_t* rowptr, const int32_t* col, const float* epochs_per_sample,
                          float* epoch_of_next_sample, const int64_t* neg, int n_neg, int negative_sample_rate,
                          uint64_t seed, int64_t n_iter, double a, double b, float lam, float repulsion, float lr,
                          int precise, float* grad_out, double* gnorm_sq, int* nan_flag, uint64_t* stats,
                          float* const* peer_out, int n_peers, tdr_stream_t stream) {
    TDR_CHECK_ARG(Z_in && Z_out && rowptr && col && epochs_per_sample && epoch_of_next_sample,
                  "tdr_umap_step_f32: null pointer");
    TDR_CHECK_ARG(Z_in != Z_out, "tdr_umap_step_f32: Z_in and Z_out must not alias (Jacobi update)");
    TDR_CHECK_ARG(n_total >= 2 && row0 >= 0 && n_local >= 0 && row0 + n_local <= n_total,
                  "tdr_umap_step_f32: bad row range");
    TDR_CHECK_ARG(n_neg >= 0 && negative_sample_rate >= 0, "tdr_umap_step_f32: bad negative sampling config");
    if (n_local == 0) return TDR_OK;
    UmapStepParams p{};
    p.Zin = reinterpret_cast<const float2*>(Z_in);
    p.Zout = reinterpret_cast<float2*>(Z_out);
    p.n_total = n_total;
    p.row0 = row0;
    p.n_local = n_local;
    p.rowptr = rowptr;
    p.col = col;
    p.eps = epochs_per_sample;
    p.eons = epoch_of_next_sample;
    p.neg = neg;
    p.n_neg = n_neg;
    p.rate = negative_sample_rate;
    p.seed = seed;
    p.n_iter = n_iter;
    fill_consts(p, (float)a, (float)b, a, b);
    p.lam = lam;
    p.rep = repulsion;
    p.lr = lr;
    p.grad_out = reinterpret_cast<float2*>(grad_out);
    p.gnorm_sq = gnorm_sq;
    p.nan_flag = nan_flag;
    p.stats = reinterpret_cast<unsigned long long*>(stats);
    TDR_CHECK_ARG(n_peers >= 0 && n_peers <= 8 && (n_peers == 0 || peer_out), "tdr_umap_step: at most 8 peer buffers");
    p.n_peers = n_peers;
    for (int q = 0; q < n_peers; ++q) {
        TDR_CHECK_ARG(peer_out[q] && (const float*)peer_out[q] != Z_in, "tdr_umap_step: bad peer buffer");
        p.peer_out[q] = reinterpret_cast<float2*>(peer_out[q]);
    }
    return launch_step(p, precise, (cudaStream_t)stream);
}

extern "C" TDR_API int tdr_umap_step_f32(const float* Z_in, float* Z_out, int64_t n_total, int64_t row0, int64_t n_local,
                                 const int64_t* rowptr, const int32_t* col, const float* epochs_per_sample,
                                 float* epoch_of_next_sample, const int64_t* neg, int n_neg,
                                 int negative_sample_rate, uint64_t seed, int64_t n_iter, double a, double b,
                                 float lam, float repulsion, float lr, int precise, float* grad_out,
                                 double* gnorm_sq, int* nan_flag, uint64_t* stats, tdr_stream_t stream) {
    return umap_step_impl(Z_in, Z_out, n_total, row0, n_local, rowptr, col, epochs_per_sample, epoch_of_next_sample, neg,
                          n_neg, negative_sample_rate, seed, n_iter, a, b, lam, repulsion, lr, precise, grad_out, gnorm_sq,
                          nan_flag, stats, nullptr, 0, stream);
}

extern "C" TDR_API int tdr_umap_step_p2p_f32(const float* Z_in, float* Z_out, int64_t n_total, int64_t row0,
                                             int64_t n_local, const int64_t* rowptr, const int32_t* col,
                                             const float* epochs_per_sample, float* epoch_of_next_sample, int n_neg,
                                             int negative_sample_rate, uint64_t seed, int64_t n_iter, double a, double b,
                                             float lam, float repulsion, float lr, double* gnorm_sq, int* nan_flag,
                                             const uint64_t* peer_out_ptrs /*host*/, int n_peers, tdr_stream_t stream) {
    float* peers[8] = {nullptr};
    TDR_CHECK_ARG(n_peers >= 0 && n_peers <= 8, "tdr_umap_step_p2p_f32: at most 8 peers");
    for (int q = 0; q < n_peers; ++q) peers[q] = reinterpret_cast<float*>(peer_out_ptrs[q]);
    return umap_step_impl(Z_in, Z_out, n_total, row0, n_local, rowptr, col, epochs_per_sample, epoch_of_next_sample,
                          nullptr, n_neg, negative_sample_rate, seed, n_iter, a, b, lam, repulsion, lr, 0, nullptr,
                          gnorm_sq, nan_flag, nullptr, peers, n_peers, stream);
}

extern "C" TDR_API int tdr_umap_run_f32(float* Z_a, float* Z_b, int64_t n_total, const int64_t* rowptr,
                                const int32_t* col, const float* epochs_per_sample, float* epoch_of_next_sample,
                                int n_neg, int negative_sample_rate, uint64_t seed, int64_t n_iter0, int n_steps,
                                const float* lrs_host, double a, double b, float lam, float repulsion, int precise,
                                double* gnorm_sq, int* nan_flag, uint64_t* stats, tdr_stream_t stream) {
    TDR_CHECK_ARG(Z_a && Z_b && Z_a != Z_b && lrs_host && n_steps >= 0, "tdr_umap_run_f32: bad arguments");
    float* src = Z_a;
    float* dst = Z_b;
    for (int t = 0; t < n_steps; ++t) {
        // gradient norm is only wanted for the last step of the batch (the host's check_interval)
        int rc = tdr_umap_step_f32(src, dst, n_total, 0, n_total, rowptr, col, epochs_per_sample,
                                   epoch_of_next_sample, nullptr, n_neg, negative_sample_rate, seed, n_iter0 + t,
                                   a, b, lam, repulsion, lrs_host[t], precise, nullptr,
                                   (t == n_steps - 1) ? gnorm_sq : nullptr, nan_flag, stats, stream);
        if (rc != TDR_OK) return rc;
        float* tmp = src;
        src = dst;
        dst = tmp;
    }
    return TDR_OK;
}

// ---- cross-GPU barrier on peer-mapped flags + the multi-step sharded loop -------------------------------
// flags: one uint32 per rank in EVERY rank's (zero-initialised, peer-mapped) flag buffer.  Rank r announces epoch
// E by storing E into slot r of every peer's buffer, then waits until every slot of its own buffer has reached E.
// Launched on the stream right after the step kernel, so the step's peer stores have completed (kernel boundary)
// before the flag goes out; the next step kernel in the stream starts only after all peers have announced.
namespace tdr {
__global__ void peer_barrier_kernel(volatile uint32_t* my_flags, const UmapStepParams peers_as_ptrs, int n_peers,
                                    int my_rank, int world, uint32_t epoch) {
    const int lane = threadIdx.x;
    if (lane < n_peers) {
        __threadfence_system();
        volatile uint32_t* remote = reinterpret_cast<volatile uint32_t*>(peers_as_ptrs.peer_out[lane]);
        remote[my_rank] = epoch;
    }
    if (lane < world && lane != my_rank) {
        while ((int32_t)(my_flags[lane] - epoch) < 0) {
        }
    }
    __threadfence_system();
}
}  // namespace tdr

extern "C" TDR_API int tdr_umap_run_p2p_f32(float* Z_a, float* Z_b, int64_t n_total, int64_t row0, int64_t n_local,
                                            const int64_t* rowptr, const int32_t* col, const float* epochs_per_sample,
                                            float* epoch_of_next_sample, int n_neg, int negative_sample_rate,
                                            uint64_t seed, int64_t n_iter0, int n_steps, const float* lrs_host, double a,
                                            double b, float lam, float repulsion, double* gnorm_sq, int* nan_flag,
                                            const uint64_t* peers_a /*host*/, const uint64_t* peers_b /*host*/,
                                            uint32_t* my_flags, const uint64_t* peer_flags /*host*/, int n_peers,
                                            int rank, int world, uint32_t epoch0, tdr_stream_t stream) {
    TDR_CHECK_ARG(Z_a && Z_b && Z_a != Z_b && lrs_host && n_steps >= 0 && my_flags && peers_a && peers_b && peer_flags,
                  "tdr_umap_run_p2p_f32: bad arguments");
    TDR_CHECK_ARG(n_peers >= 1 && n_peers <= 8 && world == n_peers + 1 && rank >= 0 && rank < world,
                  "tdr_umap_run_p2p_f32: 2..9 ranks");
    UmapStepParams fl{};
    for (int q = 0; q < n_peers; ++q) fl.peer_out[q] = reinterpret_cast<float2*>(peer_flags[q]);
    float* src = Z_a;
    float* dst = Z_b;
    const uint64_t* dst_peers = peers_b;
    const uint64_t* src_peers = peers_a;
    for (int t = 0; t < n_steps; ++t) {
        int rc = tdr_umap_step_p2p_f32(src, dst, n_total, row0, n_local, rowptr, col, epochs_per_sample,
                                       epoch_of_next_sample, n_neg, negative_sample_rate, seed, n_iter0 + t, a, b, lam,
                                       repulsion, lrs_host[t], (t == n_steps - 1) ? gnorm_sq : nullptr, nan_flag,
                                       dst_peers, n_peers, stream);
        if (rc != TDR_OK) return rc;
        tdr::peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(my_flags, fl, n_peers, rank, world,
                                                                     epoch0 + (uint32_t)t + 1u);
        TDR_LAUNCH_CHECK();
        float* tmp = src;
        src = dst;
        dst = tmp;
        const uint64_t* tp = src_peers;
        src_peers = dst_peers;
        dst_peers = tp;
    }
    return TDR_OK;
}
