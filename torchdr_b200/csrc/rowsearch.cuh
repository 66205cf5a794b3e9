// Per-row bandwidth search, one warp per row.  Device-side restatement of
// torchdr/utils/root_search.py:17-198 (bracket + bisection, tol 1e-6) specialised to
//   UMAPAffinity      torchdr/affinity/knn_normalized.py:445-468
//   EntropicAffinity  torchdr/affinity/entropic.py:272-310 (+ bounds :96-115)
// Rows are independent in the reference (masked batched updates), so each warp runs the
// same update sequence on its own row and stops when the row's state can no longer change.
//
// Arithmetic notes: every reference op is a separately rounded fp32 torch op, so the
// code below uses __f*_rn intrinsics where the compiler could otherwise contract to FMA.
#pragma once

#include "common.cuh"

namespace tdr {

constexpr int kMaxEPL = TDR_MAX_K / 32;  // elements per lane
constexpr float kRootTol = 1e-6f;        // root_search.py:13

// f is evaluated cooperatively by the warp and returns the same value in every lane.
template <class F>
__device__ __forceinline__ float bracket_bisect(F f, float lo, float hi, int max_iter) {
    // root_search.py:176-185
    for (int it = 0; it < max_iter; ++it) {
        if (!(f(lo) > 0.0f)) break;
        hi = fminf(hi, lo);
        lo = lo * 0.5f;
    }
    // root_search.py:187-196
    for (int it = 0; it < max_iter; ++it) {
        if (!(f(hi) < 0.0f)) break;
        lo = fmaxf(lo, hi);
        hi = hi * 2.0f;
    }
    // root_search.py:53-75
    float f_lo = f(lo);
    float mid = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
    float f_mid = f(mid);
    for (int it = 0; it < max_iter; ++it) {
        if (!(fabsf(f_mid) >= kRootTol)) break;
        if (__fmul_rn(f_mid, f_lo) > 0.0f) {
            lo = mid;
            f_lo = f_mid;
        } else {
            hi = mid;
        }
        const float next = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
        if (next == mid) break;  // fixed point: every later reference iteration reproduces `mid`
        mid = next;
        f_mid = f(mid);
    }
    return mid;
}

// Reductions over a group of W lanes (W = 32: the warp; W = 16: a half warp, two rows per warp).  xor-butterflies from
// W/2 down: a 32-lane butterfly whose upper 16 lanes hold the neutral element gives the same bits as the 16-lane one
// (x + 0, min(x, +inf), max(x, -inf) are exact), so results do not depend on the group width.
template <int W>
__device__ __forceinline__ float grp_sum(float v, unsigned mask) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
template <int W>
__device__ __forceinline__ float grp_max(float v, unsigned mask) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, o));
    return v;
}
template <int W>
__device__ __forceinline__ float grp_min(float v, unsigned mask) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(mask, v, o));
    return v;
}

// ---- UMAP ------------------------------------------------------------------------------
// c[e] holds C[row, lane + W e] (lane = index inside the row's group of W lanes); entries with index >= k are ignored.
template <int EPL, int W = 32>
struct UmapRow {
    float c[EPL];
    int k, lane;
    float rho, target;
    unsigned mask = 0xffffffffu;  // the lanes of this row's group

    __device__ __forceinline__ bool valid(int e) const { return lane + W * e < k; }

    __device__ __forceinline__ void init() {
        float m = INFINITY;
#pragma unroll
        for (int e = 0; e < EPL; ++e)
            if (valid(e)) m = fminf(m, c[e]);
        rho = grp_min<W>(m, mask);  // knn_normalized.py:445
    }
    // knn_normalized.py:452-454: exp(logsumexp(-(C-rho)/sigma)) - log2(k)
    __device__ __forceinline__ float gap(float sigma) const {
        float x[EPL];
        float m = -INFINITY;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            x[e] = __fdiv_rn(-__fsub_rn(c[e], rho), sigma);
            if (valid(e)) m = fmaxf(m, x[e]);
        }
        m = grp_max<W>(m, mask);
        float s = 0.0f;
#pragma unroll
        for (int e = 0; e < EPL; ++e)
            if (valid(e)) s += expf(__fsub_rn(x[e], m));
        s = grp_sum<W>(s, mask);
        return __fsub_rn(expf(__fadd_rn(logf(s), m)), target);
    }
    __device__ __forceinline__ float solve(int max_iter) {
        return bracket_bisect([&](float s) { return gap(s); }, 1.0f, 1.0f, max_iter);
    }
    __device__ __forceinline__ float p(int e, float sigma) const {  // knn_normalized.py:464-465
        return expf(__fdiv_rn(-__fsub_rn(c[e], rho), sigma));
    }
};

// ---- Entropic --------------------------------------------------------------------------
struct EntropicConsts {
    float target;       // log(perplexity) + 1, entropic.py:272
    float log_n_total;  // entropic.py:308-310
    int use_bounds;     // entropic.py:280-287
    float b_num, b_den, b_lr, b_logp1;
};

template <int EPL>
struct EntropicRow {
    float c[EPL];
    int k, lane;
    float target;

    __device__ __forceinline__ bool valid(int e) const { return lane + 32 * e < k; }

    // entropic.py:274-277 ; utils/utils.py:167-168 ; torch.logsumexp = max + log(sum(exp(x-max)))
    __device__ __forceinline__ float lse(float eps, float (&l)[EPL]) const {
        float m = -INFINITY;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            l[e] = __fdiv_rn(-c[e], eps);
            if (valid(e)) m = fmaxf(m, l[e]);
        }
        m = warp_max(m);
        float s = 0.0f;
#pragma unroll
        for (int e = 0; e < EPL; ++e)
            if (valid(e)) s += expf(__fsub_rn(l[e], m));
        s = warp_sum(s);
        return __fadd_rn(logf(s), m);
    }
    __device__ __forceinline__ float gap(float eps) const {
        float l[EPL];
        const float z = lse(eps, l);
        float h = 0.0f;
#pragma unroll
        for (int e = 0; e < EPL; ++e)
            if (valid(e)) {
                const float ln = __fsub_rn(l[e], z);
                h += __fmul_rn(expf(ln), __fsub_rn(ln, 1.0f));
            }
        h = warp_sum(h);
        return __fsub_rn(-h, target);
    }
    // entropic.py:96-115: per-row bracket from the two smallest and the largest distance
    __device__ __forceinline__ void bounds(const EntropicConsts& K, float& begin, float& end) const {
        float mx = -INFINITY, mn = INFINITY;
        int mn_pos = 0x7fffffff;
#pragma unroll
        for (int e = 0; e < EPL; ++e)
            if (valid(e)) {
                mx = fmaxf(mx, c[e]);
                mn = fminf(mn, c[e]);
            }
        mx = warp_max(mx);
        mn = warp_min(mn);
#pragma unroll
        for (int e = 0; e < EPL; ++e)
            if (valid(e) && c[e] == mn) mn_pos = min(mn_pos, lane + 32 * e);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn_pos = min(mn_pos, __shfl_xor_sync(0xffffffffu, mn_pos, o));
        float m2 = INFINITY;
#pragma unroll
        for (int e = 0; e < EPL; ++e)
            if (valid(e) && lane + 32 * e != mn_pos) m2 = fminf(m2, c[e]);
        m2 = warp_min(m2);
        const float dN = mx, d1 = mn, d2 = m2;
        const float span = __fsub_rn(dN, d1), step = __fsub_rn(d2, d1);
        const float t1 = __fdiv_rn(K.b_num, __fmul_rn(K.b_den, span));
        const float t2 = sqrtf(__fdiv_rn(K.b_lr, __fsub_rn(__fmul_rn(dN, dN), __fmul_rn(d1, d1))));
        float beta_lo = fmaxf(t1, t2);
        if (t1 != t1 || t2 != t2) beta_lo = NAN;  // torch.max propagates NaN
        const float beta_hi = __fdiv_rn(K.b_logp1, step);
        begin = __fadd_rn(__fdiv_rn(1.0f, beta_hi), 1e-6f);  // entropic.py:287
        end = __fdiv_rn(1.0f, beta_lo);
    }
    __device__ __forceinline__ float solve(const EntropicConsts& K, int max_iter) {
        float b = 1.0f, e = 1.0f;
        if (K.use_bounds) bounds(K, b, e);
        return bracket_bisect([&](float x) { return gap(x); }, b, e, max_iter);
    }
};

}  // namespace tdr
