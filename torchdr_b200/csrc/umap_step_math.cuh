// Shared pieces of the throughput ("fast") UMAP step kernels (included by umap_step.cu).
//
// ncu on the first versions (profiles/r1_step_kernel.md) showed the step is *issue*-bound, not
// bandwidth-bound: 55 % of all warp instructions were libdevice powf.  x^y is therefore evaluated as
// 2^(y log2 x) with the exponent split off exactly and the product carried in hi + lo form
// (MUFU.LG2 / MUFU.EX2 on reduced arguments): ~4e-7 relative error, ~20 instructions instead of ~120;
// attraction needs a single power: D^b = D * D^(b-1).  Reciprocals are MUFU.RCP + one Newton step
// (<= 1 ulp) instead of the IEEE-rounded __frcp_rn / __fdiv_rn sequences (~17 instructions each).
// The arithmetic that defines the result (umap.py:236-292) is otherwise the same op sequence as the
// parity kernel umap_step_kernel<true>.
#pragma once

namespace tdr {

constexpr int FU = 4;            // chunks in flight per lane
constexpr int kFastThreads = 256;

// x > 0 (or 0): 2^(y * log2 x).  log2 x = e + log2 m with m in [sqrt(1/2), sqrt(2)); y*e is carried as
// hi + lo (fma residual) so the only inexact pieces are MUFU.LG2(m) (|err| <= 2^-22) and MUFU.EX2 on
// a fraction in [-1, 1].
__device__ __forceinline__ float pow_fast(float x, float y) {
    const int ix = __float_as_int(x);
    const int e = (ix - 0x3f3504f3) >> 23;
    const float m = __int_as_float(ix - (e << 23));
    const float ef = (float)e;
    const float hi = y * ef;
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(m));  // m is normal: no denormal pre-scaling needed
    const float lo = fmaf(y, ef, -hi) + y * lg;
    float n = rintf(hi);
    const float f = (hi - n) + lo;
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f));  // |f| <= ~1: rel. error 2^-22
    n = fmaxf(n, -100.0f);                                   // x = 0 -> ~1e-30 (acts as 0 next to the +1 / +1e-3 terms)
    return __int_as_float(__float_as_int(r) + ((int)n << 23));
}

// L2 eviction-priority hints (createpolicy + ld.global.L2::cache_hint).  At 10 M points the embedding (80 MB) and
// the edge streams (3.2 GB per iteration) compete for the 126 MB L2: the streams are read once per iteration and
// are marked evict-first, the gathered rows of Z are marked evict-last.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float2 ldg_f2_hint(const float2* a, uint64_t pol) {
    float2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ float2 ldcg_f2_hint(const float2* a, uint64_t pol) {
    float2 v;
    asm volatile("ld.global.cg.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_f32_hint(const float* a, uint64_t pol) {
    float v;
    asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ float ldg_f32_hint(const float* a, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ int ldg_s32_hint(const int* a, uint64_t pol) {
    int v;
    asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
    return v;
}

// x^y = 2^(y log2 x) without the hi + lo product: 3 instructions; the rounding of y * log2 x (|.| up to ~30 for the
// distances of the first iterations) costs ~1e-6 relative instead of ~4e-7.  x = 0 gives 0 for y > 0 and +inf for
// y < 0 — the callers multiply by D or add 1e-3 / 1, as with pow_fast.  EXPERIMENTAL (TDR_STEP_CFG=7).
__device__ __forceinline__ float pow_cheap(float x, float y) {
    float lg, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y * fmaxf(lg, -100.0f)));
    return r;
}

__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

}  // namespace tdr
