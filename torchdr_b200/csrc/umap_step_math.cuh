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

__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

}  // namespace tdr
