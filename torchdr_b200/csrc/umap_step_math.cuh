// Shared pieces of the throughput ("fast") UMAP step kernels (included by umap_step.cu).
//
// ncu on the first versions (profiles/r1_step_kernel.md) showed the step is *issue*-bound, not
// bandwidth-bound: 55 % of all warp instructions were libdevice powf.  x^y is therefore evaluated as
// 2^(y log2 x) on MUFU.LG2 / MUFU.EX2 (3 instructions instead of ~120, ~1e-6 relative; the hi + lo product variant
// of round 1 — ~20 instructions, ~4e-7 — measured 4 % slower for no visible gain in the single-step parity bound);
// attraction needs a single power: D^b = D * D^(b-1).  Reciprocals are MUFU.RCP + one Newton step
// (<= 1 ulp) instead of the IEEE-rounded __frcp_rn / __fdiv_rn sequences (~17 instructions each).
// The arithmetic that defines the result (umap.py:236-292) is otherwise the same op sequence as the
// parity kernel umap_step_kernel<true>.
#pragma once

namespace tdr {

constexpr int FU = 4;            // chunks in flight per lane
constexpr int kFastThreads = 256;

// x^y = 2^(y log2 x): 3 instructions; the rounding of y * log2 x (|.| up to ~30 for the distances of the first
// iterations) costs ~1e-6 relative.  x = 0 gives ~0 for y > 0 and a large finite value for y < 0 — the callers
// multiply by D or add 1e-3 / 1.
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
