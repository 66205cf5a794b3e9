// Throughput variant of the UMAP step (included by umap_step.cu after UmapStepParams / block_flush).
//
// ncu on the first two versions (profiles/r1_step_kernel.md) showed the step is *issue*-bound, not
// bandwidth-bound: 55 % of all warp instructions were libdevice powf, executed with 5 of 32 lanes
// active in the attraction loop (only ~30 % of a row's edges are due in a given iteration) and 15 of
// 32 in the repulsion loop.  This version therefore
//   * gives 8 lanes to a row (4 rows per warp) and loads the three edge arrays together,
//     unconditionally, 4 chunks at a time (memory-level parallelism for the HBM stream);
//   * compacts the due edges of a row through a 32-entry shared-memory slot list so the attraction
//     math runs on dense lanes (typically one pass with 7 of 8 lanes busy);
//   * pools the negatives of the warp's 4 rows: quads of 4 slots (one Philox block) are dealt to the 32
//     lanes, 4 gathers in flight per lane;
//   * evaluates x^y as 2^(y log2 x) with the exponent split off exactly and the product carried in
//     hi + lo form (MUFU.LG2 / MUFU.EX2 on reduced arguments): ~4e-7 relative error, ~20 instructions
//     instead of ~120; attraction needs a single power: D^b = D * D^(b-1).
// The arithmetic that defines the result (umap.py:236-292) is otherwise the same op sequence as the
// parity kernel umap_step_kernel<true>.
#pragma once

namespace tdr {

constexpr int FG = 8;            // lanes per row
constexpr int FU = 4;            // chunks in flight per lane
constexpr int kFastThreads = 256;
constexpr int kFastGroups = kFastThreads / FG;

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

__device__ __forceinline__ float fgroup_sum(float v, unsigned mask) {
#pragma unroll
    for (int o = FG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(kFastThreads, MIN_CTAS) umap_step_kernel_fast(const UmapStepParams p) {
    __shared__ int s_col[kFastGroups][FG * FU];  // compacted columns of the due edges of the current chunk set
    const int lane = threadIdx.x & 31;
    const int l = lane & (FG - 1);
    const int gshift = lane & ~(FG - 1);
    const unsigned gmask = ((1u << FG) - 1u) << gshift;
    const int gslot = threadIdx.x / FG;
    const int64_t group_global = ((int64_t)blockIdx.x * kFastThreads + threadIdx.x) / FG;
    const int64_t n_groups = (int64_t)gridDim.x * kFastGroups;
    const float due_before = (float)(p.n_iter + 1);  // umap.py:251
    const Philox rng(p.seed);
    const uint32_t nm1 = (uint32_t)(p.n_total - 1);
    double gn_local = 0.0;
    bool saw_nan = false;
    unsigned long long n_act = 0, n_neg_used = 0;

    // warp-uniform trip count: the 4 groups of a warp take 4 consecutive rows; groups past the end idle
    const int64_t warp_first = group_global - (lane >> 3);
    for (int64_t rb = warp_first; rb < p.n_local; rb += n_groups) {
        const int64_t r = rb + (lane >> 3);
        const bool live = r < p.n_local;
        const int64_t gi = p.row0 + (live ? r : 0);
        const float2 zi = __ldg(p.Zin + gi);
        const int64_t e0 = live ? __ldg(p.rowptr + r) : 0, e1 = live ? __ldg(p.rowptr + r + 1) : 0;
        float gx = 0.0f, gy = 0.0f;
        int active = 0;
        // ---- attraction (umap.py:236-264)
        for (int64_t base = e0; base < e1; base += FG * FU) {
            float nxt[FU], ep[FU];
            int cj[FU];
#pragma unroll
            for (int u = 0; u < FU; ++u) {
                const int64_t e = base + u * FG + l;
                const bool ok = e < e1;
                nxt[u] = ok ? p.eons[e] : INFINITY;
                cj[u] = ok ? __ldg(p.col + e) : 0;
                ep[u] = ok ? __ldg(p.eps + e) : 0.0f;
            }
            int n_due = 0;
#pragma unroll
            for (int u = 0; u < FU; ++u) {
                const bool due = nxt[u] <= due_before;
                const unsigned bal = (__ballot_sync(gmask, due) >> gshift) & ((1u << FG) - 1u);
                if (due) {
                    p.eons[base + u * FG + l] = __fadd_rn(nxt[u], ep[u]);  // umap.py:253-255
                    s_col[gslot][n_due + __popc(bal & ((1u << l) - 1u))] = cj[u];
                }
                n_due += __popc(bal);
            }
            __syncwarp(gmask);
            active += n_due;
            for (int t0 = 0; t0 < n_due; t0 += FG) {
                const int t = t0 + l;
                if (t < n_due) {
                    const float2 zj = __ldg(p.Zin + s_col[gslot][t]);
                    const float dx = __fsub_rn(zi.x, zj.x), dy = __fsub_rn(zi.y, zj.y);
                    const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));  // distance/base.py:384-385
                    if (D > 0.0f) {  // umap.py:243-247
                        const float pw = pow_fast(D, p.bm1);  // D^(b-1); D^b = D * D^(b-1)
                        const float den = __fadd_rn(1.0f, __fmul_rn(p.a, __fmul_rn(pw, D)));
                        const float coef = __fdiv_rn(__fmul_rn(pw, p.two_ab), den);
                        gx = fmaf(dx, coef, gx);
                        gy = fmaf(dy, coef, gy);
                    }
                }
            }
            __syncwarp(gmask);
        }
        gx = fgroup_sum(gx, gmask);
        gy = fgroup_sum(gy, gmask);
        gx = fminf(fmaxf(gx, -4.0f), 4.0f);  // umap.py:263
        gy = fminf(fmaxf(gy, -4.0f), 4.0f);

        // ---- repulsion (umap.py:266-292) on the first rate*active negatives.
        // The four rows of the warp pool their work: a work item is a quad (row, 4 consecutive negative
        // slots = one Philox block); quads are dealt to the 32 lanes, so a row with a small quota does not
        // leave its 8 lanes idle while a neighbour with a large one loops.
        int quota = active * p.rate;
        if (quota > p.n_neg) quota = p.n_neg;
        __syncwarp();
        const int my_q = live ? quota : 0;
        const int q0 = __shfl_sync(0xffffffffu, my_q, 0), q1 = __shfl_sync(0xffffffffu, my_q, FG),
                  q2 = __shfl_sync(0xffffffffu, my_q, 2 * FG), q3 = __shfl_sync(0xffffffffu, my_q, 3 * FG);
        const int o1 = (q0 + 3) >> 2, o2 = o1 + ((q1 + 3) >> 2), o3 = o2 + ((q2 + 3) >> 2), nquad = o3 + ((q3 + 3) >> 2);
        float ax[4] = {0.f, 0.f, 0.f, 0.f}, ay[4] = {0.f, 0.f, 0.f, 0.f};
        for (int w0 = 0; w0 < nquad; w0 += 32) {  // warp-uniform trip count (usually 1)
            const int w = w0 + lane;
            const int g = (w >= o1) + (w >= o2) + (w >= o3);
            const int src = g * FG;
            const float zx = __shfl_sync(0xffffffffu, zi.x, src), zy = __shfl_sync(0xffffffffu, zi.y, src);
            const int64_t gj = __shfl_sync(0xffffffffu, (long long)gi, src);
            const int64_t rj = __shfl_sync(0xffffffffu, (long long)r, src);
            const int qg = g == 0 ? q0 : g == 1 ? q1 : g == 2 ? q2 : q3;
            const int quad = w - (g == 0 ? 0 : g == 1 ? o1 : g == 2 ? o2 : o3);
            const int nval = w < nquad ? min(4, qg - 4 * quad) : 0;
            int64_t j[4];
            if (p.neg) {
#pragma unroll
                for (int u = 0; u < 4; ++u) j[u] = (u < nval) ? __ldg(p.neg + rj * p.n_neg + 4 * quad + u) : gj;
            } else {
                const uint4 wd = rng((uint32_t)p.n_iter, (uint32_t)(p.n_iter >> 32) ^ (uint32_t)(gj >> 32), (uint32_t)gj,
                                     (uint32_t)quad);
                const uint32_t wv[4] = {wd.x, wd.y, wd.z, wd.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t t = __umulhi(wv[u], nm1);  // uniform on [0, N-2] (N < 2^31: indices are int32)
                    j[u] = (int64_t)(t + ((t >= (uint32_t)gj) ? 1u : 0u));  // NE base.py:636
                }
            }
            float2 zn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (u < nval) zn[u] = __ldg(p.Zin + j[u]);
            float sx = 0.0f, sy = 0.0f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (u < nval) {
                    const float dx = __fsub_rn(zx, zn[u].x), dy = __fsub_rn(zy, zn[u].y);
                    const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                    const float den = __fadd_rn(1.0f, __fmul_rn(p.a, pow_fast(D, p.b)));  // umap.py:273
                    const float coef = __fmul_rn(__frcp_rn(__fmul_rn(__fadd_rn(D, 1e-3f), den)), p.neg_two_b);  // :274-276
                    sx = fmaf(dx, coef, sx);
                    sy = fmaf(dy, coef, sy);
                }
            }
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
                ax[gg] += (g == gg) ? sx : 0.0f;
                ay[gg] += (g == gg) ? sy : 0.0f;
            }
        }
        // transposed reduction: 4 groups x 32 lanes -> lane L ends with the total of group L >> 3
        // (2 + 1 + 3 shuffles per component instead of 4 full butterflies)
        const bool up = lane & 16, odd = lane & 8;
        float rx, ry;
        {
            float k0 = up ? ax[2] : ax[0], k1 = up ? ax[3] : ax[1];
            k0 += __shfl_xor_sync(0xffffffffu, up ? ax[0] : ax[2], 16);
            k1 += __shfl_xor_sync(0xffffffffu, up ? ax[1] : ax[3], 16);
            rx = (odd ? k1 : k0) + __shfl_xor_sync(0xffffffffu, odd ? k0 : k1, 8);
            float m0 = up ? ay[2] : ay[0], m1 = up ? ay[3] : ay[1];
            m0 += __shfl_xor_sync(0xffffffffu, up ? ay[0] : ay[2], 16);
            m1 += __shfl_xor_sync(0xffffffffu, up ? ay[1] : ay[3], 16);
            ry = (odd ? m1 : m0) + __shfl_xor_sync(0xffffffffu, odd ? m0 : m1, 8);
        }
        rx = fgroup_sum(rx, gmask);
        ry = fgroup_sum(ry, gmask);
        rx = fminf(fmaxf(rx, -4.0f), 4.0f);  // umap.py:291
        ry = fminf(fmaxf(ry, -4.0f), 4.0f);
        const float g0 = __fadd_rn(__fmul_rn(p.lam, gx), __fmul_rn(p.rep, rx));  // NE base.py:237-241
        const float g1 = __fadd_rn(__fmul_rn(p.lam, gy), __fmul_rn(p.rep, ry));
        float2 zo;  // torch.optim.SGD: param.add_(grad, alpha=-lr)
        zo.x = fmaf(-p.lr, g0, zi.x);
        zo.y = fmaf(-p.lr, g1, zi.y);
        if (l == 0 && live) {
            if (p.grad_out) p.grad_out[r] = make_float2(g0, g1);
            gn_local += (double)g0 * g0 + (double)g1 * g1;
            saw_nan |= (zo.x != zo.x) || (zo.y != zo.y);
            n_act += active;
            n_neg_used += quota;
        }
        // The warp's 4 rows are consecutive: move the 4 results to lanes 0-3 and write ONE contiguous 32-byte
        // segment per destination (own buffer + every peer over NVLink) instead of 4 scattered 8-byte stores.
        __syncwarp();
        const float ox = __shfl_sync(0xffffffffu, zo.x, (lane & 3) * FG), oy = __shfl_sync(0xffffffffu, zo.y, (lane & 3) * FG);
        if (lane < 4 && rb + lane < p.n_local) store_row(p, p.row0 + rb + lane, make_float2(ox, oy));
    }
    block_flush(l == 0, gn_local, saw_nan, n_act, n_neg_used, p);
}

}  // namespace tdr
