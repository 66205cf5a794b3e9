// Second throughput variant of the UMAP step (included by umap_step.cu after umap_step_fast.cuh).
//
// Same work decomposition and the same arithmetic as umap_step_kernel_fast (8 lanes per row, due-edge
// compaction, pooled negative quads, pow_fast); what changes is the instruction count.  A per-line SASS
// count of the first version showed ~420 of its ~1700 instructions in shuffle / ballot plumbing: every
// __shfl_sync / __ballot_sync with a per-group (non-constant) member mask compiles to a
// WARPSYNC + SHFL/VOTE + ENDCOLLECTIVE sequence with convergence barriers, and the per-group trip counts
// of the attraction loop forced those masks.  Here
//   * every loop has a WARP-uniform trip count (max over the warp's 4 rows via REDUX / VOTE.ANY), so all
//     collectives use the full mask and are single instructions;
//   * the row's edge segment is addressed as base pointer + 32-bit offset;
//   * reciprocals are MUFU.RCP + one Newton step (<= 1 ulp) instead of the IEEE-rounded __frcp_rn /
//     __fdiv_rn sequences (~17 instructions each) — the kernel is "fast" mode: pow_fast already carries
//     4e-7 relative error, parity runs use the precise kernels;
//   * the repulsion quads take their row from the warp's base row (rows of a warp are consecutive) instead
//     of 64-bit shuffles.
#pragma once

namespace tdr {

__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(kFastThreads, MIN_CTAS) umap_step_kernel_fast2(const UmapStepParams p) {
    __shared__ int s_col[kFastThreads / 32][4 * FG * FU];  // per warp: 4 rows x 32 compacted columns
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int l = lane & (FG - 1);
    const int g = lane >> 3;          // row of the warp's quartet handled by this lane
    const int gshift = lane & ~(FG - 1);
    int* const my_col = &s_col[threadIdx.x >> 5][g * 32];
    const unsigned lt_mask = (1u << l) - 1u;
    const int64_t warp_global = ((int64_t)blockIdx.x * kFastThreads + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t)gridDim.x * (kFastThreads / 32);
    const float due_before = (float)(p.n_iter + 1);  // umap.py:251
    const Philox rng(p.seed);
    const uint32_t nm1 = (uint32_t)(p.n_total - 1);
    const uint32_t c0 = (uint32_t)p.n_iter, c1 = (uint32_t)(p.n_iter >> 32);
    double gn_local = 0.0;
    bool saw_nan = false;
    unsigned long long n_act = 0, n_neg_used = 0;

    for (int64_t rb = warp_global * 4; rb < p.n_local; rb += n_warps * 4) {
        const int64_t r = rb + g;
        const bool live = r < p.n_local;
        const int64_t gi = p.row0 + (live ? r : rb);
        const float2 zi = __ldg(p.Zin + gi);
        const int64_t e0 = live ? __ldg(p.rowptr + r) : 0;
        const int deg = live ? (int)(__ldg(p.rowptr + r + 1) - e0) : 0;
        float* const eons_r = p.eons + e0;
        const float* const eps_r = p.eps + e0;
        const int32_t* const col_r = p.col + e0;
        const int max_deg = __reduce_max_sync(FULL, deg);
        float gx = 0.0f, gy = 0.0f;
        int active = 0;
        // ---- attraction (umap.py:236-264)
        for (int base = 0; base < max_deg; base += FG * FU) {
            float nxt[FU], ep[FU];
            int cj[FU];
#pragma unroll
            for (int u = 0; u < FU; ++u) {
                const int off = base + u * FG + l;
                const bool ok = off < deg;
                nxt[u] = ok ? eons_r[off] : INFINITY;
                cj[u] = ok ? __ldg(col_r + off) : 0;
                ep[u] = ok ? __ldg(eps_r + off) : 0.0f;
            }
            int n_due = 0;
#pragma unroll
            for (int u = 0; u < FU; ++u) {
                const bool due = nxt[u] <= due_before;
                const unsigned bal = (__ballot_sync(FULL, due) >> gshift) & 0xffu;
                if (due) {
                    eons_r[base + u * FG + l] = __fadd_rn(nxt[u], ep[u]);  // umap.py:253-255
                    my_col[n_due + __popc(bal & lt_mask)] = cj[u];
                }
                n_due += __popc(bal);
            }
            __syncwarp();
            active += n_due;
            for (int t = l; __any_sync(FULL, t < n_due); t += FG) {
                if (t < n_due) {
                    const float2 zj = __ldg(p.Zin + my_col[t]);
                    const float dx = __fsub_rn(zi.x, zj.x), dy = __fsub_rn(zi.y, zj.y);
                    const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));  // distance/base.py:384-385
                    if (D > 0.0f) {  // umap.py:243-247
                        const float pw = pow_fast(D, p.bm1);  // D^(b-1); D^b = D * D^(b-1)
                        const float den = __fadd_rn(1.0f, __fmul_rn(p.a, __fmul_rn(pw, D)));
                        const float coef = __fmul_rn(__fmul_rn(pw, p.two_ab), rcp_fast(den));
                        gx = fmaf(dx, coef, gx);
                        gy = fmaf(dy, coef, gy);
                    }
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = FG / 2; o > 0; o >>= 1) {
            gx += __shfl_xor_sync(FULL, gx, o);
            gy += __shfl_xor_sync(FULL, gy, o);
        }
        gx = fminf(fmaxf(gx, -4.0f), 4.0f);  // umap.py:263
        gy = fminf(fmaxf(gy, -4.0f), 4.0f);

        // ---- repulsion (umap.py:266-292) on the first rate*active negatives, pooled over the warp's 4 rows:
        // a work item is a quad (row, 4 consecutive negative slots = one Philox block), dealt to the 32 lanes.
        int quota = active * p.rate;
        if (quota > p.n_neg) quota = p.n_neg;
        if (!live) quota = 0;
        const int q0 = __shfl_sync(FULL, quota, 0), q1 = __shfl_sync(FULL, quota, FG), q2 = __shfl_sync(FULL, quota, 2 * FG),
                  q3 = __shfl_sync(FULL, quota, 3 * FG);
        const int o1 = (q0 + 3) >> 2, o2 = o1 + ((q1 + 3) >> 2), o3 = o2 + ((q2 + 3) >> 2), nquad = o3 + ((q3 + 3) >> 2);
        float ax[4] = {0.f, 0.f, 0.f, 0.f}, ay[4] = {0.f, 0.f, 0.f, 0.f};
        for (int w = lane; w - lane < nquad; w += 32) {  // warp-uniform trip count (usually 1)
            const int gq = (w >= o1) + (w >= o2) + (w >= o3);
            const float zx = __shfl_sync(FULL, zi.x, gq * FG), zy = __shfl_sync(FULL, zi.y, gq * FG);
            const uint32_t gj = (uint32_t)(p.row0 + rb) + (uint32_t)gq;  // global row of the quad (indices are int32)
            const int qg = gq == 0 ? q0 : gq == 1 ? q1 : gq == 2 ? q2 : q3;
            const int quad = w - (gq == 0 ? 0 : gq == 1 ? o1 : gq == 2 ? o2 : o3);
            const int nval = w < nquad ? min(4, qg - 4 * quad) : 0;
            uint32_t j[4];
            if (p.neg) {
                const int64_t* nr = p.neg + (rb + gq) * p.n_neg + 4 * quad;
#pragma unroll
                for (int u = 0; u < 4; ++u) j[u] = (u < nval) ? (uint32_t)__ldg(nr + u) : gj;
            } else {
                const uint4 wd = rng(c0, c1, gj, (uint32_t)quad);
                j[0] = __umulhi(wd.x, nm1); j[1] = __umulhi(wd.y, nm1);
                j[2] = __umulhi(wd.z, nm1); j[3] = __umulhi(wd.w, nm1);
#pragma unroll
                for (int u = 0; u < 4; ++u) j[u] += (j[u] >= gj) ? 1u : 0u;  // uniform on [0, N-1] \ {i}: NE base.py:636
            }
            float2 zn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) zn[u] = (u < nval) ? __ldg(p.Zin + j[u]) : make_float2(zx, zy);
            float sx = 0.0f, sy = 0.0f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dx = __fsub_rn(zx, zn[u].x), dy = __fsub_rn(zy, zn[u].y);
                const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                const float den = __fadd_rn(1.0f, __fmul_rn(p.a, pow_fast(D, p.b)));  // umap.py:273
                float coef = __fmul_rn(rcp_fast(__fmul_rn(__fadd_rn(D, 1e-3f), den)), p.neg_two_b);  // :274-276
                coef = (u < nval) ? coef : 0.0f;
                sx = fmaf(dx, coef, sx);
                sy = fmaf(dy, coef, sy);
            }
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
                ax[gg] += (gq == gg) ? sx : 0.0f;
                ay[gg] += (gq == gg) ? sy : 0.0f;
            }
        }
        // transposed reduction: 4 groups x 32 lanes -> lane L ends with the total of group L >> 3
        const bool up = lane & 16, odd = lane & 8;
        float rx, ry;
        {
            float k0 = up ? ax[2] : ax[0], k1 = up ? ax[3] : ax[1];
            k0 += __shfl_xor_sync(FULL, up ? ax[0] : ax[2], 16);
            k1 += __shfl_xor_sync(FULL, up ? ax[1] : ax[3], 16);
            rx = (odd ? k1 : k0) + __shfl_xor_sync(FULL, odd ? k0 : k1, 8);
            float m0 = up ? ay[2] : ay[0], m1 = up ? ay[3] : ay[1];
            m0 += __shfl_xor_sync(FULL, up ? ay[0] : ay[2], 16);
            m1 += __shfl_xor_sync(FULL, up ? ay[1] : ay[3], 16);
            ry = (odd ? m1 : m0) + __shfl_xor_sync(FULL, odd ? m0 : m1, 8);
        }
#pragma unroll
        for (int o = FG / 2; o > 0; o >>= 1) {
            rx += __shfl_xor_sync(FULL, rx, o);
            ry += __shfl_xor_sync(FULL, ry, o);
        }
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
        // the warp's 4 rows are consecutive: lanes 0-3 write ONE contiguous 32-byte segment per destination
        const float ox = __shfl_sync(FULL, zo.x, (lane & 3) * FG), oy = __shfl_sync(FULL, zo.y, (lane & 3) * FG);
        if (lane < 4 && rb + lane < p.n_local) store_row(p, p.row0 + rb + lane, make_float2(ox, oy));
    }
    block_flush(l == 0, gn_local, saw_nan, n_act, n_neg_used, p);
}

}  // namespace tdr
