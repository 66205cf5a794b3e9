// Pooled-row throughput variant of the UMAP step, 8 rows per warp iteration (included by umap_step.cu;
// selected with TDR_STEP_FAST=3, kept as the measured predecessor of umap_step_fast4.cuh).
//
// The first throughput kernels gave 8 lanes to a row, so every loop ran for the LARGEST of the warp's rows
// (degree 25 +- 12, 7 +- 2.3 due edges, 37 +- 6 negative quads per 4 rows) and the per-row register
// accumulators cost 16 select+add per pass plus transposed reductions: 286 M warp instructions per
// iteration at 1 M points although the useful work is about a third of that.
// This version treats the warp's rows as ONE pool in every phase (8 consecutive rows per warp iteration):
//   scan   the rows' CSR segments are contiguous: lanes take consecutive edges of the pooled range
//          (row id = number of row offsets <= edge offset), due edges are appended in edge order to a
//          per-warp shared-memory list, per-row due counts by shared-memory integer atomics;
//   attract lanes take consecutive list entries (32 per pass, ~90 % busy), z_i by shuffle from the row's
//          owner lane, the contribution (c dx, c dy) overwrites the list entry;
//   repulse quads (row, 4 negative slots = one Philox block) are dealt the same way, results to the list;
//   sums   the 4 owner lanes of a row add the row's entries (stride 4) and finish with two butterflies.
// The order of a row's sum depends only on the row's own counts, so results do not depend on how rows
// are grouped into warps or sharded over GPUs, except when the pool overflows the list (kCap3 entries)
// and is flushed in several rounds (hub rows).  237 M warp instructions, 0.270 ms per iteration.
#pragma once

namespace tdr {

constexpr int kRows3 = 8;              // rows per warp iteration
constexpr int kOwn3 = 32 / kRows3;     // owner lanes per row
constexpr int kCap3 = 512;             // list entries per warp (8 B each)
constexpr int kWarps3 = kFastThreads / 32;

struct Warp3Smem {
    int2 ent[kCap3];   // (col, row) of a due edge, then reused for float2 contributions
    int cnt[kRows3];   // due edges per row in the current round
};

// sum of the entries [lo, lo + n) of a row over its kOwn3 owner lanes (j = lane & 3); every owner lane
// returns the total
__device__ __forceinline__ float2 row_sum3(const int2* ent, int lo, int n, int j) {
    float sx = 0.0f, sy = 0.0f;
    for (int i = j; i < n; i += kOwn3) {
        const int2 e = ent[lo + i];
        sx += __int_as_float(e.x);
        sy += __int_as_float(e.y);
    }
#pragma unroll
    for (int o = kOwn3 / 2; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    return make_float2(sx, sy);
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(kFastThreads, MIN_CTAS) umap_step_kernel_fast3(const UmapStepParams p) {
    __shared__ Warp3Smem s_all[kWarps3];
    constexpr unsigned FULL = 0xffffffffu;
    Warp3Smem& sm = s_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int own = lane / kOwn3;      // row of the warp's 8 owned by this lane
    const int j = lane % kOwn3;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t warp_global = ((int64_t)blockIdx.x * kFastThreads + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t)gridDim.x * kWarps3;
    const float due_before = (float)(p.n_iter + 1);  // umap.py:251
    const Philox rng(p.seed);
    const uint32_t nm1 = (uint32_t)(p.n_total - 1);
    const uint32_t c0 = (uint32_t)p.n_iter, c1 = (uint32_t)(p.n_iter >> 32);
    double gn_local = 0.0;
    bool saw_nan = false;
    unsigned long long n_act = 0, n_neg_used = 0;

    for (int64_t rb = warp_global * kRows3; rb < p.n_local; rb += n_warps * kRows3) {
        const int64_t r = rb + own;
        const bool live = r < p.n_local;
        const float2 zi = __ldg(p.Zin + p.row0 + (live ? r : rb));
        // row offsets of the pooled edge range: lane k <= 8 loads rowptr[rb + k]
        int64_t rp = 0;
        if (lane <= kRows3) rp = __ldg(p.rowptr + min(rb + lane, p.n_local));
        const int64_t E0 = __shfl_sync(FULL, (long long)rp, 0);
        const int my_off = (int)(rp - E0);
        int o[kRows3 + 1];
#pragma unroll
        for (int k = 1; k <= kRows3; ++k) o[k] = __shfl_sync(FULL, my_off, k);
        const int total = o[kRows3];
        float* const eons_w = p.eons + E0;
        const float* const eps_w = p.eps + E0;
        const int32_t* const col_w = p.col + E0;
        if (lane < kRows3) sm.cnt[lane] = 0;
        __syncwarp();

        // ---- attraction (umap.py:236-264)
        float gx = 0.0f, gy = 0.0f;  // row totals, replicated over the row's owner lanes
        int active = 0;
        for (int set0 = 0; set0 < total; set0 += kCap3) {  // one round unless the pool has > kCap3 edges
            const int set1 = min(total, set0 + kCap3);
            int nd = 0;
            for (int cb = set0; cb < set1; cb += 32 * FU) {
                float nxt[FU], ep[FU];
                int cj[FU];
#pragma unroll
                for (int u = 0; u < FU; ++u) {
                    const int c = cb + u * 32 + lane;
                    const bool ok = c < set1;
                    nxt[u] = ok ? eons_w[c] : INFINITY;
                    cj[u] = ok ? __ldg(col_w + c) : 0;
                    ep[u] = ok ? __ldg(eps_w + c) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < FU; ++u) {
                    if (cb + u * 32 < set1) {  // warp-uniform
                        const int c = cb + u * 32 + lane;
                        const bool due = nxt[u] <= due_before;
                        const unsigned bal = __ballot_sync(FULL, due);
                        if (due) {
                            int row = 0;
#pragma unroll
                            for (int k = 1; k < kRows3; ++k) row += (c >= o[k]) ? 1 : 0;
                            eons_w[c] = __fadd_rn(nxt[u], ep[u]);  // umap.py:253-255
                            sm.ent[nd + __popc(bal & lt_mask)] = make_int2(cj[u], row);
                            atomicAdd(&sm.cnt[row], 1);
                        }
                        nd += __popc(bal);
                    }
                }
            }
            __syncwarp();
            for (int t = lane; t - lane < nd; t += 32) {
                const bool valid = t < nd;
                // idle lanes evaluate the row-0 point against itself (D = 0 -> coefficient 0) and store nothing
                const int2 en = valid ? sm.ent[t] : make_int2((int)(p.row0 + rb), 0);
                const float zx = __shfl_sync(FULL, zi.x, en.y * kOwn3), zy = __shfl_sync(FULL, zi.y, en.y * kOwn3);
                const float2 zj = __ldg(p.Zin + en.x);
                const float dx = __fsub_rn(zx, zj.x), dy = __fsub_rn(zy, zj.y);
                const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));  // distance/base.py:384-385
                const float pw = pow_fast(D, p.bm1);                              // D^(b-1); D^b = D * D^(b-1)
                const float den = __fadd_rn(1.0f, __fmul_rn(p.a, __fmul_rn(pw, D)));
                float coef = __fmul_rn(__fmul_rn(pw, p.two_ab), rcp_fast(den));
                coef = (D > 0.0f) ? coef : 0.0f;  // umap.py:243-247
                if (valid) sm.ent[t] = make_int2(__float_as_int(__fmul_rn(dx, coef)), __float_as_int(__fmul_rn(dy, coef)));
            }
            __syncwarp();
            // per-row sums of this round's contributions (entries are in edge order, so row-sorted)
            int cn = (lane < kRows3) ? sm.cnt[lane] : 0, st = cn;
#pragma unroll
            for (int d = 1; d < kRows3; d <<= 1) {
                const int v = __shfl_up_sync(FULL, st, d);
                if (lane >= d) st += v;
            }
            const int my_cn = __shfl_sync(FULL, cn, own), my_st = __shfl_sync(FULL, st - cn, own);
            const float2 part = row_sum3(sm.ent, my_st, my_cn, j);
            gx += part.x;
            gy += part.y;
            active += my_cn;
            __syncwarp();
            if (lane < kRows3) sm.cnt[lane] = 0;
            __syncwarp();
        }
        gx = fminf(fmaxf(gx, -4.0f), 4.0f);  // umap.py:263
        gy = fminf(fmaxf(gy, -4.0f), 4.0f);

        // ---- repulsion (umap.py:266-292) on the first rate*active negatives of every row; a work item is a
        // quad (row, 4 consecutive negative slots = one Philox block), quads of the 8 rows dealt to the lanes
        int quota = active * p.rate;
        if (quota > p.n_neg) quota = p.n_neg;
        if (!live) quota = 0;
        const int nq_own = (quota + 3) >> 2;
        // exclusive prefix of the rows' quad counts over the owner-leader lanes (lane = row * kOwn3)
        int qo[kRows3 + 1];
        qo[0] = 0;
#pragma unroll
        for (int k = 0; k < kRows3; ++k) qo[k + 1] = qo[k] + __shfl_sync(FULL, nq_own, k * kOwn3);
        const int nquad = qo[kRows3];
        const int my_q0 = [&] { int v = 0;
#pragma unroll
            for (int k = 0; k < kRows3; ++k) v = (own == k) ? qo[k] : v;
            return v; }();
        float rx = 0.0f, ry = 0.0f;
        for (int wb = 0; wb < nquad; wb += kCap3) {  // one round unless n_neg is huge
            const int w1 = min(nquad, wb + kCap3);
            for (int w = wb + lane; w - lane < w1; w += 32) {
                const bool valid = w < w1;
                int gq = 0;
#pragma unroll
                for (int k = 1; k < kRows3; ++k) gq += (w >= qo[k]) ? 1 : 0;
                const float zx = __shfl_sync(FULL, zi.x, gq * kOwn3), zy = __shfl_sync(FULL, zi.y, gq * kOwn3);
                const int qrow = __shfl_sync(FULL, quota, gq * kOwn3);
                const int qst = __shfl_sync(FULL, my_q0, gq * kOwn3);
                const int quad = w - qst;
                const int nval = valid ? min(4, qrow - 4 * quad) : 0;
                const uint32_t gj = (uint32_t)(p.row0 + rb) + (uint32_t)gq;  // global row of the quad (indices are int32)
                uint32_t jn[4];
                if (p.neg) {
                    const int64_t* nr = p.neg + (rb + gq) * p.n_neg + 4 * quad;
#pragma unroll
                    for (int u = 0; u < 4; ++u) jn[u] = (u < nval) ? (uint32_t)__ldg(nr + u) : gj;
                } else {
                    const uint4 wd = rng(c0, c1, gj, (uint32_t)quad);
                    jn[0] = __umulhi(wd.x, nm1); jn[1] = __umulhi(wd.y, nm1);
                    jn[2] = __umulhi(wd.z, nm1); jn[3] = __umulhi(wd.w, nm1);
#pragma unroll
                    for (int u = 0; u < 4; ++u) jn[u] += (jn[u] >= gj) ? 1u : 0u;  // uniform on [0, N-1] \ {i}: NE base.py:636
                }
                float2 zn[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) zn[u] = (u < nval) ? __ldg(p.Zin + jn[u]) : make_float2(zx, zy);
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
                if (valid) sm.ent[w - wb] = make_int2(__float_as_int(sx), __float_as_int(sy));
            }
            __syncwarp();
            const int lo = max(my_q0, wb), hi = min(my_q0 + nq_own, w1);
            const float2 part = row_sum3(sm.ent, lo - wb, max(hi - lo, 0), j);
            rx += part.x;
            ry += part.y;
            __syncwarp();
        }
        rx = fminf(fmaxf(rx, -4.0f), 4.0f);  // umap.py:291
        ry = fminf(fmaxf(ry, -4.0f), 4.0f);
        const float g0 = __fadd_rn(__fmul_rn(p.lam, gx), __fmul_rn(p.rep, rx));  // NE base.py:237-241
        const float g1 = __fadd_rn(__fmul_rn(p.lam, gy), __fmul_rn(p.rep, ry));
        float2 zo;  // torch.optim.SGD: param.add_(grad, alpha=-lr)
        zo.x = fmaf(-p.lr, g0, zi.x);
        zo.y = fmaf(-p.lr, g1, zi.y);
        if (j == 0 && live) {
            if (p.grad_out) p.grad_out[r] = make_float2(g0, g1);
            gn_local += (double)g0 * g0 + (double)g1 * g1;
            saw_nan |= (zo.x != zo.x) || (zo.y != zo.y);
            n_act += active;
            n_neg_used += quota;
        }
        // the warp's 8 rows are consecutive: lanes 0-7 write ONE contiguous 64-byte segment per destination
        const float ox = __shfl_sync(FULL, zo.x, (lane & 7) * kOwn3), oy = __shfl_sync(FULL, zo.y, (lane & 7) * kOwn3);
        if (lane < kRows3 && rb + lane < p.n_local) store_row(p, p.row0 + rb + lane, make_float2(ox, oy));
    }
    block_flush(j == 0, gn_local, saw_nan, n_act, n_neg_used, p);
}

}  // namespace tdr
