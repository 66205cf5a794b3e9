// Throughput kernel of the UMAP step: one lane per row, 32 consecutive rows pooled per warp iteration
// (included by umap_step.cu after umap_step_fast3.cuh; the default for precise = 0).
//
// umap_step_kernel_fast3 (8 rows per warp iteration, 4 owner lanes per row) still spends ~45 % of its
// 237 M warp instructions per iteration outside the per-edge / per-negative arithmetic: row set-up, offset
// shuffles, row look-ups (7 compares per scanned edge and per quad), prefix sums, owner-lane butterflies
// and the epilogue are paid once per 8 rows, and 74 +- 8 quads fill three 32-lane passes to 77 %.
// Here a warp iteration covers 32 rows and every lane OWNS one of them:
//   scan     lanes take consecutive edges of the pooled CSR range; due edges are appended in edge order to
//            the warp's shared-memory list as (col, edge offset) — no row look-up, no atomics;
//   owner    each lane finds its row's slice of the list by binary search on the edge offsets (the list is
//            sorted) and writes its z_i next to its entries;
//   attract  lanes take consecutive entries (32 per pass, 229 +- 15 entries -> 8 passes at 90 %), read
//            (col, z_i) from the list, overwrite the entry with the contribution (c dx, c dy);
//   quads    each lane writes a descriptor (global row, quad index, valid count) + z_i for each of its
//            negative quads; lanes take consecutive descriptors (298 +- 16 quads -> 10 passes at 93 %);
//   sums     each lane adds its row's entries sequentially, in entry order.
// A row's sums are therefore a fixed left-to-right chain over its own edges / quads: results are independent
// of warp grouping, sharding and list-overflow rounds (bit-identical single- vs multi-GPU).
// Arithmetic per edge / negative is that of umap_step_kernel_fast3 (umap_step_math.cuh).
// Measured alternatives (profiles/r1_step_kernel.md): issuing the gathers of pass k+1 before the arithmetic of
// pass k (software pipelining) was 2-3 % slower; a 384-entry list at 4 CTAs/SM (192 KB of shared memory, 28 KB
// of L1 left) was 50 % slower than at 3 CTAs/SM.
#pragma once

namespace tdr {

constexpr int kWarps4 = kFastThreads / 32;

// kCap4 = list entries per warp (2 x 8 B each): 384 -> 6 KB per warp, 48 KB per CTA; 256 -> 32 KB per CTA
template <int kCap4>
struct Warp4Smem {
    int2 a[kCap4];    // (col, edge offset) | (global row, quad | nval << 28), then float2 contributions
    float2 b[kCap4];  // z_i of the entry's row
};

// NEG_CG: gather the (uniformly random) negatives with ld.global.cg so they do not evict the neighbour
// rows and edge streams from L1
// L2H: L2 eviction hints (umap_step_math.cuh) — edge streams evict-first, gathered rows of Z evict-last
// PF: persistent grid (the host launches one wave of CTAs; every warp walks several 32-row blocks) and, while a block's
//     attraction runs, L2 prefetch of the NEXT block's edge streams — the ncu capture of the default configuration
//     shows 31 % long-scoreboard stalls, the largest on the compare that consumes the epoch_of_next_sample stream
//     (profiles/r1_step_fast4_hotspots.txt).  EXPERIMENTAL (TDR_STEP_CFG=6), not yet measured.
// CHEAP: Philox4x32-7 instead of -10 and pow_cheap instead of pow_fast (fewer issue slots per negative).  EXPERIMENTAL
//     (TDR_STEP_CFG=7), not yet measured; changes the in-kernel negative stream, not its distribution.
template <int MIN_CTAS, int kCap4, bool NEG_CG, bool L2H = false, bool PF = false, bool CHEAP = false>
__global__ void __launch_bounds__(kFastThreads, MIN_CTAS) umap_step_kernel_fast4(const UmapStepParams p) {
    extern __shared__ __align__(16) unsigned char s_raw4[];
    constexpr unsigned FULL = 0xffffffffu;
    Warp4Smem<kCap4>& sm = reinterpret_cast<Warp4Smem<kCap4>*>(s_raw4)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t warp_global = ((int64_t)blockIdx.x * kFastThreads + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t)gridDim.x * kWarps4;
    const float due_before = (float)(p.n_iter + 1);  // umap.py:251
    const uint64_t pol_keep = L2H ? l2_policy_evict_last() : 0, pol_stream = L2H ? l2_policy_evict_first() : 0;
    const Philox rng(p.seed);
    const uint32_t nm1 = (uint32_t)(p.n_total - 1);
    const uint32_t c0 = (uint32_t)p.n_iter, c1 = (uint32_t)(p.n_iter >> 32);
    double gn_local = 0.0;
    bool saw_nan = false;
    unsigned long long n_act = 0, n_neg_used = 0;

    for (int64_t rb = warp_global * 32; rb < p.n_local; rb += n_warps * 32) {
        const int64_t r = rb + lane;
        const bool live = r < p.n_local;
        const int gi = (int)(p.row0 + (live ? r : rb));  // global row owned by this lane (indices are int32)
        const float2 zi = __ldg(p.Zin + gi);
        const int64_t rp = __ldg(p.rowptr + min(r, p.n_local)), rp_next = __ldg(p.rowptr + min(r + 1, p.n_local));
        const int64_t E0 = __shfl_sync(FULL, (long long)rp, 0);
        const int off = (int)(rp - E0), off_next = (int)(rp_next - E0);  // my row's slice of the pooled edge range
        const int total = __shfl_sync(FULL, off_next, 31);
        float* const eons_w = p.eons + E0;
        const float* const eps_w = p.eps + E0;
        const int32_t* const col_w = p.col + E0;

        if (PF) {
            const int64_t rbn = rb + n_warps * 32;
            if (rbn < p.n_local) {
                const int64_t e0 = __ldg(p.rowptr + rbn), e1 = __ldg(p.rowptr + min(rbn + 32, p.n_local));
                for (int64_t o = e0 + lane * 32; o < e1; o += 32 * 32) {  // one 128-byte line per lane and array
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.eons + o));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.col + o));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.eps + o));
                }
            }
        }
        // ---- attraction (umap.py:236-264)
        float gx = 0.0f, gy = 0.0f;
        int active = 0, nd = 0;
        // evaluates the nd listed edges and adds each row's contributions to its owner lane
        auto flush_attraction = [&]() {
            __syncwarp();
            int es = 0;
            {   // lower bound of my row's first edge offset among the (sorted) listed offsets
                int hi = nd;
                while (es < hi) {
                    const int mid = (es + hi) >> 1;
                    if (sm.a[mid].y < off) es = mid + 1; else hi = mid;
                }
            }
            int ee = __shfl_down_sync(FULL, es, 1);
            if (lane == 31) ee = nd;
            if (off_next == off) ee = es;  // empty row
            for (int e = es; e < ee; ++e) sm.b[e] = zi;
            __syncwarp();
            struct EdgeIn { float2 z, zj; bool valid; };
            // idle lanes evaluate their own point against itself (D = 0 -> coefficient 0) and store nothing
            auto fetch_edge = [&](int t) {
                EdgeIn in;
                in.valid = t < nd;
                const int cj = in.valid ? sm.a[t].x : gi;
                in.z = in.valid ? sm.b[t] : zi;
                in.zj = L2H ? ldg_f2_hint(p.Zin + cj, pol_keep) : __ldg(p.Zin + cj);
                return in;
            };
            auto eval_edge = [&](const EdgeIn& in, int t) {
                const float dx = __fsub_rn(in.z.x, in.zj.x), dy = __fsub_rn(in.z.y, in.zj.y);
                const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));  // distance/base.py:384-385
                const float pw = CHEAP ? pow_cheap(D, p.bm1) : pow_fast(D, p.bm1);  // D^(b-1); D^b = D * D^(b-1)
                const float den = __fadd_rn(1.0f, __fmul_rn(p.a, __fmul_rn(pw, D)));
                float coef = __fmul_rn(__fmul_rn(pw, p.two_ab), rcp_fast(den));
                coef = (D > 0.0f) ? coef : 0.0f;  // umap.py:243-247
                if (in.valid) sm.a[t] = make_int2(__float_as_int(__fmul_rn(dx, coef)), __float_as_int(__fmul_rn(dy, coef)));
            };
            for (int t = lane; t - lane < nd; t += 32) eval_edge(fetch_edge(t), t);
            __syncwarp();
            for (int e = es; e < ee; ++e) {
                const int2 v = sm.a[e];
                gx += __int_as_float(v.x);
                gy += __int_as_float(v.y);
            }
            active += ee - es;
            nd = 0;
            __syncwarp();
        };
        int cb = 0;
        do {
            // scan up to kCap4 listed edges (warp-uniform); more than one round only in dense early iterations / hub rows
            while (cb < total && nd + 32 * FU <= kCap4) {
                float nxt[FU], ep[FU];
                int cj[FU];
#pragma unroll
                for (int u = 0; u < FU; ++u) {
                    const int c = cb + u * 32 + lane;
                    const bool ok = c < total;
                    if (L2H) {
                        nxt[u] = ok ? ld_f32_hint(eons_w + c, pol_stream) : INFINITY;
                        cj[u] = ok ? ldg_s32_hint(col_w + c, pol_stream) : 0;
                        ep[u] = ok ? ldg_f32_hint(eps_w + c, pol_stream) : 0.0f;
                    } else {
                        nxt[u] = ok ? eons_w[c] : INFINITY;
                        cj[u] = ok ? __ldg(col_w + c) : 0;
                        ep[u] = ok ? __ldg(eps_w + c) : 0.0f;
                    }
                }
#pragma unroll
                for (int u = 0; u < FU; ++u) {
                    if (cb + u * 32 < total) {  // warp-uniform
                        const int c = cb + u * 32 + lane;
                        const bool due = nxt[u] <= due_before;
                        const unsigned bal = __ballot_sync(FULL, due);
                        if (due) {
                            eons_w[c] = __fadd_rn(nxt[u], ep[u]);  // umap.py:253-255
                            sm.a[nd + __popc(bal & lt_mask)] = make_int2(cj[u], c);
                        }
                        nd += __popc(bal);
                    }
                }
                cb += 32 * FU;
            }
            flush_attraction();
        } while (cb < total);
        gx = fminf(fmaxf(gx, -4.0f), 4.0f);  // umap.py:263
        gy = fminf(fmaxf(gy, -4.0f), 4.0f);

        // ---- repulsion (umap.py:266-292) on the first rate*active negatives of every row; a work item is a
        // quad (row, 4 consecutive negative slots = one Philox block)
        int quota = active * p.rate;
        if (quota > p.n_neg) quota = p.n_neg;
        if (!live) quota = 0;
        const int nq = (quota + 3) >> 2;
        int incl = nq;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += v;
        }
        const int qo = incl - nq, nquad = __shfl_sync(FULL, incl, 31);
        float rx = 0.0f, ry = 0.0f;
        for (int wb = 0; wb < nquad; wb += kCap4) {  // one round unless the rows hold > kCap4 quads
            const int w1 = min(nquad, wb + kCap4);
            const int lo = max(qo, wb), hi = min(qo + nq, w1);
            for (int w = lo; w < hi; ++w) {
                const int q = w - qo;
                sm.a[w - wb] = make_int2(gi, q | (min(4, quota - 4 * q) << 28));
                sm.b[w - wb] = zi;
            }
            __syncwarp();
            struct QuadIn { float2 z, zn[4]; bool valid; };
            auto fetch_quad = [&](int w) {
                QuadIn in;
                in.valid = w < w1;
                const int2 ds = in.valid ? sm.a[w - wb] : make_int2(gi, 0);
                in.z = in.valid ? sm.b[w - wb] : zi;
                const uint32_t gj = (uint32_t)ds.x;
                const int quad = ds.y & 0x0fffffff, nval = ds.y >> 28;
                uint32_t jn[4];
                if (p.neg) {
                    const int64_t* nr = p.neg + ((int64_t)gj - p.row0) * p.n_neg + 4 * quad;
#pragma unroll
                    for (int u = 0; u < 4; ++u) jn[u] = (u < nval) ? (uint32_t)__ldg(nr + u) : gj;
                } else {
                    const uint4 wd = CHEAP ? rng.rounds<7>(c0, c1, gj, (uint32_t)quad) : rng(c0, c1, gj, (uint32_t)quad);
                    jn[0] = __umulhi(wd.x, nm1); jn[1] = __umulhi(wd.y, nm1);
                    jn[2] = __umulhi(wd.z, nm1); jn[3] = __umulhi(wd.w, nm1);
#pragma unroll
                    for (int u = 0; u < 4; ++u) jn[u] += (jn[u] >= gj) ? 1u : 0u;  // uniform on [0, N-1] \ {i}: NE base.py:636
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)  // unused slot: dx = dy = 0
                    in.zn[u] = (u < nval) ? (L2H ? ldcg_f2_hint(p.Zin + jn[u], pol_keep)
                                                 : (NEG_CG ? __ldcg(p.Zin + jn[u]) : __ldg(p.Zin + jn[u])))
                                          : in.z;
                return in;
            };
            auto eval_quad = [&](const QuadIn& in, int w) {
                float sx = 0.0f, sy = 0.0f;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float dx = __fsub_rn(in.z.x, in.zn[u].x), dy = __fsub_rn(in.z.y, in.zn[u].y);
                    const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                    const float den = __fadd_rn(1.0f, __fmul_rn(p.a, CHEAP ? pow_cheap(D, p.b) : pow_fast(D, p.b)));  // umap.py:273
                    const float coef = __fmul_rn(rcp_fast(__fmul_rn(__fadd_rn(D, 1e-3f), den)), p.neg_two_b);  // :274-276
                    sx = fmaf(dx, coef, sx);
                    sy = fmaf(dy, coef, sy);
                }
                if (in.valid) sm.a[w - wb] = make_int2(__float_as_int(sx), __float_as_int(sy));
            };
            for (int w = wb + lane; w - lane < w1; w += 32) eval_quad(fetch_quad(w), w);
            __syncwarp();
            for (int w = lo; w < hi; ++w) {
                const int2 v = sm.a[w - wb];
                rx += __int_as_float(v.x);
                ry += __int_as_float(v.y);
            }
            __syncwarp();
        }
        rx = fminf(fmaxf(rx, -4.0f), 4.0f);  // umap.py:291
        ry = fminf(fmaxf(ry, -4.0f), 4.0f);
        const float g0 = __fadd_rn(__fmul_rn(p.lam, gx), __fmul_rn(p.rep, rx));  // NE base.py:237-241
        const float g1 = __fadd_rn(__fmul_rn(p.lam, gy), __fmul_rn(p.rep, ry));
        float2 zo;  // torch.optim.SGD: param.add_(grad, alpha=-lr)
        zo.x = fmaf(-p.lr, g0, zi.x);
        zo.y = fmaf(-p.lr, g1, zi.y);
        if (live) {
            store_row(p, gi, zo);  // 32 consecutive rows: one contiguous 256-byte segment per destination
            if (p.grad_out) p.grad_out[r] = make_float2(g0, g1);
            gn_local += (double)g0 * g0 + (double)g1 * g1;
            saw_nan |= (zo.x != zo.x) || (zo.y != zo.y);
            n_act += active;
            n_neg_used += quota;
        }
    }
    block_flush(true, gn_local, saw_nan, n_act, n_neg_used, p);
}

}  // namespace tdr
