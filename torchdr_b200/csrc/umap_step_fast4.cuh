// Throughput path of the UMAP step: one lane per row, 32 consecutive rows pooled per warp iteration
// (included by umap_step.cu; the default for precise = 0).
//
// A warp iteration covers 32 rows and every lane OWNS one of them:
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
// of warp grouping, block scheduling, sharding and list-overflow rounds (bit-identical single- vs multi-GPU,
// per-iteration launches vs the persistent loop).
//
// Two kernels share the block routine:
//   umap_step_kernel_fast4    one launch = one iteration (tdr_umap_step_f32: hooks, injected negatives, grad_out)
//   umap_run_kernel_persist   ONE cooperative launch = many iterations (tdr_umap_run_f32 / tdr_umap_run_p2p_f32):
//                             warps take 32-row blocks from a per-iteration work counter, the iterations are
//                             separated by a grid barrier on a device word, and — row-sharded over several GPUs —
//                             the last CTA to arrive also runs the cross-GPU exchange barrier on NVLink peer-mapped
//                             flag words, so the per-iteration collective of affinity_matcher.py:395-413 is entirely
//                             inside the kernel (updated rows go to the peers as plain stores from the step itself).
// Measured alternatives (profiles/r1_step_kernel.md, profiles/r2_step_variants.md): software pipelining of the
// gathers 2-3 % slower; 384-entry lists at 4 CTAs/SM 50 % slower; L2 eviction hints +3 % at 10 M / -4 % at 1 M;
// L2 prefetch of the next block's edge streams -3 % (1 M) / -15 % (10 M); Philox4x32-7 + 3-instruction pow +4 %
// (adopted).
#pragma once

namespace tdr {

constexpr int kWarps4 = kFastThreads / 32;
constexpr int kCap4 = 256;  // list entries per warp (2 x 8 B each): 4 KB per warp, 32 KB per CTA
constexpr int kOcc4 = 4;    // CTAs per SM (64 registers)

struct Warp4Smem {
    int2 a[kCap4];    // (col, edge offset) | (global row, quad | nval << 28), then float2 contributions
    float2 b[kCap4];  // z_i of the entry's row
};

// what changes from iteration to iteration
struct IterArgs {
    const float2* Zin;
    float2* Zout;
    float2* const* peer_out;  // n_peers device pointers (this iteration's destination buffer on every peer)
    int64_t n_iter;
    float lr;
};

struct Accum {
    double gn = 0.0;
    bool saw_nan = false;
    unsigned long long n_act = 0, n_neg = 0;
};

// COH: the embedding buffers are rewritten between iterations of the same launch (persistent loop), so rows of Z are
// read with ordinary (coherent) loads that the barrier's fences order, never with ld.global.nc.
template <bool COH>
__device__ __forceinline__ float2 ld_z(const float2* a) {
    if (COH) return *a;
    return __ldg(a);
}

template <bool COH>
__device__ __forceinline__ void step_block32(const UmapStepParams& p, const IterArgs& it, Warp4Smem& sm, const int lane,
                                             const int64_t rb, Accum& acc) {
    constexpr unsigned FULL = 0xffffffffu;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float due_before = (float)(it.n_iter + 1);  // umap.py:251
    const Philox rng(p.seed);
    const uint32_t nm1 = (uint32_t)(p.n_total - 1);
    const uint32_t c0 = (uint32_t)it.n_iter, c1 = (uint32_t)(it.n_iter >> 32);

    const int64_t r = rb + lane;
    const bool live = r < p.n_local;
    const int gi = (int)(p.row0 + (live ? r : rb));  // global row owned by this lane (indices are int32)
    const float2 zi = ld_z<COH>(it.Zin + gi);
    const int64_t rp = __ldg(p.rowptr + min(r, p.n_local)), rp_next = __ldg(p.rowptr + min(r + 1, p.n_local));
    const int64_t E0 = __shfl_sync(FULL, (long long)rp, 0);
    const int off = (int)(rp - E0), off_next = (int)(rp_next - E0);  // my row's slice of the pooled edge range
    const int total = __shfl_sync(FULL, off_next, 31);
    float* const eons_w = p.eons + E0;
    const float* const eps_w = p.eps + E0;
    const int32_t* const col_w = p.col + E0;

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
        // idle lanes evaluate their own point against itself (D = 0 -> coefficient 0) and store nothing
        for (int t = lane; t - lane < nd; t += 32) {
            const bool valid = t < nd;
            const int cj = valid ? sm.a[t].x : gi;
            const float2 z = valid ? sm.b[t] : zi;
            const float2 zj = ld_z<COH>(it.Zin + cj);
            const float dx = __fsub_rn(z.x, zj.x), dy = __fsub_rn(z.y, zj.y);
            const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));  // distance/base.py:384-385
            const float pw = pow_cheap(D, p.bm1);                              // D^(b-1); D^b = D * D^(b-1)
            const float den = __fadd_rn(1.0f, __fmul_rn(p.a, __fmul_rn(pw, D)));
            float coef = __fmul_rn(__fmul_rn(pw, p.two_ab), rcp_fast(den));
            coef = (D > 0.0f) ? coef : 0.0f;  // umap.py:243-247
            if (valid) sm.a[t] = make_int2(__float_as_int(__fmul_rn(dx, coef)), __float_as_int(__fmul_rn(dy, coef)));
        }
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
                nxt[u] = ok ? eons_w[c] : INFINITY;
                cj[u] = ok ? __ldg(col_w + c) : 0;
                ep[u] = ok ? __ldg(eps_w + c) : 0.0f;
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
        for (int w = wb + lane; w - lane < w1; w += 32) {
            const bool valid = w < w1;
            const int2 ds = valid ? sm.a[w - wb] : make_int2(gi, 0);
            const float2 z = valid ? sm.b[w - wb] : zi;
            const uint32_t gj = (uint32_t)ds.x;
            const int quad = ds.y & 0x0fffffff, nval = ds.y >> 28;
            uint32_t jn[4];
            if (p.neg) {
                const int64_t* nr = p.neg + ((int64_t)gj - p.row0) * p.n_neg + 4 * quad;
#pragma unroll
                for (int u = 0; u < 4; ++u) jn[u] = (u < nval) ? (uint32_t)__ldg(nr + u) : gj;
            } else {
                // Philox4x32-7 (the smallest round count of Salmon et al. that passes BigCrush): counter = (iteration, row, quad)
                const uint4 wd = rng.rounds<7>(c0, c1, gj, (uint32_t)quad);
                jn[0] = __umulhi(wd.x, nm1); jn[1] = __umulhi(wd.y, nm1);
                jn[2] = __umulhi(wd.z, nm1); jn[3] = __umulhi(wd.w, nm1);
#pragma unroll
                for (int u = 0; u < 4; ++u) jn[u] += (jn[u] >= gj) ? 1u : 0u;  // uniform on [0, N-1] \ {i}: NE base.py:636
            }
            float2 zn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)  // unused slot: dx = dy = 0; negatives go through L2 only (no L1 pollution)
                zn[u] = (u < nval) ? __ldcg(it.Zin + jn[u]) : z;
            float sx = 0.0f, sy = 0.0f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dx = __fsub_rn(z.x, zn[u].x), dy = __fsub_rn(z.y, zn[u].y);
                const float D = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                const float den = __fadd_rn(1.0f, __fmul_rn(p.a, pow_cheap(D, p.b)));                        // umap.py:273
                const float coef = __fmul_rn(rcp_fast(__fmul_rn(__fadd_rn(D, 1e-3f), den)), p.neg_two_b);  // :274-276
                sx = fmaf(dx, coef, sx);
                sy = fmaf(dy, coef, sy);
            }
            if (valid) sm.a[w - wb] = make_int2(__float_as_int(sx), __float_as_int(sy));
        }
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
    zo.x = fmaf(-it.lr, g0, zi.x);
    zo.y = fmaf(-it.lr, g1, zi.y);
    if (live) {
        // 32 consecutive rows: one contiguous 256-byte segment per destination (own buffer + every NVLink peer)
        it.Zout[gi] = zo;
        if (p.n_peers != 0) {
#pragma unroll
            for (int q = 0; q < kMaxPeers; ++q)
                if (q < p.n_peers) it.peer_out[q][gi] = zo;
        }
        if (p.grad_out) p.grad_out[r] = make_float2(g0, g1);
        acc.gn += (double)g0 * g0 + (double)g1 * g1;
        acc.saw_nan |= (zo.x != zo.x) || (zo.y != zo.y);
        acc.n_act += active;
        acc.n_neg += quota;
    }
}

// ---- one launch = one iteration -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFastThreads, kOcc4) umap_step_kernel_fast4(const UmapStepParams p) {
    extern __shared__ __align__(16) unsigned char s_raw4[];
    Warp4Smem& sm = reinterpret_cast<Warp4Smem*>(s_raw4)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * kFastThreads + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t)gridDim.x * kWarps4;
    IterArgs it;
    it.Zin = p.Zin;
    it.Zout = p.Zout;
    it.peer_out = p.peer_out;
    it.n_iter = p.n_iter;
    it.lr = p.lr;
    Accum acc;
    for (int64_t rb = warp_global * 32; rb < p.n_local; rb += n_warps * 32) step_block32<false>(p, it, sm, lane, rb, acc);
    block_flush(true, acc.gn, acc.saw_nan, acc.n_act, acc.n_neg, p, p.gnorm_sq);
}

// ---- one launch = many iterations ---------------------------------------------------------------------------------
// Device words shared by the CTAs of one rank (caller-owned, zero-initialised once, reused by every launch):
//   [0] arrive   CTAs arrived at the current barrier (reset by the last one)
//   [1] go       number of barriers completed so far (monotonic over the life of the buffer) | kAbort
//   [2],[3]      work counters of even / odd iterations (next 32-row block to hand out; reset at the barrier)
//   [4] status   0 = ok, 1 = a peer did not arrive within the time limit, 2 = a local CTA did not arrive
constexpr uint32_t kAbort = 0xffffffffu;

struct RunParams {
    float* Z[2];                      // this rank's two embedding buffers; iteration t reads Z[(cur0 + t) & 1]
    float2* peer_Z[2][kMaxPeers];     // the peers' addresses of the same two buffers
    uint32_t* peer_flags[kMaxPeers];  // peer q's flag buffer (this rank writes word [rank])
    uint32_t* my_flags;               // this rank's flag buffer: word [r] = barriers announced by rank r
    uint32_t* sync;                   // the five words above
    int peer_rank[kMaxPeers];
    int rank, n_steps, gnorm_last;
    uint32_t epoch0;                  // barriers already completed on sync / flags before this launch
    unsigned long long timeout_ns;
    float lrs[kMaxRunSteps];
};

__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* a) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* a) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* a, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(uint32_t* a, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_acq_rel_gpu(uint32_t* a, uint32_t v) {
    uint32_t old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(a), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Barrier between two iterations, executed by thread 0 of every CTA (the CTA's other threads wait at __syncthreads).
// `epoch` = number of barriers completed once this one is.  Returns false if the run was aborted.
// Release side: the CTA barrier ordered every store of this CTA (own buffer, NVLink peers, epoch_of_next_sample) before
// this point; one fence makes them visible at the scope the consumers read from (system scope when rows also went to
// peers), then the arrival is counted.  Acquire side: the flag words are polled with RELAXED loads (a spinning
// ld.acquire would issue a fence + L1 invalidation per poll from every CTA) and ONE fence after the wait orders the
// next iteration's loads behind it and drops stale L1 lines of the embedding buffers.
__device__ bool run_barrier(const UmapStepParams& p, const RunParams& rp, uint32_t epoch, uint32_t* work_used) {
    uint32_t* const arrive = rp.sync;
    uint32_t* const go = rp.sync + 1;
    uint32_t* const status = rp.sync + 4;
    const bool multi = p.n_peers != 0;
    if (multi) __threadfence_system(); else __threadfence();
    const uint32_t old = atom_add_acq_rel_gpu(arrive, 1u);
    const unsigned long long t0 = global_ns();
    bool ok = true;
    if (old == gridDim.x - 1) {  // last CTA of this rank
        *arrive = 0u;
        *work_used = 0u;
        if (multi) {
            // release: ONE system-scope fence orders this rank's stores (the other CTAs' were observed through the
            // arrival counter) before the flag stores, which can then be relaxed
            __threadfence_system();
            for (int q = 0; q < p.n_peers; ++q) st_relaxed_sys(rp.peer_flags[q] + rp.rank, epoch);
            for (int q = 0; q < p.n_peers && ok; ++q) {
                const uint32_t* f = rp.my_flags + rp.peer_rank[q];
                unsigned spins = 0;
                while ((int32_t)(ld_relaxed_sys(f) - epoch) < 0) {
                    if ((++spins & 1023u) == 0 && global_ns() - t0 > rp.timeout_ns) {
                        ok = false;
                        break;
                    }
                }
            }
            __threadfence_system();
        }
        if (!ok) *status = 1u;
        st_release_gpu(go, ok ? epoch : kAbort);
    } else {
        unsigned spins = 0;
        for (;;) {
            const uint32_t v = ld_relaxed_gpu(go);
            if (v == kAbort) {
                ok = false;
                break;
            }
            if ((int32_t)(v - epoch) >= 0) break;
            __nanosleep(40);
            if ((++spins & 1023u) == 0 && global_ns() - t0 > 2 * rp.timeout_ns) {
                *status = 2u;
                st_release_gpu(go, kAbort);
                ok = false;
                break;
            }
        }
    }
    if (multi) __threadfence_system(); else __threadfence();
    return ok;
}

__global__ void __launch_bounds__(kFastThreads, kOcc4)
umap_run_kernel_persist(const UmapStepParams p, const __grid_constant__ RunParams rp, const int cur0) {
    extern __shared__ __align__(16) unsigned char s_raw4[];
    __shared__ int s_alive;
    Warp4Smem& sm = reinterpret_cast<Warp4Smem*>(s_raw4)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int64_t n_blocks = (p.n_local + 31) >> 5;
    for (int t = 0; t < rp.n_steps; ++t) {
        const int src = (cur0 + t) & 1;
        IterArgs it;
        it.Zin = reinterpret_cast<const float2*>(rp.Z[src]);
        it.Zout = reinterpret_cast<float2*>(rp.Z[src ^ 1]);
        it.peer_out = rp.peer_Z[src ^ 1];
        it.n_iter = p.n_iter + t;
        it.lr = rp.lrs[t];
        uint32_t* const work = rp.sync + 2 + (t & 1);
        Accum acc;
        // 32-row blocks are handed out dynamically (the hardware CTA scheduler did this for the per-iteration launch);
        // the next ticket is requested before the current block is processed so its latency is hidden
        uint32_t blk = 0;
        if (lane == 0) blk = atomicAdd(work, 1u);
        blk = __shfl_sync(0xffffffffu, blk, 0);
        while ((int64_t)blk < n_blocks) {
            uint32_t nxt = 0;
            if (lane == 0) nxt = atomicAdd(work, 1u);
            step_block32<true>(p, it, sm, lane, (int64_t)blk << 5, acc);
            blk = __shfl_sync(0xffffffffu, nxt, 0);
        }
        const bool last = t == rp.n_steps - 1;
        block_flush(true, acc.gn, acc.saw_nan, acc.n_act, acc.n_neg, p, (last && rp.gnorm_last) ? p.gnorm_sq : nullptr);
        __syncthreads();
        if (threadIdx.x == 0) s_alive = run_barrier(p, rp, rp.epoch0 + (uint32_t)t + 1u, work) ? 1 : 0;
        __syncthreads();
        if (!s_alive) return;
    }
}

}  // namespace tdr
