// (iii) UMAP optimisation step: sampled attraction + repulsion + clamp + SGD in ONE kernel.
//
// Replaces, per iteration, the ~40 ATen launches of torchdr/neighbor_embedding/umap.py:236-292
// + neighbor_embedding/base.py:617-649 (negative sampling) + affinity_matcher.py:427 (SGD):
//   * graph as CSR (live edges only) instead of the 75 %-padded ELL;
//   * epoch_of_next_sample read-modify-written in place, only for due edges;
//   * negatives drawn in-kernel (Philox4x32-7, counter = iteration / row / quad of 4 slots), or read
//     from the caller's table (the reference's neg_indices_) in parity tests;
//   * Jacobi update: every gradient is computed from Z_in, results go to Z_out.
// umap_step_kernel<true> below is the parity kernel (one warp per row, fp64 pow); the throughput kernels —
// per-iteration launch and the persistent multi-iteration loop with the in-kernel grid / cross-GPU barrier —
// are in umap_step_fast4.cuh.
//
// Arithmetic mirrors the reference op by op (each torch op is one rounding): explicit
// __f*_rn intrinsics keep the compiler from contracting them into FMAs.
#include <stdlib.h>

#include "common.cuh"

namespace tdr {

constexpr int kMaxPeers = 8;       // NVLink peers a rank stores its updated rows to (world <= 9)
constexpr int kMaxRunSteps = 128;  // iterations per persistent launch (learning rates travel as kernel parameters)

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
    float2* peer_out[kMaxPeers];
    int n_peers;
};

__device__ __forceinline__ void store_row(const UmapStepParams& p, int64_t gi, float2 zo) {
    p.Zout[gi] = zo;
    if (p.n_peers == 0) return;  // uniform: single-GPU launches skip the peer loop
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
        if (q < p.n_peers) p.peer_out[q][gi] = zo;
}

// Per-block reduction of the optional diagnostics (gradient norm, NaN flag, sampled-edge counters) in
// shared memory, then one global atomic per block: thousands of same-address global atomics per launch
// would otherwise serialise in L2.
__device__ __forceinline__ void block_flush(bool leader, double gn, bool saw_nan, unsigned long long n_act,
                                            unsigned long long n_neg, const UmapStepParams& p, double* gnorm_sq) {
    if (p.nan_flag && leader && saw_nan) atomicExch(p.nan_flag, 1);  // rare: no reduction needed
    if (!gnorm_sq && !p.stats) return;  // uniform: the common per-iteration case has no block barrier
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
        if (gnorm_sq && s_gn != 0.0) atomicAdd(gnorm_sq, s_gn);
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
                const uint4 u = rng.rounds<7>((uint32_t)p.n_iter, (uint32_t)(p.n_iter >> 32) ^ (uint32_t)(gi >> 32),
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
    block_flush(lane == 0, gn_local, saw_nan, n_act, n_neg_used, p, p.gnorm_sq);
}

}  // namespace tdr
#include "umap_step_math.cuh"
#include "umap_step_fast4.cuh"
namespace tdr {

constexpr size_t kSmem4 = sizeof(Warp4Smem) * kWarps4;

template <typename K>
static cudaError_t allow_smem(K kernel) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem4);
}

// precise: 0 = throughput kernel (umap_step_fast4.cuh), 1 = parity kernel (fp64 pow, one warp per row)
static int launch_step(const UmapStepParams& p, int precise, cudaStream_t st) {
    if (precise == 0) {
        static const cudaError_t attr = allow_smem(umap_step_kernel_fast4);
        TDR_CUDA(attr);
        const int64_t cap = (int64_t)kNumSMs * kOcc4 * 8;  // grid-stride beyond ~8 waves of resident CTAs
        int64_t b4 = (p.n_local + kWarps4 * 32 - 1) / (kWarps4 * 32);
        if (b4 > cap) b4 = cap;
        umap_step_kernel_fast4<<<(unsigned)b4, kFastThreads, kSmem4, st>>>(p);
        TDR_LAUNCH_CHECK();
        return TDR_OK;
    }
    int64_t blocks = (p.n_local + kStepWarps - 1) / kStepWarps;
    const int64_t cap = (int64_t)kNumSMs * 32;
    if (blocks > cap) blocks = cap;
    umap_step_kernel<true><<<(unsigned)blocks, kStepWarps * 32, 0, st>>>(p);
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

// One cooperative launch (all CTAs co-resident: the in-kernel barrier needs it) per <= kMaxRunSteps iterations.
static int launch_persistent(UmapStepParams& p, RunParams& rp, int n_steps, const float* lrs_host, int cur0,
                             int want_gnorm, cudaStream_t st) {
    static const cudaError_t attr = allow_smem(umap_run_kernel_persist);
    TDR_CUDA(attr);
    int dev = 0, sms = 0, occ = 0;
    TDR_CUDA(cudaGetDevice(&dev));
    TDR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TDR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, umap_run_kernel_persist, kFastThreads, kSmem4));
    if (occ < 1) {
        set_error("tdr_umap_run: the persistent step kernel does not fit on an SM");
        return TDR_E_CUDA;
    }
    int64_t grid = (int64_t)sms * occ;
    const int64_t ctas_needed = (p.n_local + kWarps4 * 32 - 1) / (kWarps4 * 32);
    if (grid > ctas_needed) grid = ctas_needed < 1 ? 1 : ctas_needed;
    // (An L2 access-policy window marking the two embedding buffers persisting and the edge streams streaming was
    // measured at 10 M points: 48 MB set-aside +0.3 %, 79 MB -7 % — profiles/r2_step_variants.md; not used.)
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeCooperative;
    attrs[0].val.cooperative = 1;
    const unsigned n_attrs = 1;
    const int64_t n_iter0 = p.n_iter;
    for (int done = 0; done < n_steps; done += kMaxRunSteps) {
        const int chunk = n_steps - done < kMaxRunSteps ? n_steps - done : kMaxRunSteps;
        rp.n_steps = chunk;
        rp.gnorm_last = (want_gnorm && done + chunk == n_steps) ? 1 : 0;
        for (int t = 0; t < chunk; ++t) rp.lrs[t] = lrs_host[done + t];
        p.n_iter = n_iter0 + done;
        int cur = (cur0 + done) & 1;
        void* args[] = {(void*)&p, (void*)&rp, (void*)&cur};
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kFastThreads);
        cfg.dynamicSmemBytes = kSmem4;
        cfg.stream = st;
        cfg.attrs = attrs;
        cfg.numAttrs = n_attrs;
        TDR_CUDA(cudaLaunchKernelExC(&cfg, (const void*)umap_run_kernel_persist, args));
        rp.epoch0 += (uint32_t)chunk;
    }
    return TDR_OK;
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
    TDR_CHECK_ARG(n_total <= 0x7fffffffLL, "tdr_umap_step_f32: more than 2^31 - 1 points");
    TDR_CHECK_ARG(n_neg >= 0 && negative_sample_rate >= 0, "tdr_umap_step_f32: bad negative sampling config");
    TDR_CHECK_ARG(precise == 0 || precise == 1, "tdr_umap_step_f32: precise is 0 or 1");
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
    TDR_CHECK_ARG(n_peers >= 0 && n_peers <= kMaxPeers && (n_peers == 0 || peer_out), "tdr_umap_step: at most 8 peer buffers");
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
    float* peers[kMaxPeers] = {nullptr};
    TDR_CHECK_ARG(n_peers >= 0 && n_peers <= kMaxPeers, "tdr_umap_step_p2p_f32: at most 8 peers");
    for (int q = 0; q < n_peers; ++q) peers[q] = reinterpret_cast<float*>(peer_out_ptrs[q]);
    return umap_step_impl(Z_in, Z_out, n_total, row0, n_local, rowptr, col, epochs_per_sample, epoch_of_next_sample,
                          nullptr, n_neg, negative_sample_rate, seed, n_iter, a, b, lam, repulsion, lr, 0, nullptr,
                          gnorm_sq, nan_flag, nullptr, peers, n_peers, stream);
}

// Shared body of the two multi-iteration entry points: the persistent kernel over rows [row0, row0 + n_local).
static int umap_run_impl(float* Z_a, float* Z_b, int64_t n_total, int64_t row0, int64_t n_local, const int64_t* rowptr,
                         const int32_t* col, const float* epochs_per_sample, float* epoch_of_next_sample, int n_neg,
                         int negative_sample_rate, uint64_t seed, int64_t n_iter0, int n_steps, const float* lrs_host,
                         double a, double b, float lam, float repulsion, double* gnorm_sq, int* nan_flag, uint64_t* stats,
                         uint32_t* sync_words, const uint64_t* peers_a, const uint64_t* peers_b, uint32_t* my_flags,
                         const uint64_t* peer_flags, int n_peers, int rank, uint32_t epoch0, double timeout_s,
                         tdr_stream_t stream) {
    TDR_CHECK_ARG(Z_a && Z_b && Z_a != Z_b && lrs_host && n_steps >= 0 && rowptr && col && epochs_per_sample &&
                      epoch_of_next_sample && sync_words, "tdr_umap_run: bad arguments");
    TDR_CHECK_ARG(n_total >= 2 && n_total <= 0x7fffffffLL && row0 >= 0 && n_local >= 0 && row0 + n_local <= n_total,
                  "tdr_umap_run: bad row range");
    TDR_CHECK_ARG(n_neg >= 0 && negative_sample_rate >= 0, "tdr_umap_run: bad negative sampling config");
    TDR_CHECK_ARG(n_peers >= 0 && n_peers <= kMaxPeers, "tdr_umap_run: at most 8 peers");
    TDR_CHECK_ARG(n_peers == 0 || (peers_a && peers_b && my_flags && peer_flags && rank >= 0 && rank <= n_peers),
                  "tdr_umap_run: peer buffers / flags missing");
    if (n_steps == 0) return TDR_OK;
    UmapStepParams p{};
    p.n_total = n_total;
    p.row0 = row0;
    p.n_local = n_local;
    p.rowptr = rowptr;
    p.col = col;
    p.eps = epochs_per_sample;
    p.eons = epoch_of_next_sample;
    p.n_neg = n_neg;
    p.rate = negative_sample_rate;
    p.seed = seed;
    p.n_iter = n_iter0;
    fill_consts(p, (float)a, (float)b, a, b);
    p.lam = lam;
    p.rep = repulsion;
    p.gnorm_sq = gnorm_sq;
    p.nan_flag = nan_flag;
    p.stats = reinterpret_cast<unsigned long long*>(stats);
    p.n_peers = n_peers;
    RunParams rp{};
    rp.Z[0] = Z_a;
    rp.Z[1] = Z_b;
    rp.sync = sync_words;
    rp.my_flags = my_flags;
    rp.rank = rank;
    rp.epoch0 = epoch0;
    rp.timeout_ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 60.0) * 1e9);
    for (int q = 0; q < n_peers; ++q) {
        rp.peer_Z[0][q] = reinterpret_cast<float2*>(peers_a[q]);
        rp.peer_Z[1][q] = reinterpret_cast<float2*>(peers_b[q]);
        rp.peer_flags[q] = reinterpret_cast<uint32_t*>(peer_flags[q]);
        rp.peer_rank[q] = q < rank ? q : q + 1;  // ranks ascending without this rank
        TDR_CHECK_ARG(rp.peer_Z[0][q] && rp.peer_Z[1][q] && rp.peer_flags[q], "tdr_umap_run: null peer address");
    }
    return launch_persistent(p, rp, n_steps, lrs_host, 0, gnorm_sq != nullptr, (cudaStream_t)stream);
}

extern "C" TDR_API int tdr_umap_run_f32(float* Z_a, float* Z_b, int64_t n_total, const int64_t* rowptr,
                                const int32_t* col, const float* epochs_per_sample, float* epoch_of_next_sample,
                                int n_neg, int negative_sample_rate, uint64_t seed, int64_t n_iter0, int n_steps,
                                const float* lrs_host, double a, double b, float lam, float repulsion, int precise,
                                double* gnorm_sq, int* nan_flag, uint64_t* stats, uint32_t* sync_words,
                                uint32_t epoch0, tdr_stream_t stream) {
    if (precise == 0)
        return umap_run_impl(Z_a, Z_b, n_total, 0, n_total, rowptr, col, epochs_per_sample, epoch_of_next_sample, n_neg,
                             negative_sample_rate, seed, n_iter0, n_steps, lrs_host, a, b, lam, repulsion, gnorm_sq,
                             nan_flag, stats, sync_words, nullptr, nullptr, nullptr, nullptr, 0, 0, epoch0, 0.0, stream);
    // parity kernel: one launch per iteration
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

extern "C" TDR_API int tdr_umap_run_p2p_f32(float* Z_a, float* Z_b, int64_t n_total, int64_t row0, int64_t n_local,
                                            const int64_t* rowptr, const int32_t* col, const float* epochs_per_sample,
                                            float* epoch_of_next_sample, int n_neg, int negative_sample_rate,
                                            uint64_t seed, int64_t n_iter0, int n_steps, const float* lrs_host, double a,
                                            double b, float lam, float repulsion, double* gnorm_sq, int* nan_flag,
                                            uint64_t* stats, uint32_t* sync_words, const uint64_t* peers_a /*host*/,
                                            const uint64_t* peers_b /*host*/, uint32_t* my_flags,
                                            const uint64_t* peer_flags /*host*/, int n_peers, int rank, int world,
                                            uint32_t epoch0, double timeout_s, tdr_stream_t stream) {
    TDR_CHECK_ARG(n_peers >= 1 && n_peers <= kMaxPeers && world == n_peers + 1 && rank >= 0 && rank < world,
                  "tdr_umap_run_p2p_f32: 2..9 ranks");
    return umap_run_impl(Z_a, Z_b, n_total, row0, n_local, rowptr, col, epochs_per_sample, epoch_of_next_sample, n_neg,
                         negative_sample_rate, seed, n_iter0, n_steps, lrs_host, a, b, lam, repulsion, gnorm_sq, nan_flag,
                         stats, sync_words, peers_a, peers_b, my_flags, peer_flags, n_peers, rank, epoch0, timeout_s, stream);
}
