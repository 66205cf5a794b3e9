// (i) Tiled pairwise distance + exact top-k, fp32 SIMT mainloop (v0 of "kernel 1").
//
// Replaces torchdr/distance/torch.py:81-122 (expanded-form distance, 1e12 diagonal,
// topk) without ever materialising the N x N matrix.  One CTA owns BM=128 query rows and
// sweeps the whole database in BN=128-column tiles: register-tiled 8x8 FFMA mainloop over
// K chunks of 16 (register-staged double buffer in shared memory), then each thread
// filters its 64 distances against the per-row running k-th best held in shared memory.
// Survivors (rare after the first tiles: ~k ln(N/k) per row over the whole sweep) go
// through a small per-row queue and are merged into the row's sorted top-k list by one
// warp.  Total order on (distance, index): ties resolve to the lower index.
//
// MODE_KNN   : out_dist/out_idx [nq,k]
// MODE_FUSED : + UMAP rho/sigma search on the finished rows (rowsearch.cuh) — the
//              "fused distance + sigma-bisection" kernel of BASELINE.json config 2
// MODE_FULL  : one (query tile, db tile) per CTA, writes the dense C tile (k=None path)
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "rowsearch.cuh"

namespace tdr {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int LDS = BM + 4;  // padded leading dimension of the K-major smem tiles
constexpr int QCAP = 16;     // per-row candidate queue

enum { MODE_KNN = 0, MODE_FUSED = 1, MODE_FULL = 2 };

// knn_tc.cu
bool knn_tc_supported(int d, int k);
bool knn_tc_full_supported(int d);
int knn_tc_full_launch(const float* X, int64_t n, const float* Y, int64_t m, int d, bool same, int metric,
                       int exclude_diag, float* C, void* ws, size_t ws_bytes, cudaStream_t st);
size_t knn_tc_workspace_bytes(int64_t nq, int64_t ndb, int d, int k, bool same);
int knn_tc_launch(const float* Xq, int64_t nq, int64_t q_row0, const float* Xdb, int64_t ndb, int d, int k,
                  bool same, int exclude_self, int metric, int fused, int max_iter, float* out_dist, int32_t* out_idx,
                  float* P, float* rho, float* sigma, int prune, unsigned long long* sweep_stats,
                  const int32_t* db_labels, void* ws, size_t ws_bytes, cudaStream_t st);

// fp32 SIMT path with labels: the search ranks ties by row index; the reported ids are re-labelled afterwards
__global__ void __launch_bounds__(256) relabel_kernel(int32_t* __restrict__ idx, int64_t n, const int32_t* __restrict__ labels) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = __ldg(labels + idx[i]);
}

struct KnnParams {
    const float* Xq;   // [nq, ld]
    const float* Xdb;  // [ndb, ld]
    const float* qn;   // [nq]
    const float* dbn;  // [ndb]
    int64_t nq, ndb, q_row0;
    int ld;            // row stride, multiple of BK
    int k, kpad;
    int exclude_self, metric;
    float* out_dist;
    int32_t* out_idx;
    // fused epilogue
    int max_iter;
    float* P;
    float* rho;
    float* sigma;
    // full mode
    float* Cfull;
    int exclude_diag;
};

__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float* __restrict__ X, int64_t n, int d,
                                                         float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* x = X + row * d;
    // fp64 accumulation: the norm enters every distance of the row/column, keep it at 0.5 ulp
    double s = 0.0;
    for (int j = lane; j < d; j += 32) s = fma((double)x[j], (double)x[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = (float)s;
}

__global__ void __launch_bounds__(256) pad_rows_kernel(const float* __restrict__ X, int64_t n, int d, int ld,
                                                       float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * ld) return;
    const int64_t r = i / ld;
    const int c = (int)(i - r * ld);
    out[i] = c < d ? X[r * d + c] : 0.0f;
}

__device__ __forceinline__ bool cand_before(float d, int i, float td, int ti) {
    return d < td || (d == td && i < ti);
}

// Insert queue entries of `row` into its sorted list (one warp).
__device__ __forceinline__ void merge_row(float* __restrict__ ld_s, int* __restrict__ li_s, int k,
                                          const float* __restrict__ qd, const int* __restrict__ qi,
                                          int cnt, int lane) {
    for (int c = 0; c < cnt; ++c) {
        const float cd = qd[c];
        const int ci = qi[c];
        if (!cand_before(cd, ci, ld_s[k - 1], li_s[k - 1])) continue;  // warp-uniform
        int before = 0;
        for (int p = lane; p < k; p += 32) before += cand_before(ld_s[p], li_s[p], cd, ci) ? 1 : 0;
        const int pos = warp_sum_int(before);
        // new[p] = old[p] (p < pos) | cand (p == pos) | old[p-1] (p > pos)
        float nv[kMaxEPL];
        int ni[kMaxEPL];
#pragma unroll
        for (int e = 0; e < kMaxEPL; ++e) {
            const int p = lane + 32 * e;
            if (p > pos && p < k) {
                nv[e] = ld_s[p - 1];
                ni[e] = li_s[p - 1];
            }
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < kMaxEPL; ++e) {
            const int p = lane + 32 * e;
            if (p < k) {
                if (p == pos) {
                    ld_s[p] = cd;
                    li_s[p] = ci;
                } else if (p > pos) {
                    ld_s[p] = nv[e];
                    li_s[p] = ni[e];
                }
            }
        }
        __syncwarp();
    }
}

template <int MODE>
__global__ void __launch_bounds__(NT, 2) knn_tile_kernel(const KnnParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);  // [2][BK][LDS]
    float* Bs = As + 2 * BK * LDS;                   // [2][BK][LDS]
    float* tail = Bs + 2 * BK * LDS;
    // kNN state (unused in MODE_FULL)
    float* ld_s = tail;                                          // [BM][kpad]
    int* li_s = reinterpret_cast<int*>(ld_s + BM * prm.kpad);    // [BM][kpad]
    float* qd_s = reinterpret_cast<float*>(li_s + BM * prm.kpad);  // [BM][QCAP]
    int* qi_s = reinterpret_cast<int*>(qd_s + BM * QCAP);        // [BM][QCAP]
    int* qcnt_s = qi_s + BM * QCAP;                              // [BM]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int k = prm.k, kpad = prm.kpad;
    const int64_t q0 = (int64_t)blockIdx.x * BM;  // first query row of this CTA
    const int nK = prm.ld / BK;
    const int64_t n_db_tiles = (prm.ndb + BN - 1) / BN;
    const int64_t jt_begin = (MODE == MODE_FULL) ? (int64_t)blockIdx.y : 0;
    const int64_t jt_end = (MODE == MODE_FULL) ? jt_begin + 1 : n_db_tiles;

    if (MODE != MODE_FULL) {
        for (int i = tid; i < BM * kpad; i += NT) {
            ld_s[i] = INFINITY;
            li_s[i] = 0x7fffffff;
        }
        for (int i = tid; i < BM; i += NT) qcnt_s[i] = 0;
    }

    // rows / columns of this thread's 8x8 micro-tile
    int rloc[8], cloc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        rloc[i] = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
        cloc[i] = (i < 4) ? tx * 4 + i : 64 + tx * 4 + (i - 4);
    }
    float qn_r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t gr = q0 + rloc[i];
        qn_r[i] = gr < prm.nq ? __ldg(prm.qn + gr) : 0.0f;
    }

    // global -> register staging: 2 float4 of A and 2 of B per thread per K chunk
    const int lrow0 = tid >> 2, lq = tid & 3;  // rows lrow0 and lrow0 + 64, k-quad lq
    const float* a_src[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        int64_t gr = q0 + lrow0 + 64 * s;
        if (gr >= prm.nq) gr = prm.nq - 1;
        a_src[s] = prm.Xq + gr * prm.ld + lq * 4;
    }
    float4 a_reg[2], b_reg[2];
    auto load_chunk = [&](int64_t jt, int kc) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            a_reg[s] = __ldg(reinterpret_cast<const float4*>(a_src[s] + kc * BK));
            int64_t gc = jt * BN + lrow0 + 64 * s;
            if (gc >= prm.ndb) gc = prm.ndb - 1;
            b_reg[s] = __ldg(reinterpret_cast<const float4*>(prm.Xdb + gc * prm.ld + lq * 4 + kc * BK));
        }
    };
    auto store_chunk = [&](int buf) {
        float* a = As + buf * BK * LDS;
        float* b = Bs + buf * BK * LDS;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int r = lrow0 + 64 * s;
            a[(lq * 4 + 0) * LDS + r] = a_reg[s].x;
            a[(lq * 4 + 1) * LDS + r] = a_reg[s].y;
            a[(lq * 4 + 2) * LDS + r] = a_reg[s].z;
            a[(lq * 4 + 3) * LDS + r] = a_reg[s].w;
            b[(lq * 4 + 0) * LDS + r] = b_reg[s].x;
            b[(lq * 4 + 1) * LDS + r] = b_reg[s].y;
            b[(lq * 4 + 2) * LDS + r] = b_reg[s].z;
            b[(lq * 4 + 3) * LDS + r] = b_reg[s].w;
        }
    };

    load_chunk(jt_begin, 0);
    store_chunk(0);
    __syncthreads();
    int buf = 0;

    for (int64_t jt = jt_begin; jt < jt_end; ++jt) {
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

        for (int kc = 0; kc < nK; ++kc) {
            const bool last_chunk = (kc + 1 == nK);
            const bool has_next = !(last_chunk && jt + 1 == jt_end);
            if (has_next) load_chunk(last_chunk ? jt + 1 : jt, last_chunk ? 0 : kc + 1);
            const float* a = As + buf * BK * LDS;
            const float* b = Bs + buf * BK * LDS;
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(a + kk * LDS + ty * 4);
                const float4 a1 = *reinterpret_cast<const float4*>(a + kk * LDS + 64 + ty * 4);
                const float4 b0 = *reinterpret_cast<const float4*>(b + kk * LDS + tx * 4);
                const float4 b1 = *reinterpret_cast<const float4*>(b + kk * LDS + 64 + tx * 4);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            if (has_next) store_chunk(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }

        // ---- tile epilogue: distances (torch.py:89-95), filter, queue, merge ----
        float dbn_r[8];
        int gcol[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            gcol[j] = (int)(jt * BN) + cloc[j];
            dbn_r[j] = gcol[j] < prm.ndb ? __ldg(prm.dbn + gcol[j]) : 0.0f;
        }
        auto dist = [&](int i, int j) {
            float d = __fsub_rn(__fadd_rn(qn_r[i], dbn_r[j]), 2.0f * acc[i][j]);
            if (prm.metric == TDR_METRIC_EUCLIDEAN) d = sqrtf(fmaxf(d, 0.0f));
            return d;
        };

        if (MODE == MODE_FULL) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t gr = q0 + rloc[i];
                if (gr >= prm.nq) continue;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (gcol[j] >= prm.ndb) continue;
                    float d = dist(i, j);
                    if (prm.exclude_diag && gr == (int64_t)gcol[j]) d = __fadd_rn(d, 1e12f);  // torch.py:111-116
                    prm.Cfull[gr * prm.ndb + gcol[j]] = d;
                }
            }
        } else {
            unsigned long long pending = 0ull;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rloc[i];
                const float td = ld_s[r * kpad + k - 1];
                const int ti = li_s[r * kpad + k - 1];
                const int64_t self = prm.q_row0 + q0 + r;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = dist(i, j);
                    const bool ok = gcol[j] < prm.ndb && !(prm.exclude_self && (int64_t)gcol[j] == self) &&
                                    cand_before(d, gcol[j], td, ti);
                    if (ok) pending |= 1ull << (i * 8 + j);
                }
            }
            while (true) {
                if (pending) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (pending & (1ull << (i * 8 + j))) {
                                const int r = rloc[i];
                                const int slot = atomicAdd(&qcnt_s[r], 1);
                                if (slot < QCAP) {
                                    qd_s[r * QCAP + slot] = dist(i, j);
                                    qi_s[r * QCAP + slot] = gcol[j];
                                    pending &= ~(1ull << (i * 8 + j));
                                }
                            }
                }
                __syncthreads();
                // warp w merges rows w*16 .. w*16+15
                {
                    const int r = warp * 16 + (lane & 15);
                    const unsigned has = __ballot_sync(0xffffffffu, lane < 16 && qcnt_s[r] > 0);
                    unsigned m = has;
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        const int row = warp * 16 + b;
                        const int cnt = min(qcnt_s[row], QCAP);
                        merge_row(ld_s + row * kpad, li_s + row * kpad, k, qd_s + row * QCAP,
                                  qi_s + row * QCAP, cnt, lane);
                        if (lane == 0) qcnt_s[row] = 0;
                    }
                }
                const int any = __syncthreads_or(pending != 0ull);
                if (!any) break;
                // overflowed candidates: drop those that no longer qualify, retry the rest
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rloc[i];
                    const float td = ld_s[r * kpad + k - 1];
                    const int ti = li_s[r * kpad + k - 1];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if ((pending & (1ull << (i * 8 + j))) && !cand_before(dist(i, j), gcol[j], td, ti))
                            pending &= ~(1ull << (i * 8 + j));
                }
            }
        }
    }

    if (MODE == MODE_FULL) return;

    // ---- write back (and, fused, run the sigma/rho search on the finished rows) ----
    for (int rr = 0; rr < 16; ++rr) {
        const int row = warp * 16 + rr;
        const int64_t gr = q0 + row;
        if (gr >= prm.nq) continue;  // warp-uniform
        const float* ldr = ld_s + row * kpad;
        const int* lir = li_s + row * kpad;
        for (int p = lane; p < k; p += 32) {
            if (prm.out_dist) prm.out_dist[gr * k + p] = ldr[p];
            prm.out_idx[gr * k + p] = lir[p];
        }
        if (MODE == MODE_FUSED) {
            auto run = [&](auto tag) {
                constexpr int EPL = decltype(tag)::value;
                UmapRow<EPL> u;
                u.k = k;
                u.lane = lane;
                u.target = log2f((float)k);
#pragma unroll
                for (int e = 0; e < EPL; ++e) u.c[e] = u.valid(e) ? ldr[lane + 32 * e] : INFINITY;
                u.init();
                const float s = u.solve(prm.max_iter);
#pragma unroll
                for (int e = 0; e < EPL; ++e)
                    if (u.valid(e)) prm.P[gr * k + lane + 32 * e] = u.p(e, s);
                if (lane == 0) {
                    prm.rho[gr] = u.rho;
                    prm.sigma[gr] = s;
                }
            };
            switch ((k + 31) / 32) {
                case 1: run(std::integral_constant<int, 1>{}); break;
                case 2: run(std::integral_constant<int, 2>{}); break;
                case 3: run(std::integral_constant<int, 3>{}); break;
                case 4: run(std::integral_constant<int, 4>{}); break;
                default: run(std::integral_constant<int, 5>{}); break;
            }
        }
    }
}

static size_t knn_smem_bytes(int kpad, int mode) {
    size_t b = (size_t)2 * 2 * BK * LDS * sizeof(float);
    if (mode != MODE_FULL) b += (size_t)BM * kpad * 8 + (size_t)BM * QCAP * 8 + BM * sizeof(int);
    return b;
}

struct Prepared {
    const float* Xq;
    const float* Xdb;
    float* qn;
    float* dbn;
    int ld;
};

static size_t prepare_bytes(int64_t nq, int64_t ndb, int d, bool same) {
    const int ld = (int)align_up((size_t)d, BK);
    size_t b = align_up((size_t)ndb * 4, 256);
    if (!same) b += align_up((size_t)nq * 4, 256);
    if (ld != d) {
        b += align_up((size_t)ndb * ld * 4, 256);
        if (!same) b += align_up((size_t)nq * ld * 4, 256);
    }
    return b;
}

// Row norms (torch.py:81-86) and, when d is not a multiple of BK, zero-padded copies.
static int prepare(const float* Xq, int64_t nq, const float* Xdb, int64_t ndb, int d, bool same,
                   void* ws, size_t ws_bytes, cudaStream_t st, Prepared* out) {
    const size_t need = prepare_bytes(nq, ndb, d, same);
    if (ws_bytes < need || (need && !ws)) {
        set_error("workspace too small: need %zu bytes, got %zu", need, ws_bytes);
        return TDR_E_WORKSPACE;
    }
    if ((uintptr_t)ws % 256 != 0 || (uintptr_t)Xq % 16 != 0 || (uintptr_t)Xdb % 16 != 0) {
        set_error("workspace must be 256-byte aligned and inputs 16-byte aligned");
        return TDR_E_WORKSPACE;
    }
    const int ld = (int)align_up((size_t)d, BK);
    char* p = (char*)ws;
    float* dbn = (float*)p;
    p += align_up((size_t)ndb * 4, 256);
    float* qn = dbn;
    if (!same) {
        qn = (float*)p;
        p += align_up((size_t)nq * 4, 256);
    }
    row_sqnorm_kernel<<<(unsigned)((ndb + 7) / 8), 256, 0, st>>>(Xdb, ndb, d, dbn);
    if (!same) row_sqnorm_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(Xq, nq, d, qn);
    const float* xq = Xq;
    const float* xdb = Xdb;
    if (ld != d) {
        float* pdb = (float*)p;
        p += align_up((size_t)ndb * ld * 4, 256);
        pad_rows_kernel<<<(unsigned)((ndb * ld + 255) / 256), 256, 0, st>>>(Xdb, ndb, d, ld, pdb);
        xdb = pdb;
        if (same) {
            xq = pdb;
        } else {
            float* pq = (float*)p;
            pad_rows_kernel<<<(unsigned)((nq * ld + 255) / 256), 256, 0, st>>>(Xq, nq, d, ld, pq);
            xq = pq;
        }
    }
    TDR_LAUNCH_CHECK();
    out->Xq = xq;
    out->Xdb = xdb;
    out->qn = qn;
    out->dbn = dbn;
    out->ld = ld;
    return TDR_OK;
}

template <int MODE>
static int launch_knn(const KnnParams& prm, dim3 grid, cudaStream_t st) {
    const size_t smem = knn_smem_bytes(prm.kpad, MODE);
    // per call: the attribute belongs to the current device's context (a process may drive several GPUs)
    TDR_CUDA(cudaFuncSetAttribute(knn_tile_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    knn_tile_kernel<MODE><<<grid, NT, smem, st>>>(prm);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

static int knn_common(int mode, const float* Xq, int64_t nq, int64_t q_row0, const float* Xdb, int64_t ndb,
                      int d, int k, int exclude_self, int metric, int max_iter, float* out_dist,
                      int32_t* out_idx, float* P, float* rho, float* sigma, int path, int prune,
                      uint64_t* sweep_stats, const int32_t* db_labels, void* ws, size_t ws_bytes, cudaStream_t st) {
    TDR_CHECK_ARG(Xq && Xdb && out_idx, "knn: null pointer");
    TDR_CHECK_ARG(path >= TDR_KNN_PATH_AUTO && path <= TDR_KNN_PATH_TC, "knn: path must be 0 (auto), 1 (SIMT fp32) or 2 (tcgen05)");
    TDR_CHECK_ARG(prune >= TDR_KNN_PRUNE_DEFAULT && prune <= TDR_KNN_PRUNE_CERTIFIED, "knn: prune must be -1 .. 2");
    if (prune == TDR_KNN_PRUNE_DEFAULT) prune = TDR_KNN_PRUNE_ON;
    TDR_CHECK_ARG(nq >= 0 && ndb >= 1 && d >= 1, "knn: bad shape nq=%lld ndb=%lld d=%d", (long long)nq,
                  (long long)ndb, d);
    TDR_CHECK_ARG(k >= 1 && k <= TDR_MAX_K, "knn: k=%d outside [1,%d]", k, TDR_MAX_K);
    TDR_CHECK_ARG(k <= ndb - (exclude_self ? 1 : 0), "knn: k=%d exceeds the %lld available neighbours", k,
                  (long long)(ndb - (exclude_self ? 1 : 0)));
    TDR_CHECK_ARG(ndb < 0x7fffffffLL, "knn: ndb must fit int32 indices");
    TDR_CHECK_ARG(metric == TDR_METRIC_SQEUCLIDEAN || metric == TDR_METRIC_EUCLIDEAN,
                  "[TorchDR] ERROR : metric id %d is not supported.", metric);
    if (nq == 0) return TDR_OK;
    // tensor-core path (knn_tc.cu) when the tile shapes allow it; the SIMT kernel below otherwise
    if (path != TDR_KNN_PATH_SIMT && knn_tc_supported(d, k)) {
        const bool inside = (Xq == Xdb + q_row0 * d) && q_row0 + nq <= ndb;
        return knn_tc_launch(Xq, nq, q_row0, Xdb, ndb, d, k, inside, exclude_self, metric, mode == MODE_FUSED,
                             max_iter, out_dist, out_idx, P, rho, sigma, prune,
                             reinterpret_cast<unsigned long long*>(sweep_stats), db_labels, ws, ws_bytes, st);
    }
    if (path == TDR_KNN_PATH_TC) {
        set_error("knn: tensor-core path forced but unsupported for d=%d k=%d", d, k);
        return TDR_E_UNSUPPORTED;
    }
    const bool same = (Xq == Xdb && nq == ndb);
    Prepared pr;
    int rc = prepare(Xq, nq, Xdb, ndb, d, same, ws, ws_bytes, st, &pr);
    if (rc != TDR_OK) return rc;
    KnnParams prm{};
    prm.Xq = pr.Xq;
    prm.Xdb = pr.Xdb;
    prm.qn = pr.qn;
    prm.dbn = pr.dbn;
    prm.nq = nq;
    prm.ndb = ndb;
    prm.q_row0 = q_row0;
    prm.ld = pr.ld;
    prm.k = k;
    prm.kpad = (k + 31) / 32 * 32;
    prm.exclude_self = exclude_self;
    prm.metric = metric;
    prm.out_dist = out_dist;
    prm.out_idx = out_idx;
    prm.max_iter = max_iter;
    prm.P = P;
    prm.rho = rho;
    prm.sigma = sigma;
    dim3 grid((unsigned)((nq + BM - 1) / BM));
    rc = mode == MODE_FUSED ? launch_knn<MODE_FUSED>(prm, grid, st) : launch_knn<MODE_KNN>(prm, grid, st);
    if (rc == TDR_OK && db_labels) {
        relabel_kernel<<<(unsigned)((nq * k + 255) / 256), 256, 0, st>>>(out_idx, nq * k, db_labels);
        TDR_LAUNCH_CHECK();
    }
    return rc;
}

}  // namespace tdr

using namespace tdr;

extern "C" TDR_API size_t tdr_knn_workspace_bytes(int64_t nq, int64_t ndb, int d, int k) {
    // conservative: assume distinct query / database buffers; covers both kernel paths
    size_t b = prepare_bytes(nq, ndb, d, false) + 256;
    if (knn_tc_supported(d, k)) b = std::max(b, knn_tc_workspace_bytes(nq, ndb, d, k, false) + 256);
    return b;
}

extern "C" TDR_API int tdr_knn_f32(const float* Xq, int64_t nq, int64_t q_row0, const float* Xdb, int64_t ndb, int d,
                           int k, int exclude_self, int metric, float* out_dist, int32_t* out_idx, int path,
                           int prune, uint64_t* sweep_stats, const int32_t* db_labels, void* ws, size_t ws_bytes,
                           tdr_stream_t stream) {
    TDR_CHECK_ARG(out_dist, "tdr_knn_f32: out_dist is null");
    return knn_common(MODE_KNN, Xq, nq, q_row0, Xdb, ndb, d, k, exclude_self, metric, 0, out_dist, out_idx,
                      nullptr, nullptr, nullptr, path, prune, sweep_stats, db_labels, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" TDR_API int tdr_knn_umap_fused_f32(const float* Xq, int64_t nq, int64_t q_row0, const float* Xdb,
                                      int64_t ndb, int d, int k, int exclude_self, int max_iter,
                                      float* out_dist, int32_t* out_idx, float* P, float* rho, float* sigma,
                                      int path, int prune, uint64_t* sweep_stats, const int32_t* db_labels, void* ws,
                                      size_t ws_bytes, tdr_stream_t stream) {
    TDR_CHECK_ARG(P && rho && sigma, "tdr_knn_umap_fused_f32: null output");
    return knn_common(MODE_FUSED, Xq, nq, q_row0, Xdb, ndb, d, k, exclude_self, TDR_METRIC_SQEUCLIDEAN,
                      max_iter, out_dist, out_idx, P, rho, sigma, path, prune, sweep_stats, db_labels, ws, ws_bytes,
                      (cudaStream_t)stream);
}

extern "C" TDR_API int tdr_pairwise_full_f32(const float* X, int64_t n, const float* Y, int64_t m, int d, int metric,
                                     int exclude_diag, float* C, int path, void* ws, size_t ws_bytes,
                                     tdr_stream_t stream) {
    TDR_CHECK_ARG(X && Y && C, "tdr_pairwise_full_f32: null pointer");
    TDR_CHECK_ARG(path >= TDR_KNN_PATH_AUTO && path <= TDR_KNN_PATH_TC, "tdr_pairwise_full_f32: path must be 0, 1 or 2");
    TDR_CHECK_ARG(n >= 1 && m >= 1 && d >= 1, "tdr_pairwise_full_f32: bad shape");
    TDR_CHECK_ARG(m < 0x7fffffffLL, "tdr_pairwise_full_f32: m must fit int32");
    TDR_CHECK_ARG(metric == TDR_METRIC_SQEUCLIDEAN || metric == TDR_METRIC_EUCLIDEAN,
                  "[TorchDR] ERROR : metric id %d is not supported.", metric);
    cudaStream_t st = (cudaStream_t)stream;
    const bool same = (X == Y && n == m);
    if (path != TDR_KNN_PATH_SIMT && knn_tc_full_supported(d))
        return knn_tc_full_launch(X, n, Y, m, d, same, metric, exclude_diag, C, ws, ws_bytes, st);
    if (path == TDR_KNN_PATH_TC) {
        set_error("tdr_pairwise_full_f32: tensor-core path forced but unsupported for d=%d", d);
        return TDR_E_UNSUPPORTED;
    }
    Prepared pr;
    int rc = prepare(X, n, Y, m, d, same, ws, ws_bytes, st, &pr);
    if (rc != TDR_OK) return rc;
    KnnParams prm{};
    prm.Xq = pr.Xq;
    prm.Xdb = pr.Xdb;
    prm.qn = pr.qn;
    prm.dbn = pr.dbn;
    prm.nq = n;
    prm.ndb = m;
    prm.ld = pr.ld;
    prm.k = 1;
    prm.kpad = 32;
    prm.metric = metric;
    prm.Cfull = C;
    prm.exclude_diag = exclude_diag && same;
    dim3 grid((unsigned)((n + BM - 1) / BM), (unsigned)((m + BN - 1) / BN));
    return launch_knn<MODE_FULL>(prm, grid, st);
}
