// Graph stage between the affinity rows and the optimisation loop (one-off, O(N k)).
//
//  * fuzzy-union symmetrisation Q = P + P^T - P o P^T of torchdr/utils/sparse.py:170-206,
//    emitted as CSR with ascending columns instead of the reference's -1-padded ELL
//    (same entries, same order, no padding);
//  * UMAP edge schedule of torchdr/neighbor_embedding/umap.py:215-234;
//  * compaction of never-sampled edges, CSR->ELL export for the SparseAffinity seam,
//    and the edge export used by the multi-GPU all_to_all (sparse.py:259-309).
//
// The (row, col) merge is a radix sort of 2 N k 64-bit keys; it uses CUB's device-wide
// sort/scan primitives (header-only, part of the CUDA toolkit) — this is a build-once
// stage, not the per-iteration hot path.
#include <cub/cub.cuh>

#include "common.cuh"

namespace tdr {

__host__ __device__ inline int key_bits(uint64_t max_key) {
    int b = 1;
    while (b < 64 && (max_key >> b) != 0) ++b;
    return b;
}

// keys: ((local_row * n_total + col) << 1) | from_transpose ; invalid slots get `sentinel`.
__global__ void __launch_bounds__(256)
sym_emit_kernel(const float* __restrict__ P, const int32_t* __restrict__ idx, int64_t n_local, int k,
                int64_t row0, int64_t n_total, const int64_t* __restrict__ ext_row,
                const int32_t* __restrict__ ext_col, const float* __restrict__ ext_val, int64_t n_ext,
                int transpose_local, uint64_t sentinel, uint64_t* __restrict__ keys,
                float* __restrict__ vals) {
    const int64_t nk = n_local * k;
    const int64_t total = 2 * nk + n_ext;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    if (i < nk) {
        const int64_t r = i / k;
        const int64_t c = idx[i];
        keys[i] = ((uint64_t)(r * n_total + c)) << 1;
        vals[i] = P[i];
    } else if (i < 2 * nk) {
        const int64_t e = i - nk;
        const int64_t src = row0 + e / k;  // global row of the edge
        const int64_t tgt = idx[e];        // becomes the row of the transposed entry
        const bool local = transpose_local && tgt >= row0 && tgt < row0 + n_local;
        keys[i] = local ? ((((uint64_t)((tgt - row0) * n_total + src)) << 1) | 1ull) : sentinel;
        vals[i] = P[e];
    } else {
        const int64_t e = i - 2 * nk;
        const int64_t r = ext_row[e] - row0;
        const bool ok = r >= 0 && r < n_local;
        keys[i] = ok ? ((((uint64_t)(r * n_total + ext_col[e])) << 1) | 1ull) : sentinel;
        vals[i] = ext_val[e];
    }
}

__global__ void __launch_bounds__(256)
sym_heads_kernel(const uint64_t* __restrict__ sk, int64_t m, uint64_t sentinel, int* __restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > m) return;
    if (i == m) {
        head[i] = 0;
        return;
    }
    const uint64_t key = sk[i];
    head[i] = (key < sentinel && (i == 0 || (sk[i - 1] >> 1) != (key >> 1))) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
sym_combine_kernel(const uint64_t* __restrict__ sk, const float* __restrict__ sv, int64_t m,
                   uint64_t sentinel, int64_t n_total, const int* __restrict__ head,
                   const int64_t* __restrict__ pos, int mode, int32_t* __restrict__ col, float* __restrict__ val) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m || !head[i]) return;
    const uint64_t cell = sk[i] >> 1;
    float from_p = 0.0f, from_pt = 0.0f;  // sparse.py:78-81 (scatter_add of each side)
    for (int64_t j = i; j < m && sk[j] < sentinel && (sk[j] >> 1) == cell; ++j) {
        if (sk[j] & 1ull) from_pt = __fadd_rn(from_pt, sv[j]);
        else from_p = __fadd_rn(from_p, sv[j]);
    }
    const int64_t o = pos[i];
    col[o] = (int32_t)(cell % (uint64_t)n_total);
    // sparse.py:163-164: vP + vPT - vP * vPT, three separately rounded ops; mode SUM (sparse.py:159-160): vP + vPT
    const float sum = __fadd_rn(from_p, from_pt);
    val[o] = mode == TDR_SYM_SUM ? sum : __fsub_rn(sum, __fmul_rn(from_p, from_pt));
}

__global__ void __launch_bounds__(256)
sym_rowptr_kernel(const uint64_t* __restrict__ sk, int64_t m, int64_t n_local, int64_t n_total,
                  const int64_t* __restrict__ pos, int64_t* __restrict__ rowptr, int64_t* __restrict__ nnz_out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_local) return;
    const uint64_t want = ((uint64_t)(r * n_total)) << 1;
    int64_t lo = 0, hi = m;  // first position with key >= want
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sk[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    rowptr[r] = pos[lo];
    if (r == n_local && nnz_out) *nnz_out = pos[lo];
}

struct SymWs {
    uint64_t *k_in, *k_out;
    float *v_in, *v_out;
    int* head;
    int64_t* pos;
    void* cub_tmp;
    size_t cub_bytes;
    size_t total;
};

static SymWs sym_layout(void* ws, int64_t m) {
    SymWs w{};
    size_t sort_b = 0, scan_b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (uint64_t*)nullptr, (uint64_t*)nullptr, (float*)nullptr,
                                    (float*)nullptr, m, 0, 64);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_b, (int*)nullptr, (int64_t*)nullptr, m + 1);
    w.cub_bytes = sort_b > scan_b ? sort_b : scan_b;
    char* p = (char*)ws;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* q = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return q;
    };
    w.k_in = (uint64_t*)take((size_t)m * 8);
    w.k_out = (uint64_t*)take((size_t)m * 8);
    w.v_in = (float*)take((size_t)m * 4);
    w.v_out = (float*)take((size_t)m * 4);
    w.head = (int*)take((size_t)(m + 1) * 4);
    w.pos = (int64_t*)take((size_t)(m + 1) * 8);
    w.cub_tmp = take(w.cub_bytes);
    w.total = off;
    return w;
}

__global__ void __launch_bounds__(256)
max_kernel(const float* __restrict__ v, int64_t n, int* __restrict__ out_bits) {
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, v[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_int(m));  // values are >= 0
}

__global__ void __launch_bounds__(256)
schedule_kernel(const float* __restrict__ val, int64_t nnz, float a_max, float thr,
                float* __restrict__ eps, float* __restrict__ eons) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const float v = val[i];
    // umap.py:227: add_(1e-3).reciprocal_().mul_(A_max) ; :228-230 masked_fill_(small, inf)
    float e = __fmul_rn(__frcp_rn(__fadd_rn(v, 1e-3f)), a_max);
    if (v <= thr) e = INFINITY;
    eps[i] = e;
    eons[i] = e;
}

__global__ void __launch_bounds__(256)
finite_flag_kernel(const float* __restrict__ eps, int64_t nnz, int* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nnz) return;
    flag[i] = (i < nnz && eps[i] < INFINITY) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
compact_kernel(const int32_t* __restrict__ col, const float* __restrict__ eps, int64_t nnz,
               const int* __restrict__ flag, const int64_t* __restrict__ pos, int32_t* __restrict__ out_col,
               float* __restrict__ out_eps, float* __restrict__ out_eons) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz || !flag[i]) return;
    const int64_t o = pos[i];
    out_col[o] = col[i];
    out_eps[o] = eps[i];
    out_eons[o] = eps[i];
}

__global__ void __launch_bounds__(256)
compact_rowptr_kernel(const int64_t* __restrict__ rowptr, int64_t n_local, const int64_t* __restrict__ pos,
                      int64_t* __restrict__ out_rowptr, int64_t* __restrict__ nnz_out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_local) return;
    out_rowptr[r] = pos[rowptr[r]];
    if (r == n_local && nnz_out) *nnz_out = pos[rowptr[r]];
}

__global__ void __launch_bounds__(256)
csr_to_ell_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                  const float* __restrict__ val, int64_t n_local, int64_t width, float pad_val,
                  float* __restrict__ ell_val, int64_t* __restrict__ ell_idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local * width) return;
    const int64_t r = i / width, s = i - r * width;
    const int64_t b = rowptr[r], deg = rowptr[r + 1] - b;
    if (s < deg) {
        ell_val[i] = val[b + s];
        ell_idx[i] = col[b + s];
    } else {
        ell_val[i] = pad_val;
        ell_idx[i] = -1;
    }
}

__device__ __forceinline__ int owner_of(int64_t i, int64_t n, int world) {
    // torchdr/distributed/__init__.py:251-267
    const int64_t base = n / world, extra = n % world, cut = extra * (base + 1);
    int64_t r = i < cut ? i / (base + 1) : extra + (base ? (i - cut) / base : 0);
    if (r < 0) r = 0;
    if (r > world - 1) r = world - 1;
    return (int)r;
}

__global__ void __launch_bounds__(256)
export_count_kernel(const int32_t* __restrict__ idx, int64_t nk, int64_t n_total, int world, int rank,
                    unsigned long long* __restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nk) return;
    const int o = owner_of(idx[i], n_total, world);
    if (o != rank) atomicAdd(&counts[o], 1ull);
}

__global__ void export_offsets_kernel(unsigned long long* counts, int world) {
    // counts[0..world) -> cursors[world..2world) = exclusive prefix
    unsigned long long acc = 0;
    for (int r = 0; r < world; ++r) {
        counts[world + r] = acc;
        acc += counts[r];
    }
}

__global__ void __launch_bounds__(256)
export_fill_kernel(const float* __restrict__ P, const int32_t* __restrict__ idx, int64_t nk, int k,
                   int64_t row0, int64_t n_total, int world, int rank, unsigned long long* __restrict__ cursors,
                   int64_t* __restrict__ out_row, int32_t* __restrict__ out_col, float* __restrict__ out_val) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nk) return;
    const int64_t tgt = idx[i];
    const int o = owner_of(tgt, n_total, world);
    if (o == rank) return;
    const unsigned long long slot = atomicAdd(&cursors[o], 1ull);
    out_row[slot] = tgt;
    out_col[slot] = (int32_t)(row0 + i / k);
    out_val[slot] = P[i];
}

}  // namespace tdr

using namespace tdr;

static inline unsigned blocks_for(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

extern "C" TDR_API size_t tdr_symmetrize_workspace_bytes(int64_t n_local, int k, int64_t n_ext) {
    const int64_t m = 2 * n_local * k + n_ext;
    return sym_layout(nullptr, m).total + 256;
}

extern "C" TDR_API int tdr_symmetrize_csr_f32(const float* P, const int32_t* idx, int64_t n_local, int k, int64_t row0,
                                      int64_t n_total, const int64_t* ext_row, const int32_t* ext_col,
                                      const float* ext_val, int64_t n_ext, int transpose_local, int mode,
                                      int64_t* rowptr, int32_t* col, float* val, int64_t* nnz_out, void* ws,
                                      size_t ws_bytes, tdr_stream_t stream) {
    TDR_CHECK_ARG(P && idx && rowptr && col && val, "tdr_symmetrize_csr_f32: null pointer");
    TDR_CHECK_ARG(mode == TDR_SYM_SUM_MINUS_PROD || mode == TDR_SYM_SUM, "tdr_symmetrize_csr_f32: unknown mode %d", mode);
    TDR_CHECK_ARG(n_local >= 1 && k >= 1 && n_total >= n_local && row0 >= 0 && row0 + n_local <= n_total,
                  "tdr_symmetrize_csr_f32: bad shape");
    TDR_CHECK_ARG(n_ext == 0 || (ext_row && ext_col && ext_val), "tdr_symmetrize_csr_f32: null ext arrays");
    const int64_t m = 2 * n_local * k + n_ext;
    TDR_CHECK_ARG(m < 0x7fffffffLL, "tdr_symmetrize_csr_f32: %lld edges exceed the 2^31 limit", (long long)m);
    TDR_CHECK_ARG((double)n_local * (double)n_total < 4.0e18, "tdr_symmetrize_csr_f32: key overflow");
    SymWs w = sym_layout(ws, m);
    if (!ws || ws_bytes < w.total || (uintptr_t)ws % 256) {
        set_error("symmetrize workspace: need %zu bytes (256-aligned), got %zu", w.total, ws_bytes);
        return TDR_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t sentinel = ((uint64_t)n_local * (uint64_t)n_total) << 1;
    sym_emit_kernel<<<blocks_for(m), 256, 0, st>>>(P, idx, n_local, k, row0, n_total, ext_row, ext_col, ext_val,
                                                   n_ext, transpose_local, sentinel, w.k_in, w.v_in);
    TDR_LAUNCH_CHECK();
    size_t tmp = w.cub_bytes;
    TDR_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.k_in, w.k_out, w.v_in, w.v_out, m, 0,
                                             key_bits(sentinel), st));
    sym_heads_kernel<<<blocks_for(m + 1), 256, 0, st>>>(w.k_out, m, sentinel, w.head);
    tmp = w.cub_bytes;
    TDR_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tmp, w.head, w.pos, m + 1, st));
    sym_combine_kernel<<<blocks_for(m), 256, 0, st>>>(w.k_out, w.v_out, m, sentinel, n_total, w.head, w.pos, mode, col, val);
    sym_rowptr_kernel<<<blocks_for(n_local + 1), 256, 0, st>>>(w.k_out, m, n_local, n_total, w.pos, rowptr, nnz_out);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

extern "C" TDR_API int tdr_symmetrize_export_f32(const float* P, const int32_t* idx, int64_t n_local, int k, int64_t row0,
                                         int64_t n_total, int world, int rank, int64_t* send_counts,
                                         int64_t* out_row, int32_t* out_col, float* out_val,
                                         tdr_stream_t stream) {
    TDR_CHECK_ARG(P && idx && send_counts && out_row && out_col && out_val, "tdr_symmetrize_export_f32: null pointer");
    TDR_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "tdr_symmetrize_export_f32: bad rank/world");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nk = n_local * k;
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(send_counts);  // [2*world]
    TDR_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 2 * world, st));
    export_count_kernel<<<blocks_for(nk), 256, 0, st>>>(idx, nk, n_total, world, rank, counts);
    export_offsets_kernel<<<1, 1, 0, st>>>(counts, world);
    export_fill_kernel<<<blocks_for(nk), 256, 0, st>>>(P, idx, nk, k, row0, n_total, world, rank, counts + world,
                                                      out_row, out_col, out_val);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

extern "C" TDR_API int tdr_csr_to_ell_f32(const int64_t* rowptr, const int32_t* col, const float* val, int64_t n_local,
                                  int64_t width, float pad_val, float* ell_val, int64_t* ell_idx,
                                  tdr_stream_t stream) {
    TDR_CHECK_ARG(rowptr && col && val && ell_val && ell_idx, "tdr_csr_to_ell_f32: null pointer");
    if (n_local * width == 0) return TDR_OK;
    csr_to_ell_kernel<<<blocks_for(n_local * width), 256, 0, (cudaStream_t)stream>>>(rowptr, col, val, n_local, width,
                                                                                    pad_val, ell_val, ell_idx);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

extern "C" TDR_API int tdr_max_f32(const float* val, int64_t nnz, float* out, tdr_stream_t stream) {
    TDR_CHECK_ARG(val && out, "tdr_max_f32: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    TDR_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
    if (nnz == 0) return TDR_OK;
    const unsigned grid = (unsigned)min((int64_t)kNumSMs * 8, (nnz + 255) / 256);
    max_kernel<<<grid, 256, 0, st>>>(val, nnz, reinterpret_cast<int*>(out));
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

extern "C" TDR_API int tdr_umap_schedule_f32(const float* val, int64_t nnz, float a_max, int max_iter,
                                     float* epochs_per_sample, float* epoch_of_next_sample, tdr_stream_t stream) {
    TDR_CHECK_ARG(val && epochs_per_sample && epoch_of_next_sample, "tdr_umap_schedule_f32: null pointer");
    TDR_CHECK_ARG(max_iter >= 1, "tdr_umap_schedule_f32: max_iter must be >= 1");
    if (nnz == 0) return TDR_OK;
    const float thr = a_max / (float)max_iter;  // umap.py:219-220, fp32 division
    schedule_kernel<<<blocks_for(nnz), 256, 0, (cudaStream_t)stream>>>(val, nnz, a_max, thr, epochs_per_sample,
                                                                      epoch_of_next_sample);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

static size_t compact_layout(int64_t nnz, size_t* flag_off, size_t* pos_off, size_t* cub_off, size_t* cub_bytes) {
    size_t scan_b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_b, (int*)nullptr, (int64_t*)nullptr, nnz + 1);
    size_t off = 0;
    *flag_off = off;
    off += align_up((size_t)(nnz + 1) * 4, 256);
    *pos_off = off;
    off += align_up((size_t)(nnz + 1) * 8, 256);
    *cub_off = off;
    off += align_up(scan_b, 256);
    *cub_bytes = scan_b;
    return off;
}

extern "C" TDR_API size_t tdr_compact_workspace_bytes(int64_t n_local, int64_t nnz) {
    (void)n_local;
    size_t a, b, c, d;
    return compact_layout(nnz, &a, &b, &c, &d) + 256;
}

extern "C" TDR_API int tdr_umap_compact_f32(const int64_t* rowptr, const int32_t* col, const float* eps, int64_t n_local,
                                    int64_t nnz, int64_t* out_rowptr, int32_t* out_col, float* out_eps,
                                    float* out_eons, int64_t* nnz_out, void* ws, size_t ws_bytes,
                                    tdr_stream_t stream) {
    TDR_CHECK_ARG(rowptr && col && eps && out_rowptr && out_col && out_eps && out_eons, "tdr_umap_compact_f32: null pointer");
    TDR_CHECK_ARG(nnz < 0x7fffffffLL, "tdr_umap_compact_f32: nnz exceeds 2^31");
    size_t fo, po, co, cb;
    const size_t need = compact_layout(nnz, &fo, &po, &co, &cb);
    if (!ws || ws_bytes < need || (uintptr_t)ws % 256) {
        set_error("compact workspace: need %zu bytes (256-aligned), got %zu", need, ws_bytes);
        return TDR_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int* flag = (int*)((char*)ws + fo);
    int64_t* pos = (int64_t*)((char*)ws + po);
    finite_flag_kernel<<<blocks_for(nnz + 1), 256, 0, st>>>(eps, nnz, flag);
    TDR_CUDA(cub::DeviceScan::ExclusiveSum((char*)ws + co, cb, flag, pos, nnz + 1, st));
    compact_kernel<<<blocks_for(nnz ? nnz : 1), 256, 0, st>>>(col, eps, nnz, flag, pos, out_col, out_eps, out_eons);
    compact_rowptr_kernel<<<blocks_for(n_local + 1), 256, 0, st>>>(rowptr, n_local, pos, out_rowptr, nnz_out);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}
