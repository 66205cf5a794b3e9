// (i) Exact kNN on the 5th-gen tensor cores: tcgen05.mma (kind::f16) + TMEM + TMA, sm_100a.
//
// The only dense contraction of the path is the -2 X Y^T term of torchdr/distance/torch.py:89-91.
// fp32 accuracy is kept by an error-compensated split: every input is pre-scaled by a power of
// two s (so |x s| < 2^10) and written as x s = hi + lo with hi, lo in fp16 (11 significant bits
// each, 22 together).  Three MMA passes give  hi.hi  (exact products, accumulated in one TMEM
// accumulator) and  hi.lo + lo.hi  (2^-11 smaller, a second accumulator); the dropped lo.lo term
// is 2^-22 relative — below the fp32 rounding of the reference's own sgemm.  fp16 MMAs run at
// the full 16-bit tensor rate, i.e. 3 passes cost what 1.5 TF32 passes would.
//
// One CTA (256 or 384 threads, 1 per SM) owns 128 query rows: their hi/lo tiles stay resident in
// shared memory (TMA, 128B swizzle) for the whole sweep over the database, which streams through
// a ring of TMA stages.  A stage holds a whole 128-row database tile when two such stages fit next to the
// resident queries and the top-k lists (d = 128 / k <= 33, d <= 64), else ONE K-ATOM of a tile
// ([128 x 64] hi + lo = 32 KB): that is what extends the kernel to d = 256 / k = 33 and d = 128 / k = 96
// (the entropic k = 3 * perplexity of t-SNE / LargeVis).  Warp 0 = TMA producer, warp 1 = MMA issuer (one
// thread), warp 2 = TMEM allocator, warps 4-11 = two epilogue warpgroups: thread t owns TMEM lane t = query
// row t, reads its accumulator columns with tcgen05.ld, forms the distance and compares it with its row's
// running k-th best held in a register; the rare survivor goes into the row's UNSORTED list of 64-bit
// (distance, index) keys in shared memory, rank-sorted once after the sweep.  Two accumulator stages
// (2 x 256 TMEM columns) overlap the MMAs of tile t+1 with the filter of tile t.  Fused mode runs the UMAP
// rho/sigma search (rowsearch.cuh) on the finished rows before they leave the SM; dense mode
// (tdr_pairwise_full_f32) writes every distance of the sweep instead of filtering.
#include <cuda.h>
#include <cuda_fp16.h>

#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "rowsearch.cuh"

namespace tdr {
namespace tc {

constexpr int BM = 128, BN = 128, KATOM = 64;  // 64 fp16 = one 128-byte swizzle row
constexpr int TILE_BYTES = BM * KATOM * 2;  // 16 KB: one [128 x 64] fp16 box
constexpr int MAX_ATOMS = 4;                // d <= 256
constexpr int STAGE_BYTES = 2 * TILE_BYTES; // one ring stage: hi + lo of one atom of a database tile
constexpr int MAX_STAGES = 8;
constexpr size_t SMEM_LIMIT = 227 * 1024;
constexpr size_t SMEM_MISC = 512 + 1024;    // barriers + TMEM slot, 1024-byte alignment slack
constexpr int MAX_K = 96;                   // top-k lists in shared memory next to the operand tiles
constexpr int kMaxUnion = 2 * MAX_K / 32;  // entries per lane of the union of two lists

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// wait::ld with the destination registers as in/out operands, so that no use of them can be
// scheduled above the wait
__device__ __forceinline__ void tc_ld_wait(uint32_t (&a)[32], uint32_t (&b)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]),
                   "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(a[16]),
                   "+r"(a[17]), "+r"(a[18]), "+r"(a[19]), "+r"(a[20]), "+r"(a[21]), "+r"(a[22]), "+r"(a[23]), "+r"(a[24]),
                   "+r"(a[25]), "+r"(a[26]), "+r"(a[27]), "+r"(a[28]), "+r"(a[29]), "+r"(a[30]), "+r"(a[31]), "+r"(b[0]),
                   "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(b[8]), "+r"(b[9]),
                   "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15]), "+r"(b[16]), "+r"(b[17]),
                   "+r"(b[18]), "+r"(b[19]), "+r"(b[20]), "+r"(b[21]), "+r"(b[22]), "+r"(b[23]), "+r"(b[24]), "+r"(b[25]),
                   "+r"(b[26]), "+r"(b[27]), "+r"(b[28]), "+r"(b[29]), "+r"(b[30]), "+r"(b[31])
                 :
                 : "memory");
}

// K-major, 128B-swizzled [rows x 64 fp16] tile: SBO = 8 rows * 128 B, version 1 (sm_100), layout 2 (SW128)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16, A/B = F16 (0), D = F32 (1), both K-major, N = 128, M = 128
constexpr uint32_t kInstrDesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t kInstrDescN256 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ------------------------------------------------------------------ pre-pass: scale + fp16 split
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ X, int64_t n, int* __restrict__ out_bits) {
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(X[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_int(m));
}

__device__ __forceinline__ int scale_exponent(float absmax) {
    if (!(absmax > 0.0f) || absmax == INFINITY) return 0;
    return 9 - ilogbf(absmax);  // absmax * 2^e in [2^9, 2^10)
}

__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ X, int64_t n, int d, int dp, const int* __restrict__ absmax_bits,
             __half* __restrict__ hi, __half* __restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * dp) return;
    const int64_t r = i / dp;
    const int c = (int)(i - r * dp);
    float v = 0.0f;
    if (c < d) v = ldexpf(X[r * d + c], scale_exponent(__int_as_float(*absmax_bits)));
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
}

// norms in fp64 -> fp32; entries beyond n (up to the 128-padded length) are +inf so padded columns never qualify
__global__ void __launch_bounds__(256) sqnorm_pad_kernel(const float* __restrict__ X, int64_t n, int64_t n_pad, int d,
                                                         float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_pad) return;
    if (row >= n) {
        if (lane == 0) out[row] = INFINITY;
        return;
    }
    const float* x = X + row * d;
    double s = 0.0;
    for (int j = lane; j < d; j += 32) s = fma((double)x[j], (double)x[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = (float)s;
}

__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u64(uint32_t a, unsigned long long v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
// List entries are 64-bit keys  (order-preserving image of the fp32 distance) << 32 | column index : one unsigned
// compare orders entries by (distance, index) — the tie rule of every kernel of this engine.
__device__ __forceinline__ uint32_t ord_f32(float f) {
    const uint32_t b = __float_as_uint(f + 0.0f);  // -0 -> +0
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(uint32_t u) { return __uint_as_float(u ^ (((u >> 31) - 1u) | 0x80000000u)); }
constexpr unsigned long long kEmptyKey = ~0ull;  // sorts after every real entry

// Lists of k <= kListScanMaxK entries are kept UNSORTED in shared memory during the sweep.  While the list is not full a
// candidate is appended (two stores); once it is full the candidate overwrites the current worst entry (position
// amax) and one pass over the list finds the new worst by (distance, index).  The loads of that pass are
// independent, whereas a sorted insert walks a chain of dependent shared-memory round trips (~750 cycles per call at
// k = 15, ~4500 at k = 90 — round 1 measured the k = 90 search of t-SNE / LargeVis at 28x the k = 15 one): in the
// pruned sweep, where every tile is a near one, the inserts are what the epilogue spends its time on.
// Returns (ordered image of the worst distance << 32 | its position).  The lists are rank-sorted once after the sweep.
__device__ __noinline__ unsigned long long list_scan_max(uint32_t my_k, int k) {
    unsigned long long m = 0ull;
    int mp = 0;
#pragma unroll 4
    for (int p = 0; p < k; ++p) {
        const unsigned long long v = lds_u64(my_k + 8u * (uint32_t)p);
        const bool gt = v > m;
        m = gt ? v : m;
        mp = gt ? p : mp;
    }
    return (m & 0xffffffff00000000ull) | (unsigned long long)(uint32_t)mp;
}

// Longer lists (HEAP instantiations of the kernel): while the list is not full a candidate is appended (one store);
// the k-th append turns the list into an implicit binary MAX-heap of keys (slot 0 = the list's worst entry, Floyd's
// bottom-up construction, O(k)); from then on a candidate replaces the root and sifts down: <= floor(log2 k) levels
// of two independent loads each (6 at k = 90) instead of a pass over all k entries.  Measured on one box (`profiles/r2_knn_orders.txt`): k = 90,
// 1 M x 128: pruned sweep 147 -> 63 ms, full sweep 1262 -> 1012 ms; k = 15: the heap's 3 DEPENDENT round trips are
// slower than 15 independent loads (11.1 -> 12.4 ms), hence the split.  The lists are rank-sorted once after the
// sweep, so their internal order is free.
constexpr int kListScanMaxK = 32;
// sift_down: `key` is to be placed in the sub-heap rooted at hole i; returns what ends up in slot i.
__device__ __forceinline__ unsigned long long sift_down(uint32_t my_k, int k, int i, unsigned long long key) {
    unsigned long long top = key;
    const int i0 = i;
    for (;;) {
        const int l = 2 * i + 1;
        if (l >= k) break;
        const unsigned long long vl = lds_u64(my_k + 8u * (uint32_t)l);
        const unsigned long long vr = l + 1 < k ? lds_u64(my_k + 8u * (uint32_t)(l + 1)) : 0ull;
        const bool right = vr > vl;
        const unsigned long long vc = right ? vr : vl;
        if (vc <= key) break;
        sts_u64(my_k + 8u * (uint32_t)i, vc);
        if (i == i0) top = vc;
        i = right ? l + 1 : l;
    }
    sts_u64(my_k + 8u * (uint32_t)i, key);
    return top;
}
// Both return the list's worst key (the heap's root) after the operation.
__device__ __noinline__ unsigned long long list_replace_worst(uint32_t my_k, int k, unsigned long long key) {
    return sift_down(my_k, k, 0, key);
}
__device__ __noinline__ unsigned long long list_make_heap(uint32_t my_k, int k) {
    unsigned long long top = lds_u64(my_k);
    for (int s = (k - 2) >> 1; s >= 0; --s) top = sift_down(my_k, k, s, lds_u64(my_k + 8u * (uint32_t)s));
    return top;
}

// ------------------------------------------------------------------ main kernel
struct Params {
    int64_t nq, ndb, q_row0;  // q_row0: global id of query row 0 (self exclusion)
    int64_t q_tile_row0;      // row of query 0 inside the query split arrays
    const float* qn;
    const float* dbn;  // padded to a multiple of 128 with +inf
    const int* absmax_bits;
    int k, kpad, atoms, stages;
    int aps;  // K-atoms per ring stage: `atoms` (a stage = a whole database tile) when two such stages fit, else 1
    int exclude_self, metric, fused, max_iter;
    int minima_a;  // phase A keeps 32 running minima per thread instead of lists (k <= 32, see below)
    int dual;   // 1: two epilogue warpgroups (384 threads), each with its own top-k lists, on alternate tiles
    int debug;  // ablation bits for timing experiments (always 0 in the library): 1 = skip filter, 2 = skip MMAs, 4 = skip database TMA
    // tile-pruned sweep (see the "pruned sweep" section below): which database tiles this CTA visits
    int win;                 // > 0: only the tiles within +-win of the CTA's own rows (phase A)
    const int* tile_list;    // [gridDim.x][list_cap] ascending tile ids (phase B), or null = all tiles
    const int* tile_count;   // [gridDim.x]; a count above list_cap means "list overflowed: sweep everything"
    int list_cap;
    float* kth_out;          // phase A: only a bound on the k-th distance of every row is written (squared domain)
    const float* tau_seed;   // phase B: that bound; the row's threshold starts there instead of at +inf
    const int32_t* col_label;  // optional [ndb]: the id reported for a database row AND the tie-break key (null: the row index)
    unsigned long long* sweep_stats;  // optional: [0] += tiles swept, [1] += tiles of a full sweep
    float* out_dist;
    int32_t* out_idx;
    float* P;
    float* rho;
    float* sigma;
    // dense mode (tdr_pairwise_full_f32 on the tensor cores): every distance of the sweep is written to
    // c_full[row * ndb + col] instead of being filtered; no lists
    float* c_full;
    int full_exclude_diag;  // add 1e12 where the global query row equals the column (torch.py:111-116)
    // re-sweep of selected query tiles (robust mode, knn_tc_kernel<true> only): CTA b works on query tile
    // qtile_map[b]; CTAs >= *qtile_count exit.  Kept at the end so that the default kernel's argument layout is
    // the one that was verified on hardware.
    const int* qtile_map;
    const int* qtile_count;
};

// REDO: the robust mode's second sweep (CTA -> query tile through prm.qtile_map); a separate instantiation so that
// the default kernel keeps its register allocation (168, no spills)
// LAB: database labels (prm.col_label) are reported and rank distance ties; a separate instantiation so that the default
// kernel's filter carries none of it
// HEAP: candidate lists are max-heaps (k > kListScanMaxK), else unsorted lists with a re-scan
template <bool REDO, bool LAB = false, bool HEAP = false>
__global__ void __launch_bounds__(384, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
              const __grid_constant__ CUtensorMap map_db_hi, const __grid_constant__ CUtensorMap map_db_lo,
              const Params prm) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int atoms = prm.atoms, stages = prm.stages, k = prm.k, kpad = prm.kpad, aps = prm.aps;
    const int stage_bytes = aps * STAGE_BYTES;
    const int a_bytes = atoms * STAGE_BYTES;  // hi + lo
    unsigned char* a_tiles = smem;                    // [atoms][hi, lo][16 KB]
    unsigned char* b_tiles = smem + a_bytes;          // [stages][hi, lo][16 KB]: one atom of a database tile per stage
    const int nl = 1 + prm.dual;  // list sets (epilogue warpgroups)
    const int n_lists = prm.c_full ? 0 : nl;
    unsigned long long* lk_s = reinterpret_cast<unsigned long long*>(b_tiles + (size_t)stages * stage_bytes);  // [n_lists][128][kpad] keys
    uint64_t* bars = reinterpret_cast<uint64_t*>(lk_s + n_lists * BM * kpad);
    // barrier slots: 0 a_full | 1..S full | 1+S..2S empty | 2S+1, 2S+2 tmem_full | 2S+3, 2S+4 tmem_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 5);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int qtile_redo = 0;
    if (REDO) {  // uniform over the CTA, before any barrier or allocation
        if ((int)blockIdx.x >= __ldg(prm.qtile_count)) return;
        qtile_redo = __ldg(prm.qtile_map + blockIdx.x);
    }
#define TDR_QTILE (REDO ? (unsigned)qtile_redo : blockIdx.x)
    const int64_t q0 = (int64_t)TDR_QTILE * BM;
    const int64_t n_tiles = (prm.ndb + BN - 1) / BN;
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    // The sweep of this CTA: n_sweep database tiles, the t-th of which is tile_of(t) (ascending, so equal
    // distances still arrive in index order).  Every role derives it from the same kernel arguments.
    int64_t t_first = 0, n_sweep = n_tiles;
    const int* my_list = nullptr;
    if (prm.win > 0) {
        const int64_t r0 = prm.q_tile_row0 + q0;
        t_first = max((int64_t)0, r0 / BN - prm.win);
        n_sweep = min(n_tiles, (r0 + BM - 1) / BN + prm.win + 1) - t_first;
    } else if (prm.tile_count) {
        const int c = __ldg(prm.tile_count + TDR_QTILE);
        if (c <= prm.list_cap) {
            my_list = prm.tile_list + (int64_t)TDR_QTILE * prm.list_cap;
            n_sweep = c;
        }
    }
#undef TDR_QTILE
    auto tile_of = [&](int64_t t) -> int64_t { return my_list ? (int64_t)__ldg(my_list + t) : t_first + t; };
    if (prm.sweep_stats && tid == 0 && prm.win == 0) {
        atomicAdd(prm.sweep_stats + 0, (unsigned long long)n_sweep);
        atomicAdd(prm.sweep_stats + 1, (unsigned long long)n_tiles);
    }
    const int B_FULL = 1, B_EMPTY = 1 + stages, T_FULL = 1 + 2 * stages, T_EMPTY = 3 + 2 * stages;

    if (warp == 1 && lane == 0) {
        mbar_init(BAR(0), 1);
        for (int s = 0; s < stages; ++s) {
            mbar_init(BAR(B_FULL + s), 1);
            mbar_init(BAR(B_EMPTY + s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(BAR(T_FULL + s), 1);
            mbar_init(BAR(T_EMPTY + s), 128u * (uint32_t)nl);  // every epilogue thread of every warpgroup arrives
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < n_lists * BM * kpad; i += (int)blockDim.x) {
        lk_s[i] = kEmptyKey;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(BAR(0), (uint32_t)a_bytes);
            for (int a = 0; a < atoms; ++a) {
                tma_load_2d(smem_u32(a_tiles + (a * 2 + 0) * TILE_BYTES), &map_q_hi, BAR(0), a * KATOM,
                            (int)(prm.q_tile_row0 + q0));
                tma_load_2d(smem_u32(a_tiles + (a * 2 + 1) * TILE_BYTES), &map_q_lo, BAR(0), a * KATOM,
                            (int)(prm.q_tile_row0 + q0));
            }
            int64_t c = 0;  // ring slot counter: one slot per (tile, group of aps atoms)
            for (int64_t t = 0; t < n_sweep; ++t) {
                const int row_db = (int)(tile_of(t) * BN);
                for (int a0 = 0; a0 < atoms; a0 += aps, ++c) {
                    const int s = (int)(c % stages);
                    const uint32_t ph = (uint32_t)((c / stages) & 1);
                    mbar_wait(BAR(B_EMPTY + s), ph ^ 1u);
                    if ((prm.debug & 4) && c >= stages) {
                        mbar_arrive(BAR(B_FULL + s));
                        continue;
                    }
                    mbar_expect_tx(BAR(B_FULL + s), (uint32_t)stage_bytes);
                    unsigned char* dst = b_tiles + (size_t)s * stage_bytes;
                    for (int a = 0; a < aps; ++a) {
                        tma_load_2d(smem_u32(dst + (a * 2 + 0) * TILE_BYTES), &map_db_hi, BAR(B_FULL + s), (a0 + a) * KATOM, row_db);
                        tma_load_2d(smem_u32(dst + (a * 2 + 1) * TILE_BYTES), &map_db_lo, BAR(B_FULL + s), (a0 + a) * KATOM, row_db);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            mbar_wait(BAR(0), 0);
            int64_t c = 0;
            for (int64_t t = 0; t < n_sweep; ++t) {
                const int as = (int)(t & 1);
                const uint32_t aph = (uint32_t)((t >> 1) & 1);
                mbar_wait(BAR(T_EMPTY + as), aph ^ 1u);
                const uint32_t d_big = tmem_base + (uint32_t)(as * 256);
                const uint32_t d_small = d_big + 128u;
                for (int a0 = 0; a0 < atoms; a0 += aps, ++c) {
                    const int s = (int)(c % stages);
                    const uint32_t ph = (uint32_t)((c / stages) & 1);
                    mbar_wait(BAR(B_FULL + s), ph);
                    tc_fence_after();
                    for (int a = a0; a < a0 + aps && !(prm.debug & 2); ++a) {
                        const unsigned char* bt = b_tiles + (size_t)s * stage_bytes + (size_t)(a - a0) * STAGE_BYTES;
                        const uint64_t a_hi = make_smem_desc(smem_u32(a_tiles + (a * 2 + 0) * TILE_BYTES));
                        const uint64_t a_lo = make_smem_desc(smem_u32(a_tiles + (a * 2 + 1) * TILE_BYTES));
                        const uint64_t b_hi = make_smem_desc(smem_u32(bt));
#pragma unroll
                        for (int kk = 0; kk < KATOM / 16; ++kk) {
                            const uint64_t adv = (uint64_t)(kk * 2);  // 16 fp16 = 32 B -> +2 in the >>4 address field
                            const uint32_t acc = (a | kk) ? 1u : 0u;
                            // The kernel is bound by shared-memory operand bandwidth (A + B are both read from smem for
                            // every MMA).  hi.hi and hi.lo share the A operand: the database hi and lo tiles of an atom
                            // are adjacent in smem (16 KB + 16 KB = 256 K-major rows), so ONE N = 256 MMA produces
                            // [hi.hi | hi.lo] into TMEM columns [0,128) | [128,256) with a single read of A_hi.
                            tc_mma_f16(d_big, a_hi + adv, b_hi + adv, kInstrDescN256, acc);
                            tc_mma_f16(d_small, a_lo + adv, b_hi + adv, kInstrDesc, 1u);
                        }
                    }
                    tc_commit(BAR(B_EMPTY + s));  // ring slot reusable once these MMAs retire
                }
                tc_commit(BAR(T_FULL + as));  // accumulators ready for the epilogue
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: thread = query row =====================
        // One warp per scheduler, so latency is hidden by instruction-level parallelism: the 32
        // distances of a TMEM chunk are formed as independent chains (2 FADD + 1 FFMA + 1 FMNMX each),
        // only their minimum is compared with the row's running k-th best, and the next chunk's
        // tcgen05.ld is in flight meanwhile.  The per-element test runs only when that minimum wins.
        // warps 4-7 = warpgroup 0, warps 8-11 = warpgroup 1 (dual mode): a warp may only touch TMEM lanes
        // 32*(warp % 4)..+31, so both groups see all 128 rows; they take alternate tiles (= accumulator stages)
        // and keep separate lists, merged after the sweep.
        const int wg = (warp - 4) >> 2;
        const int row = (tid - 128) & 127;  // == TMEM lane
        const int64_t gq = q0 + row;
        const int64_t self = prm.exclude_self ? prm.q_row0 + gq : -1;
        const float qn = gq < prm.nq ? __ldg(prm.qn + gq) : 0.0f;
        const int e = scale_exponent(__int_as_float(__ldg(prm.absmax_bits)));
        const float neg2s = -2.0f * ldexpf(1.0f, -2 * e);  // power of two: the FFMA below rounds once, like sub(add, 2*dot)
        const uint32_t my_k = smem_u32(lk_s + (wg * BM + row) * kpad);
        // Threshold of the row.  Phase B of the pruned sweep starts it one ulp above the bound of phase A (so the
        // strict test below admits every candidate <= bound); the lists start empty either way, and since at least
        // k candidates of the swept tiles lie within the bound the union of the row's lists fills up.
        float tau = INFINITY;
        if (prm.tau_seed && gq < prm.nq) {
            const float sd = __ldg(prm.tau_seed + gq);
            if (sd < INFINITY)
                tau = sd >= 0.0f ? __uint_as_float(__float_as_uint(sd + 0.0f) + 1u) : __uint_as_float(__float_as_uint(sd) - 1u);
        }
        int amax = 0, cnt = 0;  // position of the list's worst entry (unsorted lists), entries inserted so far
        const int32_t* const labels = LAB ? prm.col_label : nullptr;
        int worst_lab = 0x7fffffff;  // label of the list's worst entry once the list is full (labels only)
        const uint32_t lane_addr = (uint32_t)(((warp - 4) & 3) * 32) << 16;

        auto filter = [&](uint32_t(&big)[32], uint32_t(&small)[32], int col_base) {
            float m0 = INFINITY, m1 = INFINITY, m2 = INFINITY, m3 = INFINITY;
#pragma unroll
            for (int j4 = 0; j4 < 32; j4 += 4) {
                const float4 nb = __ldg(reinterpret_cast<const float4*>(prm.dbn + col_base + j4));
                // torch.py:89-91: (|x|^2 + |y|^2) - 2 x.y
                const float d0 = fmaf(__fadd_rn(__uint_as_float(big[j4 + 0]), __uint_as_float(small[j4 + 0])), neg2s, __fadd_rn(qn, nb.x));
                const float d1 = fmaf(__fadd_rn(__uint_as_float(big[j4 + 1]), __uint_as_float(small[j4 + 1])), neg2s, __fadd_rn(qn, nb.y));
                const float d2 = fmaf(__fadd_rn(__uint_as_float(big[j4 + 2]), __uint_as_float(small[j4 + 2])), neg2s, __fadd_rn(qn, nb.z));
                const float d3 = fmaf(__fadd_rn(__uint_as_float(big[j4 + 3]), __uint_as_float(small[j4 + 3])), neg2s, __fadd_rn(qn, nb.w));
                big[j4 + 0] = __float_as_uint(d0);
                big[j4 + 1] = __float_as_uint(d1);
                big[j4 + 2] = __float_as_uint(d2);
                big[j4 + 3] = __float_as_uint(d3);
                m0 = fminf(m0, d0);
                m1 = fminf(m1, d1);
                m2 = fminf(m2, d2);
                m3 = fminf(m3, d3);
            }
            // With labels (a search in a re-ordered database whose result must not depend on that order) candidates that
            // TIE with the list's worst distance are ranked by label, so the kept set is the k smallest by
            // (distance, label) whatever the order of the columns; without labels the column index plays that role and
            // ascending arrival makes the strict test sufficient.
            const float mn = fminf(fminf(m0, m1), fminf(m2, m3));
            if (mn < tau || (LAB && cnt >= k && mn == tau)) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float dist = __uint_as_float(big[j]);
                    bool adm = dist < tau;
                    if (LAB && !adm && cnt >= k && dist == tau && col_base + j < (int)prm.ndb)
                        adm = __ldg(labels + col_base + j) < worst_lab;
                    if (adm && (int64_t)(col_base + j) != self) {
                        const int lab = LAB ? __ldg(labels + col_base + j) : col_base + j;
                        const unsigned long long key = ((unsigned long long)ord_f32(dist) << 32) | (unsigned long long)(uint32_t)lab;
                        if constexpr (HEAP) {
                            unsigned long long root;
                            if (cnt < k) {  // filling: append; the k-th entry completes the list and orders it as a heap
                                sts_u64(my_k + 8u * (uint32_t)cnt, key);
                                root = cnt + 1 == k ? list_make_heap(my_k, k) : kEmptyKey;
                            } else {
                                root = list_replace_worst(my_k, k, key);
                            }
                            if (++cnt >= k) {  // full: the root is a real entry, the row's running k-th best
                                tau = fminf(tau, unord_f32((uint32_t)(root >> 32)));
                                if (LAB) worst_lab = (int)(uint32_t)root;
                            }
                        } else {
                            sts_u64(my_k + 8u * (uint32_t)(cnt < k ? cnt : amax), key);
                            if (++cnt >= k) {
                                const unsigned long long r = list_scan_max(my_k, k);
                                tau = fminf(tau, unord_f32((uint32_t)(r >> 32)));
                                amax = (int)(uint32_t)r;
                                if (LAB) worst_lab = (int)(uint32_t)lds_u64(my_k + 8u * (uint32_t)amax);
                            }
                        }
                    }
                }
            }
        };

        const int c_beg = prm.dual ? 64 * wg : 0;   // first column of this warpgroup's share
        const int n_ch = prm.dual ? 2 : 4;          // 32-column chunks in the share
        if (prm.c_full) {
            // ---- dense mode: the whole tile of distances goes to global memory (row-major C[nq, ndb]); a thread
            // owns one row and writes 32 consecutive columns (128 B) per chunk
            const bool vec_ok = (prm.ndb & 3) == 0;
            const int64_t grow = prm.q_row0 + gq;  // global id of the query row (diagonal rule)
            for (int64_t t = 0; t < n_sweep; ++t) {
                const int as = (int)(t & 1);
                const uint32_t aph = (uint32_t)((t >> 1) & 1);
                mbar_wait(BAR(T_FULL + as), aph);
                tc_fence_after();
                const int col0 = (int)(tile_of(t) * BN) + c_beg;
                const uint32_t tb = tmem_base + lane_addr + (uint32_t)(as * 256 + c_beg);
#pragma unroll 1
                for (int c = 0; c < n_ch; ++c) {
                    uint32_t big[32], small[32];
                    tc_ld32(tb + 32 * c, big);
                    tc_ld32(tb + 128 + 32 * c, small);
                    tc_ld_wait(big, small);
                    const int col_base = col0 + 32 * c;
                    float dv[32];
#pragma unroll
                    for (int j4 = 0; j4 < 32; j4 += 4) {
                        const float4 nb = __ldg(reinterpret_cast<const float4*>(prm.dbn + col_base + j4));
                        const float nbv[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = j4 + u;
                            // torch.py:89-91: (|x|^2 + |y|^2) - 2 x.y
                            float dd = fmaf(__fadd_rn(__uint_as_float(big[j]), __uint_as_float(small[j])), neg2s,
                                            __fadd_rn(qn, nbv[u]));
                            if (prm.metric == TDR_METRIC_EUCLIDEAN) dd = sqrtf(fmaxf(dd, 0.0f));  // torch.py:92-95
                            if (prm.full_exclude_diag && (int64_t)(col_base + j) == grow) dd = __fadd_rn(dd, 1e12f);  // :111-116
                            dv[j] = dd;
                        }
                    }
                    if (gq < prm.nq) {
                        float* out = prm.c_full + gq * prm.ndb + col_base;
                        if (vec_ok && (int64_t)col_base + 32 <= prm.ndb) {
#pragma unroll
                            for (int j4 = 0; j4 < 32; j4 += 4)
                                *reinterpret_cast<float4*>(out + j4) = make_float4(dv[j4], dv[j4 + 1], dv[j4 + 2], dv[j4 + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if ((int64_t)col_base + j < prm.ndb) out[j] = dv[j];
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(BAR(T_EMPTY + as));
            }
        } else if (prm.kth_out && prm.minima_a) {
            // ---- phase A of the pruned sweep, k <= 32: no lists.  gm[j] = smallest distance among the columns this
            // thread sees at chunk position j; the row's 32 (x 2 warpgroups) minima belong to disjoint column sets,
            // so their k-th smallest bounds the k-th neighbour distance (k columns at most that far) — and it is
            // nearly tight: with 18+ columns per group the k-th of 64 group minima sits at about the same quantile
            // as the exact k-th of the window.
            float gm[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) gm[j] = INFINITY;
            for (int64_t t = 0; t < n_sweep; ++t) {
                const int as = (int)(t & 1);
                const uint32_t aph = (uint32_t)((t >> 1) & 1);
                mbar_wait(BAR(T_FULL + as), aph);
                tc_fence_after();
                const int col0 = (int)(tile_of(t) * BN) + c_beg;
                const uint32_t tb = tmem_base + lane_addr + (uint32_t)(as * 256 + c_beg);
#pragma unroll 1
                for (int c = 0; c < n_ch; ++c) {
                    uint32_t big[32], small[32];
                    tc_ld32(tb + 32 * c, big);
                    tc_ld32(tb + 128 + 32 * c, small);
                    tc_ld_wait(big, small);
                    const int col_base = col0 + 32 * c;
#pragma unroll
                    for (int j4 = 0; j4 < 32; j4 += 4) {
                        const float4 nb = __ldg(reinterpret_cast<const float4*>(prm.dbn + col_base + j4));
                        const float nbv[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = j4 + u;
                            float dd = fmaf(__fadd_rn(__uint_as_float(big[j]), __uint_as_float(small[j])), neg2s,
                                            __fadd_rn(qn, nbv[u]));
                            if ((int64_t)(col_base + j) == self) dd = INFINITY;
                            gm[j] = fminf(gm[j], dd);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(BAR(T_EMPTY + as));
            }
            // the last tile's MMAs have retired (T_FULL), so the TMA ring is free: park the minima there,
            // [j][warpgroup * 128 + row] (conflict-free), for the write-back warps
            float* scratch = reinterpret_cast<float*>(b_tiles);
#pragma unroll
            for (int j = 0; j < 32; ++j) scratch[j * (nl * BM) + wg * BM + row] = gm[j];
        } else {
        // Dual mode: BOTH warpgroups work on every tile, each on half of its columns.  With only two accumulator
        // stages in TMEM the MMAs of tile t+2 wait for the epilogue of tile t, so what matters is the epilogue
        // LATENCY per tile (measured: alternate-tile assignment gave (T_mma + 2 T_epi)/2 per tile).
        for (int64_t t = 0; t < n_sweep; ++t) {
            const int as = (int)(t & 1);
            const uint32_t aph = (uint32_t)((t >> 1) & 1);
            // database norms of the NEXT tile: pull this warpgroup's lines into L1 now, so that the filter's
            // broadcast loads do not each pay an L2 round trip
            if (lane < n_ch && t + 1 < n_sweep && !(prm.debug & 16))
                asm volatile("prefetch.global.L1 [%0];" ::"l"(prm.dbn + tile_of(t + 1) * BN + c_beg + lane * 32));
            mbar_wait(BAR(T_FULL + as), aph);
            tc_fence_after();
            if (prm.debug & 1) {
                tc_fence_before();
                mbar_arrive(BAR(T_EMPTY + as));
                continue;
            }
            const int col0 = (int)(tile_of(t) * BN) + c_beg;
            const uint32_t tb = tmem_base + lane_addr + (uint32_t)(as * 256 + c_beg);
            uint32_t bigA[32], smallA[32], bigB[32], smallB[32];
            tc_ld32(tb + 0, bigA);
            tc_ld32(tb + 128, smallA);
            tc_ld_wait(bigA, smallA);
#pragma unroll 1
            for (int c = 0; c < n_ch; c += 2) {
                tc_ld32(tb + 32 * (c + 1), bigB);
                tc_ld32(tb + 128 + 32 * (c + 1), smallB);
                filter(bigA, smallA, col0 + 32 * c);
                tc_ld_wait(bigB, smallB);
                if (c + 2 < n_ch) {
                    tc_ld32(tb + 32 * (c + 2), bigA);
                    tc_ld32(tb + 128 + 32 * (c + 2), smallA);
                }
                filter(bigB, smallB, col0 + 32 * (c + 1));
                if (c + 2 < n_ch) tc_ld_wait(bigA, smallA);
            }
            tc_fence_before();
            mbar_arrive(BAR(T_EMPTY + as));
        }
        }  // lists
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }

    // ---- write back; fused: rho/sigma search on the finished rows (one warp per row) ----
    const int n_warps = (int)blockDim.x >> 5;
    for (int row = warp; row < (prm.c_full ? 0 : BM); row += n_warps) {
        const int64_t gr = q0 + row;
        if (gr >= prm.nq) continue;
        // the row's storage in list 0: keys during the sweep, (fp32 distances | int32 indices), sorted, after the rank sort
        float* ldr = reinterpret_cast<float*>(lk_s + row * kpad);
        int* lir = reinterpret_cast<int*>(ldr + kpad);
        if (prm.kth_out && prm.minima_a) {
            // phase A, k <= 32: k-th smallest of the row's nl * 32 group minima (ranked by (value, group id))
            const float* scratch = reinterpret_cast<const float*>(b_tiles);
            float v[2];
            int rk[2] = {0, 0};
#pragma unroll
            for (int e = 0; e < 2; ++e) v[e] = e < nl ? scratch[lane * (nl * BM) + e * BM + row] : INFINITY;
            for (int q = 0; q < nl * 32; ++q) {
                const int qe = q >> 5, qj = q & 31;
                const float x = scratch[qj * (nl * BM) + qe * BM + row];
#pragma unroll
                for (int e = 0; e < 2; ++e) rk[e] += (x < v[e] || (x == v[e] && q < e * 32 + lane)) ? 1 : 0;
            }
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (e < nl && rk[e] == k - 1) prm.kth_out[gr] = v[e];
            continue;
        }
        {
            // rank-sort the union of the row's (unsorted) lists by (distance, index) into list 0: at most
            // 2 x 96 = 192 entries = kMaxUnion per lane.  Unused slots are (+inf, INT_MAX) and rank last.
            const int total = nl * k;
            const int epl = (total + 31) >> 5;
            unsigned long long kv[kMaxUnion];
            int rk[kMaxUnion];
#pragma unroll
            for (int e = 0; e < kMaxUnion; ++e) {
                const int q = lane + 32 * e;  // entry q: list q / k, slot q % k
                const bool ok = e < epl && q < total;
                const int li = ok ? q / k : 0, sl = ok ? q - li * k : 0;
                kv[e] = ok ? lk_s[(li * BM + row) * kpad + sl] : kEmptyKey;
                rk[e] = 0;
            }
            for (int li = 0; li < nl; ++li) {
                const unsigned long long* xs = lk_s + (li * BM + row) * kpad;
                for (int sl = 0; sl < k; ++sl) {
                    const unsigned long long x = xs[sl];
#pragma unroll
                    for (int e = 0; e < kMaxUnion; ++e)
                        if (e < epl) rk[e] += (x < kv[e]) ? 1 : 0;  // keys are distinct (distinct columns) except empty slots
                }
            }
            __syncwarp();  // every entry of the row is in registers: its storage can change layout
#pragma unroll
            for (int e = 0; e < kMaxUnion; ++e)
                if (e < epl && lane + 32 * e < total && rk[e] < k && kv[e] != kEmptyKey) {
                    ldr[rk[e]] = unord_f32((uint32_t)(kv[e] >> 32));
                    lir[rk[e]] = (int)(uint32_t)kv[e];
                }
            // slots that stay empty (fewer than k candidates: phase A windows) read as +inf
            {
                int filled = 0;
#pragma unroll
                for (int e = 0; e < kMaxUnion; ++e) filled += (e < epl && kv[e] != kEmptyKey) ? 1 : 0;
                filled = warp_sum_int(filled);
                for (int p = filled + lane; p < k; p += 32) {
                    ldr[p] = INFINITY;
                    lir[p] = 0x7fffffff;
                }
            }
            __syncwarp();
        }
        if (prm.kth_out) {  // phase A of the pruned sweep: the row's bound, nothing else
            if (lane == 0) prm.kth_out[gr] = ldr[k - 1];
            continue;
        }
        for (int p = lane; p < k; p += 32) {
            // lists are kept in the squared domain; euclidean = sqrt(clamp(., 0)) (torch.py:92-95) is monotone
            float dv = ldr[p];
            if (prm.metric == TDR_METRIC_EUCLIDEAN) dv = sqrtf(fmaxf(dv, 0.0f));
            if (prm.out_dist) prm.out_dist[gr * k + p] = dv;
            prm.out_idx[gr * k + p] = lir[p];
        }
        if (prm.fused) {
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
                default: run(std::integral_constant<int, 3>{}); break;
            }
        }
    }
}

// ------------------------------------------------------------------ pruned sweep
// Exact kNN without sweeping database tiles that cannot hold a neighbour.  Every 128-row tile gets an
// axis-aligned bounding box; phase A runs the kernel above over the +-kWindow tiles around each query tile,
// which gives every row an upper bound tau_i on its k-th neighbour distance; tile_prune_kernel keeps, per
// query tile A, the database tiles B with
//        sum_d max(0, lo_B[d] - hi_A[d], lo_A[d] - hi_B[d])^2  <=  max_{i in A} tau_i  (+ margin),
// i.e. drops B only if every point of B is farther from every point of A than a bound on all k-th neighbour
// distances of A; phase B runs the kernel over the surviving tiles (ascending order).  A dropped tile cannot
// change any list, so the result is bit-identical to the full sweep: each (row, column) distance is computed by
// the same instructions on the same operands, and the margin covers the fp32 error of those distances
// (documented gap 4e-6 (|x|^2 + |y|^2)).  Boxes, not balls: a tile that straddles two clusters has a
// huge ball but a box that is still thin in most dimensions (scripts/prune_sim.py: 9 of 782 tiles survive per
// query tile at 100 k x 128 with 100-point clusters; balls keep 780).  Data without index locality keeps
// every tile and costs phase A + the box test extra (measured in profiles/).
constexpr int kWindow = 4;       // phase A: own tile +- 4 (>= 5 full tiles = 640 rows >= k + 1)
constexpr int kPruneMinTiles = 64;
constexpr int kListCap = 1024;   // surviving tiles kept per query tile; beyond that the CTA sweeps everything

// lo/hi over rows [row0 + 128 b, row0 + 128 b + 128) /\ [0, row_limit).  ld_t > 0: transposed output [d][ld_t]
// (tiles b >= the last real one are written as empty boxes), ld_t == 0: row-major [n_tiles][d].
__global__ void __launch_bounds__(128) tile_box_kernel(const float* __restrict__ X, int64_t row0, int64_t row_limit,
                                                       int d, float* __restrict__ lo, float* __restrict__ hi,
                                                       int64_t ld_t) {
    const int64_t b = blockIdx.x;
    const int64_t r_beg = row0 + b * BM;
    const int64_t r_end = min(r_beg + BM, row_limit);
    for (int j = threadIdx.x; j < d; j += (int)blockDim.x) {
        float mn = INFINITY, mx = -INFINITY;
        for (int64_t r = r_beg; r < r_end; ++r) {
            const float v = __ldg(X + r * d + j);
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
        if (ld_t) {
            lo[(int64_t)j * ld_t + b] = mn;
            hi[(int64_t)j * ld_t + b] = mx;
        } else {
            lo[b * d + j] = mn;
            hi[b * d + j] = mx;
        }
    }
}

__global__ void __launch_bounds__(256) maxnorm_kernel(const float* __restrict__ nrm, int64_t n, int* __restrict__ out_bits) {
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, nrm[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_int(m));
}

// boxes of super-tiles (32 consecutive tiles = 4096 rows), same transposed layout [d][ld_s]
__global__ void __launch_bounds__(256) super_box_kernel(const float* __restrict__ lo_t, const float* __restrict__ hi_t,
                                                        int64_t ld_t, int64_t n_super, int64_t ld_s, int d,
                                                        float* __restrict__ slo_t, float* __restrict__ shi_t) {
    // one warp per (dimension, super-tile): lanes = the 32 member tiles (empty boxes beyond the last real tile)
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (int64_t)d * ld_s) return;
    const int64_t j = w / ld_s, sidx = w - j * ld_s;
    float mn = INFINITY, mx = -INFINITY;
    if (sidx < n_super) {
        mn = __ldg(lo_t + j * ld_t + sidx * 32 + lane);
        mx = __ldg(hi_t + j * ld_t + sidx * 32 + lane);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) {
        slo_t[j * ld_s + sidx] = mn;
        shi_t[j * ld_s + sidx] = mx;
    }
}

// One warp per query tile.  Two levels: lanes first test 32 consecutive SUPER-tiles (32 tiles = 4096 rows each)
// against the query box, then the member tiles of every surviving super-tile (lanes = its 32 tiles).  16
// dimensions at a time with a warp-uniform early exit (a far box is settled by its first few dimensions).
// Survivors are appended in ascending tile order.
// EXT: the robust mode's extras (strided / square-rooted bounds, outlier rejection, bound output, query-tile queue);
// EXT = false is the default path exactly as it was verified on hardware.
template <bool EXT>
__global__ void __launch_bounds__(256)
tile_prune_kernel(const float* __restrict__ qlo, const float* __restrict__ qhi, const float* __restrict__ dlo_t,
                  const float* __restrict__ dhi_t, const float* __restrict__ slo_t, const float* __restrict__ shi_t,
                  int64_t n_qtiles, int64_t n_tiles, int64_t ld_t, int64_t n_super, int64_t ld_s, int d,
                  const float* __restrict__ tau, int64_t nq, const int* __restrict__ maxnorm_bits,
                  int* __restrict__ list, int* __restrict__ count, int cap,
                  // robust mode (all zero / null in the default path):
                  int tau_stride, int tau_is_sqrt, int robust, float* __restrict__ bound_out,
                  const int* __restrict__ qtile_map, const int* __restrict__ qtile_count) {
    __shared__ float s_box[8][2][MAX_ATOMS * KATOM];
    __shared__ float s_tau[EXT ? 8 : 1][EXT ? BM : 1];
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t qt = (int64_t)blockIdx.x * 8 + warp;
    if (EXT && qtile_map) {
        if (qt >= (int64_t)__ldg(qtile_count)) return;  // whole warp
        qt = __ldg(qtile_map + qt);
    }
    if (qt >= n_qtiles) return;  // whole warp
    for (int j = lane; j < d; j += 32) {
        s_box[warp][0][j] = __ldg(qlo + qt * d + j);
        s_box[warp][1][j] = __ldg(qhi + qt * d + j);
    }
    float tm = 0.0f;
    if (!EXT) {
        for (int r = lane; r < BM; r += 32) {
            const int64_t gr = qt * BM + r;
            if (gr < nq) tm = fmaxf(tm, __ldg(tau + gr));
        }
    } else {
        const int64_t stride = tau_stride > 0 ? tau_stride : 1;
        for (int r = lane; r < BM; r += 32) {
            const int64_t gr = qt * BM + r;
            float v = -1.0f;  // rows beyond nq never decide anything
            if (gr < nq) {
                v = __ldg(tau + gr * stride);
                if (tau_is_sqrt) v = v * v;
                tm = fmaxf(tm, v);
            }
            s_tau[warp][r] = v;
        }
    }
    tm = warp_max(tm);
    if (EXT) __syncwarp();
    if (EXT && robust) {
        // Outlier rejection: a row whose bound is far above the tile's median (its neighbours are not inside the
        // phase-A window) must not decide what the other 127 rows sweep.  tm = largest bound <= 4 x the median; rows
        // above it are caught by knn_certify_kernel after the sweep and their tile is swept again.
        const int n_valid = (int)min((int64_t)BM, nq - qt * BM);
        float med = 0.0f;
        for (int e = 0; e < BM / 32; ++e) {
            const int me = lane + 32 * e;
            const float v = s_tau[warp][me];
            int rank = 0;
            for (int q = 0; q < n_valid; ++q) {
                const float x = s_tau[warp][q];
                rank += (x < v || (x == v && q < me)) ? 1 : 0;
            }
            const unsigned hit = __ballot_sync(FULL, me < n_valid && rank == n_valid / 2);
            if (hit) med = __shfl_sync(FULL, v, __ffs(hit) - 1);
        }
        float t2 = 0.0f;
        for (int r = lane; r < n_valid; r += 32) {
            const float v = s_tau[warp][r];
            if (v <= 4.0f * med) t2 = fmaxf(t2, v);
        }
        tm = warp_max(t2);
    }
    // margin: 1e-3 relative + 5 x the documented fp32 gap of the kernel's distances on the scale of the norms
    const float bound = fmaf(tm, 1.001f, 2e-5f * 2.0f * __int_as_float(__ldg(maxnorm_bits)));
    if (EXT && bound_out && lane == 0) bound_out[qt] = bound;
    __syncwarp();
    const float* lo_a = s_box[warp][0];
    const float* hi_a = s_box[warp][1];
    const unsigned lt_mask = (1u << lane) - 1u;
    // squared box distance of this lane's box (column `col` of the transposed arrays) from the query box; true =
    // "cannot hold a neighbour".  The early exit is warp-uniform, so all lanes run the same number of chunks.
    auto too_far = [&](const float* __restrict__ lo_t, const float* __restrict__ hi_t, int64_t ld, int64_t col) {
        float acc = 0.0f;
        for (int c = 0; c < d; c += 16) {
            const int c_end = min(c + 16, d);
#pragma unroll 4
            for (int j = c; j < c_end; ++j) {
                const float lb = __ldg(lo_t + (int64_t)j * ld + col), hb = __ldg(hi_t + (int64_t)j * ld + col);
                const float g = fmaxf(fmaxf(lb - hi_a[j], lo_a[j] - hb), 0.0f);
                acc = fmaf(g, g, acc);
            }
            if (__all_sync(FULL, acc * 0.9999f > bound)) break;
        }
        return acc * 0.9999f > bound;
    };
    int cnt = 0;
    int* my_list = list + qt * cap;
    for (int64_t s0 = 0; s0 < n_super; s0 += 32) {
        const int64_t sidx = s0 + lane;  // < ld_s (multiple of 32); super-tiles >= n_super hold empty boxes
        const bool s_far = too_far(slo_t, shi_t, ld_s, sidx);  // every lane takes part in its votes
        unsigned sm = __ballot_sync(FULL, sidx < n_super && !s_far);
        while (sm) {
            const int b = __ffs(sm) - 1;
            sm &= sm - 1;
            const int64_t t = (s0 + b) * 32 + lane;  // < ld_t; tiles >= n_tiles hold empty boxes
            const bool t_far = too_far(dlo_t, dhi_t, ld_t, t);
            const bool keep = t < n_tiles && !t_far;
            const unsigned m = __ballot_sync(FULL, keep);
            if (keep) {
                const int pos = cnt + __popc(m & lt_mask);
                if (pos < cap) my_list[pos] = (int)t;
            }
            cnt += __popc(m);
        }
    }
    if (lane == 0) count[qt] = cnt;
}

// Robust mode, after the sweep: row i is certified exact iff kth_i * 1.001 + margin <= the bound its tile was swept
// with — every database tile within that box distance of the query TILE was swept, and the distance of a point to a
// box is at least the distance between the boxes.  Query tiles with an uncertified row are queued for a second sweep.
__global__ void __launch_bounds__(256)
knn_certify_kernel(const float* __restrict__ out_dist, int k, int dist_is_sqrt, int64_t nq, int64_t n_qtiles,
                   const float* __restrict__ bound_used, const int* __restrict__ maxnorm_bits,
                   const int* __restrict__ tile_count, int cap, int* __restrict__ redo_map,
                   int* __restrict__ redo_count) {
    const int lane = threadIdx.x & 31;
    const int64_t qt = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (qt >= n_qtiles) return;
    if (__ldg(tile_count + qt) > cap) return;  // the list overflowed: this tile swept everything, it is exact
    const float margin = 2e-5f * 2.0f * __int_as_float(__ldg(maxnorm_bits));
    const float used = __ldg(bound_used + qt);
    bool bad = false;
    for (int r = lane; r < BM; r += 32) {
        const int64_t gr = qt * BM + r;
        if (gr >= nq) continue;
        float v = __ldg(out_dist + gr * k + (k - 1));
        if (dist_is_sqrt) v = v * v;
        bad |= !(fmaf(v, 1.001f, margin) <= used);  // NaN / inf count as uncertified
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) redo_map[atomicAdd(redo_count, 1)] = (int)qt;
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int make_map(CUtensorMap* m, const __half* base, int64_t rows, int dp) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return TDR_E_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)dp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)dp * 2};
    cuuint32_t box[2] = {(cuuint32_t)KATOM, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with code %d", (int)r);
        return TDR_E_CUDA;
    }
    return TDR_OK;
}

}  // namespace tc

// Shared-memory plan: resident query tile (atoms x 32 KB) + ring of >= 2 atom stages (32 KB each) + list sets + misc.
static bool tc_smem_plan(int d, int k, bool dense, int* dual_out, int* stages_out, size_t* smem_out, int* aps_out = nullptr) {
    using namespace tc;
    if (d > MAX_ATOMS * KATOM || (!dense && k > MAX_K)) return false;
    const int atoms = (d + KATOM - 1) / KATOM;
    const size_t q_bytes = (size_t)atoms * STAGE_BYTES;
    const size_t one_list = dense ? 0 : (size_t)BM * k * 8;
    auto fits = [&](int nl, int st) { return q_bytes + (size_t)st * STAGE_BYTES + nl * one_list + SMEM_MISC <= SMEM_LIMIT; };
    if (!fits(1, 2)) return false;
    // two epilogue warpgroups (each with its own list set) when that still leaves a ring of >= 3 atom stages
    const int dual = (dense || fits(2, 3)) ? 1 : 0;
    int atom_stages = (int)((SMEM_LIMIT - q_bytes - (1 + dual) * one_list - SMEM_MISC) / STAGE_BYTES);
    if (atom_stages > MAX_STAGES) atom_stages = MAX_STAGES;
    // a stage = a whole database tile (one barrier round trip per tile, as in round 1) when two of them fit; else K-atoms
    const int aps = atom_stages >= 2 * atoms ? atoms : 1;
    const int stages = atom_stages / aps;
    if (dual_out) *dual_out = dual;
    if (stages_out) *stages_out = stages;
    if (aps_out) *aps_out = aps;
    if (smem_out) *smem_out = q_bytes + (size_t)stages * aps * STAGE_BYTES + (1 + dual) * one_list + SMEM_MISC;
    return true;
}

bool knn_tc_supported(int d, int k) { return tc_smem_plan(d, k, false, nullptr, nullptr, nullptr); }
bool knn_tc_full_supported(int d) { return tc_smem_plan(d, 1, true, nullptr, nullptr, nullptr); }

namespace tc {
struct PruneLayout {
    int64_t n_tiles, ld_t, n_super, ld_s, n_qtiles;
    int cap;
    size_t box_t, sbox_t, qbox, tau, count, list, dist, bound, redo, total;
};
static PruneLayout prune_layout(int64_t nq, int64_t ndb, int d, int k) {
    PruneLayout L;
    L.n_tiles = (ndb + BN - 1) / BN;
    L.ld_t = (int64_t)align_up((size_t)L.n_tiles, 32);
    L.n_super = L.ld_t / 32;
    L.ld_s = (int64_t)align_up((size_t)L.n_super, 32);
    L.n_qtiles = (nq + BM - 1) / BM;
    L.cap = (int)std::min<int64_t>(L.n_tiles, kListCap);
    L.box_t = align_up((size_t)d * L.ld_t * 4, 256);
    L.sbox_t = align_up((size_t)d * L.ld_s * 4, 256);
    L.qbox = align_up((size_t)L.n_qtiles * d * 4, 256);
    L.tau = align_up((size_t)nq * 4, 256);
    L.count = align_up((size_t)L.n_qtiles * 4, 256);
    L.list = align_up((size_t)L.n_qtiles * L.cap * 4, 256);
    L.dist = align_up((size_t)nq * k * 4, 256);  // distances for the sigma/rho kernel when the caller wants none
    L.bound = align_up((size_t)L.n_qtiles * 4, 256);       // robust mode: bound every query tile was swept with
    L.redo = 256 + align_up((size_t)L.n_qtiles * 4, 256);  // robust mode: counter + query tiles to sweep again
    L.total = 256 + 2 * L.box_t + 2 * L.sbox_t + 2 * L.qbox + L.tau + L.count + L.list + L.dist + L.bound + L.redo;
    return L;
}
}  // namespace tc

size_t knn_tc_workspace_bytes(int64_t nq, int64_t ndb, int d, int k, bool same) {
    const int dp = (int)align_up((size_t)d, tc::KATOM);
    const int64_t ndb_pad = (int64_t)align_up((size_t)ndb, 128);
    const int64_t nq_pad = (int64_t)align_up((size_t)nq, 128);
    size_t b = 256;                                        // absmax
    b += align_up((size_t)ndb_pad * 4, 256);               // dbn (padded with +inf)
    b += 2 * align_up((size_t)ndb * dp * 2, 256);          // db hi, lo
    if (!same) {
        b += align_up((size_t)nq_pad * 4, 256);
        b += 2 * align_up((size_t)nq * dp * 2, 256);
    }
    // pruned sweep (queries inside the database only); part of the size whether or not the call asks for it
    if ((ndb + tc::BN - 1) / tc::BN >= tc::kPruneMinTiles) b += tc::prune_layout(nq, ndb, d, k).total;
    return b;
}

// one launch of the sweep kernel: instantiation by (second sweep of the certified mode, labels, list structure)
static cudaError_t launch_sweep(bool redo, const tc::Params& prm, unsigned grid, unsigned threads, size_t smem, cudaStream_t st,
                                const CUtensorMap& mq_hi, const CUtensorMap& mq_lo, const CUtensorMap& mdb_hi,
                                const CUtensorMap& mdb_lo) {
    using namespace tc;
    auto go = [&](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        kernel<<<grid, threads, smem, st>>>(mq_hi, mq_lo, mdb_hi, mdb_lo, prm);
        return cudaGetLastError();
    };
    if (prm.k > kListScanMaxK) {
        if (prm.col_label) return redo ? go(knn_tc_kernel<true, true, true>) : go(knn_tc_kernel<false, true, true>);
        return redo ? go(knn_tc_kernel<true, false, true>) : go(knn_tc_kernel<false, false, true>);
    }
    if (prm.col_label) return redo ? go(knn_tc_kernel<true, true>) : go(knn_tc_kernel<false, true>);
    return redo ? go(knn_tc_kernel<true, false>) : go(knn_tc_kernel<false, false>);
}

// `same`: the query rows are rows [q_row0, q_row0+nq) of the database buffer itself.
int knn_tc_launch(const float* Xq, int64_t nq, int64_t q_row0, const float* Xdb, int64_t ndb, int d, int k,
                  bool same, int exclude_self, int metric, int fused, int max_iter, float* out_dist, int32_t* out_idx,
                  float* P, float* rho, float* sigma, int prune, unsigned long long* sweep_stats,
                  const int32_t* db_labels, void* ws, size_t ws_bytes, cudaStream_t st) {
    using namespace tc;
    const size_t need = knn_tc_workspace_bytes(nq, ndb, d, k, same);
    if (!ws || ws_bytes < need || (uintptr_t)ws % 256) {
        set_error("knn (tensor-core path) workspace: need %zu bytes (256-aligned), got %zu", need, ws_bytes);
        return TDR_E_WORKSPACE;
    }
    const int dp = (int)align_up((size_t)d, KATOM);
    const int atoms = dp / KATOM;
    const int64_t ndb_pad = (int64_t)align_up((size_t)ndb, 128);
    const int64_t nq_pad = (int64_t)align_up((size_t)nq, 128);
    char* p = (char*)ws;
    int* absmax = (int*)p;
    p += 256;
    float* dbn = (float*)p;
    p += align_up((size_t)ndb_pad * 4, 256);
    __half* db_hi = (__half*)p;
    p += align_up((size_t)ndb * dp * 2, 256);
    __half* db_lo = (__half*)p;
    p += align_up((size_t)ndb * dp * 2, 256);
    float* qn = dbn + (same ? q_row0 : 0);
    __half *q_hi = db_hi, *q_lo = db_lo;
    if (!same) {
        qn = (float*)p;
        p += align_up((size_t)nq_pad * 4, 256);
        q_hi = (__half*)p;
        p += align_up((size_t)nq * dp * 2, 256);
        q_lo = (__half*)p;
    }
    TDR_CUDA(cudaMemsetAsync(absmax, 0, 4, st));
    const unsigned rgrid = (unsigned)std::min<int64_t>((int64_t)kNumSMs * 16, (ndb * d + 255) / 256);
    absmax_kernel<<<rgrid, 256, 0, st>>>(Xdb, ndb * d, absmax);
    if (!same) {
        const unsigned qgrid = (unsigned)std::min<int64_t>((int64_t)kNumSMs * 16, (nq * d + 255) / 256);
        absmax_kernel<<<qgrid, 256, 0, st>>>(Xq, nq * d, absmax);
    }
    split_kernel<<<(unsigned)((ndb * dp + 255) / 256), 256, 0, st>>>(Xdb, ndb, d, dp, absmax, db_hi, db_lo);
    sqnorm_pad_kernel<<<(unsigned)((ndb_pad + 7) / 8), 256, 0, st>>>(Xdb, ndb, ndb_pad, d, dbn);
    if (!same) {
        split_kernel<<<(unsigned)((nq * dp + 255) / 256), 256, 0, st>>>(Xq, nq, d, dp, absmax, q_hi, q_lo);
        sqnorm_pad_kernel<<<(unsigned)((nq_pad + 7) / 8), 256, 0, st>>>(Xq, nq, nq_pad, d, qn);
    }
    TDR_LAUNCH_CHECK();

    CUtensorMap mq_hi, mq_lo, mdb_hi, mdb_lo;
    int rc;
    const int64_t q_rows = same ? ndb : nq;
    if ((rc = make_map(&mq_hi, q_hi, q_rows, dp)) || (rc = make_map(&mq_lo, q_lo, q_rows, dp)) ||
        (rc = make_map(&mdb_hi, db_hi, ndb, dp)) || (rc = make_map(&mdb_lo, db_lo, ndb, dp)))
        return rc;

    Params prm{};
    prm.nq = nq;
    prm.ndb = ndb;
    prm.q_row0 = q_row0;
    prm.q_tile_row0 = same ? q_row0 : 0;
    prm.qn = qn;
    prm.dbn = dbn;
    prm.absmax_bits = absmax;
    prm.k = k;
    prm.kpad = k;
    prm.atoms = atoms;
    prm.exclude_self = exclude_self;
    prm.metric = metric;
    prm.fused = fused;
    prm.max_iter = max_iter;
    prm.debug = 0;  // ablation bits (skip filter / MMAs / TMA), set by hand in timing experiments only
    prm.col_label = db_labels;
    prm.out_dist = out_dist;
    prm.out_idx = out_idx;
    prm.P = P;
    prm.rho = rho;
    prm.sigma = sigma;
    prm.minima_a = k <= 32 ? 1 : 0;
    int stages = 0;
    size_t smem = 0;
    if (!tc_smem_plan(d, k, false, &prm.dual, &stages, &smem, &prm.aps)) {
        set_error("knn (tensor-core path): shared memory budget exceeded for d=%d k=%d", d, k);
        return TDR_E_UNSUPPORTED;
    }
    if ((prm.debug & 32) && prm.dual) {  // ablation: one epilogue warpgroup
        prm.dual = 0;
        smem -= (size_t)BM * k * 8;
    }
    prm.stages = stages;
    // per call: the attribute belongs to the current device's context (a process may drive several GPUs)
    const unsigned grid = (unsigned)((nq + BM - 1) / BM);
    const unsigned threads = prm.dual ? 384 : 256;

    // ---- pruned sweep: boxes -> phase A (window) -> surviving-tile lists; the launch below is phase B
    const int64_t n_tiles = (ndb + BN - 1) / BN;
    if (prune && same && n_tiles >= kPruneMinTiles && !prm.debug) {
        const PruneLayout L = prune_layout(nq, ndb, d, k);
        char* w = p;  // same: p is the end of the split buffers
        if ((size_t)(w - (char*)ws) + L.total > ws_bytes) {
            set_error("knn (tensor-core path) workspace: pruned sweep needs %zu more bytes", L.total);
            return TDR_E_WORKSPACE;
        }
        int* maxnorm = (int*)w;
        w += 256;
        float* dlo_t = (float*)w;
        w += L.box_t;
        float* dhi_t = (float*)w;
        w += L.box_t;
        float* slo_t = (float*)w;
        w += L.sbox_t;
        float* shi_t = (float*)w;
        w += L.sbox_t;
        float* qlo = (float*)w;
        w += L.qbox;
        float* qhi = (float*)w;
        w += L.qbox;
        float* tau = (float*)w;
        w += L.tau;
        int* count = (int*)w;
        w += L.count;
        int* list = (int*)w;
        w += L.list;
        float* dist_scratch = (float*)w;
        w += L.dist;
        float* bound_used = (float*)w;
        w += L.bound;
        int* redo_count = (int*)w;
        int* redo_map = (int*)(w + 256);
        const int robust = prune == 2 ? 1 : 0;
        TDR_CUDA(cudaMemsetAsync(maxnorm, 0, 4, st));
        tile_box_kernel<<<(unsigned)L.ld_t, 128, 0, st>>>(Xdb, 0, ndb, d, dlo_t, dhi_t, L.ld_t);
        super_box_kernel<<<(unsigned)(((int64_t)d * L.ld_s * 32 + 255) / 256), 256, 0, st>>>(dlo_t, dhi_t, L.ld_t, L.n_super,
                                                                                             L.ld_s, d, slo_t, shi_t);
        tile_box_kernel<<<(unsigned)L.n_qtiles, 128, 0, st>>>(Xdb, q_row0, q_row0 + nq, d, qlo, qhi, 0);
        maxnorm_kernel<<<(unsigned)std::min<int64_t>((int64_t)kNumSMs * 8, (ndb + 255) / 256), 256, 0, st>>>(dbn, ndb, maxnorm);
        Params pa = prm;
        pa.win = kWindow;
        pa.kth_out = tau;
        pa.fused = 0;
        pa.out_dist = nullptr;
        pa.out_idx = nullptr;
        TDR_CUDA(launch_sweep(false, pa, grid, threads, smem, st, mq_hi, mq_lo, mdb_hi, mdb_lo));
        const unsigned pgrid = (unsigned)((L.n_qtiles + 7) / 8);
        if (robust)
            tile_prune_kernel<true><<<pgrid, 256, 0, st>>>(qlo, qhi, dlo_t, dhi_t, slo_t, shi_t, L.n_qtiles, L.n_tiles, L.ld_t,
                                                           L.n_super, L.ld_s, d, tau, nq, maxnorm, list, count, L.cap, 1, 0, 1,
                                                           bound_used, nullptr, nullptr);
        else
            tile_prune_kernel<false><<<pgrid, 256, 0, st>>>(qlo, qhi, dlo_t, dhi_t, slo_t, shi_t, L.n_qtiles, L.n_tiles,
                                                            L.ld_t, L.n_super, L.ld_s, d, tau, nq, maxnorm, list, count, L.cap,
                                                            1, 0, 0, nullptr, nullptr, nullptr);
        TDR_LAUNCH_CHECK();
        prm.tau_seed = tau;
        prm.tile_list = list;
        prm.tile_count = count;
        prm.list_cap = L.cap;
        prm.sweep_stats = sweep_stats;
        if (robust) {
            // TDR_KNN_PRUNE_CERTIFIED (for inputs brought into a locality-creating order, torchdr_b200/reorder.py): thresholds that ignore outlier rows, then
            // certify every row against the bound its tile was swept with and sweep the uncertified tiles again with
            // the (by then tight) k-th distances found.  Device-side queue, no host synchronisation.
            const int fused_req = prm.fused;
            prm.fused = 0;
            if (!prm.out_dist) prm.out_dist = dist_scratch;
            TDR_CUDA(launch_sweep(false, prm, grid, threads, smem, st, mq_hi, mq_lo, mdb_hi, mdb_lo));
            TDR_CUDA(cudaMemsetAsync(redo_count, 0, 4, st));
            const int is_sqrt = metric == TDR_METRIC_EUCLIDEAN ? 1 : 0;
            knn_certify_kernel<<<pgrid, 256, 0, st>>>(prm.out_dist, k, is_sqrt, nq, L.n_qtiles, bound_used, maxnorm, count,
                                                      L.cap, redo_map, redo_count);
            tile_prune_kernel<true><<<pgrid, 256, 0, st>>>(qlo, qhi, dlo_t, dhi_t, slo_t, shi_t, L.n_qtiles, L.n_tiles, L.ld_t,
                                                           L.n_super, L.ld_s, d, prm.out_dist + (k - 1), nq, maxnorm, list,
                                                           count, L.cap, k, is_sqrt, 0, nullptr, redo_map, redo_count);
            Params pr = prm;
            pr.qtile_map = redo_map;
            pr.qtile_count = redo_count;
            pr.tau_seed = nullptr;
            TDR_CUDA(launch_sweep(true, pr, grid, threads, smem, st, mq_hi, mq_lo, mdb_hi, mdb_lo));
            TDR_LAUNCH_CHECK();
            if (fused_req) return tdr_umap_affinity_f32(prm.out_dist, nq, k, max_iter, P, rho, sigma, (tdr_stream_t)st);
            return TDR_OK;
        }
        if (prm.fused) {
            // With a handful of tiles per CTA the in-kernel sigma/rho search (12 warps per SM walking dependent
            // bisection chains) would take longer than the sweep itself: 8.4 ms of 18.8 ms at 1 M x 128, against
            // 3.5 ms for the standalone row kernel at full occupancy.  Same arithmetic, bit-identical rows.
            prm.fused = 0;
            if (!prm.out_dist) prm.out_dist = dist_scratch;
            TDR_CUDA(launch_sweep(false, prm, grid, threads, smem, st, mq_hi, mq_lo, mdb_hi, mdb_lo));
            return tdr_umap_affinity_f32(prm.out_dist, nq, k, max_iter, P, rho, sigma, (tdr_stream_t)st);
        }
    }
    TDR_CUDA(launch_sweep(false, prm, grid, threads, smem, st, mq_hi, mq_lo, mdb_hi, mdb_lo));
    return TDR_OK;
}

// Dense distance matrix C[n, m] on the tensor cores (tdr_pairwise_full_f32 for d <= 256): same mainloop, the epilogue
// writes every distance instead of filtering.  Workspace = knn_tc_workspace_bytes(n, m, d, 1, same).
int knn_tc_full_launch(const float* X, int64_t n, const float* Y, int64_t m, int d, bool same, int metric,
                       int exclude_diag, float* C, void* ws, size_t ws_bytes, cudaStream_t st) {
    using namespace tc;
    const size_t need = knn_tc_workspace_bytes(n, m, d, 1, same);
    if (!ws || ws_bytes < need || (uintptr_t)ws % 256) {
        set_error("pairwise (tensor-core path) workspace: need %zu bytes (256-aligned), got %zu", need, ws_bytes);
        return TDR_E_WORKSPACE;
    }
    const int dp = (int)align_up((size_t)d, KATOM);
    const int64_t m_pad = (int64_t)align_up((size_t)m, 128);
    const int64_t n_pad = (int64_t)align_up((size_t)n, 128);
    char* p = (char*)ws;
    int* absmax = (int*)p;
    p += 256;
    float* dbn = (float*)p;
    p += align_up((size_t)m_pad * 4, 256);
    __half* db_hi = (__half*)p;
    p += align_up((size_t)m * dp * 2, 256);
    __half* db_lo = (__half*)p;
    p += align_up((size_t)m * dp * 2, 256);
    float* qn = dbn;
    __half *q_hi = db_hi, *q_lo = db_lo;
    if (!same) {
        qn = (float*)p;
        p += align_up((size_t)n_pad * 4, 256);
        q_hi = (__half*)p;
        p += align_up((size_t)n * dp * 2, 256);
        q_lo = (__half*)p;
    }
    TDR_CUDA(cudaMemsetAsync(absmax, 0, 4, st));
    absmax_kernel<<<(unsigned)std::min<int64_t>((int64_t)kNumSMs * 16, (m * d + 255) / 256), 256, 0, st>>>(Y, m * d, absmax);
    if (!same)
        absmax_kernel<<<(unsigned)std::min<int64_t>((int64_t)kNumSMs * 16, (n * d + 255) / 256), 256, 0, st>>>(X, n * d, absmax);
    split_kernel<<<(unsigned)((m * dp + 255) / 256), 256, 0, st>>>(Y, m, d, dp, absmax, db_hi, db_lo);
    sqnorm_pad_kernel<<<(unsigned)((m_pad + 7) / 8), 256, 0, st>>>(Y, m, m_pad, d, dbn);
    if (!same) {
        split_kernel<<<(unsigned)((n * dp + 255) / 256), 256, 0, st>>>(X, n, d, dp, absmax, q_hi, q_lo);
        sqnorm_pad_kernel<<<(unsigned)((n_pad + 7) / 8), 256, 0, st>>>(X, n, n_pad, d, qn);
    }
    TDR_LAUNCH_CHECK();
    CUtensorMap mq_hi, mq_lo, mdb_hi, mdb_lo;
    int rc;
    if ((rc = make_map(&mq_hi, q_hi, same ? m : n, dp)) || (rc = make_map(&mq_lo, q_lo, same ? m : n, dp)) ||
        (rc = make_map(&mdb_hi, db_hi, m, dp)) || (rc = make_map(&mdb_lo, db_lo, m, dp)))
        return rc;
    Params prm{};
    prm.nq = n;
    prm.ndb = m;
    prm.qn = qn;
    prm.dbn = dbn;
    prm.absmax_bits = absmax;
    prm.k = 1;
    prm.kpad = 1;
    prm.atoms = dp / KATOM;
    prm.metric = metric;
    prm.c_full = C;
    prm.full_exclude_diag = (exclude_diag && same) ? 1 : 0;
    size_t smem = 0;
    if (!tc_smem_plan(d, 1, true, &prm.dual, &prm.stages, &smem, &prm.aps)) {
        set_error("pairwise (tensor-core path): d=%d exceeds %d", d, MAX_ATOMS * KATOM);
        return TDR_E_UNSUPPORTED;
    }
    TDR_CUDA(cudaFuncSetAttribute(knn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    knn_tc_kernel<false><<<(unsigned)((n + BM - 1) / BM), prm.dual ? 384 : 256, smem, st>>>(mq_hi, mq_lo, mdb_hi, mdb_lo, prm);
    TDR_LAUNCH_CHECK();
    return TDR_OK;
}

}  // namespace tdr
