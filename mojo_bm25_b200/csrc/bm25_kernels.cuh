// bm25_kernels.cuh -- hand-written sm_100a kernels of the BM25 query hot path.
//
// Path (reference bm25_native.py:129-158): for every query, gather the posting slices of its
// terms from the CSC index, accumulate the per-(term, doc) weights into per-document scores in
// query-term order (fp32, one add per posting -- bit-identical to the reference's csc mat-vec),
// and select the top-k by (score desc, doc id asc).
//
// Kernels:
//   k_relayout           load time: CSC postings -> padded, 16-byte aligned, sentinel-terminated posting lists
//   k_build_table        load time: per (heavy term, document tile) first-posting table
//   k_quantize / k_pack  compressed index: bf16-rounded weights, 4-byte packed postings (16-bit tile-local slot | bf16)
//   k_term_bounds[_exact] load time: per-term weight order statistics (threshold priming)
//   k_segments           per (query, term): threshold priming; cursor starts of the light terms for
//                        every document chunk (binary search on doc id inside the term's posting list)
//   k_query_order        per batch: queries with the same heaviest term become neighbours (L2 sharing);
//                        normally launched together with k_segments as k_segments_order
//   k_score_topk_s       (query width <= 32) / k_score_topk (any width): one warp per (query, document
//                        chunk): private shared-memory score tile, in-order accumulation (heavy terms:
//                        table-addressed 16-byte vector loads, PTX read-modify-write; light terms:
//                        cursors with cp.async-prefetched heads), hot-list tile end with st.bulk
//                        clear, threshold-pruned top-k into the CTA's candidate buffer; emits k
//                        64-bit keys per (query, CTA)
//   k_merge[_large]      per query: merge candidate lists (tile ranges or GPU shards), zero-score
//                        fill, unpack to (doc id, score); _large: k above BM25_SMALL_K, in global memory
//   k_scores_dense       parity/debug: dense [Q, n_docs] score slab
//   k_validate_*         load-time canonical-form checks of the CSC arrays
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bm25 {

typedef unsigned long long u64;

constexpr int kThreads = 512;          // threads per CTA of k_merge / k_scores_dense
constexpr unsigned kFull = 0xffffffffu;
constexpr int kBoundLevels = 11;       // per-term weight order statistics at ranks 1, 2, 4, ..., 1024
constexpr int kBoundSample = 1024;     // postings sampled per term for those statistics

// ---------------------------------------------------------------------------------------------
// 64-bit candidate keys: high word = order-preserving image of the fp32 score, low word =
// ~doc id, so that a larger key means (higher score, then lower doc id).  0 is never a valid key.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_to_ord(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord_to_f32(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ u64 make_key(float score, uint32_t doc) {
    return ((u64)f32_to_ord(score) << 32) | (u64)(0xffffffffu - doc);
}
__host__ __device__ __forceinline__ uint32_t key_doc(u64 key) { return 0xffffffffu - (uint32_t)key; }
__host__ __device__ __forceinline__ float key_score(u64 key) { return ord_to_f32((uint32_t)(key >> 32)); }

constexpr int kDocNone = 0x7fffffff;  // sentinel doc id: padding postings, exhausted cursors
constexpr int kSelectMin = 256;       // candidate sets larger than this are compacted by radix select
constexpr int kStateInts = 8;         // ints of cursor state per (warp, query term) in k_score_topk
constexpr int kPkMaxTileDocs = 8192;  // compressed postings: the tile-local byte offset 4 * (doc mod S) is a uint16

// ---------------------------------------------------------------------------------------------
// The index in HBM ("re-bucketed into document-range tiles", built at load time):
//   ids / w     posting arrays in the PADDED layout: every term's posting list starts at a multiple
//               of 4 elements (16 bytes) and is followed by 1..4 sentinel postings (kDocNone, 0.0)
//               up to the next multiple of 4, so that 16-byte vector loads never straddle two
//               terms and a cursor that runs off a list reads a sentinel instead of needing an end.
//   tptr[t]     {start, end} of term t's list in the padded arrays (end - start = df_t).
//   term_row[t] row of term t in the tile table, or -1.  "Heavy" terms (df_t >= heavy_min postings
//               per document tile on average) own a row; "light" terms are walked with cursors.
//   tab[r][j]   for heavy row r and tile j in [0, n_tiles]: index (padded layout) of the first
//               posting whose doc id is >= j * tile_docs.  Tile j's postings of that term are
//               [tab[r][j], tab[r][j+1]) -- known WITHOUT looking at any doc id, so the score
//               kernel's loads never depend on previously loaded postings.
// ---------------------------------------------------------------------------------------------

// k_relayout (load time): copy CSC postings into the padded layout.  One CTA per term (grid-stride).
__global__ void __launch_bounds__(128) k_relayout(const int32_t* __restrict__ indptr, const int32_t* __restrict__ ids_in,
                                                  const float* __restrict__ w_in, const int2* __restrict__ tptr,
                                                  int n_terms, int32_t* __restrict__ ids_out,
                                                  float* __restrict__ w_out, float2* __restrict__ wrange) {
    __shared__ float s_mn[4], s_mx[4];
    for (int t = blockIdx.x; t < n_terms; t += gridDim.x) {
        const int s = indptr[t];
        const int n = indptr[t + 1] - s;
        const int o = tptr[t].x;
        const int padded = (n + 4) & ~3;  // >= 1 sentinel after the last posting
        float mn = INFINITY, mx = -INFINITY;
        for (int i = threadIdx.x; i < padded; i += blockDim.x) {
            const float wv = i < n ? w_in[s + i] : 0.f;
            ids_out[o + i] = i < n ? ids_in[s + i] : kDocNone;
            w_out[o + i] = wv;
            if (i < n) {
                mn = fminf(mn, wv);
                mx = fmaxf(mx, wv);
            }
        }
        // wrange[t] = {smallest, largest} weight of the term ({+inf, -inf} for an empty list)
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(kFull, mn, o2));
            mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o2));
        }
        if ((threadIdx.x & 31) == 0) {
            s_mn[threadIdx.x >> 5] = mn;
            s_mx[threadIdx.x >> 5] = mx;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            wrange[t] = make_float2(fminf(fminf(s_mn[0], s_mn[1]), fminf(s_mn[2], s_mn[3])),
                                    fmaxf(fmaxf(s_mx[0], s_mx[1]), fmaxf(s_mx[2], s_mx[3])));
        __syncthreads();
    }
}

// k_build_table (load time / when tile_docs changes): tab[r][j] by binary search on doc id.
__global__ void __launch_bounds__(256) k_build_table(const int2* __restrict__ tptr,
                                                     const int32_t* __restrict__ heavy_terms,
                                                     const int32_t* __restrict__ ids, int64_t n_entries,
                                                     int n_tiles, int tile_docs, int32_t* __restrict__ tab) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries) return;
    const int64_t row = i / (n_tiles + 1);
    const int j = (int)(i - row * (n_tiles + 1));
    const int2 se = tptr[heavy_terms[row]];
    int lo = se.x, hi = se.y;
    if (j == 0) hi = lo;
    else if (j == n_tiles) lo = hi;
    else {
        const int64_t target = (int64_t)j * tile_docs;
        while (lo < hi) {
            const int mid = lo + ((hi - lo) >> 1);
            if ((int64_t)__ldg(ids + mid) < target) lo = mid + 1; else hi = mid;
        }
    }
    tab[i] = lo;
}

// ---------------------------------------------------------------------------------------------
// Compressed index (16-bit tile-local doc ids + 16-bit weights, 4 bytes per posting).
// k_quantize (once, bm25_index_compress): rounds every weight to bf16 (round to nearest even) IN
//   PLACE -- from then on the handle IS the quantised index: every kernel, the order statistics and
//   the weight ranges see the rounded weights, and results are bit-identical to the reference
//   run on the rounded CSC matrix.  bf16 rather than fp16: widening is a 16-bit shift (exact, full
//   rate) and the exponent range is fp32's.  Recomputes wrange; flags[0] += weights that are no
//   longer > 0, flags[1] += weights that are no longer finite.
// k_pack (with the tile table): pk[p] = (4 * (doc mod S)) << 16 | bf16 bits, padding -> 0xffff0000.
//   The tile a posting belongs to is implied by its position (tile table), so the doc id needs only
//   its offset inside the tile; stored as the byte offset of the fp32 slot.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t bf16_round_bits(uint32_t u) {  // finite input
    return (u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u;
}
__global__ void __launch_bounds__(128) k_quantize(const int2* __restrict__ tptr, int n_terms, float* __restrict__ w,
                                                  float2* __restrict__ wrange, unsigned long long* flags) {
    __shared__ float s_mn[4], s_mx[4];
    unsigned long long bad0 = 0, bad1 = 0;
    for (int t = blockIdx.x; t < n_terms; t += gridDim.x) {
        const int2 se = tptr[t];
        float mn = INFINITY, mx = -INFINITY;
        for (int i = se.x + threadIdx.x; i < se.y; i += blockDim.x) {
            const float r = __uint_as_float(bf16_round_bits(__float_as_uint(w[i])));
            w[i] = r;
            if (!(r > 0.f)) ++bad0;
            if (!(fabsf(r) <= 3.402823466e38f)) ++bad1;
            mn = fminf(mn, r);
            mx = fmaxf(mx, r);
        }
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(kFull, mn, o2));
            mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o2));
        }
        if ((threadIdx.x & 31) == 0) {
            s_mn[threadIdx.x >> 5] = mn;
            s_mx[threadIdx.x >> 5] = mx;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            wrange[t] = make_float2(fminf(fminf(s_mn[0], s_mn[1]), fminf(s_mn[2], s_mn[3])),
                                    fmaxf(fmaxf(s_mx[0], s_mx[1]), fmaxf(s_mx[2], s_mx[3])));
        __syncthreads();
    }
    if (bad0) atomicAdd(flags + 0, bad0);
    if (bad1) atomicAdd(flags + 1, bad1);
}
__global__ void __launch_bounds__(256) k_pack(const int32_t* __restrict__ ids, const float* __restrict__ w, int64_t n,
                                              int tile_docs, uint32_t* __restrict__ pk) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int d = ids[i];
        uint32_t v = 0xffff0000u;
        if (d != kDocNone) v = ((uint32_t)(4 * (d % tile_docs)) << 16) | (__float_as_uint(w[i]) >> 16);
        pk[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// k_segments: per (query, term) one warp.
//  * threshold priming (lane 0): theta_q[q] = max over the query's terms of the 2^level-th largest
//    weight of the term (see k_term_bounds).
//  * cursor starts: seg[(q*(n_rows+1) + c)*T + t] = first posting index of term queries[q,t] whose
//    doc id is >= c*row_docs (index into the padded ids/w); row n_rows is the end of the list.
//    Row-major so that one row's T boundaries are contiguous.  Lanes stride over the rows (binary
//    search on doc id inside the term's list).  Terms with a row in the tile table are skipped
//    when `term_row` is given: the score kernel never reads their cursor starts.
// ---------------------------------------------------------------------------------------------
struct SegArgs {
    const int2* __restrict__ tptr;
    const int32_t* __restrict__ term_row;
    const int32_t* __restrict__ ids;
    const int32_t* __restrict__ queries;
    int64_t n_qt;
    int T, n_terms, row_docs, n_rows;
    int32_t* __restrict__ seg;
    const float* __restrict__ bounds;
    int level;
    u64* __restrict__ theta_q;
};
__device__ __forceinline__ void segments_warp(const SegArgs& g, int64_t warp, int lane) {
    const int2* __restrict__ tptr = g.tptr;
    const int32_t* __restrict__ term_row = g.term_row;
    const int32_t* __restrict__ ids = g.ids;
    const int32_t* __restrict__ queries = g.queries;
    const int T = g.T, n_terms = g.n_terms, row_docs = g.row_docs, n_rows = g.n_rows, level = g.level;
    int32_t* __restrict__ seg = g.seg;
    const float* __restrict__ bounds = g.bounds;
    u64* __restrict__ theta_q = g.theta_q;
    const int term = queries[warp];
    const int64_t q = warp / T;
    const int t = (int)(warp - q * T);
    int lo0 = 0, hi0 = 0;
    bool heavy = false;
    if (term >= 0 && term < n_terms) {
        const int2 se = tptr[term];
        lo0 = se.x;
        hi0 = se.y;
        heavy = term_row != nullptr && term_row[term] >= 0;
    }
    // threshold priming: the 2^level-th largest weight of this term belongs to 2^level distinct
    // documents whose score is at least that weight (all weights > 0), so it bounds the k-th best
    // score of the query from below for every k <= 2^level.
    if (lane == 0 && bounds != nullptr && hi0 > lo0) {
        const float b = __ldg(bounds + (int64_t)term * kBoundLevels + level);
        if (b > 0.f) atomicMax(theta_q + q, make_key(b, 0xffffffffu) - 1ull);
    }
    if (heavy) return;
    int32_t* out = seg + q * (int64_t)(n_rows + 1) * T + t;
    for (int j = lane; j <= n_rows; j += 32) {
        int res;
        if (j == 0) res = lo0;
        else if (j == n_rows) res = hi0;
        else {
            const int64_t target = (int64_t)j * row_docs;
            int lo = lo0, hi = hi0;
            while (lo < hi) {
                const int mid = lo + ((hi - lo) >> 1);
                if ((int64_t)__ldg(ids + mid) < target) lo = mid + 1; else hi = mid;
            }
            res = lo;
        }
        out[(int64_t)j * T] = res;
    }
}

__global__ void __launch_bounds__(256) k_segments(const SegArgs g) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= g.n_qt) return;
    segments_warp(g, warp, threadIdx.x & 31);
}

// ---------------------------------------------------------------------------------------------
// group barriers: __syncthreads (barrier 0), or named barrier 1 over the warps of k_score_topk --
// its warps reach the (rare) candidate-compaction rounds from different places in their tile loops
// ---------------------------------------------------------------------------------------------
struct CtaGroup {
    int size, rank;
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
struct WarpsGroup {
    int size, rank;
    __device__ __forceinline__ void sync() const {
        asm volatile("bar.sync 1, %0;" ::"r"(size) : "memory");
    }
};

// group-wide bitonic sort (descending) of P = 2^m keys in shared memory.  A power-of-two number of
// warps each own a contiguous slice of >= 64 keys: every stage whose compare distance stays inside
// a slice needs only __syncwarp, so a sort of 512 keys by 8 warps meets at ~8 group barriers
// instead of 45.
template <typename G>
__device__ __forceinline__ void bitonic_sort_desc(u64* buf, int P, const G& g) {
    const int lane = g.rank & 31, warp = g.rank >> 5;
    const int nw = g.size >> 5;
    int nwu = 1;
    while (nwu * 2 <= nw && P / (nwu * 2) >= 64) nwu <<= 1;
    const int slice = P / nwu;  // keys per participating warp
    const int half = slice >> 1;
    bool prev_global = true;
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const bool global = stride >= slice;
            if (global || prev_global) g.sync(); else __syncwarp();
            prev_global = global;
            if (warp < nwu) {
                for (int m = lane; m < half; m += 32) {
                    const int i = warp * half + m;
                    const int a = 2 * i - (i & (stride - 1));
                    const int b = a + stride;
                    const bool desc = ((a & size) == 0);
                    const u64 x = buf[a], y = buf[b];
                    if ((x < y) == desc) { buf[a] = y; buf[b] = x; }
                }
            }
        }
    }
    g.sync();
}

// Keep the k best of the n = min(*s_ncand, cap) candidates (sorted, best first) and raise the
// threshold.  Called by every thread of the group with no push in flight.
template <typename G>
__device__ __forceinline__ void compact_candidates(u64* cand, int cap, int k, u64 theta0, int* s_ncand,
                                                   u64* s_theta, const G& g) {
    const int n = min(*s_ncand, cap);
    int P = 2;
    while (P < n) P <<= 1;
    for (int i = n + g.rank; i < P; i += g.size) cand[i] = 0;
    g.sync();
    bitonic_sort_desc(cand, P, g);
    if (g.rank == 0) {
        *s_ncand = n < k ? n : k;
        *s_theta = (n >= k) ? cand[k - 1] : theta0;
    }
    g.sync();
}

// Candidate sets above kSelectMin keys: radix select instead of a full sort.  Finds the k-th largest of
// the n > k distinct keys in cand[0..n) (most significant byte first, one 256-bin shared-memory
// histogram per pass, stops as soon as the bucket holding the k-th key has a single member), then
// moves the k keys >= that threshold -- unsorted -- through `scratch` (k keys of global memory owned
// by this CTA) back to cand[0..k).  O(n) per pass instead of O(n log^2 n) compare-exchanges.
// hist: 264 ints of shared memory.  Called by every thread of the group with no push in flight.
template <typename G>
__device__ __forceinline__ void select_candidates(u64* cand, int n, int k, u64* scratch, int* hist, int* s_ncand,
                                                  u64* s_theta, bool copy_back, const G& g) {
    const int lane = g.rank & 31;
    u64 prefix = 0, known = 0;  // known = mask of the key bits fixed so far
    int need = k;
    int* ctl = hist + 256;  // {digit, need, bucket, keeper count}
    if (g.rank == 0) ctl[3] = 0;  // ordered before its first use by the barriers of the passes
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = g.rank; i < 256; i += g.size) hist[i] = 0;
        g.sync();
        for (int i = g.rank; i < n; i += g.size) {
            const u64 key = cand[i];
            if ((key & known) == prefix) atomicAdd(&hist[(int)(key >> shift) & 255], 1);
        }
        g.sync();
        if (g.rank < 32) {  // lane L owns bins 255-8L .. 248-8L (descending)
            int c[8], s = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j] = hist[255 - 8 * lane - j]; s += c[j]; }
            int cum = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kFull, cum, o);
                if (lane >= o) cum += v;
            }
            const unsigned hit = __ballot_sync(kFull, cum >= need);
            if (lane == __ffs(hit) - 1) {
                int above = cum - s;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (above + c[j] >= need) {
                        ctl[0] = 255 - 8 * lane - j;
                        ctl[1] = need - above;
                        ctl[2] = c[j];
                        break;
                    }
                    above += c[j];
                }
            }
        }
        g.sync();
        prefix |= (u64)ctl[0] << shift;
        known |= 0xffull << shift;
        need = ctl[1];
        if (ctl[2] == 1) break;  // the k-th key is the only one left with this prefix
    }
    // the threshold itself, and the move of the keepers
    u64 theta = 0;
    for (int i = g.rank; i < n; i += g.size) {
        const u64 key = cand[i];
        if ((key & known) == prefix) theta = key;  // exactly one thread sees it (or all 64 bits are known)
    }
    if (theta) *s_theta = theta;
    g.sync();
    theta = *s_theta;
    for (int i = g.rank; i < n; i += g.size) {
        const u64 key = cand[i];
        if (key >= theta) scratch[atomicAdd(&ctl[3], 1)] = key;
    }
    g.sync();
    if (copy_back)
        for (int i = g.rank; i < k; i += g.size) cand[i] = scratch[i];
    if (g.rank == 0) *s_ncand = k;
    g.sync();
}

// ---------------------------------------------------------------------------------------------
// k_query_order: cross-query sharing of posting lists through the L2.  Queries whose heaviest
// (largest-df) term is the same are given neighbouring CTA indices, so that they are resident at
// the same time and walk that term's posting list together: the second and later readers of a
// list segment hit the 126 MB L2 instead of HBM (batched scoring as a sparse query-by-term times
// term-by-doc product, reference bm25_native.py:160-192, without changing the per-query sums).
// One CTA; key = (heaviest term id << 32 | query); bitonic sort in shared memory (P <= 4096) or
// in global memory; perm[i] = query run by the i-th group of CTAs.
// ---------------------------------------------------------------------------------------------
struct OrderArgs {
    const int2* __restrict__ tptr;
    const int32_t* __restrict__ queries;
    int Q, T, n_terms, P;
    u64* __restrict__ keys_g;
    int32_t* __restrict__ perm;
};
__device__ __forceinline__ void query_order_cta(const OrderArgs& o) {  // one CTA of 1024 threads
    const int2* __restrict__ tptr = o.tptr;
    const int32_t* __restrict__ queries = o.queries;
    const int Q = o.Q, T = o.T, n_terms = o.n_terms, P = o.P;
    u64* __restrict__ keys_g = o.keys_g;
    int32_t* __restrict__ perm = o.perm;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* keys = (P <= 4096) ? reinterpret_cast<u64*>(smem_raw) : keys_g;
    const int tid = threadIdx.x;
    for (int q = tid; q < P; q += 1024) {
        u64 key = ~0ull;
        if (q < Q) {
            int best = 0x7fffffff, best_df = -1;
            for (int t = 0; t < T; ++t) {
                const int term = queries[(int64_t)q * T + t];
                if (term < 0 || term >= n_terms) continue;
                const int2 se = tptr[term];
                const int df = se.y - se.x;
                if (df > best_df || (df == best_df && term < best)) { best_df = df; best = term; }
            }
            key = ((u64)(uint32_t)best << 32) | (uint32_t)q;
        }
        keys[q] = key;
    }
    __syncthreads();
    if (P <= 1024) {
        // one key per thread (threads beyond P hold the largest key): compare distances below 32 are
        // warp shuffles, only the 15 steps with distance >= 32 go through shared memory and a barrier
        u64 x = tid < P ? keys[tid] : ~0ull;
        __syncthreads();
        for (int size = 2; size <= 1024; size <<= 1) {
            const bool asc = (tid & size) == 0;
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                u64 y;
                if (stride >= 32) {
                    keys[tid] = x;
                    __syncthreads();
                    y = keys[tid ^ stride];
                    __syncthreads();
                } else {
                    y = __shfl_xor_sync(kFull, x, stride);
                }
                const bool lower = (tid & stride) == 0;
                x = (lower == asc) ? (x < y ? x : y) : (x < y ? y : x);
            }
        }
        if (tid < Q) perm[tid] = (int32_t)(uint32_t)x;
        return;
    }
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (P >> 1); i += 1024) {
                const int a = 2 * i - (i & (stride - 1));
                const int b = a + stride;
                const bool asc = ((a & size) == 0);
                const u64 x = keys[a], y = keys[b];
                if ((x > y) == asc) { keys[a] = y; keys[b] = x; }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < Q; i += 1024) perm[i] = (int32_t)(uint32_t)keys[i];
}

// k_segments and the batch ordering in ONE launch (they are independent): the last CTA orders the batch
// while the others (32 warps each) prime the thresholds and search the cursor starts.
__global__ void __launch_bounds__(1024) k_segments_order(const SegArgs g, const OrderArgs o) {
    if (blockIdx.x == gridDim.x - 1) {
        query_order_cta(o);
        return;
    }
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= g.n_qt) return;
    segments_warp(g, warp, threadIdx.x & 31);
}

struct SearchArgs {
    const int32_t* __restrict__ ids;       // [nnz_padded] doc ids (padded layout, see above)
    const float* __restrict__ w;           // [nnz_padded] weights
    const int32_t* __restrict__ term_row;  // [V] tile-table row of a heavy term, -1 for a light term
    const int32_t* __restrict__ tab;       // [n_heavy, n_tiles+1] tile table
    const int32_t* __restrict__ queries;   // [Q,T]
    const int32_t* __restrict__ qperm;     // [Q] query run by each group of `splits` CTAs (k_query_order), or NULL
    const int32_t* __restrict__ seg;       // [Q,n_chunks+1,T] cursor starts of the light terms (k_scores_dense: [Q,n_tiles+1,T], all terms)
    u64* __restrict__ partial;             // [Q,splits,k] keys (0 = none)
    u64* theta_q;                          // [Q] best known k-th key per query (shared by its CTAs)
    u64* cand_global;                      // [Q*splits, cap] candidate buffers in global memory (large k), or NULL
    float* __restrict__ dense_out;         // [Q,n_docs]  (k_scores_dense only)
    u64 theta0;                            // initial threshold: key must be > theta0 to compete
    int Q, T, k;
    int n_terms;
    int n_docs, tile_docs, n_tiles;        // tile_docs = S, documents per warp tile
    int tiles_per_chunk, n_chunks;         // a chunk = the tiles one warp walks
    int splits, tiles_per_split, cap;      // splits = CTAs per query (tiles_per_split: k_scores_dense)
    int general;                           // 1: zero-score docs compete (weights may be <= 0)
    int no_hot;                            // 1: always use the dense tile scan (A/B switch)
    int poison;                            // 1 (debug): fill the dynamic shared memory with 0xff before use
    int sp_major;                          // 1: CTA index = split * Q + query slot (else query slot * splits + split)
    int bulk_clear;                        // 1: clear the score tile with st.bulk (UMEMSETS) instead of vector stores
    const float2* __restrict__ wrange;     // [n_terms] {smallest, largest} weight per term (k_relayout)
    const uint32_t* __restrict__ pk;       // [nnz_padded] 4-byte packed postings (compressed index, k_pack), or NULL
    int no_epoch;                          // 1: zero the score tile after every tile (no exponent epochs)
};

// ---------------------------------------------------------------------------------------------
// k_scores_dense (parity/debug): CTA = (query, range of document tiles); direct global loads,
// shared-memory score tile, terms strictly in query order, writes the dense [Q, n_docs] slab.
// shared memory: float score[tile_docs] | int seg_lo[T] | int seg_hi[T]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2) k_scores_dense(const SearchArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sc = reinterpret_cast<float*>(smem_raw);
    int* s_lo = reinterpret_cast<int*>(smem_raw + (size_t)a.tile_docs * sizeof(float));
    int* s_hi = s_lo + a.T;
    const int tid = threadIdx.x;
    const int q = blockIdx.x / a.splits;
    const int sp = blockIdx.x - q * a.splits;
    const int j0 = sp * a.tiles_per_split;
    const int j1 = min(a.n_tiles, j0 + a.tiles_per_split);
    for (int i = tid; i < a.tile_docs; i += kThreads) sc[i] = 0.f;
    const int32_t* segq = a.seg + (int64_t)q * a.T * (a.n_tiles + 1);
    __syncthreads();
    for (int j = j0; j < j1; ++j) {
        const int base = j * a.tile_docs;
        const int nd = min(a.tile_docs, a.n_docs - base);
        for (int t = tid; t < a.T; t += kThreads) {
            s_lo[t] = __ldg(segq + (int64_t)j * a.T + t);
            s_hi[t] = __ldg(segq + (int64_t)(j + 1) * a.T + t);
        }
        __syncthreads();
        for (int t = 0; t < a.T; ++t) {
            const int lo = s_lo[t], hi = s_hi[t];
            if (lo >= hi) continue;  // uniform
            for (int i = lo + tid; i < hi; i += kThreads) sc[__ldg(a.ids + i) - base] += __ldg(a.w + i);
            __syncthreads();
        }
        float* out = a.dense_out + (int64_t)q * a.n_docs + base;
        for (int i = tid; i < nd; i += kThreads) { out[i] = sc[i]; sc[i] = 0.f; }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// k_score_topk (v5): the hot kernel.  Document-at-a-time streaming, one WARP per worker, over the
// re-bucketed index.
//
// CTA = (query, group of NCW document chunks); warp w of the CTA owns chunk sp*NCW + w, a
// contiguous range of `tiles_per_chunk` document tiles of S = tile_docs documents, and a private
// fp32 score tile of S slots in shared memory.  For every tile the warp walks the query's terms
// strictly in query order and adds each posting's weight into the tile: one fp32 add per posting,
// a term has at most one posting per document and one warp handles every posting of a document,
// so __syncwarp between terms is all the ordering needed -- no atomics, no CTA barriers,
// bit-identical to the reference's csc mat-vec (bm25_native.py:152).
//
//  * heavy terms (a row in the tile table): the tile's posting range [lo, hi) comes from the table,
//    whose entries for the next tiles are prefetched into the warp's state ring with 4-byte
//    cp.async (LDGSTS) two tiles ahead.  The range is streamed in pieces of 128 postings, four
//    consecutive postings per lane through two 16-byte loads (ids, weights); the piece after the
//    current one -- of this term or of the next heavy term of the tile -- is requested before the
//    current one is added.  No load address depends on a loaded doc id.  Postings of the
//    neighbouring tiles (or padding) that the 16-byte granularity drags in fail the range test
//    (unsigned)(doc - base) < S.
//  * light terms: a cursor with the next two postings (doc id, weight) already resident in the
//    state ring (refilled with cp.async when the cursor moves), so deciding that a light term has
//    nothing in the tile costs one shared-memory read, and its postings are added by lane 0
//    without waiting for global memory.
//
// After the last term the warp pushes the documents that beat the running k-th best key into the
// CTA's candidate buffer -- from the short list of slots whose running score reached the threshold
// during the adds, followed by a write-only clear of the tile, or (list overflow / non-positive
// weights) by a 16-byte vector scan + zero of all S slots.  The buffer is shared by the NCW warps;
// when it overflows they meet in a (rare) named-barrier round, keep the k best and raise the
// threshold, which is also published per query in global memory so that the other CTAs of the
// query prune with it.
// shared memory (dynamic):
//   float score[NCW][S] | u64 cand[cap] (unless in global memory) | int state[NCW][kStateInts][T]
//   | uint16 hot[NCW][kHotCap]
// state (per warp, SoA over the T query terms):
//   heavy term: rows 0..3 = ring of tile-table entries (entry of tile j in row j & 3), row 4 = table row
//   light term: row 0 = cursor, 1 = end, 2/3 = next two doc ids, 4/5 = their weights
//   row 7 = 1 for a heavy term, 0 for a light one
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct TopkState {
    u64* cand;
    int* s_ncand;
    int* s_overflow;
    int cap;
    int general;    // zero / negative scores compete: the raw-score pre-filter must not clamp at 0
    u64 theta;      // key must be > theta to compete
    float theta_f;  // cheap pre-filter on the raw score (never stricter than theta)
    float scale;    // the tile holds score * scale (exponent epochs, a power of two; 1 when unused)
    float floor;    // scaled tile values below floor * scale are leftovers of earlier epochs, i.e. zero
    __device__ __forceinline__ void set_theta(u64 t) {
        theta = t;
        if (t == 0ull) theta_f = -INFINITY;
        else theta_f = general ? key_score(t) : fmaxf(key_score(t), 1.401298464e-45f);
    }
    // returns false when the candidate buffer is full (the caller keeps the score in place)
    __device__ __forceinline__ bool push(float v, uint32_t doc) {
        const u64 key = make_key(v, doc);
        if (key > theta) {
            const int pos = atomicAdd(s_ncand, 1);
            if (pos < cap) cand[pos] = key;
            else { *reinterpret_cast<volatile int*>(s_overflow) = 1; return false; }
        }
        return true;
    }
};

// Per-warp list of the tile slots whose running score reached the pre-filter threshold while
// postings were being added ("hot" documents; weights are > 0 on this path so scores only grow and
// every document that ends above the threshold is recorded, possibly more than once).  While the
// list fits, the tile epilogue visits only these slots and then clears the tile with plain
// 16-byte stores instead of reading and testing all S scores.
constexpr int kHotCap = 64;
struct HotList {
    unsigned short* slots;  // [kHotCap] tile-local ids
    int n;                  // warp-uniform; > kHotCap = overflowed or disabled -> dense scan
    unsigned lt;            // lanemask_lt
    __device__ __forceinline__ void reset(bool enabled) { n = enabled ? 0 : kHotCap + 1; }
    __device__ __forceinline__ bool active() const { return n <= kHotCap; }
    __device__ __forceinline__ void add(bool hot, int slot) {  // warp-collective
        const unsigned m = __ballot_sync(kFull, hot);
        if (m == 0u) return;
        const int c = __popc(m);
        if (n + c <= kHotCap) {
            if (hot) slots[n + __popc(m & lt)] = (unsigned short)slot;
            n += c;
        } else {
            n = kHotCap + 1;
        }
    }
};

// ---- explicit shared-memory accessors (32-bit shared-window addresses): the hot loop never
// ---- re-derives a generic pointer
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lds_i32(unsigned a) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds_f32(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_i32(unsigned a, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_zero16(unsigned a) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0) : "memory");
}
// Blackwell bulk zero-fill of shared memory (PTX st.bulk, SASS UMEMSETS): one warp-uniform
// instruction instead of bytes/512 vector stores through the LSU pipe.  bytes: multiple of 8.
__device__ __forceinline__ void smem_bulk_zero(unsigned a, unsigned bytes) {
    asm volatile("st.bulk.weak.shared::cta [%0], %1, 0;" ::"r"(a), "l"((unsigned long long)bytes) : "memory");
}
__device__ __forceinline__ void cp_async4_s(unsigned smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

// A piece of one heavy term's tile range: 128 consecutive postings starting at a multiple of 4,
// four per lane (lane L holds postings p0 + 4L .. p0 + 4L + 3), fetched with two 16-byte loads.
struct PostingPiece {
    int4 d;
    float4 w;
    __device__ __forceinline__ void load(const int32_t* __restrict__ ids, const float* __restrict__ wts, int p0, int hi,
                                         int lane) {
        const int idx = p0 + 4 * lane;
        d = make_int4(kDocNone, kDocNone, kDocNone, kDocNone);
        w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < hi) {
            d = __ldg(reinterpret_cast<const int4*>(ids + idx));
            w = __ldg(reinterpret_cast<const float4*>(wts + idx));
        }
    }
    // Adds the postings whose doc id lies in [base, base + S) into the tile at shared address
    // `tile`.  One term has at most one posting per document, so the four slots are distinct and
    // the four read-modify-writes are independent: predicated (never branching), loads first.
    // n0..n3 = the new scores (0 for postings outside the tile).
    __device__ __forceinline__ void add_into(unsigned tile, int base, unsigned S, float& n0, float& n1, float& n2,
                                             float& n3) const {
        asm volatile(
            "{\n\t"
            ".reg .pred p0, p1, p2, p3;\n\t"
            ".reg .u32 s0, s1, s2, s3;\n\t"
            ".reg .f32 v0, v1, v2, v3;\n\t"
            "sub.u32 s0, %4, %12;\n\t"
            "sub.u32 s1, %5, %12;\n\t"
            "sub.u32 s2, %6, %12;\n\t"
            "sub.u32 s3, %7, %12;\n\t"
            "setp.lt.u32 p0, s0, %13;\n\t"
            "setp.lt.u32 p1, s1, %13;\n\t"
            "setp.lt.u32 p2, s2, %13;\n\t"
            "setp.lt.u32 p3, s3, %13;\n\t"
            "mad.lo.u32 s0, s0, 4, %14;\n\t"
            "mad.lo.u32 s1, s1, 4, %14;\n\t"
            "mad.lo.u32 s2, s2, 4, %14;\n\t"
            "mad.lo.u32 s3, s3, 4, %14;\n\t"
            "mov.f32 %0, 0f00000000;\n\t"
            "mov.f32 %1, 0f00000000;\n\t"
            "mov.f32 %2, 0f00000000;\n\t"
            "mov.f32 %3, 0f00000000;\n\t"
            "@p0 ld.shared.f32 v0, [s0];\n\t"
            "@p1 ld.shared.f32 v1, [s1];\n\t"
            "@p2 ld.shared.f32 v2, [s2];\n\t"
            "@p3 ld.shared.f32 v3, [s3];\n\t"
            "@p0 add.f32 %0, v0, %8;\n\t"
            "@p1 add.f32 %1, v1, %9;\n\t"
            "@p2 add.f32 %2, v2, %10;\n\t"
            "@p3 add.f32 %3, v3, %11;\n\t"
            "@p0 st.shared.f32 [s0], %0;\n\t"
            "@p1 st.shared.f32 [s1], %1;\n\t"
            "@p2 st.shared.f32 [s2], %2;\n\t"
            "@p3 st.shared.f32 [s3], %3;\n\t"
            "}"
            : "=&f"(n0), "=&f"(n1), "=&f"(n2), "=&f"(n3)
            : "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w), "r"(base), "r"(S),
              "r"(tile)
            : "memory");
    }
};

// Vectorised scan of one warp's tile (S documents, the first nd_w of them real): pushes every
// document that beats the threshold and marks its slot as done; a slot whose push did not fit into
// the candidate buffer keeps its score for the rescan of the next overflow round (returns true).
// Positive path: done = 0.0 (never competitive, theta_f > 0), so a clean scan leaves the tile zeroed.
// General path (zero and negative scores compete): done = -inf, which no rescan pushes again; the
// caller zeroes the tile once the scan is clean (tile_clear).
__device__ __forceinline__ bool tile_scan(float* scw, int S, int nd_w, uint32_t doc0, int lane, TopkState& tk) {
    bool left = false;
    const float done = tk.general ? -INFINITY : 0.f;
    const float th = fmaxf(tk.theta_f, tk.floor) * tk.scale;  // threshold in the tile's scale
    const float inv = 1.0f / tk.scale;                        // exact: scale is a power of two
#pragma unroll 4
    for (int idx = lane * 4; idx < S; idx += 128) {
        const float4 v = *reinterpret_cast<const float4*>(scw + idx);
        float4 z = make_float4(done, done, done, done);
        if (fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) >= th) {
            const float vv[4] = {v.x, v.y, v.z, v.w};
            float zz[4] = {done, done, done, done};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (vv[e] >= th && vv[e] != -INFINITY && idx + e < nd_w) {
                    if (!tk.push(vv[e] * inv, doc0 + (uint32_t)(idx + e))) { zz[e] = vv[e]; left = true; }
                }
            }
            z = make_float4(zz[0], zz[1], zz[2], zz[3]);
        }
        *reinterpret_cast<float4*>(scw + idx) = z;
    }
    return left;
}
__device__ __forceinline__ void tile_clear(float* scw, int S, int lane) {
    for (int idx = lane * 4; idx < S; idx += 128) *reinterpret_cast<float4*>(scw + idx) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------
// Rare paths of k_score_topk, kept out of line (__noinline__) so that their state -- candidate
// buffer, thresholds, output pointers -- does not occupy registers in the hot tile loop.
// ---------------------------------------------------------------------------------------------
struct CtaShared {
    int ncand, overflow;
    u64 theta;
    // exponent epochs of this CTA's query (k_score_topk_s): scale of the first epoch, factor between
    // epochs, scale of the last epoch, smallest possible non-zero score (all powers of two)
    float scale0, step, top, floor;
    float wscale[16];  // per warp: scale of the tile handed to tile_finish
    int hist[264];  // select_candidates scratch
};

struct ColdCtx {
    u64* cand;       // candidate buffer of this CTA (shared or global memory)
    u64* out;        // this CTA's k output keys (also select scratch)
    u64* theta_q;    // per-query shared threshold slot, or NULL
    u64 theta0;
    CtaShared* sh;
    float* scw;      // this warp's score tile
    unsigned short* hot;
    int cap, k, S, general, nthreads, tid;
};

// candidate-buffer overflow round: every warp of the CTA takes part.  `leftover`: this warp's tile
// (documents left_doc0 .. left_doc0 + left_nd) still holds scores that did not fit in the buffer.
__device__ __forceinline__ void overflow_round(const ColdCtx& c, TopkState& tk, bool leftover, uint32_t left_doc0,
                                               int left_nd) {
    const WarpsGroup grp{c.nthreads, c.tid};
    const int lane = c.tid & 31;
    for (;;) {
        if (c.cap > kSelectMin)  // a round is only entered with a full buffer (n = cap > k)
            select_candidates(c.cand, c.cap, c.k, c.out, c.sh->hist, &c.sh->ncand, &c.sh->theta, true, grp);
        else
            compact_candidates(c.cand, c.cap, c.k, c.theta0, &c.sh->ncand, &c.sh->theta, grp);
        if (c.tid == 0) {
            c.sh->overflow = 0;
            if (c.theta_q) {  // share the threshold with the other CTAs of this query
                const u64 mine = c.sh->theta;
                const u64 old = atomicMax(c.theta_q, mine);
                if (old > mine) c.sh->theta = old;
            }
        }
        grp.sync();
        tk.set_theta(c.sh->theta);
        if (leftover) {
            leftover = __any_sync(kFull, tile_scan(c.scw, c.S, left_nd, left_doc0, lane, tk));
            if (!leftover && c.general) tile_clear(c.scw, c.S, lane);  // the scan left -inf marks
        }
        grp.sync();
        const int again = ld_volatile(&c.sh->overflow);
        grp.sync();
        if (!again) break;
    }
}

// End of a tile that needs more than the write-only clear: pushes the documents that beat the k-th
// best so far into the candidate buffer -- from the hot list (hl_n <= kHotCap entries) or by a
// full scan -- clears the tile, and serves a candidate-buffer overflow round if one is pending.
// Returns the (possibly raised) raw-score pre-filter threshold.
__device__ __noinline__ float tile_finish(const ColdCtx c, int base, int nd_w, int hl_n, int use_list) {
    const int lane = c.tid & 31;
    TopkState tk{c.cand, &c.sh->ncand, &c.sh->overflow, c.cap, c.general, 0ull, 0.f, c.sh->wscale[c.tid >> 5], c.sh->floor};
    tk.set_theta(c.sh->theta);  // thresholds only change inside rounds, which every warp attends
    bool left = false;
    if (!use_list) {
        left = __any_sync(kFull, tile_scan(c.scw, c.S, nd_w, (uint32_t)base, lane, tk));  // reads and tests all S slots
        if (!left && c.general) tile_clear(c.scw, c.S, lane);
    } else {
        __syncwarp();
        for (int i0 = 0; i0 < hl_n; i0 += 32) {
            const int i = i0 + lane;
            const int x = i < hl_n ? (int)c.hot[i] : -1 - lane;
            const unsigned same = __match_any_sync(kFull, x);  // a slot may be listed twice
            if (x >= 0 && (__ffs(same) - 1) == lane) {
                const float v = c.scw[x];
                if (v > 0.f) {  // not yet taken by an earlier step of this loop
                    if (tk.push(v * (1.0f / tk.scale), (uint32_t)(base + x))) c.scw[x] = 0.f;
                    else left = true;
                }
            }
            __syncwarp();
        }
        left = __any_sync(kFull, left);
        if (!left) smem_bulk_zero(smem_u32(c.scw), (unsigned)c.S * 4u);  // one st.bulk instead of 16 vector stores per lane
    }
    left = __any_sync(kFull, left);
    if (left || ld_volatile(&c.sh->overflow)) {
        const WarpsGroup grp{c.nthreads, c.tid};
        grp.sync();
        overflow_round(c, tk, left, (uint32_t)base, nd_w);
    }
    return tk.theta_f;
}

// After a warp's last tile: keep serving overflow rounds until every warp of the CTA is here, then
// write the CTA's k best keys.
__device__ __noinline__ void cta_finish(const ColdCtx c, int q_has_theta) {
    const WarpsGroup grp{c.nthreads, c.tid};
    TopkState tk{c.cand, &c.sh->ncand, &c.sh->overflow, c.cap, c.general, 0ull, 0.f, 1.f, 0.f};
    tk.set_theta(c.sh->theta);
    for (;;) {
        grp.sync();
        if (!ld_volatile(&c.sh->overflow)) break;
        overflow_round(c, tk, false, 0u, 0);
    }
    const int n_end = min(ld_volatile(&c.sh->ncand), c.cap);
    if (n_end > kSelectMin && n_end > c.k) {
        // the k keepers go straight to the output slot, unsorted (k_merge sorts what it loads)
        select_candidates(c.cand, n_end, c.k, c.out, c.sh->hist, &c.sh->ncand, &c.sh->theta, false, grp);
        if (c.tid == 0 && c.theta_q) atomicMax(c.theta_q, c.sh->theta);
        return;
    }
    if (q_has_theta /* candidate buffer in global memory */) {
        // here n_end <= k (a larger set went through the select above): everything is a keeper, unsorted
        for (int i = c.tid; i < c.k; i += c.nthreads) c.out[i] = (i < n_end) ? c.cand[i] : 0ull;
        return;
    }
    compact_candidates(c.cand, c.cap, c.k, c.theta0, &c.sh->ncand, &c.sh->theta, grp);
    if (c.tid == 0 && c.theta_q && c.sh->ncand >= c.k) atomicMax(c.theta_q, c.sh->theta);
    const int n = c.sh->ncand;
    for (int i = c.tid; i < c.k; i += c.nthreads) c.out[i] = (i < n) ? c.cand[i] : 0ull;
}

// MAXT = 256: up to 8 warps, three CTAs per SM (<= 80 registers); MAXT = 512: up to 16 warps, two CTAs per SM
#ifndef BM25_LB_T
#define BM25_LB_T 256
#endif
#ifndef BM25_LB_B
#define BM25_LB_B 3
#endif
template <int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == BM25_LB_T ? BM25_LB_B : 2) k_score_topk(const SearchArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NCW = blockDim.x >> 5;
    const int T = a.T, S = a.tile_docs;
    float* sc = reinterpret_cast<float*>(smem_raw);
    // candidate buffer: shared memory, or (large k) this CTA's slice of a global, L2-resident array --
    // pushes are rare once the threshold has settled, and the freed shared memory buys a third CTA per SM
    u64* cand_smem = reinterpret_cast<u64*>(sc + (size_t)NCW * S);
    int* st_all = reinterpret_cast<int*>(cand_smem + (a.cand_global ? 0 : a.cap));
    unsigned short* st_hot = reinterpret_cast<unsigned short*>(st_all + (size_t)NCW * kStateInts * T);
    __shared__ CtaShared sh;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    // split-major CTA order: the CTAs of the first document range of EVERY query come first, so the
    // later ranges of a query start from the threshold its earlier ranges published (theta_q), and
    // neighbouring CTAs walk the same document range of neighbouring (same heaviest term) queries
    const int sp = a.sp_major ? blockIdx.x / a.Q : blockIdx.x % a.splits;
    const int qslot = a.sp_major ? blockIdx.x - sp * a.Q : blockIdx.x / a.splits;
    const int q = a.qperm ? __ldg(a.qperm + qslot) : qslot;
    const int chunk = sp * NCW + warp;

    if (a.poison) {  // debug: no read of uninitialised shared memory may go unnoticed
        unsigned* all = reinterpret_cast<unsigned*>(smem_raw);
        const size_t words = ((size_t)NCW * S * 4 + (a.cand_global ? 0 : (size_t)a.cap * 8) +
                              (size_t)NCW * kStateInts * T * 4 + (size_t)NCW * kHotCap * 2) / 4;
        for (size_t i = tid; i < words; i += blockDim.x) all[i] = 0xffffffffu;
        __syncthreads();
    }
    float* scw = sc + (size_t)warp * S;
    for (int i = lane * 4; i < S; i += 128) *reinterpret_cast<float4*>(scw + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        sh.ncand = 0;
        sh.overflow = 0;
        const u64 shared_theta = a.theta_q ? *reinterpret_cast<volatile u64*>(a.theta_q + q) : 0ull;
        sh.theta = shared_theta > a.theta0 ? shared_theta : a.theta0;
        sh.floor = a.general ? -INFINITY : 1.401298464e-45f;  // this kernel keeps unscaled scores in its tiles
    }
    if (tid < 16) sh.wscale[tid] = 1.f;
    __syncthreads();

    // gathered only where a rare path is entered
    auto cold_ctx = [&]() {
        ColdCtx c;
        c.cand = a.cand_global ? a.cand_global + ((size_t)q * a.splits + sp) * a.cap : cand_smem;
        c.out = a.partial + ((int64_t)q * a.splits + sp) * a.k;
        c.theta_q = a.theta_q ? a.theta_q + q : nullptr;
        c.theta0 = a.theta0;
        c.sh = &sh;
        c.scw = scw;
        c.hot = st_hot + warp * kHotCap;
        c.cap = a.cap;
        c.k = a.k;
        c.S = S;
        c.general = a.general;
        c.nthreads = (int)blockDim.x;
        c.tid = tid;
        return c;
    };

    if (chunk < a.n_chunks) {
        const unsigned T4 = (unsigned)T * 4u;                                 // bytes per state row
        const unsigned st = smem_u32(st_all + (size_t)warp * kStateInts * T);  // this warp's state
        const unsigned tile = smem_u32(scw);                                  // this warp's score tile
        const unsigned hot_s = smem_u32(st_hot + warp * kHotCap);             // this warp's hot list
        const unsigned uS = (unsigned)S;
        const int NB = a.n_tiles;
        const int j0 = chunk * a.tiles_per_chunk;
        const int j1 = min(NB, j0 + a.tiles_per_chunk);
        const bool hot_enabled = !a.general && !a.no_hot;
        float theta_f;
        {
            const u64 t0 = sh.theta;
            theta_f = (t0 == 0ull) ? -INFINITY : (a.general ? key_score(t0) : fmaxf(key_score(t0), 1.401298464e-45f));
        }
        {
            const int32_t* seg0 = a.seg + ((int64_t)q * (a.n_chunks + 1) + chunk) * T;
            for (int t = lane; t < T; t += 32) {
                const int term = __ldg(a.queries + (int64_t)q * T + t);
                const bool valid = term >= 0 && term < a.n_terms;
                const int row = valid ? __ldg(a.term_row + term) : -1;
                const unsigned s = st + 4u * t;
                if (row >= 0) {
                    const int32_t* tr = a.tab + (int64_t)row * (NB + 1);
                    sts_i32(s + ((j0 + 0) & 3) * T4, __ldg(tr + j0));
                    sts_i32(s + ((j0 + 1) & 3) * T4, __ldg(tr + min(j0 + 1, NB)));
                    sts_i32(s + ((j0 + 2) & 3) * T4, __ldg(tr + min(j0 + 2, NB)));
                    sts_i32(s + ((j0 + 3) & 3) * T4, 0);
                    sts_i32(s + 4 * T4, row);
                    sts_i32(s + 7 * T4, 1);
                } else {
                    int p = 0, e = 0;
                    if (valid) {
                        p = __ldg(seg0 + t);
                        e = __ldg(seg0 + T + t);
                    }
                    sts_i32(s + 0 * T4, p);
                    sts_i32(s + 1 * T4, e);
                    sts_i32(s + 2 * T4, p < e ? __ldg(a.ids + p) : kDocNone);
                    sts_i32(s + 4 * T4, p < e ? __float_as_int(__ldg(a.w + p)) : 0);
                    sts_i32(s + 3 * T4, p + 1 < e ? __ldg(a.ids + p + 1) : kDocNone);
                    sts_i32(s + 5 * T4, p + 1 < e ? __float_as_int(__ldg(a.w + p + 1)) : 0);
                    sts_i32(s + 7 * T4, 0);
                }
            }
        }
        for (int j = j0; j < j1; ++j) {
            cp_async_wait_all();  // table entries / cursor refills requested during the previous tiles
            __syncwarp();
            const int base = j * S;
            const int tile_end = base + S;
            bool touched = false;
            bool refill_inflight = false;  // a light-term refill was requested during THIS tile
            int hl_n = hot_enabled ? 0 : kHotCap + 1;  // hot-list length; > kHotCap: overflowed / disabled
            const unsigned r0 = st + (j & 3) * T4, r1 = st + ((j + 1) & 3) * T4, r3 = st + ((j + 3) & 3) * T4;
            const int jn = min(j + 3, NB);

            // warp-collective: append the tile slots of the lanes with `hot` to the hot list
            auto hot_add = [&](bool hot, int slot) {
                const unsigned m = __ballot_sync(kFull, hot);
                if (m == 0u) return;
                const int c = __popc(m);
                if (hl_n + c <= kHotCap) {
                    unsigned lt;
                    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
                    if (hot) asm volatile("st.shared.u16 [%0], %1;" ::"r"(hot_s + 2u * (hl_n + __popc(m & lt))), "h"((unsigned short)slot) : "memory");
                    hl_n += c;
                } else {
                    hl_n = kHotCap + 1;
                }
            };

            // light terms of `mask` (bit b = term g0 + b), in query order; lane 0 adds, from the
            // cursor's resident postings
            auto light_visits = [&](unsigned mask, int g0) {
                while (mask) {
                    const unsigned s = st + 4u * (g0 + __ffs(mask) - 1);
                    mask &= mask - 1;
                    for (;;) {
                        const int hd = lds_i32(s + 2 * T4);
                        if (hd >= tile_end) break;
                        bool hot = false;
                        if (lane == 0) {
                            const unsigned slot = tile + 4u * (unsigned)(hd - base);
                            const float nw = lds_f32(slot) + __int_as_float(lds_i32(s + 4 * T4));
                            sts_f32(slot, nw);
                            hot = nw >= theta_f;
                        }
                        if (hl_n <= kHotCap) hot_add(hot, hd - base);
                        if (refill_inflight) {  // the second resident posting may still be on its way
                            cp_async_wait_all();
                            refill_inflight = false;
                        }
                        __syncwarp();
                        if (lane == 0) {
                            const int pos = lds_i32(s) + 1, e = lds_i32(s + T4);
                            sts_i32(s, pos);
                            sts_i32(s + 2 * T4, lds_i32(s + 3 * T4));
                            sts_i32(s + 4 * T4, lds_i32(s + 5 * T4));
                            if (pos + 1 < e) {
                                cp_async4_s(s + 3 * T4, a.ids + pos + 1);
                                cp_async4_s(s + 5 * T4, a.w + pos + 1);
                            } else {
                                sts_i32(s + 3 * T4, kDocNone);
                            }
                        }
                        refill_inflight = true;
                        __syncwarp();
                    }
                }
            };

            // ---- accumulate: terms strictly in query order ------------------------------------
            for (int g0 = 0; g0 < T; g0 += 32) {
                const int t = g0 + lane;
                int lo = 0, hi = 0, hd = kDocNone;
                if (t < T) {
                    const unsigned s = 4u * t;
                    if (lds_i32(st + 7 * T4 + s)) {
                        lo = lds_i32(r0 + s);
                        hi = lds_i32(r1 + s);
                        cp_async4_s(r3 + s, a.tab + (int64_t)lds_i32(st + 4 * T4 + s) * (NB + 1) + jn);
                    } else {
                        hd = lds_i32(st + 2 * T4 + s);
                    }
                }
                unsigned hm = __ballot_sync(kFull, hi > lo);        // heavy terms with postings in this tile
                unsigned lm = __ballot_sync(kFull, hd < tile_end);  // light terms with postings in this tile
                if ((hm | lm) == 0u) continue;
                touched = true;
                // piece generator over the heavy terms of the group (warp-uniform state)
                int gp = 0, ghi = 0, gt = 0;
                auto next_piece = [&](bool& first) -> bool {
                    if (gp + 128 < ghi) {
                        gp += 128;
                        first = false;
                        return true;
                    }
                    if (hm == 0u) return false;
                    gt = __ffs(hm) - 1;
                    hm &= hm - 1;
                    gp = __shfl_sync(kFull, lo, gt) & ~3;
                    ghi = __shfl_sync(kFull, hi, gt);
                    first = true;
                    return true;
                };
                auto consume = [&](const PostingPiece& P, bool first, int tt) {
                    if (first) {  // a new term: first the light terms that precede it in the query
                        const unsigned pl = lm & ((1u << tt) - 1u);
                        if (pl) {
                            __syncwarp();
                            lm &= ~pl;
                            light_visits(pl, g0);
                        }
                        __syncwarp();
                    }
                    float n0, n1, n2, n3;
                    P.add_into(tile, base, uS, n0, n1, n2, n3);
                    if (hl_n <= kHotCap) {  // postings outside the tile give n = 0 < theta_f
                        if (__any_sync(kFull, fmaxf(fmaxf(n0, n1), fmaxf(n2, n3)) >= theta_f)) {
                            hot_add(n0 >= theta_f, P.d.x - base);
                            hot_add(n1 >= theta_f, P.d.y - base);
                            hot_add(n2 >= theta_f, P.d.z - base);
                            hot_add(n3 >= theta_f, P.d.w - base);
                        }
                    }
                };
                // ping-pong: the piece after the current one is requested before the current one is added
                PostingPiece A, B;
                bool fA = false, fB = false;
                bool more = next_piece(fA);
                int tA = gt, tB = 0;
                if (more) A.load(a.ids, a.w, gp, ghi, lane);
                while (more) {
                    more = next_piece(fB);
                    tB = gt;
                    if (more) B.load(a.ids, a.w, gp, ghi, lane);
                    consume(A, fA, tA);
                    if (!more) break;
                    more = next_piece(fA);
                    tA = gt;
                    if (more) A.load(a.ids, a.w, gp, ghi, lane);
                    consume(B, fB, tB);
                }
                __syncwarp();
                if (lm) light_visits(lm, g0);
            }
            // ---- end of tile: the common case is a write-only clear ---------------------------
            const bool cold = a.general || (touched && hl_n != 0) || ld_volatile(&sh.overflow);
            if (!cold) {
                if (touched) {
                    if (a.bulk_clear) {
                        smem_bulk_zero(tile, uS * 4u);
                    } else {
                        for (unsigned off = lane * 16u; off < uS * 4u; off += 512u) sts_zero16(tile + off);
                    }
                }
            } else {
                theta_f = tile_finish(cold_ctx(), base, min(S, a.n_docs - base), hl_n,
                                      (hl_n <= kHotCap && !a.general) ? 1 : 0);
            }
        }
        cp_async_wait_all();  // nothing of this warp may still be in flight towards shared memory
    }
    cta_finish(cold_ctx(), a.cand_global != nullptr);
}

// ---------------------------------------------------------------------------------------------
// k_score_topk_s: the same kernel specialised for queries of at most 32 term slots (every
// BASELINE config except the 64-term stress).  Lane t owns query term t for the whole chunk:
//   * its class (heavy / light / none), its tile-table row pointer or its cursor live in registers;
//   * its 16 bytes of shared state are {ring of 4 table entries} (heavy) or {doc0, w0, doc1, w1},
//     the two resident postings of the cursor (light) -- cp.async targets, one LDS per tile;
//   * heavy terms with at most 32 postings in the tile are fetched as NARROW pieces (one posting
//     per lane, 4-byte loads, one read-modify-write) instead of 128-posting pieces;
//   * a light term's postings are added by its own lane; posting lists are sentinel-terminated
//     (>= 1 kDocNone after the last posting), so a cursor needs no end pointer.
// shared memory: same carve-up as k_score_topk (state area: 16 bytes per term are used).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lds_v2(unsigned a, int& x, int& y) {
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts_v2(unsigned a, int x, int y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}

struct PieceRegs {
    int4 d;
    float4 w;
    // 128 postings from p0 (multiple of 4), four per lane
    __device__ __forceinline__ void load_wide(const int32_t* __restrict__ ids, const float* __restrict__ wts, int p0, int hi,
                                              int lane) {
        const int idx = p0 + 4 * lane;
        d = make_int4(kDocNone, kDocNone, kDocNone, kDocNone);
        w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < hi) {
            d = __ldg(reinterpret_cast<const int4*>(ids + idx));
            w = __ldg(reinterpret_cast<const float4*>(wts + idx));
        }
    }
    // up to 32 postings [lo, hi), one per lane (component x)
    __device__ __forceinline__ void load_narrow(const int32_t* __restrict__ ids, const float* __restrict__ wts, int lo, int hi,
                                                int lane) {
        const int idx = lo + lane;
        d.x = kDocNone;
        w.x = 0.f;
        if (idx < hi) {
            d.x = __ldg(ids + idx);
            w.x = __ldg(wts + idx);
        }
    }
    __device__ __forceinline__ void add_wide(unsigned tile, int base, unsigned S, float scale, float& n0, float& n1,
                                             float& n2, float& n3) const {
        asm volatile(
            "{\n\t"
            ".reg .pred p0, p1, p2, p3;\n\t"
            ".reg .u32 s0, s1, s2, s3;\n\t"
            "sub.u32 s0, %4, %12;\n\t"
            "sub.u32 s1, %5, %12;\n\t"
            "sub.u32 s2, %6, %12;\n\t"
            "sub.u32 s3, %7, %12;\n\t"
            "setp.lt.u32 p0, s0, %13;\n\t"
            "setp.lt.u32 p1, s1, %13;\n\t"
            "setp.lt.u32 p2, s2, %13;\n\t"
            "setp.lt.u32 p3, s3, %13;\n\t"
            "mad.lo.u32 s0, s0, 4, %14;\n\t"
            "mad.lo.u32 s1, s1, 4, %14;\n\t"
            "mad.lo.u32 s2, s2, 4, %14;\n\t"
            "mad.lo.u32 s3, s3, 4, %14;\n\t"
            "mov.f32 %0, 0f00000000;\n\t"
            "mov.f32 %1, 0f00000000;\n\t"
            "mov.f32 %2, 0f00000000;\n\t"
            "mov.f32 %3, 0f00000000;\n\t"
            "@p0 ld.shared.f32 %0, [s0];\n\t"
            "@p1 ld.shared.f32 %1, [s1];\n\t"
            "@p2 ld.shared.f32 %2, [s2];\n\t"
            "@p3 ld.shared.f32 %3, [s3];\n\t"
            "@p0 fma.rn.f32 %0, %8, %15, %0;\n\t"
            "@p1 fma.rn.f32 %1, %9, %15, %1;\n\t"
            "@p2 fma.rn.f32 %2, %10, %15, %2;\n\t"
            "@p3 fma.rn.f32 %3, %11, %15, %3;\n\t"
            "@p0 st.shared.f32 [s0], %0;\n\t"
            "@p1 st.shared.f32 [s1], %1;\n\t"
            "@p2 st.shared.f32 [s2], %2;\n\t"
            "@p3 st.shared.f32 [s3], %3;\n\t"
            "}"
            : "=&f"(n0), "=&f"(n1), "=&f"(n2), "=&f"(n3)
            : "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w), "r"(base), "r"(S),
              "r"(tile), "f"(scale)
            : "memory");
    }
    // ---- compressed index (k_pack): one 32-bit word per posting, high half = byte offset of the
    // ---- document's slot in its tile (4 * (doc mod S)), low half = bf16 weight ------------------
    // 128 packed postings from p0 (multiple of 4), four per lane, ONE 16-byte load (component d).
    // The 16-byte granularity drags in up to 3 postings of other tiles on either side of the
    // tile's range [lo, hi): validity is decided by position (w.x carries idx - lo as an int).
    __device__ __forceinline__ void load_wide_pk(const uint32_t* __restrict__ pk, int p0, int lo, int hi, int lane) {
        const int idx = p0 + 4 * lane;
        w.x = __int_as_float(idx - lo);
        if (idx < hi) d = __ldg(reinterpret_cast<const int4*>(pk + idx));
    }
    // up to 32 postings [lo, hi), one per lane; lanes beyond hi hold the invalid word (slot offset 0xffff)
    __device__ __forceinline__ void load_narrow_pk(const uint32_t* __restrict__ pk, int lo, int hi, int lane) {
        const int idx = lo + lane;
        d.x = (int)0xffff0000u;
        if (idx < hi) d.x = (int)__ldg(pk + idx);
    }
    static __device__ __forceinline__ int pk_doc(int word) { return (int)((unsigned)word >> 18); }  // tile-local doc id
    // n = hi - lo (postings of the term in this tile)
    __device__ __forceinline__ void add_wide_pk(unsigned tile, unsigned n, float scale, float& n0, float& n1,
                                                float& n2, float& n3) const {
        asm volatile(
            "{\n\t"
            ".reg .pred p0, p1, p2, p3;\n\t"
            ".reg .u32 s0, s1, s2, s3, v1, v2, v3;\n\t"
            ".reg .b32 w0, w1, w2, w3;\n\t"
            "add.u32 v1, %8, 1;\n\t"
            "add.u32 v2, %8, 2;\n\t"
            "add.u32 v3, %8, 3;\n\t"
            "setp.lt.u32 p0, %8, %9;\n\t"
            "setp.lt.u32 p1, v1, %9;\n\t"
            "setp.lt.u32 p2, v2, %9;\n\t"
            "setp.lt.u32 p3, v3, %9;\n\t"
            "shr.u32 s0, %4, 16;\n\t"
            "shr.u32 s1, %5, 16;\n\t"
            "shr.u32 s2, %6, 16;\n\t"
            "shr.u32 s3, %7, 16;\n\t"
            "add.u32 s0, s0, %10;\n\t"
            "add.u32 s1, s1, %10;\n\t"
            "add.u32 s2, s2, %10;\n\t"
            "add.u32 s3, s3, %10;\n\t"
            "shl.b32 w0, %4, 16;\n\t"
            "shl.b32 w1, %5, 16;\n\t"
            "shl.b32 w2, %6, 16;\n\t"
            "shl.b32 w3, %7, 16;\n\t"
            "mov.f32 %0, 0f00000000;\n\t"
            "mov.f32 %1, 0f00000000;\n\t"
            "mov.f32 %2, 0f00000000;\n\t"
            "mov.f32 %3, 0f00000000;\n\t"
            "@p0 ld.shared.f32 %0, [s0];\n\t"
            "@p1 ld.shared.f32 %1, [s1];\n\t"
            "@p2 ld.shared.f32 %2, [s2];\n\t"
            "@p3 ld.shared.f32 %3, [s3];\n\t"
            "@p0 fma.rn.f32 %0, w0, %11, %0;\n\t"
            "@p1 fma.rn.f32 %1, w1, %11, %1;\n\t"
            "@p2 fma.rn.f32 %2, w2, %11, %2;\n\t"
            "@p3 fma.rn.f32 %3, w3, %11, %3;\n\t"
            "@p0 st.shared.f32 [s0], %0;\n\t"
            "@p1 st.shared.f32 [s1], %1;\n\t"
            "@p2 st.shared.f32 [s2], %2;\n\t"
            "@p3 st.shared.f32 [s3], %3;\n\t"
            "}"
            : "=&f"(n0), "=&f"(n1), "=&f"(n2), "=&f"(n3)
            : "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w), "r"(__float_as_int(w.x)), "r"(n), "r"(tile), "f"(scale)
            : "memory");
    }
    __device__ __forceinline__ float add_narrow_pk(unsigned tile, unsigned S4, float scale) const {
        float n0;
        asm volatile(
            "{\n\t"
            ".reg .pred p0;\n\t"
            ".reg .u32 s0;\n\t"
            ".reg .b32 w0;\n\t"
            "shr.u32 s0, %1, 16;\n\t"
            "setp.lt.u32 p0, s0, %2;\n\t"
            "add.u32 s0, s0, %3;\n\t"
            "shl.b32 w0, %1, 16;\n\t"
            "mov.f32 %0, 0f00000000;\n\t"
            "@p0 ld.shared.f32 %0, [s0];\n\t"
            "@p0 fma.rn.f32 %0, w0, %4, %0;\n\t"
            "@p0 st.shared.f32 [s0], %0;\n\t"
            "}"
            : "=&f"(n0)
            : "r"(d.x), "r"(S4), "r"(tile), "f"(scale)
            : "memory");
        return n0;
    }
    __device__ __forceinline__ float add_narrow(unsigned tile, int base, unsigned S, float scale) const {
        float n0;
        asm volatile(
            "{\n\t"
            ".reg .pred p0;\n\t"
            ".reg .u32 s0;\n\t"
            "sub.u32 s0, %1, %3;\n\t"
            "setp.lt.u32 p0, s0, %4;\n\t"
            "mad.lo.u32 s0, s0, 4, %5;\n\t"
            "mov.f32 %0, 0f00000000;\n\t"
            "@p0 ld.shared.f32 %0, [s0];\n\t"
            "@p0 fma.rn.f32 %0, %2, %6, %0;\n\t"
            "@p0 st.shared.f32 [s0], %0;\n\t"
            "}"
            : "=&f"(n0)
            : "r"(d.x), "f"(w.x), "r"(base), "r"(S), "r"(tile), "f"(scale)
            : "memory");
        return n0;
    }
};

template <int MAXT, bool PK>
__global__ void __launch_bounds__(MAXT, MAXT == BM25_LB_T ? BM25_LB_B : 2) k_score_topk_s(const SearchArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NCW = blockDim.x >> 5;
    const int T = a.T, S = a.tile_docs;
    float* sc = reinterpret_cast<float*>(smem_raw);
    u64* cand_smem = reinterpret_cast<u64*>(sc + (size_t)NCW * S);
    int* st_all = reinterpret_cast<int*>(cand_smem + (a.cand_global ? 0 : a.cap));
    unsigned short* st_hot = reinterpret_cast<unsigned short*>(st_all + (size_t)NCW * kStateInts * T);
    __shared__ CtaShared sh;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    // split-major CTA order: the CTAs of the first document range of EVERY query come first, so the
    // later ranges of a query start from the threshold its earlier ranges published (theta_q), and
    // neighbouring CTAs walk the same document range of neighbouring (same heaviest term) queries
    const int sp = a.sp_major ? blockIdx.x / a.Q : blockIdx.x % a.splits;
    const int qslot = a.sp_major ? blockIdx.x - sp * a.Q : blockIdx.x / a.splits;
    const int q = a.qperm ? __ldg(a.qperm + qslot) : qslot;
    const int chunk = sp * NCW + warp;

    if (a.poison) {  // debug: no read of uninitialised shared memory may go unnoticed
        unsigned* all = reinterpret_cast<unsigned*>(smem_raw);
        const size_t words = ((size_t)NCW * S * 4 + (a.cand_global ? 0 : (size_t)a.cap * 8) +
                              (size_t)NCW * kStateInts * T * 4 + (size_t)NCW * kHotCap * 2) / 4;
        for (size_t i = tid; i < words; i += blockDim.x) all[i] = 0xffffffffu;
        __syncthreads();
    }
    float* scw = sc + (size_t)warp * S;
    for (int i = lane * 4; i < S; i += 128) *reinterpret_cast<float4*>(scw + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 0) {
        // ---- exponent epochs ------------------------------------------------------------------
        // With positive weights the tile need not be zeroed after every tile: tile e of a run of E
        // tiles accumulates weight * 2^(s0 + c*e) (fma, exact: a power of two), with c so large
        // that whatever an earlier tile of the run left in a slot is less than half an ulp of the
        // first weight added on top of it -- the addition returns that weight exactly, as if the
        // slot had been zero.  The bound: every score of this query is < 2^ls (sum of the terms'
        // largest weights) and every weight is >= 2^lw, so c = ls - lw + 26 suffices; the run
        // length E is what the fp32 exponent range allows.  Scores are unscaled (exactly) when
        // they leave the tile; values below floor * scale are leftovers, i.e. zero.
        float mn = INFINITY, sum = 0.f;
        if (lane < T) {
            const int term = __ldg(a.queries + (int64_t)q * T + lane);
            if (term >= 0 && term < a.n_terms) {
                const float2 r = __ldg(a.wrange + term);
                if (r.y >= r.x) {  // not an empty list
                    mn = r.x;
                    sum = r.y;
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(kFull, mn, o));
            sum += __shfl_xor_sync(kFull, sum, o);
        }
        if (lane == 0) {
            float scale0 = 1.f, step = 1.f, top = 1.f, floor = a.general ? -INFINITY : 1.401298464e-45f;
            if (!a.general && !a.no_epoch && mn >= 7.8886090522e-31f /* 2^-100 */ && sum >= mn &&
                sum < 1.2676506002e30f /* 2^100 */) {
                const int lw = (__float_as_int(mn) >> 23) - 127;         // 2^lw <= every weight
                const int ls = (__float_as_int(sum) >> 23) - 127 + 2;    // every score < 2^ls
                const int c = ls - lw + 26;
                const int s0 = max(-125 - lw, -120);
                const int E = 1 + (125 - ls - s0) / c;
                if (E >= 2) {
                    scale0 = __int_as_float((s0 + 127) << 23);
                    step = __int_as_float((c + 127) << 23);
                    top = __int_as_float((s0 + c * (E - 1) + 127) << 23);
                    floor = __int_as_float((lw + 127) << 23);
                }
            }
            sh.scale0 = scale0;
            sh.step = step;
            sh.top = top;
            sh.floor = floor;
        }
    }
    if (tid == 0) {
        sh.ncand = 0;
        sh.overflow = 0;
        const u64 shared_theta = a.theta_q ? *reinterpret_cast<volatile u64*>(a.theta_q + q) : 0ull;
        sh.theta = shared_theta > a.theta0 ? shared_theta : a.theta0;
    }
    __syncthreads();

    float scale = sh.scale0;  // scale of the current tile
    auto cold_ctx = [&]() {
        ColdCtx c;
        c.cand = a.cand_global ? a.cand_global + ((size_t)q * a.splits + sp) * a.cap : cand_smem;
        c.out = a.partial + ((int64_t)q * a.splits + sp) * a.k;
        c.theta_q = a.theta_q ? a.theta_q + q : nullptr;
        c.theta0 = a.theta0;
        c.sh = &sh;
        c.scw = scw;
        c.hot = st_hot + warp * kHotCap;
        c.cap = a.cap;
        c.k = a.k;
        c.S = S;
        c.general = a.general;
        c.nthreads = (int)blockDim.x;
        c.tid = tid;
        return c;
    };

    if (chunk < a.n_chunks) {
        const unsigned st = smem_u32(st_all + (size_t)warp * kStateInts * T) + 16u * lane;  // this lane's 16 bytes
        const unsigned tile = smem_u32(scw);
        const unsigned hot_s = smem_u32(st_hot + warp * kHotCap);
        const unsigned uS = (unsigned)S;
        const int NB = a.n_tiles;
        const int j0 = chunk * a.tiles_per_chunk;
        const int j1 = min(NB, j0 + a.tiles_per_chunk);
        const bool hot_enabled = !a.general && !a.no_hot;
        float theta_f;  // pre-filter threshold IN THE CURRENT TILE'S SCALE
        {
            const u64 t0 = sh.theta;
            theta_f = (t0 == 0ull) ? -INFINITY : fmaxf(key_score(t0), sh.floor) * scale;
        }
        // ---- this lane's term -------------------------------------------------------------------
        const int32_t* tabp = nullptr;  // heavy: row of the tile table
        int hi_prev = 0;                // heavy: table entry of the current tile (= end of the previous one)
        int lpos = -1;                  // light: posting index of the first resident posting (-1: no term)
        if (lane < T) {
            const int term = __ldg(a.queries + (int64_t)q * T + lane);
            if (term >= 0 && term < a.n_terms) {
                const int row = __ldg(a.term_row + term);
                if (row >= 0) {
                    tabp = a.tab + (int64_t)row * (NB + 1);
                    hi_prev = __ldg(tabp + j0);
                    sts_i32(st + 4u * ((j0 + 1) & 3), __ldg(tabp + min(j0 + 1, NB)));
                    sts_i32(st + 4u * ((j0 + 2) & 3), __ldg(tabp + min(j0 + 2, NB)));
                } else {
                    lpos = __ldg(a.seg + ((int64_t)q * (a.n_chunks + 1) + chunk) * T + lane);
                    sts_i32(st + 0, __ldg(a.ids + lpos));
                    sts_i32(st + 4, __float_as_int(__ldg(a.w + lpos)));
                    sts_i32(st + 8, __ldg(a.ids + lpos + 1));
                    sts_i32(st + 12, __float_as_int(__ldg(a.w + lpos + 1)));
                }
            }
        }
        const bool heavy = tabp != nullptr;
        const bool light = lpos >= 0;
        const unsigned st0 = st - 16u * lane;  // state of term 0 (warp-uniform)

        for (int j = j0; j < j1; ++j) {
            cp_async_wait_all();  // table entries / cursor refills requested during the previous tiles
            __syncwarp();
            const int base = j * S;
            const int tile_end = base + S;
            bool refill_inflight = false;  // a light-term refill was requested during THIS tile
            int hl_n = hot_enabled ? 0 : kHotCap + 1;  // hot-list length; > kHotCap: overflowed / disabled
            int lo = 0, hi = 0, hd = kDocNone;
            if (heavy) {
                lo = hi_prev;
                hi = lds_i32(st + 4u * ((j + 1) & 3));
                cp_async4_s(st + 4u * ((j + 3) & 3), tabp + min(j + 3, NB));
                hi_prev = hi;
            } else if (light) {
                hd = lds_i32(st);
            }
            unsigned hm = __ballot_sync(kFull, hi > lo);        // heavy terms with postings in this tile
            unsigned lm = __ballot_sync(kFull, hd < tile_end);  // light terms with postings in this tile
            const bool touched = (hm | lm) != 0u;

            // warp-collective: append the tile slots of the lanes with `hot` to the hot list
            auto hot_add = [&](bool hot, int slot) {
                const unsigned m = __ballot_sync(kFull, hot);
                if (m == 0u) return;
                const int c = __popc(m);
                if (hl_n + c <= kHotCap) {
                    unsigned lt;
                    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
                    if (hot) asm volatile("st.shared.u16 [%0], %1;" ::"r"(hot_s + 2u * (hl_n + __popc(m & lt))), "h"((unsigned short)slot) : "memory");
                    hl_n += c;
                } else {
                    hl_n = kHotCap + 1;
                }
            };

            // light terms of `mask` (bit = lane = term), in query order; the owner lane adds its
            // cursor's resident postings and moves the cursor
            auto light_visits = [&](unsigned mask) {
                while (mask) {
                    const int tl = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const unsigned s = st0 + 16u * tl;
                    for (;;) {
                        int d0, w0;
                        lds_v2(s, d0, w0);
                        if (d0 >= tile_end) break;
                        bool hot = false;
                        if (lane == tl) {
                            const unsigned slot = tile + 4u * (unsigned)(d0 - base);
                            const float nw = fmaf(__int_as_float(w0), scale, lds_f32(slot));
                            sts_f32(slot, nw);
                            hot = nw >= theta_f;
                        }
                        if (hl_n <= kHotCap) hot_add(hot, d0 - base);
                        if (refill_inflight) {  // the second resident posting may still be on its way
                            cp_async_wait_all();
                            refill_inflight = false;
                        }
                        __syncwarp();
                        if (lane == tl) {
                            int d1, w1;
                            lds_v2(s + 8, d1, w1);
                            sts_v2(s, d1, w1);
                            lpos += 1;
                            cp_async4_s(s + 8, a.ids + lpos + 1);
                            cp_async4_s(s + 12, a.w + lpos + 1);
                        }
                        refill_inflight = true;
                        __syncwarp();
                    }
                }
            };

            if (touched) {
                // ---- accumulate: terms strictly in query order --------------------------------
                // piece generator over the heavy terms (warp-uniform state)
                int gp = 0, ghi = 0, gt = 0, gl = 0;
                auto next_piece = [&](bool& first, bool& narrow) -> bool {
                    if (gp + 128 < ghi) {  // only reached for wide terms
                        gp += 128;
                        first = false;
                        narrow = false;
                        return true;
                    }
                    if (hm == 0u) return false;
                    gt = __ffs(hm) - 1;
                    hm &= hm - 1;
                    const int glo = __shfl_sync(kFull, lo, gt);
                    ghi = __shfl_sync(kFull, hi, gt);
                    narrow = ghi - glo <= 32;
                    gp = narrow ? glo : (glo & ~3);
                    if (PK) gl = glo;
                    first = true;
                    return true;
                };
                auto fetch = [&](PieceRegs& P, bool narrow) {
                    if (PK) {
                        P.w.y = __int_as_float(ghi - gl);  // warp-uniform: postings of the term in this tile
                        if (narrow) P.load_narrow_pk(a.pk, gp, ghi, lane);
                        else P.load_wide_pk(a.pk, gp, gl, ghi, lane);
                    } else {
                        if (narrow) P.load_narrow(a.ids, a.w, gp, ghi, lane);
                        else P.load_wide(a.ids, a.w, gp, ghi, lane);
                    }
                };
                auto consume = [&](const PieceRegs& P, bool first, bool narrow, int tt) {
                    if (first) {  // a new term: first the light terms that precede it in the query
                        const unsigned pl = lm & ((1u << tt) - 1u);
                        if (pl) {
                            __syncwarp();
                            lm &= ~pl;
                            light_visits(pl);
                        }
                        __syncwarp();
                    }
                    if (PK) {  // compressed index: tile-local slot + bf16 weight in one word
                        if (narrow) {
                            const float n0 = P.add_narrow_pk(tile, uS * 4u, scale);
                            if (hl_n <= kHotCap) hot_add(n0 >= theta_f, PieceRegs::pk_doc(P.d.x));  // invalid lanes: n0 = 0 < theta_f
                        } else {
                            float n0, n1, n2, n3;
                            P.add_wide_pk(tile, (unsigned)__float_as_int(P.w.y), scale, n0, n1, n2, n3);
                            if (hl_n <= kHotCap) {
                                if (__any_sync(kFull, fmaxf(fmaxf(n0, n1), fmaxf(n2, n3)) >= theta_f)) {
                                    hot_add(n0 >= theta_f, PieceRegs::pk_doc(P.d.x));
                                    hot_add(n1 >= theta_f, PieceRegs::pk_doc(P.d.y));
                                    hot_add(n2 >= theta_f, PieceRegs::pk_doc(P.d.z));
                                    hot_add(n3 >= theta_f, PieceRegs::pk_doc(P.d.w));
                                }
                            }
                        }
                    } else if (narrow) {
                        const float n0 = P.add_narrow(tile, base, uS, scale);
                        if (hl_n <= kHotCap) hot_add(n0 >= theta_f, P.d.x - base);  // outside the tile: n0 = 0 < theta_f
                    } else {
                        float n0, n1, n2, n3;
                        P.add_wide(tile, base, uS, scale, n0, n1, n2, n3);
                        if (hl_n <= kHotCap) {
                            if (__any_sync(kFull, fmaxf(fmaxf(n0, n1), fmaxf(n2, n3)) >= theta_f)) {
                                hot_add(n0 >= theta_f, P.d.x - base);
                                hot_add(n1 >= theta_f, P.d.y - base);
                                hot_add(n2 >= theta_f, P.d.z - base);
                                hot_add(n3 >= theta_f, P.d.w - base);
                            }
                        }
                    }
                };
                // ping-pong: the piece after the current one is requested before the current one is added
                PieceRegs A, B;
                A.d = B.d = make_int4(kDocNone, kDocNone, kDocNone, kDocNone);
                A.w = B.w = make_float4(0.f, 0.f, 0.f, 0.f);
                bool fA = false, fB = false, nA = false, nB = false;
                bool more = next_piece(fA, nA);
                int tA = gt, tB = 0;
                if (more) fetch(A, nA);
                while (more) {
                    more = next_piece(fB, nB);
                    tB = gt;
                    if (more) fetch(B, nB);
                    consume(A, fA, nA, tA);
                    if (!more) break;
                    more = next_piece(fA, nA);
                    tA = gt;
                    if (more) fetch(A, nA);
                    consume(B, fB, nB, tB);
                }
                __syncwarp();
                if (lm) light_visits(lm);
            }
            // ---- end of tile: the common case is a write-only clear ---------------------------
            const bool cold = a.general || (touched && hl_n != 0) || ld_volatile(&sh.overflow);
            if (!cold) {
                if (touched) {
                    if (scale == sh.top) {  // last epoch of the run (always, when epochs are off): zero the tile
                        if (a.bulk_clear) {
                            smem_bulk_zero(tile, uS * 4u);
                        } else {
                            for (unsigned off = lane * 16u; off < uS * 4u; off += 512u) sts_zero16(tile + off);
                        }
                        const float s0 = sh.scale0;
                        theta_f = theta_f * (1.0f / scale) * s0;
                        scale = s0;
                    } else {  // next epoch: what this tile left behind is below half an ulp of the next tile's weights
                        const float up = sh.step;
                        scale *= up;
                        theta_f *= up;
                    }
                }
            } else {
                // returns the unscaled threshold and leaves the tile zeroed: a new run starts
                if (lane == 0) sh.wscale[warp] = scale;
                __syncwarp();
                const float th = tile_finish(cold_ctx(), base, min(S, a.n_docs - base), hl_n,
                                             (hl_n <= kHotCap && !a.general) ? 1 : 0);
                scale = sh.scale0;
                theta_f = fmaxf(th, sh.floor) * scale;
            }
        }
        cp_async_wait_all();  // nothing of this warp may still be in flight towards shared memory
    }
    cta_finish(cold_ctx(), a.cand_global != nullptr);
}

// ---------------------------------------------------------------------------------------------
// k_merge: one CTA per query.  Input either `keys` [Q, n_lists, k_in] (tile-range partials of
// this device) or (ids, scores) [n_lists, Q, k_in] (all-gathered shard results).  Selects the
// k_out best, fills with zero-score documents when fewer candidates exist (fill_base/fill_docs
// describe the local document range the fill ids are taken from), writes ids (+id_offset), scores.
// shared memory: u64 buf[P]  with P = pow2 >= min(n_lists*k_in, cap) ; uint8 present[k_out]
// ---------------------------------------------------------------------------------------------
struct MergeArgs {
    const u64* __restrict__ keys;
    const int32_t* __restrict__ in_ids;
    const float* __restrict__ in_scores;
    int32_t* __restrict__ out_ids;
    float* __restrict__ out_scores;
    int64_t Q;
    int64_t list_stride; // elements between consecutive lists of in_ids / in_scores
    int n_lists, k_in, k_out, P;
    int64_t id_offset;   // added to key doc ids on output (doc_id_base); 0 for shard merges
    int fill;            // 1: pad with zero-score docs 0,1,2.. not already present
    // k_merge_large only: per query [n_pad] candidate keys followed by [P] keepers; k_out bytes of flags
    u64* scratch;
    unsigned char* present_g;
    int64_t n_pad;
};

__global__ void __launch_bounds__(kThreads) k_merge(const MergeArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u64* buf = reinterpret_cast<u64*>(smem_raw);
    unsigned char* present = reinterpret_cast<unsigned char*>(buf + a.P);
    const int tid = threadIdx.x;
    const int nt = blockDim.x;  // 64 .. kThreads: small merges run in small CTAs (more of them per SM)
    const int64_t q = blockIdx.x;
    const int total = a.n_lists * a.k_in;
    const int keep = a.k_out;

    // rounds: buf[0..have) holds the best so far (sorted); append up to P - have new keys, sort.
    int have = 0, next = 0;
    while (next < total || have == 0) {
        const int room = a.P - have;
        const int take = min(room, total - next);
        for (int i = tid; i < room; i += nt) {
            u64 key = 0;
            if (i < take) {
                const int e = next + i;
                if (a.keys) {
                    key = a.keys[q * total + e];
                } else {
                    const int l = e / a.k_in, r = e - l * a.k_in;
                    const int64_t off = (int64_t)l * a.list_stride + q * a.k_in + r;
                    const int32_t id = a.in_ids[off];  // negative: padding of a shard with fewer than k_in documents
                    key = id < 0 ? 0ull : make_key(a.in_scores[off], (uint32_t)id);
                }
            }
            buf[have + i] = key;
        }
        __syncthreads();
        bitonic_sort_desc(buf, a.P, CtaGroup{nt, tid});
        next += take;
        have = min(keep, a.P);
        if (take == 0) break;
    }

    // count valid among the first keep
    __shared__ int s_valid;
    if (tid == 0) s_valid = 0;
    for (int i = tid; i < keep; i += nt) present[i] = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < keep; i += nt) {
        const u64 key = buf[i];
        if (key != 0) {
            ++local;
            const uint32_t d = key_doc(key);
            a.out_ids[q * keep + i] = (int32_t)((int64_t)d + a.id_offset);
            a.out_scores[q * keep + i] = key_score(key);
            if (a.fill && d < (uint32_t)keep) present[d] = 1;
        }
    }
    if (local) atomicAdd(&s_valid, local);
    __syncthreads();
    if (a.fill && tid == 0 && s_valid < keep) {
        int pos = s_valid;
        for (int d = 0; pos < keep; ++d) {
            if (!present[d]) {
                a.out_ids[q * keep + pos] = (int32_t)((int64_t)d + a.id_offset);
                a.out_scores[q * keep + pos] = 0.f;
                ++pos;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_merge_large: the same merge for k_out above what k_merge sorts in shared memory (BM25_SMALL_K).
// One CTA per query, everything in global memory (L2): the candidates are gathered into
// scratch, a radix select isolates the k_out best (select_candidates), the keepers are sorted
// with a bitonic sort in place, then unpacked / zero-filled like k_merge.  O(n) + O(k log^2 k) per
// query -- the rare large-k call (the reference's _topk takes any k <= D, bm25_native.py:204-214).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_merge_large(const MergeArgs a) {
    __shared__ int hist[264];
    __shared__ int s_n, s_valid;
    __shared__ u64 s_theta;
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    const int total = a.n_lists * a.k_in;
    const int keep = a.k_out;
    const CtaGroup grp{kThreads, tid};
    u64* cand = a.scratch + q * (a.n_pad + a.P);
    u64* kept = cand + a.n_pad;
    unsigned char* present = a.present_g + q * keep;
    if (tid == 0) {
        s_valid = 0;
        s_n = 0;
    }
    for (int i = tid; i < keep; i += kThreads) present[i] = 0;
    __syncthreads();
    int local = 0;
    for (int e = tid; e < total; e += kThreads) {
        u64 key;
        if (a.keys) {
            key = a.keys[q * total + e];
        } else {
            const int l = e / a.k_in, r = e - l * a.k_in;
            const int64_t off = (int64_t)l * a.list_stride + q * a.k_in + r;
            const int32_t id = a.in_ids[off];
            key = id < 0 ? 0ull : make_key(a.in_scores[off], (uint32_t)id);
        }
        cand[e] = key;
        if (key) ++local;
    }
    if (local) atomicAdd(&s_valid, local);
    __syncthreads();
    const int n_valid = s_valid;
    int nk;
    if (n_valid > keep) {
        select_candidates(cand, total, keep, kept, hist, &s_n, &s_theta, false, grp);
        nk = keep;
    } else {  // everything valid is a keeper
        for (int e = tid; e < total; e += kThreads) {
            const u64 key = cand[e];
            if (key) kept[atomicAdd(&s_n, 1)] = key;
        }
        nk = n_valid;
    }
    __syncthreads();
    for (int i = nk + tid; i < a.P; i += kThreads) kept[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(kept, a.P, grp);
    int cnt = 0;
    for (int i = tid; i < keep; i += kThreads) {
        const u64 key = kept[i];
        if (key != 0) {
            ++cnt;
            const uint32_t d = key_doc(key);
            a.out_ids[q * keep + i] = (int32_t)((int64_t)d + a.id_offset);
            a.out_scores[q * keep + i] = key_score(key);
            if (a.fill && d < (uint32_t)keep) present[d] = 1;
        }
    }
    __syncthreads();
    if (a.fill && tid == 0 && nk < keep) {
        int pos = nk;
        for (int d = 0; pos < keep; ++d) {
            if (!present[d]) {
                a.out_ids[q * keep + pos] = (int32_t)((int64_t)d + a.id_offset);
                a.out_scores[q * keep + pos] = 0.f;
                ++pos;
            }
        }
    }
    (void)cnt;
}

// ---------------------------------------------------------------------------------------------
// k_term_bounds (load time): bounds[t][l] = the 2^l-th largest weight among the first
// min(df_t, kBoundSample) postings of term t, or 0 when the term has fewer postings.  One CTA of
// 128 threads per term; the sample is sorted in shared memory (as order-preserving keys).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_term_bounds(const int2* __restrict__ tptr, const float* __restrict__ w,
                                                     int n_terms, float* __restrict__ bounds) {
    __shared__ u64 buf[kBoundSample];
    const int t = blockIdx.x;
    const int tid = threadIdx.x;
    const int lo = tptr[t].x;
    const int m = min(tptr[t].y - lo, kBoundSample);
    float* out = bounds + (int64_t)t * kBoundLevels;
    if (m < 1) {
        if (tid < kBoundLevels) out[tid] = 0.f;
        return;
    }
    int P = 2;
    while (P < m) P <<= 1;
    for (int i = tid; i < P; i += 128) buf[i] = i < m ? ((u64)f32_to_ord(w[lo + i]) << 32) : 0ull;
    __syncthreads();
    bitonic_sort_desc(buf, P, CtaGroup{128, tid});
    if (tid < kBoundLevels) {
        const int r = 1 << tid;
        out[tid] = r <= m ? ord_to_f32((uint32_t)(buf[r - 1] >> 32)) : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// k_term_bounds_exact (load time): the same order statistics for the terms with MORE than
// kBoundSample postings, over ALL their postings (k_term_bounds only sees the first kBoundSample,
// whose 2^l-th largest weight is a percentile of the term, not its 2^l-th largest weight -- a much
// looser threshold for the long lists that matter).  One CTA per term: a 4-pass 8-bit radix select
// finds the kBoundSample-th largest weight exactly (weights are > 0: their bit patterns order like
// the values), one more pass collects the larger ones, a bitonic sort orders them.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_term_bounds_exact(const int2* __restrict__ tptr, const float* __restrict__ w,
                                                           const int32_t* __restrict__ big_terms, int n_big,
                                                           float* __restrict__ bounds) {
    __shared__ u64 buf[kBoundSample];
    __shared__ int hist[256];
    __shared__ unsigned s_prefix;
    __shared__ int s_need, s_cnt;
    const int tid = threadIdx.x;
    for (int bi = blockIdx.x; bi < n_big; bi += gridDim.x) {
        const int t = big_terms[bi];
        const int lo = tptr[t].x, hi = tptr[t].y;
        if (tid == 0) {
            s_prefix = 0u;
            s_need = kBoundSample;
            s_cnt = 0;
        }
        unsigned known = 0u;
        __syncthreads();
        for (int shift = 24; shift >= 0; shift -= 8) {
            hist[tid] = 0;
            __syncthreads();
            const unsigned prefix = s_prefix;
            for (int i = lo + tid; i < hi; i += 256) {
                const unsigned b = __float_as_uint(__ldg(w + i));
                if ((b & known) == prefix) atomicAdd(&hist[(b >> shift) & 255], 1);
            }
            __syncthreads();
            if (tid == 0) {  // the digit whose bucket holds the s_need-th largest of the remaining candidates
                int need = s_need, d = 255;
                for (; d > 0; --d) {
                    if (hist[d] >= need) break;
                    need -= hist[d];
                }
                s_need = need;
                s_prefix = prefix | ((unsigned)d << shift);
            }
            known |= 0xffu << shift;
            __syncthreads();
        }
        const unsigned tstar = s_prefix;  // the kBoundSample-th largest weight (bit pattern)
        for (int i = lo + tid; i < hi; i += 256) {
            const unsigned b = __float_as_uint(__ldg(w + i));
            if (b > tstar) buf[atomicAdd(&s_cnt, 1)] = (u64)b << 32;  // fewer than kBoundSample of them
        }
        __syncthreads();
        for (int i = s_cnt + tid; i < kBoundSample; i += 256) buf[i] = (u64)tstar << 32;  // ties at the cut
        __syncthreads();
        bitonic_sort_desc(buf, kBoundSample, CtaGroup{256, tid});
        if (tid < kBoundLevels) bounds[(int64_t)t * kBoundLevels + tid] = __uint_as_float((unsigned)(buf[(1 << tid) - 1] >> 32));
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// load-time validation of device-resident CSC arrays
// flags[0] += out-of-range doc ids, flags[1] += non-finite weights, flags[2] += weights <= 0,
// flags[3] += adjacent inversions (ids[i] <= ids[i-1]), flags[4] += inversions at column starts
// ---------------------------------------------------------------------------------------------
__global__ void k_validate_postings(const int32_t* __restrict__ ids, const float* __restrict__ w,
                                    int64_t nnz, int64_t n_docs, unsigned long long* flags) {
    unsigned long long f0 = 0, f1 = 0, f2 = 0, f3 = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int d = ids[i];
        const float x = w[i];
        if (d < 0 || d >= n_docs) ++f0;
        if (!(fabsf(x) <= 3.402823466e38f)) ++f1;
        if (!(x > 0.f)) ++f2;
        if (i > 0 && d <= ids[i - 1]) ++f3;
    }
    if (f0) atomicAdd(flags + 0, f0);
    if (f1) atomicAdd(flags + 1, f1);
    if (f2) atomicAdd(flags + 2, f2);
    if (f3) atomicAdd(flags + 3, f3);
}

// flags[4] += column starts that look like inversions; flags[5] += indptr defects
__global__ void k_validate_indptr(const int32_t* __restrict__ indptr, const int32_t* __restrict__ ids,
                                  int64_t n_terms, int64_t nnz, unsigned long long* flags) {
    unsigned long long f4 = 0, f5 = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_terms;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = indptr[t], e = indptr[t + 1];
        if (t == 0 && s != 0) ++f5;
        if (t == n_terms - 1 && e != nnz) ++f5;
        if (e < s || s < 0 || e > nnz) { ++f5; continue; }
        if (e > s && s > 0 && ids[s] <= ids[s - 1]) ++f4;
    }
    if (f4) atomicAdd(flags + 4, f4);
    if (f5) atomicAdd(flags + 5, f5);
}

}  // namespace bm25
