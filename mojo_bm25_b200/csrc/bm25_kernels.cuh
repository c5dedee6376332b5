// bm25_kernels.cuh -- hand-written sm_100a kernels of the BM25 query hot path.
//
// Path (reference bm25_native.py:129-158): for every query, gather the posting slices of its
// terms from the CSC index, accumulate the per-(term, doc) weights into per-document scores in
// query-term order (fp32, one add per posting -- bit-identical to the reference's csc mat-vec),
// and select the top-k by (score desc, doc id asc).
//
// Kernels:
//   k_segments      per (query, term): posting-index boundaries of every document chunk
//                   (binary search on doc id inside the term's posting slice)
//   k_score_topk    one warp per (query, document chunk): private shared-memory score tile,
//                   in-order accumulation from per-term cursors, fused threshold-pruned scan
//                   into the CTA's candidate buffer; emits k sorted 64-bit keys per (query, CTA)
//   k_merge         per query: merge candidate lists (tile ranges or GPU shards), zero-score
//                   fill, unpack to (doc id, score)
//   k_validate_*    load-time canonical-form checks of the CSC arrays
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

// ---------------------------------------------------------------------------------------------
// k_segments: seg[(q*(n_tiles+1) + j)*T + t] = first posting index of term queries[q,t] whose doc
// id is >= j*tile_docs (absolute index into ids/w); row n_tiles is the end of the slice.
// Tile-major so that one tile's T boundaries are contiguous.  One warp per (query, term); lanes
// stride over tile boundaries (binary search on doc id inside the term's posting slice).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_segments(const int32_t* __restrict__ indptr,
                                                  const int32_t* __restrict__ ids,
                                                  const int32_t* __restrict__ queries, int64_t n_qt,
                                                  int T, int n_terms, int tile_docs, int n_tiles,
                                                  int32_t* __restrict__ seg,
                                                  const float* __restrict__ bounds, int level,
                                                  u64* __restrict__ theta_q) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_qt) return;
    const int term = queries[warp];
    const int64_t q = warp / T;
    const int t = (int)(warp - q * T);
    int lo0 = 0, hi0 = 0;
    if (term >= 0 && term < n_terms) { lo0 = indptr[term]; hi0 = indptr[term + 1]; }
    // threshold priming: the 2^level-th largest weight of this term belongs to 2^level distinct
    // documents whose score is at least that weight (all weights > 0), so it bounds the k-th best
    // score of the query from below for every k <= 2^level.
    if (lane == 0 && bounds != nullptr && hi0 > lo0) {
        const float b = __ldg(bounds + (int64_t)term * kBoundLevels + level);
        if (b > 0.f) atomicMax(theta_q + q, make_key(b, 0xffffffffu) - 1ull);
    }
    int32_t* out = seg + q * (int64_t)(n_tiles + 1) * T + t;
    for (int j = lane; j <= n_tiles; j += 32) {
        int res;
        if (j == 0) res = lo0;
        else if (j == n_tiles) res = hi0;
        else {
            const int64_t target = (int64_t)j * tile_docs;
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

struct SearchArgs {
    const int32_t* __restrict__ ids;      // [nnz]   doc ids, columns sorted ascending
    const float* __restrict__ w;          // [nnz]   weights
    const int32_t* __restrict__ queries;  // [Q,T]
    const int32_t* __restrict__ seg;      // [Q,n_chunks+1,T]  (k_scores_dense: [Q,n_tiles+1,T])
    u64* __restrict__ partial;            // [Q,splits,k] sorted keys (0 = none)
    u64* theta_q;                         // [Q] best known k-th key per query (shared by its CTAs)
    u64* cand_global;                     // [Q*splits, cap] candidate buffers in global memory (large k), or NULL
    float* __restrict__ dense_out;        // [Q,n_docs]  (k_scores_dense only)
    u64 theta0;                           // initial threshold: key must be > theta0 to compete
    int Q, T, k;
    int n_docs, tile_docs, n_tiles;       // tile_docs = S, documents per warp tile
    int tiles_per_chunk, n_chunks;        // a chunk = the tiles one warp walks
    int splits, tiles_per_split, cap;     // splits = CTAs per query (tiles_per_split: k_scores_dense)
    int general;                          // 1: zero-score docs compete (weights may be <= 0)
    int no_hot;                           // 1: always use the dense tile scan (A/B switch)
    int wide_min;                         // average postings per tile from which a term takes the 128-wide path
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
// k_score_topk (v4): the hot kernel.  Document-at-a-time streaming, one WARP per worker.
//
// CTA = (query, group of NCW document chunks); warp w of the CTA owns chunk sp*NCW + w, a
// contiguous range of `tiles_per_chunk` document tiles of S = tile_docs documents, and a private
// fp32 score tile of S slots in shared memory.  For every tile the warp walks the query's terms
// strictly in query order; per term it keeps a cursor into the term's posting slice (the start
// comes from the segment table, one entry per (query, chunk, term)) and pulls postings with
// coalesced 128-byte loads -- 32 (narrow) or 128 (wide, for dense terms) at a time -- until the
// first posting beyond the tile; each lane adds its in-tile postings into the score tile (one fp32
// add per posting; a term has at most one posting per document and one warp handles every
// posting of a document, so __syncwarp between terms is all the ordering needed: no atomics, no
// CTA barriers, bit-identical to the reference's csc mat-vec).  A per-term "next doc id" lets the
// warp skip terms with nothing in the tile without touching global memory.  After the last term
// the warp pushes the documents that beat the running k-th best key into the CTA's candidate
// buffer -- from the short list of slots whose running score reached the threshold during the
// adds, followed by a write-only clear of the tile, or (list overflow / non-positive weights) by a
// 16-byte vector scan + zero of all S slots.  The buffer is shared by the
// NCW warps; when it overflows they meet in a (rare) named-barrier round, keep the k best and
// raise the threshold, which is also published per query in global memory so that the other
// CTAs of the query prune with it.
// shared memory (dynamic):
//   float score[NCW][S] | u64 cand[cap] (unless in global memory) | int pos[NCW][T] | int cend[NCW][T] | int nxt[NCW][T]
//   | uint16 hot[NCW][kHotCap]
// ---------------------------------------------------------------------------------------------
constexpr int kDocNone = 0x7fffffff;
constexpr int kSelectMin = 256;  // candidate sets larger than this are compacted by radix select

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

struct TopkState {
    u64* cand;
    int* s_ncand;
    int* s_overflow;
    int cap;
    u64 theta;      // key must be > theta to compete
    float theta_f;  // cheap pre-filter on the raw score
    __device__ __forceinline__ void set_theta(u64 t) {
        theta = t;
        theta_f = (t == 0ull) ? -INFINITY : fmaxf(key_score(t), 1.401298464e-45f);
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

// 128 consecutive postings of one term, four per lane (lane, lane+32, lane+64, lane+96)
struct PostingChunk {
    int d0, d1, d2, d3;
    float w0, w1, w2, w3;
    __device__ __forceinline__ void load(const int32_t* __restrict__ ids, const float* __restrict__ w, int p, int r, int lane) {
        const int l0 = lane, l1 = lane + 32, l2 = lane + 64, l3 = lane + 96;
        d0 = l0 < r ? __ldg(ids + p + l0) : kDocNone;
        d1 = l1 < r ? __ldg(ids + p + l1) : kDocNone;
        d2 = l2 < r ? __ldg(ids + p + l2) : kDocNone;
        d3 = l3 < r ? __ldg(ids + p + l3) : kDocNone;
        w0 = l0 < r ? __ldg(w + p + l0) : 0.f;
        w1 = l1 < r ? __ldg(w + p + l1) : 0.f;
        w2 = l2 < r ? __ldg(w + p + l2) : 0.f;
        w3 = l3 < r ? __ldg(w + p + l3) : 0.f;
    }
    __device__ __forceinline__ int doc(int i) const { return i == 0 ? d0 : (i == 1 ? d1 : (i == 2 ? d2 : d3)); }
    // adds the postings with doc < tile_end into the tile; returns how many (a prefix: ids ascend)
    __device__ __forceinline__ int add_into(float* scw, int base, int tile_end, HotList& hl, float theta_f) const {
        const bool in0 = d0 < tile_end, in1 = d1 < tile_end, in2 = d2 < tile_end, in3 = d3 < tile_end;
        // one term has at most one posting per document: the four slots are distinct
        const float n0 = in0 ? scw[d0 - base] + w0 : 0.f;
        const float n1 = in1 ? scw[d1 - base] + w1 : 0.f;
        const float n2 = in2 ? scw[d2 - base] + w2 : 0.f;
        const float n3 = in3 ? scw[d3 - base] + w3 : 0.f;
        if (in0) scw[d0 - base] = n0;
        if (in1) scw[d1 - base] = n1;
        if (in2) scw[d2 - base] = n2;
        if (in3) scw[d3 - base] = n3;
        if (hl.active()) {
            const bool h0 = in0 && n0 >= theta_f, h1 = in1 && n1 >= theta_f;
            const bool h2 = in2 && n2 >= theta_f, h3 = in3 && n3 >= theta_f;
            if (__any_sync(kFull, h0 | h1 | h2 | h3)) {
                hl.add(h0, d0 - base);
                hl.add(h1, d1 - base);
                hl.add(h2, d2 - base);
                hl.add(h3, d3 - base);
            }
        }
        return __popc(__ballot_sync(kFull, in0)) + __popc(__ballot_sync(kFull, in1)) +
               __popc(__ballot_sync(kFull, in2)) + __popc(__ballot_sync(kFull, in3));
    }
};

// vectorised scan + zero of one warp's tile (S documents, the first nd_w of them real)
__device__ __forceinline__ bool tile_scan(float* scw, int S, int nd_w, uint32_t doc0, int lane, TopkState& tk) {
    bool left = false;
#pragma unroll 4
    for (int idx = lane * 4; idx < S; idx += 128) {
        const float4 v = *reinterpret_cast<const float4*>(scw + idx);
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) >= tk.theta_f) {
            const float vv[4] = {v.x, v.y, v.z, v.w};
            float zz[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (vv[e] >= tk.theta_f && idx + e < nd_w) {
                    if (!tk.push(vv[e], doc0 + (uint32_t)(idx + e))) { zz[e] = vv[e]; left = true; }
                }
            }
            z = make_float4(zz[0], zz[1], zz[2], zz[3]);
        }
        *reinterpret_cast<float4*>(scw + idx) = z;
    }
    return left;
}

// MAXT = 256: up to 8 warps, three CTAs per SM (<= 80 registers); MAXT = 512: up to 16 warps, two CTAs per SM
template <int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? 3 : 2) k_score_topk(const SearchArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NCW = blockDim.x >> 5;
    const int T = a.T, S = a.tile_docs, cap = a.cap;
    float* sc = reinterpret_cast<float*>(smem_raw);
    // candidate buffer: shared memory, or (large k) this CTA's slice of a global, L2-resident array --
    // pushes are rare once the threshold has settled, and the freed shared memory buys a third CTA per SM
    u64* cand_smem = reinterpret_cast<u64*>(sc + (size_t)NCW * S);
    u64* cand = a.cand_global ? a.cand_global + (size_t)blockIdx.x * cap : cand_smem;
    int* st_pos = reinterpret_cast<int*>(cand_smem + (a.cand_global ? 0 : cap));
    int* st_end = st_pos + NCW * T;
    int* st_nxt = st_end + NCW * T;
    unsigned short* st_hot = reinterpret_cast<unsigned short*>(st_nxt + NCW * T);
    __shared__ int s_ncand, s_overflow;
    __shared__ u64 s_theta;
    __shared__ int s_hist[264];  // select_candidates scratch

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int q = blockIdx.x / a.splits;
    const int sp = blockIdx.x - q * a.splits;
    const int chunk = sp * NCW + warp;
    const WarpsGroup grp{(int)blockDim.x, tid};

    float* scw = sc + (size_t)warp * S;
    for (int i = lane * 4; i < S; i += 128) *reinterpret_cast<float4*>(scw + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        s_ncand = 0;
        s_overflow = 0;
        const u64 shared_theta = a.theta_q ? *reinterpret_cast<volatile u64*>(a.theta_q + q) : 0ull;
        s_theta = shared_theta > a.theta0 ? shared_theta : a.theta0;
    }
    __syncthreads();

    TopkState tk{cand, &s_ncand, &s_overflow, cap, 0ull, 0.f};
    tk.set_theta(s_theta);
    bool leftover = false;       // this warp's tile still holds scores that did not fit in cand
    uint32_t left_doc0 = 0;
    int left_nd = 0;
    HotList hl{st_hot + warp * kHotCap, kHotCap + 1, (1u << lane) - 1u};

    u64* out = a.partial + ((int64_t)q * a.splits + sp) * a.k;  // this CTA's k output keys (also select scratch)

    // candidate-buffer overflow round: every warp of the CTA takes part
    auto overflow_round = [&]() {
        for (;;) {
            if (cap > kSelectMin)  // a round is only entered with a full buffer (n = cap > k)
                select_candidates(cand, cap, a.k, out, s_hist, &s_ncand, &s_theta, true, grp);
            else
                compact_candidates(cand, cap, a.k, a.theta0, &s_ncand, &s_theta, grp);
            if (tid == 0) {
                s_overflow = 0;
                if (a.theta_q) {  // share the threshold with the other CTAs of this query
                    const u64 mine = s_theta;
                    const u64 old = atomicMax(a.theta_q + q, mine);
                    if (old > mine) s_theta = old;
                }
            }
            grp.sync();
            tk.set_theta(s_theta);
            if (leftover) leftover = __any_sync(kFull, tile_scan(scw, S, left_nd, left_doc0, lane, tk));
            grp.sync();
            const int again = ld_volatile(&s_overflow);
            grp.sync();
            if (!again) break;
        }
    };

    if (chunk < a.n_chunks) {
        int* pos_w = st_pos + warp * T;
        int* end_w = st_end + warp * T;
        int* nxt_w = st_nxt + warp * T;
        const int32_t* seg0 = a.seg + ((int64_t)q * (a.n_chunks + 1) + chunk) * T;
        for (int t = lane; t < T; t += 32) {
            const int p = __ldg(seg0 + t), e = __ldg(seg0 + T + t);
            pos_w[t] = p;
            end_w[t] = e;
            nxt_w[t] = (p < e) ? __ldg(a.ids + p) : kDocNone;
        }
        __syncwarp();
        const int j0 = chunk * a.tiles_per_chunk;
        const int j1 = min(a.n_tiles, j0 + a.tiles_per_chunk);
        for (int j = j0; j < j1; ++j) {
            const int base = j * S;
            const int tile_end = base + S;
            bool touched = false;
            hl.reset(!a.general && !a.no_hot);
            // ---- accumulate: terms strictly in query order ------------------------------------
            // Dense terms (>= wide_min postings per remaining tile on average) stream 128 postings
            // per step with the next step's loads already in flight.  Runs of sparse terms are taken
            // up to four at a time: the first 32 postings of each are requested together
            // (memory-level parallelism across terms), then added one term after the other.
            const int wide_cut = a.wide_min * (j1 - j);
            for (int g0 = 0; g0 < T; g0 += 32) {
                const int tl = g0 + lane;
                unsigned act = __ballot_sync(kFull, tl < T && nxt_w[min(tl, T - 1)] < tile_end);
                if (act) touched = true;
                while (act) {
                    const int t0 = g0 + __ffs(act) - 1;
                    int p = pos_w[t0];
                    int e = end_w[t0];
                    if (e - p >= wide_cut) {
                        act &= act - 1;
                        int nx = kDocNone;
                        PostingChunk A, B;
                        A.load(a.ids, a.w, p, e - p, lane);
                        for (;;) {
                            B.load(a.ids, a.w, p + 128, e - p - 128, lane);
                            const int c = A.add_into(scw, base, tile_end, hl, tk.theta_f);
                            p += c;
                            if (c < 128) {  // the first posting beyond the tile (if any) is element c
                                nx = __shfl_sync(kFull, A.doc(c >> 5), c & 31);
                                break;
                            }
                            if (p >= e) break;
                            A = B;
                        }
                        if (lane == 0) {
                            pos_w[t0] = p;
                            nxt_w[t0] = nx;
                        }
                        __syncwarp();
                        continue;
                    }
                    int tt[4] = {-1, -1, -1, -1}, pp[4], ee[4], dd[4];
                    float ww[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (act) {
                            const int t = g0 + __ffs(act) - 1;
                            const int pt = pos_w[t], et = end_w[t];
                            const int r = et - pt;
                            if (g == 0 || r < wide_cut) {  // a dense term ends the run (handled above next)
                                act &= act - 1;
                                tt[g] = t;
                                pp[g] = pt;
                                ee[g] = et;
                                dd[g] = lane < r ? __ldg(a.ids + pt + lane) : kDocNone;
                                ww[g] = lane < r ? __ldg(a.w + pt + lane) : 0.f;
                            } else {
                                break;
                            }
                        }
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (tt[g] < 0) break;  // warp-uniform
                        p = pp[g];
                        e = ee[g];
                        int nx = kDocNone;
                        int d = dd[g];
                        float w = ww[g];
                        for (;;) {
                            const bool in = d < tile_end;
                            float nw = 0.f;
                            if (in) {
                                nw = scw[d - base] + w;
                                scw[d - base] = nw;
                            }
                            if (hl.active()) hl.add(in && nw >= tk.theta_f, d - base);
                            const int c = __popc(__ballot_sync(kFull, in));
                            p += c;
                            if (c < 32) {
                                nx = __shfl_sync(kFull, d, c);
                                break;
                            }
                            if (p >= e) break;
                            const int r = e - p;
                            d = lane < r ? __ldg(a.ids + p + lane) : kDocNone;
                            w = lane < r ? __ldg(a.w + p + lane) : 0.f;
                        }
                        if (lane == 0) {
                            pos_w[tt[g]] = p;
                            nxt_w[tt[g]] = nx;
                        }
                        __syncwarp();
                    }
                }
            }
            // ---- epilogue: push the documents that beat the k-th best so far, clear the tile ---
            if (touched || a.general) {
                const int nd_w = min(S, a.n_docs - base);
                bool left = false;
                if (!hl.active()) {
                    left = tile_scan(scw, S, nd_w, (uint32_t)base, lane, tk);  // reads, tests and zeroes all S slots
                } else {
                    __syncwarp();
                    for (int i0 = 0; i0 < hl.n; i0 += 32) {
                        const int i = i0 + lane;
                        const int x = i < hl.n ? (int)hl.slots[i] : -1 - lane;
                        const unsigned same = __match_any_sync(kFull, x);  // a slot may be listed twice
                        if (x >= 0 && (__ffs(same) - 1) == lane) {
                            const float v = scw[x];
                            if (v > 0.f) {  // not yet taken by an earlier step of this loop
                                if (tk.push(v, (uint32_t)(base + x))) scw[x] = 0.f;
                                else left = true;
                            }
                        }
                        __syncwarp();
                    }
                    left = __any_sync(kFull, left);
                    if (!left) {
                        for (int idx = lane * 4; idx < S; idx += 128)
                            *reinterpret_cast<float4*>(scw + idx) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                if (__any_sync(kFull, left)) {  // candidate buffer full: the overflow round rescans this tile
                    leftover = true;
                    left_doc0 = (uint32_t)base;
                    left_nd = nd_w;
                }
            }
            if (leftover || ld_volatile(&s_overflow)) {
                grp.sync();
                overflow_round();
            }
        }
    }
    for (;;) {  // finished: keep serving overflow rounds until every warp of the CTA is here
        grp.sync();
        if (!ld_volatile(&s_overflow)) break;
        overflow_round();
    }
    const int n_end = min(ld_volatile(&s_ncand), cap);
    if (n_end > kSelectMin && n_end > a.k) {
        // the k keepers go straight to the output slot, unsorted (k_merge sorts what it loads)
        select_candidates(cand, n_end, a.k, out, s_hist, &s_ncand, &s_theta, false, grp);
        if (tid == 0 && a.theta_q) atomicMax(a.theta_q + q, s_theta);
        return;
    }
    if (a.cand_global) {
        // here n_end <= k (a larger set went through the select above): everything is a keeper, unsorted
        for (int i = tid; i < a.k; i += blockDim.x) out[i] = (i < n_end) ? cand[i] : 0ull;
        return;
    }
    compact_candidates(cand, cap, a.k, a.theta0, &s_ncand, &s_theta, grp);
    if (tid == 0 && a.theta_q && s_ncand >= a.k) atomicMax(a.theta_q + q, s_theta);
    const int n = s_ncand;
    for (int i = tid; i < a.k; i += blockDim.x) out[i] = (i < n) ? cand[i] : 0ull;
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
};

__global__ void __launch_bounds__(kThreads) k_merge(const MergeArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u64* buf = reinterpret_cast<u64*>(smem_raw);
    unsigned char* present = reinterpret_cast<unsigned char*>(buf + a.P);
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    const int total = a.n_lists * a.k_in;
    const int keep = a.k_out;

    // rounds: buf[0..have) holds the best so far (sorted); append up to P - have new keys, sort.
    int have = 0, next = 0;
    while (next < total || have == 0) {
        const int room = a.P - have;
        const int take = min(room, total - next);
        for (int i = tid; i < room; i += kThreads) {
            u64 key = 0;
            if (i < take) {
                const int e = next + i;
                if (a.keys) {
                    key = a.keys[q * total + e];
                } else {
                    const int l = e / a.k_in, r = e - l * a.k_in;
                    const int64_t off = (int64_t)l * a.list_stride + q * a.k_in + r;
                    key = make_key(a.in_scores[off], (uint32_t)a.in_ids[off]);
                }
            }
            buf[have + i] = key;
        }
        __syncthreads();
        bitonic_sort_desc(buf, a.P, CtaGroup{kThreads, tid});
        next += take;
        have = min(keep, a.P);
        if (take == 0) break;
    }

    // count valid among the first keep
    __shared__ int s_valid;
    if (tid == 0) s_valid = 0;
    for (int i = tid; i < keep; i += kThreads) present[i] = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < keep; i += kThreads) {
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
// k_term_bounds (load time): bounds[t][l] = the 2^l-th largest weight among the first
// min(df_t, kBoundSample) postings of term t, or 0 when the term has fewer postings.  One CTA of
// 128 threads per term; the sample is sorted in shared memory (as order-preserving keys).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_term_bounds(const int32_t* __restrict__ indptr, const float* __restrict__ w,
                                                     int n_terms, float* __restrict__ bounds) {
    __shared__ u64 buf[kBoundSample];
    const int t = blockIdx.x;
    const int tid = threadIdx.x;
    const int lo = indptr[t];
    const int m = min(indptr[t + 1] - lo, kBoundSample);
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
