// bm25_kernels.cuh -- hand-written sm_100a kernels of the BM25 query hot path.
//
// Path (reference bm25_native.py:129-158): for every query, gather the posting slices of its
// terms from the CSC index, accumulate the per-(term, doc) weights into per-document scores in
// query-term order (fp32, one add per posting -- bit-identical to the reference's csc mat-vec),
// and select the top-k by (score desc, doc id asc).
//
// Kernels:
//   k_segments      per (query, term): posting-index boundaries of every document tile
//                   (binary search on doc id inside the term's posting slice)
//   k_score_topk    one CTA per (query, tile range): shared-memory score tile, in-order
//                   accumulation, fused threshold-pruned scan into a candidate buffer,
//                   block-wide select; emits k sorted 64-bit keys per (query, range)
//   k_merge         per query: merge candidate lists (tile ranges or GPU shards), zero-score
//                   fill, unpack to (doc id, score)
//   k_validate_*    load-time canonical-form checks of the CSC arrays
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bm25 {

typedef unsigned long long u64;

constexpr int kThreads = 512;          // threads per CTA of k_score_topk / k_merge
constexpr int kChunk = kThreads * 4;   // documents scanned per block-wide step (float4 / thread)
constexpr unsigned kFull = 0xffffffffu;

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
                                                  int32_t* __restrict__ seg) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_qt) return;
    const int term = queries[warp];
    const int64_t q = warp / T;
    const int t = (int)(warp - q * T);
    int lo0 = 0, hi0 = 0;
    if (term >= 0 && term < n_terms) { lo0 = indptr[term]; hi0 = indptr[term + 1]; }
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
// group barriers: the whole CTA (barrier 0) or the consumer warps of k_score_topk (barrier 1)
// ---------------------------------------------------------------------------------------------
struct CtaGroup {
    int size, rank;
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
struct ConsumerGroup {
    int size, rank;
    __device__ __forceinline__ void sync() const {
        asm volatile("bar.sync 1, %0;" ::"r"(size) : "memory");
    }
};

// group-wide bitonic sort (descending) of P = 2^m keys in shared memory
template <typename G>
__device__ __forceinline__ void bitonic_sort_desc(u64* buf, int P, const G& g) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = g.rank; i < (P >> 1); i += g.size) {
                const int a = 2 * i - (i & (stride - 1));
                const int b = a + stride;
                const bool desc = ((a & size) == 0);
                const u64 x = buf[a], y = buf[b];
                if ((x < y) == desc) { buf[a] = y; buf[b] = x; }
            }
            g.sync();
        }
    }
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

struct SearchArgs {
    const int32_t* __restrict__ ids;      // [nnz]   doc ids, columns sorted ascending
    const float* __restrict__ w;          // [nnz]   weights
    const int32_t* __restrict__ queries;  // [Q,T]
    const int32_t* __restrict__ seg;      // [Q,n_tiles+1,T]
    u64* __restrict__ partial;            // [Q,splits,k] sorted keys (0 = none)
    float* __restrict__ dense_out;        // [Q,n_docs]  (k_scores_dense only)
    u64 theta0;                           // initial threshold: key must be > theta0 to compete
    int Q, T, k;
    int n_docs, tile_docs, n_tiles;
    int splits, tiles_per_split, cap;
    int stage_postings;                   // capacity of one staging buffer (multiple of 4)
    int n_stages;                         // staging ring depth (2..4)
    int sparse_max;                       // tiles with at most this many postings use the sparse epilogue
    int general;                          // 1: zero-score docs compete (weights may be <= 0)
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
// mbarrier / bulk-copy (TMA 1-D) primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(u64* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk async copy; completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// k_score_topk (v3): the hot kernel.  CTA = (query, range of document tiles), warp-specialised,
// no CTA-wide barrier on the steady-state path:
//   producer warp : walks the tiles, packs the query terms' posting segments of a tile into
//                   "rounds" that fit one staging buffer and issues them as 1-D bulk async copies
//                   (TMA, UBLKCP) into an NS-stage shared-memory ring; completion on mbarriers.
//   searcher warp : for every staged piece longer than kFilterMax postings, binary-searches (in
//                   shared memory) the boundaries of the NCW document stripes of the tile, then
//                   hands the stage to the consumers (second mbarrier).
//   consumer warps: warp w OWNS stripe w of the score tile (tile_docs / NCW documents).  It adds
//                   the postings of its stripe, piece by piece in query-term order (one fp32 add
//                   per posting; a term has at most one posting per document and every posting of
//                   a document is handled by the same warp, so __syncwarp between pieces is the
//                   only ordering needed -- no atomics, no named barriers).  Epilogue per tile:
//                     sparse tile (all postings resident in one stage, few of them): re-walk the
//                       staged postings, test the final score against the running k-th best,
//                       push survivors into the candidate buffer and zero the touched slots;
//                     dense tile: vectorised scan + zero of the warp's stripe.
//                   The candidate buffer is shared by the CTA; when it overflows the consumers
//                   meet at a (rare) named-barrier round, keep the k best and raise the threshold.
// shared memory (dynamic):
//   float score[tile_docs] | int32 st_ids[NS][stg] | float st_w[NS][stg] | u64 cand[cap]
//   | u64 full[4], ready[4], empty[4] | int rd[4][8] | int pc_so[NS][T] | int pc_cnt[NS][T]
//   | int bnd[NS][T][NCW+1] | int p_lo[T], p_hi[T]
// ---------------------------------------------------------------------------------------------
constexpr int kFilterMax = 64;   // pieces up to this many postings are filtered, not searched
constexpr int kMaxStages = 4;
constexpr int kFlagTileEnd = 1, kFlagSparse = 2;

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

// vectorised scan + zero of one warp's stripe (S documents, the first nd_w of them real)
__device__ __forceinline__ bool stripe_scan(float* scw, int S, int nd_w, uint32_t doc0, int lane, TopkState& tk) {
    bool left = false;
#pragma unroll 2
    for (int idx = lane * 4; idx < S; idx += 128) {
        const float4 v = *reinterpret_cast<const float4*>(scw + idx);
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v.x >= tk.theta_f || v.y >= tk.theta_f || v.z >= tk.theta_f || v.w >= tk.theta_f) {
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

template <int NCW>
__global__ void __launch_bounds__((NCW + 2) * 32) k_score_topk(const SearchArgs a) {
    constexpr int NC = NCW * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int T = a.T, stg = a.stage_postings, cap = a.cap, NS = a.n_stages;
    float* sc = reinterpret_cast<float*>(smem_raw);
    int32_t* st_ids = reinterpret_cast<int32_t*>(smem_raw + (size_t)a.tile_docs * 4);
    float* st_w = reinterpret_cast<float*>(st_ids + NS * stg);
    u64* cand = reinterpret_cast<u64*>(st_w + NS * stg);
    u64* bar_full = cand + cap;
    u64* bar_ready = bar_full + kMaxStages;
    u64* bar_empty = bar_ready + kMaxStages;
    int* rd = reinterpret_cast<int*>(bar_empty + kMaxStages);  // [4][8] = {n_pieces, base, nd, flags}
    int* pc_so = rd + kMaxStages * 8;
    int* pc_cnt = pc_so + NS * T;
    int* bnd = pc_cnt + NS * T;
    int* p_lo = bnd + NS * T * (NCW + 1);
    int* p_hi = p_lo + T;
    __shared__ int s_ncand, s_overflow;
    __shared__ u64 s_theta;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int q = blockIdx.x / a.splits;
    const int sp = blockIdx.x - q * a.splits;
    const int j0 = sp * a.tiles_per_split;
    const int j1 = min(a.n_tiles, j0 + a.tiles_per_split);
    const int S = a.tile_docs / NCW;  // documents per consumer stripe (multiple of 128)

    for (int i = tid * 4; i < a.tile_docs; i += (NC + 64) * 4)
        *reinterpret_cast<float4*>(sc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        s_ncand = 0;
        s_overflow = 0;
        s_theta = a.theta0;
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_ready + s, 32);
            mbar_init(bar_empty + s, NCW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NCW) {
        // =============================== producer warp ========================================
        const int32_t* segq = a.seg + (int64_t)q * (a.n_tiles + 1) * T;
        int stage = 0;
        uint32_t phase = 0;
        for (int t = lane; t < T; t += 32) p_hi[t] = __ldg(segq + (int64_t)j0 * T + t);
        for (int j = j0; j < j1; ++j) {
            int tot = 0, npost = 0;
            for (int t = lane; t < T; t += 32) {
                const int lo = p_hi[t];
                const int hi = __ldg(segq + (int64_t)(j + 1) * T + t);
                p_lo[t] = lo;
                p_hi[t] = hi;
                if (hi > lo) {
                    tot += ((hi + 3) & ~3) - (lo & ~3);
                    npost += hi - lo;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                tot += __shfl_xor_sync(kFull, tot, o);
                npost += __shfl_xor_sync(kFull, npost, o);
            }
            __syncwarp();
            if (npost == 0 && !a.general) continue;  // no posting of this query in the tile
            int t = 0;
            while (t < T && p_hi[t] <= p_lo[t]) ++t;
            int pos = (t < T) ? p_lo[t] : 0;
            const int base = j * a.tile_docs;
            const int nd = min(a.tile_docs, a.n_docs - base);
            // sparse epilogue only when the whole tile is resident in ONE round (tot + 8 <= stg guarantees it)
            const int sparse = (!a.general && tot + 8 <= stg && npost <= a.sparse_max) ? kFlagSparse : 0;
            do {  // one round = one staging buffer
                mbar_wait(bar_empty + stage, phase ^ 1);
                int used = 0, np = 0;
                uint32_t bytes = 0;
                while (t < T) {
                    const int room = stg - used;
                    if (room < 8) break;
                    const int a0 = pos & ~3;
                    const int skip = pos - a0;
                    const int take = min(p_hi[t] - pos, room - skip);
                    const int ncopy = ((pos + take + 3) & ~3) - a0;
                    if (lane == 0) {
                        pc_so[stage * T + np] = stage * stg + used + skip;
                        pc_cnt[stage * T + np] = take;
                        bulk_copy_g2s(st_ids + stage * stg + used, a.ids + a0, (uint32_t)ncopy * 4u, bar_full + stage);
                        bulk_copy_g2s(st_w + stage * stg + used, a.w + a0, (uint32_t)ncopy * 4u, bar_full + stage);
                    }
                    bytes += (uint32_t)ncopy * 8u;
                    used += ncopy;
                    ++np;
                    pos += take;
                    if (pos < p_hi[t]) break;  // buffer full, the term continues in the next round
                    ++t;
                    while (t < T && p_hi[t] <= p_lo[t]) ++t;
                    if (t < T) pos = p_lo[t];
                }
                if (lane == 0) {
                    rd[stage * 8 + 0] = np;
                    rd[stage * 8 + 1] = base;
                    rd[stage * 8 + 2] = nd;
                    rd[stage * 8 + 3] = ((t >= T) ? kFlagTileEnd : 0) | sparse;
                    mbar_arrive_expect_tx(bar_full + stage, bytes);
                }
                __syncwarp();
                if (++stage == NS) { stage = 0; phase ^= 1; }
            } while (t < T);
        }
        mbar_wait(bar_empty + stage, phase ^ 1);
        if (lane == 0) {
            rd[stage * 8 + 0] = -1;  // end of stream
            mbar_arrive(bar_full + stage);
        }
        return;
    }

    if (warp == NCW + 1) {
        // =============================== searcher warp ========================================
        int stage = 0;
        uint32_t phase = 0;
        for (;;) {
            mbar_wait(bar_full + stage, phase);
            const int np = rd[stage * 8 + 0];
            if (np > 0) {
                const int base = rd[stage * 8 + 1];
                const int total = np * (NCW + 1);
                for (int idx = lane; idx < total; idx += 32) {
                    const int i = idx / (NCW + 1);
                    const int b = idx - i * (NCW + 1);
                    const int cnt = pc_cnt[stage * T + i];
                    if (cnt <= kFilterMax) continue;
                    const int so = pc_so[stage * T + i];
                    int lo = 0, hi = cnt;
                    if (b == 0) hi = 0;
                    else if (b == NCW) lo = cnt;
                    else {
                        const int target = base + b * S;
                        const int32_t* p = st_ids + so;
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (p[mid] < target) lo = mid + 1; else hi = mid;
                        }
                    }
                    bnd[(stage * T + i) * (NCW + 1) + b] = so + lo;
                }
            }
            mbar_arrive(bar_ready + stage);  // all 32 lanes arrive (release of the bnd writes)
            if (np < 0) return;
            if (++stage == NS) { stage = 0; phase ^= 1; }
        }
    }

    // ================================= consumer warps =========================================
    const ConsumerGroup grp{NC, tid};
    TopkState tk{cand, &s_ncand, &s_overflow, cap, 0ull, 0.f};
    tk.set_theta(a.theta0);
    float* scw = sc + warp * S;
    bool leftover = false;       // this warp's stripe still holds scores that did not fit in cand
    uint32_t left_doc0 = 0;
    int left_nd = 0;

    // candidate-buffer overflow round: every consumer warp takes part (see the protocol above)
    auto overflow_round = [&]() {
        for (;;) {
            compact_candidates(cand, cap, a.k, a.theta0, &s_ncand, &s_theta, grp);
            if (tid == 0) s_overflow = 0;
            grp.sync();
            tk.set_theta(s_theta);
            if (leftover) leftover = __any_sync(kFull, stripe_scan(scw, S, left_nd, left_doc0, lane, tk));
            grp.sync();
            const int again = ld_volatile(&s_overflow);
            grp.sync();
            if (!again) break;
        }
    };

    int stage = 0;
    uint32_t phase = 0;
    for (;;) {
        mbar_wait(bar_ready + stage, phase);
        mbar_wait(bar_full + stage, phase);  // already complete; orders the bulk-copy writes for this thread
        const int np = rd[stage * 8 + 0];
        if (np < 0) break;
        const int base = rd[stage * 8 + 1];
        const int nd = rd[stage * 8 + 2];
        const int flags = rd[stage * 8 + 3];
        const int sbase = base + warp * S;  // first document of this warp's stripe
        // ---- accumulate the round's pieces: terms strictly in query order ---------------------
        for (int i = 0; i < np; ++i) {
            const int cnt = pc_cnt[stage * T + i];
            const int so = pc_so[stage * T + i];
            if (cnt <= kFilterMax) {
                for (int e = so + lane; e < so + cnt; e += 32) {
                    const unsigned d = (unsigned)(st_ids[e] - sbase);
                    if (d < (unsigned)S) scw[d] += st_w[e];
                }
            } else {
                const int lo = bnd[(stage * T + i) * (NCW + 1) + warp];
                const int hi = bnd[(stage * T + i) * (NCW + 1) + warp + 1];
                for (int e = lo + lane; e < hi; e += 128) {
                    const bool v1 = e + 32 < hi, v2 = e + 64 < hi, v3 = e + 96 < hi;
                    const int i0 = st_ids[e] - sbase;
                    const int i1 = v1 ? st_ids[e + 32] - sbase : 0;
                    const int i2 = v2 ? st_ids[e + 64] - sbase : 0;
                    const int i3 = v3 ? st_ids[e + 96] - sbase : 0;
                    const float w0 = st_w[e];
                    const float w1 = v1 ? st_w[e + 32] : 0.f;
                    const float w2 = v2 ? st_w[e + 64] : 0.f;
                    const float w3 = v3 ? st_w[e + 96] : 0.f;
                    // one term has at most one posting per document: the four slots are distinct
                    const float s0 = scw[i0];
                    const float s1 = v1 ? scw[i1] : 0.f;
                    const float s2 = v2 ? scw[i2] : 0.f;
                    const float s3 = v3 ? scw[i3] : 0.f;
                    scw[i0] = s0 + w0;
                    if (v1) scw[i1] = s1 + w1;
                    if (v2) scw[i2] = s2 + w2;
                    if (v3) scw[i3] = s3 + w3;
                }
            }
            __syncwarp();
        }
        bool left = false;
        const bool sparse = (flags & (kFlagSparse | kFlagTileEnd)) == (kFlagSparse | kFlagTileEnd);
        if (sparse) {
            // ---- sparse epilogue: re-walk the staged postings, test + zero the touched slots ---
            for (int i = 0; i < np; ++i) {
                const int cnt = pc_cnt[stage * T + i];
                const int so = pc_so[stage * T + i];
                if (cnt <= kFilterMax) {
                    for (int e = so + lane; e < so + cnt; e += 32) {
                        const unsigned d = (unsigned)(st_ids[e] - sbase);
                        if (d < (unsigned)S) {
                            const float v = scw[d];
                            if (v >= tk.theta_f && !tk.push(v, (uint32_t)sbase + d)) left = true;
                            else scw[d] = 0.f;
                        }
                    }
                } else {
                    const int lo = bnd[(stage * T + i) * (NCW + 1) + warp];
                    const int hi = bnd[(stage * T + i) * (NCW + 1) + warp + 1];
                    for (int e = lo + lane; e < hi; e += 128) {
                        const bool v1 = e + 32 < hi, v2 = e + 64 < hi, v3 = e + 96 < hi;
                        const int i0 = st_ids[e] - sbase;
                        const int i1 = v1 ? st_ids[e + 32] - sbase : i0;
                        const int i2 = v2 ? st_ids[e + 64] - sbase : i0;
                        const int i3 = v3 ? st_ids[e + 96] - sbase : i0;
                        const float s0 = scw[i0], s1 = scw[i1], s2 = scw[i2], s3 = scw[i3];
                        bool k0 = false, k1 = false, k2 = false, k3 = false;  // keep (buffer full)
                        if (s0 >= tk.theta_f) k0 = !tk.push(s0, (uint32_t)(sbase + i0));
                        if (v1 && s1 >= tk.theta_f) k1 = !tk.push(s1, (uint32_t)(sbase + i1));
                        if (v2 && s2 >= tk.theta_f) k2 = !tk.push(s2, (uint32_t)(sbase + i2));
                        if (v3 && s3 >= tk.theta_f) k3 = !tk.push(s3, (uint32_t)(sbase + i3));
                        if (!k0) scw[i0] = 0.f;
                        if (v1 && !k1) scw[i1] = 0.f;
                        if (v2 && !k2) scw[i2] = 0.f;
                        if (v3 && !k3) scw[i3] = 0.f;
                        left |= k0 | k1 | k2 | k3;
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + stage);  // staging buffer may be refilled
        if (++stage == NS) { stage = 0; phase ^= 1; }
        if (flags & kFlagTileEnd) {
            const int nd_w = min(max(nd - warp * S, 0), S);
            if (!sparse) left = stripe_scan(scw, S, nd_w, (uint32_t)sbase, lane, tk);
            if (__any_sync(kFull, left)) {
                leftover = true;
                left_doc0 = (uint32_t)sbase;
                left_nd = nd_w;
            }
        }
        // checked after EVERY round (not only at tile ends): a warp that is ahead in the ring must be
        // able to join the barrier without waiting for a stage the blocked warps have not released
        if (leftover || ld_volatile(&s_overflow)) {
            grp.sync();
            overflow_round();
        }
    }
    for (;;) {  // finished: keep serving overflow rounds until every consumer warp is here
        grp.sync();
        if (!ld_volatile(&s_overflow)) break;
        overflow_round();
    }
    compact_candidates(cand, cap, a.k, a.theta0, &s_ncand, &s_theta, grp);
    const int n = s_ncand;
    u64* out = a.partial + ((int64_t)q * a.splits + sp) * a.k;
    for (int i = tid; i < a.k; i += NC) out[i] = (i < n) ? cand[i] : 0ull;
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
