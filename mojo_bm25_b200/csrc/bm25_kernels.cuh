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
// k_segments: seg[(q*T+t)*(n_tiles+1) + j] = first posting index of term queries[q,t] whose doc
// id is >= j*tile_docs (absolute index into ids/w); entry n_tiles is the end of the slice.
// One warp per (query, term); lanes stride over tile boundaries.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_segments(const int32_t* __restrict__ indptr,
                                                  const int32_t* __restrict__ ids,
                                                  const int32_t* __restrict__ queries, int64_t n_qt,
                                                  int n_terms, int tile_docs, int n_tiles,
                                                  int32_t* __restrict__ seg) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_qt) return;
    const int term = queries[warp];
    int lo0 = 0, hi0 = 0;
    if (term >= 0 && term < n_terms) { lo0 = indptr[term]; hi0 = indptr[term + 1]; }
    int32_t* out = seg + warp * (int64_t)(n_tiles + 1);
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
        out[j] = res;
    }
}

// ---------------------------------------------------------------------------------------------
// block-wide bitonic sort (descending) of P = 2^m keys in shared memory
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_sort_desc(u64* buf, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
                const int a = 2 * i - (i & (stride - 1));
                const int b = a + stride;
                const bool desc = ((a & size) == 0);
                const u64 x = buf[a], y = buf[b];
                if ((x < y) == desc) { buf[a] = y; buf[b] = x; }
            }
            __syncthreads();
        }
    }
}

// Keep the k best of the n candidates in cand[0..n) (sorted, best first), raise the threshold.
// Must be called by all threads of the CTA with no push in flight.
__device__ __forceinline__ void compact_candidates(u64* cand, int k, u64 theta0, int* s_ncand,
                                                   u64* s_theta) {
    const int n = *s_ncand;
    int P = 2;
    while (P < n) P <<= 1;
    for (int i = n + threadIdx.x; i < P; i += blockDim.x) cand[i] = 0;
    __syncthreads();
    bitonic_sort_desc(cand, P);
    if (threadIdx.x == 0) {
        *s_ncand = n < k ? n : k;
        *s_theta = (n >= k) ? cand[k - 1] : theta0;
    }
    __syncthreads();
}

struct SearchArgs {
    const int32_t* __restrict__ ids;      // [nnz]   doc ids, columns sorted ascending
    const float* __restrict__ w;          // [nnz]   weights
    const int32_t* __restrict__ queries;  // [Q,T]
    const int32_t* __restrict__ seg;      // [Q,T,n_tiles+1]
    u64* __restrict__ partial;            // [Q,splits,k] sorted keys (0 = none)
    float* __restrict__ dense_out;        // [Q,n_docs]  (dense-output variant only)
    u64 theta0;                           // initial threshold: key must be > theta0 to compete
    int Q, T, k;
    int n_docs, tile_docs, n_tiles;
    int splits, tiles_per_split, cap;
};

// ---------------------------------------------------------------------------------------------
// k_score_topk: CTA = (query, range of document tiles).
// shared memory: float score[tile_docs] | u64 cand[cap] | int seg_lo[T] | int seg_hi[T]
// ---------------------------------------------------------------------------------------------
template <bool kDenseOut>
__global__ void __launch_bounds__(kThreads, 2) k_score_topk(const SearchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sc = reinterpret_cast<float*>(smem_raw);
    u64* cand = reinterpret_cast<u64*>(smem_raw + (size_t)a.tile_docs * sizeof(float));
    int* s_lo = reinterpret_cast<int*>(cand + (kDenseOut ? 0 : a.cap));
    int* s_hi = s_lo + a.T;
    __shared__ int s_ncand;
    __shared__ u64 s_theta;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int q = blockIdx.x / a.splits;
    const int sp = blockIdx.x - q * a.splits;
    const int j0 = sp * a.tiles_per_split;
    const int j1 = min(a.n_tiles, j0 + a.tiles_per_split);

    for (int i = tid * 4; i < a.tile_docs; i += kChunk)
        *reinterpret_cast<float4*>(sc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) { s_ncand = 0; s_theta = a.theta0; }
    u64 theta = a.theta0;
    const int32_t* segq = a.seg + (int64_t)q * a.T * (a.n_tiles + 1);
    __syncthreads();

    for (int j = j0; j < j1; ++j) {
        const int base = j * a.tile_docs;
        const int nd = min(a.tile_docs, a.n_docs - base);
        for (int t = tid; t < a.T; t += kThreads) {
            const int32_t* p = segq + (int64_t)t * (a.n_tiles + 1) + j;
            s_lo[t] = __ldg(p);
            s_hi[t] = __ldg(p + 1);
        }
        __syncthreads();

        // ---- accumulate: terms strictly in query order, one fp32 add per posting -------------
        for (int t = 0; t < a.T; ++t) {
            const int lo = s_lo[t], hi = s_hi[t];
            if (lo >= hi) continue;  // uniform
            for (int i = lo + tid; i < hi; i += kChunk) {
                const int i1 = i + kThreads, i2 = i + 2 * kThreads, i3 = i + 3 * kThreads;
                const bool v1 = i1 < hi, v2 = i2 < hi, v3 = i3 < hi;
                const int d0 = __ldg(a.ids + i);
                const float w0 = __ldg(a.w + i);
                int d1 = 0, d2 = 0, d3 = 0;
                float w1 = 0.f, w2 = 0.f, w3 = 0.f;
                if (v1) { d1 = __ldg(a.ids + i1); w1 = __ldg(a.w + i1); }
                if (v2) { d2 = __ldg(a.ids + i2); w2 = __ldg(a.w + i2); }
                if (v3) { d3 = __ldg(a.ids + i3); w3 = __ldg(a.w + i3); }
                sc[d0 - base] += w0;
                if (v1) sc[d1 - base] += w1;
                if (v2) sc[d2 - base] += w2;
                if (v3) sc[d3 - base] += w3;
            }
            __syncthreads();
        }

        if (kDenseOut) {
            float* out = a.dense_out + (int64_t)q * a.n_docs + base;
            for (int i = tid; i < nd; i += kThreads) { out[i] = sc[i]; sc[i] = 0.f; }
            __syncthreads();
            continue;
        }

        // ---- fused scan + zero: push every doc whose key beats the running k-th best ---------
        for (int c0 = 0; c0 < nd; c0 += kChunk) {
            const int idx = c0 + tid * 4;
            const float4 v = *reinterpret_cast<const float4*>(sc + idx);
            *reinterpret_cast<float4*>(sc + idx) = make_float4(0.f, 0.f, 0.f, 0.f);
            const uint32_t doc = (uint32_t)(base + idx);
            const u64 k0 = make_key(v.x, doc), k1 = make_key(v.y, doc + 1);
            const u64 k2 = make_key(v.z, doc + 2), k3 = make_key(v.w, doc + 3);
            const bool p0 = (idx < nd) && k0 > theta, p1 = (idx + 1 < nd) && k1 > theta;
            const bool p2 = (idx + 2 < nd) && k2 > theta, p3 = (idx + 3 < nd) && k3 > theta;
            const int cnt = (int)p0 + (int)p1 + (int)p2 + (int)p3;
            bool risk = false;
            if (__ballot_sync(kFull, cnt > 0)) {
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += y;
                }
                const int total = __shfl_sync(kFull, incl, 31);
                int wbase = 0;
                if (lane == 31) wbase = atomicAdd(&s_ncand, total);
                wbase = __shfl_sync(kFull, wbase, 31);
                int pos = wbase + incl - cnt;
                if (p0) cand[pos++] = k0;
                if (p1) cand[pos++] = k1;
                if (p2) cand[pos++] = k2;
                if (p3) cand[pos++] = k3;
                risk = (wbase + total > a.cap - kChunk);
            }
            if (__syncthreads_or(risk)) {
                compact_candidates(cand, a.k, a.theta0, &s_ncand, &s_theta);
                theta = s_theta;
            }
        }
    }

    if (kDenseOut) return;
    compact_candidates(cand, a.k, a.theta0, &s_ncand, &s_theta);
    const int n = s_ncand;
    u64* out = a.partial + ((int64_t)q * a.splits + sp) * a.k;
    for (int i = tid; i < a.k; i += kThreads) out[i] = (i < n) ? cand[i] : 0ull;
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
    extern __shared__ __align__(16) unsigned char smem_raw[];
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
        bitonic_sort_desc(buf, a.P);
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
