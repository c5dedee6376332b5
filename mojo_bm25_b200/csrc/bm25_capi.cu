// bm25_capi.cu -- host side of libbm25_b200.so: the C ABI declared in include/bm25_b200.h.
// Owns the HBM-resident index, the per-handle workspace (no cudaMalloc on the steady-state search
// path) and the launch logic of the kernels in bm25_kernels.cuh.
#include "../../include/bm25_b200.h"
#include "bm25_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

using namespace bm25;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(e__ == cudaErrorMemoryAllocation ? BM25_ERR_OOM : BM25_ERR_CUDA,      \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,    \
                        __LINE__);                                                            \
    } while (0)

int next_pow2(int64_t x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    int reserve(size_t n) {
        if (n <= cap) return BM25_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(BM25_ERR_OOM, "cudaMalloc of %zu bytes failed: %s", want * sizeof(T),
                        cudaGetErrorString(e));
        }
        cap = want;
        return BM25_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    size_t bytes() const { return cap * sizeof(T); }
};

template <typename T>
struct PinnedBuf {
    T* p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return BM25_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMallocHost(&p, want * sizeof(T));
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(BM25_ERR_OOM, "cudaMallocHost of %zu bytes failed: %s", want * sizeof(T),
                        cudaGetErrorString(e));
        }
        cap = want;
        return BM25_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

int check_device(int device, cudaDeviceProp* prop) {
    static std::mutex mu;
    static cudaDeviceProp cache[64];
    static bool cached[64] = {false};
    {
        std::lock_guard<std::mutex> lock(mu);
        if (device >= 0 && device < 64 && cached[device]) {
            *prop = cache[device];
            return BM25_OK;
        }
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(BM25_ERR_NO_DEVICE,
                    "no CUDA device available (%s); libbm25_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count)
        return fail(BM25_ERR_INVALID, "device %d out of range (have %d)", device, count);
    CU(cudaGetDeviceProperties(prop, device));
    if (prop->major < 10)
        return fail(BM25_ERR_NO_DEVICE,
                    "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop->major, prop->minor);
    if (device < 64) {
        std::lock_guard<std::mutex> lock(mu);
        cache[device] = *prop;
        cached[device] = true;
    }
    return BM25_OK;
}

}  // namespace

struct TileTableSet {
    int32_t* d_term_row = nullptr;  // [n_terms] tile-table row or -1
    int32_t* d_tab = nullptr;       // [n_heavy, tab_tiles + 1]
    int64_t n_heavy = 0, tab_bytes = 0;
    int tab_tile_docs = 0, tab_heavy_min = 0;  // what the table was built for (0 = not built)
    uint32_t* d_pk = nullptr;       // compressed handles: [nnz_padded] packed postings for pk_tile_docs documents per tile
    int pk_tile_docs = 0;
    void release() {
        if (d_term_row) cudaFree(d_term_row);
        if (d_tab) cudaFree(d_tab);
        if (d_pk) cudaFree(d_pk);
        *this = TileTableSet{};
    }
};

struct bm25_index {
    int device = 0;
    int sm_count = 148;
    int64_t n_terms = 0, n_docs = 0, nnz = 0, doc_id_base = 0;
    bool all_positive = false, was_sorted = true;
    // the re-bucketed index (layout documented in bm25_kernels.cuh)
    int64_t nnz_padded = 0;
    int2* d_tptr = nullptr;         // [n_terms] {start, end} of every term in the padded arrays
    float2* d_wrange = nullptr;     // [n_terms] {smallest, largest} weight of the term
    int32_t* d_ids = nullptr;       // [nnz_padded]
    float* d_w = nullptr;           // [nnz_padded]
    // everything that depends on the tile size: the active set and one stashed set, so that a handle
    // serving two query shapes with different tile sizes (launch plan) does not rebuild on every switch
    TileTableSet tt, tt_stash;
    float* d_bounds = nullptr;  // [n_terms][kBoundLevels] per-term weight order statistics (threshold priming)
    // compressed index (bm25_index_compress): weights rounded to bf16, 4-byte packed postings
    int weight_format = BM25_WEIGHTS_FP32;
    std::vector<int32_t> h_indptr;  // host copy for byte accounting / heavy-term selection
    // options (0 = auto)
    int opt_tile_docs = 0, opt_splits = 0, opt_force_general = 0, opt_timing = 0;
    int opt_warps = 0, opt_cap = 0, opt_waves = 0, opt_no_theta_share = 0, opt_no_priming = 0, opt_no_hot = 0, opt_heavy_min = 0, opt_cand_smem = 0;
    int opt_poison = 0, opt_no_bulk_clear = 0, opt_no_query_sort = 0, opt_generic_kernel = 0, opt_q_major = 0, opt_no_epoch = 0, opt_no_packed = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // seg | score | merge boundaries
    bool ev_valid = false;
    // workspace: one per handle; searches on different streams are ordered through ws_done
    std::mutex mu;
    cudaEvent_t ws_done = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_used = false;
    DevBuf<int32_t> ws_seg;
    DevBuf<u64> ws_partial, ws_theta, ws_cand, ws_qkey;
    DevBuf<int32_t> ws_qperm;
    DevBuf<int32_t> ws_queries, ws_out_ids;
    DevBuf<float> ws_out_scores;
    PinnedBuf<int32_t> pin_queries, pin_out_ids;
    PinnedBuf<float> pin_out_scores;
    cudaStream_t own_stream = nullptr;
    size_t smem_optin = 0, smem_per_sm = 0;

    int warps() const { return opt_warps > 0 ? opt_warps : 8; }
    int tile_docs() const {  // S: documents per warp tile (multiple of 128)
        int t = opt_tile_docs > 0 ? opt_tile_docs : 2048;
        const int64_t need = std::max<int64_t>(((n_docs + 127) / 128) * 128, 128);
        if (need < t) t = (int)need;
        return ((t + 127) / 128) * 128;
    }
    int n_tiles() const { return (int)std::max<int64_t>(1, (n_docs + tile_docs() - 1) / tile_docs()); }
    int64_t device_bytes() const {
        int64_t b = n_terms * 8 + nnz_padded * 8;
        for (const TileTableSet* t : {&tt, &tt_stash})
            b += t->tab_bytes + (t->d_term_row ? n_terms * 4 : 0) + (t->d_pk ? nnz_padded * 4 : 0);
        if (d_bounds) b += n_terms * kBoundLevels * 4;
        b += ws_seg.bytes() + ws_partial.bytes() + ws_theta.bytes() + ws_cand.bytes() + ws_qkey.bytes() + ws_qperm.bytes() + ws_queries.bytes() + ws_out_ids.bytes() +
             ws_out_scores.bytes();
        return b;
    }
};

namespace {

// Per-term weight order statistics for threshold priming (k_term_bounds / k_term_bounds_exact),
// from the handle's current weights; dropped when the index holds non-positive weights.
int compute_bounds(bm25_index* ix) {
    const int64_t V = ix->n_terms;
    if (!(ix->all_positive && V > 0 && ix->nnz > 0)) {
        if (ix->d_bounds) cudaFree(ix->d_bounds);
        ix->d_bounds = nullptr;
        return BM25_OK;
    }
    if (!ix->d_bounds && cudaMalloc(&ix->d_bounds, (size_t)V * kBoundLevels * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        ix->d_bounds = nullptr;
        return fail(BM25_ERR_OOM, "cudaMalloc of the term-bound table failed");
    }
    k_term_bounds<<<(unsigned)V, 128>>>(ix->d_tptr, ix->d_w, (int)V, ix->d_bounds);
    ++g_launches;
    CU(cudaGetLastError());
    // long lists: exact order statistics over all their postings
    std::vector<int32_t> big;
    for (int64_t t = 0; t < V; ++t)
        if (ix->h_indptr[t + 1] - ix->h_indptr[t] > kBoundSample) big.push_back((int32_t)t);
    if (!big.empty() && !getenv("BM25_B200_SAMPLED_BOUNDS")) {
        int32_t* d_big = nullptr;
        if (cudaMalloc(&d_big, big.size() * 4) != cudaSuccess) {
            cudaGetLastError();
            return fail(BM25_ERR_OOM, "cudaMalloc of the long-term list failed");
        }
        cudaMemcpy(d_big, big.data(), big.size() * 4, cudaMemcpyHostToDevice);
        const int grid = (int)std::min<size_t>(big.size(), (size_t)ix->sm_count * 8);
        k_term_bounds_exact<<<grid, 256>>>(ix->d_tptr, ix->d_w, d_big, (int)big.size(), ix->d_bounds);
        ++g_launches;
        cudaError_t e = cudaDeviceSynchronize();
        cudaFree(d_big);
        if (e != cudaSuccess) return fail(BM25_ERR_CUDA, "k_term_bounds_exact failed: %s", cudaGetErrorString(e));
    }
    return BM25_OK;
}

// Builds the re-bucketed index from canonical CSC arrays resident on the device (h_indptr is the
// host copy of the column pointers): padded posting arrays + {start, end} per term, then the
// per-term weight order statistics.  The tile table is built lazily (ensure_table) because it
// depends on the tile size.
int finish_create(bm25_index* ix, const cudaDeviceProp& prop, const int32_t* d_indptr_raw, const int32_t* d_ids_raw,
                  const float* d_w_raw) {
    ix->sm_count = prop.multiProcessorCount;
    ix->smem_optin = prop.sharedMemPerBlockOptin;
    ix->smem_per_sm = prop.sharedMemPerMultiprocessor;
    CU(cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ix->ws_done, cudaEventDisableTiming));
    const int64_t V = ix->n_terms;
    std::vector<int2> tptr((size_t)std::max<int64_t>(V, 1));
    int64_t pos = 0;
    for (int64_t t = 0; t < V; ++t) {
        const int64_t df = ix->h_indptr[t + 1] - ix->h_indptr[t];
        tptr[t] = make_int2((int)pos, (int)(pos + df));
        pos += (df + 4) & ~(int64_t)3;  // >= 1 sentinel posting after every list, starts 16-byte aligned
        if (pos > 0x7fffffffLL - 8)
            return fail(BM25_ERR_UNSUPPORTED, "index too large for int32 posting offsets after padding (%lld)",
                        (long long)pos);
    }
    ix->nnz_padded = pos;
    if (cudaMalloc(&ix->d_tptr, (size_t)std::max<int64_t>(V, 1) * sizeof(int2)) != cudaSuccess ||
        cudaMalloc(&ix->d_ids, (size_t)(pos + 4) * 4) != cudaSuccess ||
        cudaMalloc(&ix->d_w, (size_t)(pos + 4) * 4) != cudaSuccess ||
        cudaMalloc(&ix->d_wrange, (size_t)std::max<int64_t>(V, 1) * sizeof(float2)) != cudaSuccess) {
        cudaGetLastError();
        return fail(BM25_ERR_OOM, "cudaMalloc of the index (%lld postings) failed", (long long)ix->nnz);
    }
    CU(cudaMemcpy(ix->d_tptr, tptr.data(), (size_t)V * sizeof(int2), cudaMemcpyHostToDevice));
    if (V > 0 && pos > 0) {
        const int grid = (int)std::min<int64_t>(V, (int64_t)ix->sm_count * 16);
        k_relayout<<<grid, 128>>>(d_indptr_raw, d_ids_raw, d_w_raw, ix->d_tptr, (int)V, ix->d_ids, ix->d_w, ix->d_wrange);
        ++g_launches;
        CU(cudaGetLastError());
    }
    if (int rc = compute_bounds(ix)) return rc;
    CU(cudaDeviceSynchronize());
    return BM25_OK;
}

// The tile table for S = tile_docs documents per tile (built on first use, rebuilt when the tile
// size or the heavy-term threshold changes).  heavy_min is in 1/16 postings per tile: a term is
// heavy when df * 16 >= heavy_min * n_tiles.
int ensure_table(bm25_index* ix, int S) {
    int hm = ix->opt_heavy_min > 0 ? ix->opt_heavy_min : 16;
    if (ix->tt.tab_tile_docs == S && ix->tt.tab_heavy_min == hm) return BM25_OK;
    if (ix->tt_stash.tab_tile_docs == S && ix->tt_stash.tab_heavy_min == hm) {
        std::swap(ix->tt, ix->tt_stash);  // nothing is freed: searches in flight keep valid pointers
        return BM25_OK;
    }
    CU(cudaDeviceSynchronize());  // no search may still be reading the set that is evicted
    ix->tt_stash.release();
    if (ix->tt.tab_tile_docs) std::swap(ix->tt, ix->tt_stash);
    TileTableSet& tt = ix->tt;
    tt.release();
    const int64_t V = ix->n_terms;
    const int64_t NB = std::max<int64_t>(1, (ix->n_docs + S - 1) / S);
    std::vector<int32_t> row((size_t)std::max<int64_t>(V, 1), -1), heavy;
    const int64_t limit = std::max<int64_t>(4 * ix->nnz, (int64_t)1 << 22);  // table entries
    int eff = hm;
    for (;;) {
        heavy.clear();
        for (int64_t t = 0; t < V; ++t) {
            const int64_t df = ix->h_indptr[t + 1] - ix->h_indptr[t];
            if (df > 0 && df * 16 >= (int64_t)eff * NB) heavy.push_back((int32_t)t);
        }
        if ((int64_t)heavy.size() * (NB + 1) <= limit || eff > (1 << 28)) break;
        eff *= 2;
    }
    for (size_t r = 0; r < heavy.size(); ++r) row[heavy[r]] = (int32_t)r;
    const int64_t entries = (int64_t)heavy.size() * (NB + 1);
    if (cudaMalloc(&tt.d_term_row, (size_t)std::max<int64_t>(V, 1) * 4) != cudaSuccess) {
        cudaGetLastError();
        tt.d_term_row = nullptr;
        return fail(BM25_ERR_OOM, "cudaMalloc of the term-row map failed");
    }
    if (V > 0) CU(cudaMemcpy(tt.d_term_row, row.data(), (size_t)V * 4, cudaMemcpyHostToDevice));
    if (entries > 0) {
        int32_t* d_heavy = nullptr;
        if (cudaMalloc(&tt.d_tab, (size_t)entries * 4) != cudaSuccess ||
            cudaMalloc(&d_heavy, heavy.size() * 4) != cudaSuccess) {
            cudaGetLastError();
            tt.release();
            return fail(BM25_ERR_OOM, "cudaMalloc of the tile table (%lld entries) failed", (long long)entries);
        }
        cudaMemcpy(d_heavy, heavy.data(), heavy.size() * 4, cudaMemcpyHostToDevice);
        k_build_table<<<(unsigned)((entries + 255) / 256), 256>>>(ix->d_tptr, d_heavy, ix->d_ids, entries, (int)NB, S,
                                                                  tt.d_tab);
        ++g_launches;
        cudaError_t e = cudaDeviceSynchronize();
        cudaFree(d_heavy);
        if (e != cudaSuccess) {
            tt.release();
            return fail(BM25_ERR_CUDA, "building the tile table failed: %s", cudaGetErrorString(e));
        }
    }
    tt.n_heavy = (int64_t)heavy.size();
    tt.tab_bytes = entries * 4;
    tt.tab_tile_docs = S;
    tt.tab_heavy_min = hm;
    return BM25_OK;
}

// Compressed handles: the 4-byte packed postings for S documents per tile (rebuilt with the tile size).
int ensure_packed(bm25_index* ix, int S) {  // after ensure_table(ix, S): packs for the active set
    TileTableSet& tt = ix->tt;
    if (ix->weight_format == BM25_WEIGHTS_FP32 || tt.pk_tile_docs == S) return BM25_OK;
    if (S > kPkMaxTileDocs)
        return fail(BM25_ERR_UNSUPPORTED, "a compressed index needs tile_docs <= %d (16-bit tile-local slots), got %d",
                    kPkMaxTileDocs, S);
    CU(cudaDeviceSynchronize());  // no search may still be reading the old array
    if (!tt.d_pk && cudaMalloc(&tt.d_pk, (size_t)(ix->nnz_padded + 4) * 4) != cudaSuccess) {
        cudaGetLastError();
        tt.d_pk = nullptr;
        return fail(BM25_ERR_OOM, "cudaMalloc of the packed postings (%lld) failed", (long long)ix->nnz_padded);
    }
    tt.pk_tile_docs = 0;
    if (ix->nnz_padded > 0) {
        k_pack<<<ix->sm_count * 8, 256>>>(ix->d_ids, ix->d_w, ix->nnz_padded, S, tt.d_pk);
        ++g_launches;
        CU(cudaGetLastError());
        CU(cudaDeviceSynchronize());
    }
    tt.pk_tile_docs = S;
    return BM25_OK;
}

// Host-side canonicalisation: returns false if any column needed sorting / merging.
int canonicalise_host(const int32_t* indptr, const int32_t* indices, const float* data, int64_t n_terms,
                      int64_t n_docs, int64_t nnz, std::vector<int32_t>& o_ptr,
                      std::vector<int32_t>& o_idx, std::vector<float>& o_dat, bool* was_sorted,
                      bool* all_positive) {
    if (indptr[0] != 0) return fail(BM25_ERR_INVALID, "indptr[0] must be 0 (got %d)", indptr[0]);
    if (indptr[n_terms] != nnz)
        return fail(BM25_ERR_INVALID, "indptr[n_terms] (%d) != nnz (%lld)", indptr[n_terms], (long long)nnz);
    bool sorted = true, positive = true;
    for (int64_t t = 0; t < n_terms; ++t) {
        const int64_t s = indptr[t], e = indptr[t + 1];
        if (e < s) return fail(BM25_ERR_INVALID, "indptr is not monotone at term %lld", (long long)t);
        for (int64_t p = s; p < e; ++p) {
            const int32_t d = indices[p];
            if (d < 0 || d >= n_docs)
                return fail(BM25_ERR_INVALID, "doc id %d of posting %lld is outside [0, %lld)", d,
                            (long long)p, (long long)n_docs);
            const float x = data[p];
            if (!std::isfinite(x))
                return fail(BM25_ERR_INVALID, "weight of posting %lld is not finite", (long long)p);
            if (!(x > 0.f)) positive = false;
            if (p > s && d <= indices[p - 1]) sorted = false;
        }
    }
    *was_sorted = sorted;
    *all_positive = positive;
    if (sorted) return BM25_OK;
    // sort each column by doc id (stable) and merge duplicate rows by summing in original order
    o_ptr.assign(n_terms + 1, 0);
    o_idx.clear();
    o_dat.clear();
    o_idx.reserve(nnz);
    o_dat.reserve(nnz);
    std::vector<int64_t> perm;
    for (int64_t t = 0; t < n_terms; ++t) {
        const int64_t s = indptr[t], e = indptr[t + 1];
        perm.resize(e - s);
        std::iota(perm.begin(), perm.end(), s);
        std::stable_sort(perm.begin(), perm.end(),
                         [&](int64_t a, int64_t b) { return indices[a] < indices[b]; });
        for (size_t i = 0; i < perm.size(); ++i) {
            const int32_t d = indices[perm[i]];
            const float x = data[perm[i]];
            if (!o_idx.empty() && (int64_t)o_idx.size() > o_ptr[t] && o_idx.back() == d)
                o_dat.back() += x;
            else {
                o_idx.push_back(d);
                o_dat.push_back(x);
            }
        }
        o_ptr[t + 1] = (int32_t)o_idx.size();
    }
    for (float x : o_dat)
        if (!(x > 0.f)) positive = false;
    *all_positive = positive;
    return BM25_OK;
}

struct LaunchPlan {
    int tile_docs, n_tiles, splits, tiles_per_split, cap, warps, general;
    int tiles_per_chunk, n_chunks;  // k_score_topk: a chunk = the tiles one warp walks
    int cand_global;                // candidate buffers live in global memory (k > kSelectMin)
    int seg_docs, seg_rows;         // granularity / row count of the segment table
    size_t smem;
    u64 theta0;
};

// dynamic shared memory of k_score_topk (layout documented at the kernel)
size_t score_smem(int tile_docs, int cap, int64_t T, int warps) {  // cap = 0: candidates in global memory
    return (size_t)warps * tile_docs * 4 + (size_t)cap * 8 + (size_t)warps * T * kStateInts * 4 + (size_t)warps * kHotCap * 2 + 128;
}

void plan_mode(const bm25_index* ix, LaunchPlan* lp) {
    const bool positive = ix->all_positive && !ix->opt_force_general;
    // positive index: only strictly positive scores compete, zero-score docs are filled in by
    // k_merge; general index: every document competes (theta0 = 0 admits all keys).
    lp->theta0 = positive ? make_key(0.0f, 0u) : 0ull;
    lp->general = positive ? 0 : 1;
}

// k_scores_dense (parity/debug): CTA = (query, range of 16K-document tiles)
int make_plan_dense(bm25_index* ix, int64_t Q, int64_t T, LaunchPlan* lp) {
    *lp = LaunchPlan{};
    lp->warps = kThreads / 32;
    int t = 16384;
    const int64_t need = std::max<int64_t>(((ix->n_docs + 1023) / 1024) * 1024, 1024);
    if (need < t) t = (int)need;
    lp->tile_docs = t;
    lp->smem = (size_t)lp->tile_docs * 4 + (size_t)T * 8 + 16;
    if (lp->smem > ix->smem_optin - 1024)
        return fail(BM25_ERR_UNSUPPORTED, "query width T=%lld does not fit in shared memory", (long long)T);
    lp->n_tiles = (int)std::max<int64_t>(1, (ix->n_docs + lp->tile_docs - 1) / lp->tile_docs);
    const int64_t slots = (int64_t)ix->sm_count * 2 * 4;
    int splits = (int)std::max<int64_t>(1, (slots + Q - 1) / std::max<int64_t>(Q, 1));
    splits = std::max(1, std::min(splits, lp->n_tiles));
    lp->tiles_per_split = (lp->n_tiles + splits - 1) / splits;
    lp->splits = (lp->n_tiles + lp->tiles_per_split - 1) / lp->tiles_per_split;
    lp->seg_docs = lp->tile_docs;
    lp->seg_rows = lp->n_tiles;
    plan_mode(ix, lp);
    return BM25_OK;
}

int make_plan(bm25_index* ix, int64_t Q, int64_t T, int k, LaunchPlan* lp) {
    *lp = LaunchPlan{};
    lp->warps = ix->warps();
    lp->tile_docs = ix->tile_docs();
    lp->cap = ix->opt_cap > 0 ? next_pow2(ix->opt_cap)
                              : next_pow2(std::max((k <= 1024 ? 4 : 2) * (int64_t)k, (int64_t)512));
    if (lp->cap < k + 64) lp->cap = next_pow2((int64_t)k + 64);
    // large k: the candidate buffer moves to global memory (see the kernel); it is then compacted by
    // radix select only, which needs cap > kSelectMin
    lp->cand_global = (k > kSelectMin && !ix->opt_cand_smem) ? 1 : 0;
    // wide queries (T ~ 64) need so much cursor state that only two 8-warp CTAs with 2048-document tiles
    // fit on an SM: take the warp count (8, 7 or 6) and, when the tile size is automatic, the tile size
    // (2048 or 1792 documents) that keep the most warps resident (E: 3 x 8 warps on 1792-document tiles
    // instead of 3 x 7 on 2048: -3 %); ties go to the larger tile, then to more warps per CTA
    if (ix->opt_warps <= 0) {
        int best_w = lp->warps, best_s = lp->tile_docs, best_res = 0;
        const int s_lo = (ix->opt_tile_docs <= 0 && lp->tile_docs == 2048) ? 1792 : lp->tile_docs;
        for (int sd = lp->tile_docs; sd >= s_lo; sd -= 256) {
            for (int w = lp->warps; w >= 6; --w) {
                const size_t sm = score_smem(sd, lp->cand_global ? 0 : lp->cap, T, w) + 1024 + 1152;
                // CTAs per SM: shared memory, thread slots, and the register file (80 registers per thread)
                const size_t ctas = std::min<size_t>(std::min<size_t>(ix->smem_per_sm / sm, 2048 / (w * 32)), 65536 / (80 * 32 * w));
                const int res = (int)ctas * w;
                if (res > best_res) {
                    best_res = res;
                    best_w = w;
                    best_s = sd;
                }
            }
        }
        lp->warps = best_w;
        lp->tile_docs = best_s;
    }
    // shrink the CTA (fewer warps, then smaller tiles) until it fits into shared memory
    const size_t hard = ix->smem_optin - 1024;
    for (;;) {
        lp->smem = score_smem(lp->tile_docs, lp->cand_global ? 0 : lp->cap, T, lp->warps);
        if (lp->smem <= hard) break;
        if (lp->warps > 4) { lp->warps -= 1; continue; }
        if (lp->tile_docs > 512) { lp->tile_docs -= 128; continue; }
        return fail(BM25_ERR_UNSUPPORTED, "query shape (T=%lld, k=%d) does not fit in shared memory",
                    (long long)T, k);
    }
    lp->n_tiles = (int)std::max<int64_t>(1, (ix->n_docs + lp->tile_docs - 1) / lp->tile_docs);
    const int max_splits = (lp->n_tiles + lp->warps - 1) / lp->warps;  // one tile per warp
    int splits = ix->opt_splits;
    if (splits <= 0) {
        const int64_t per_sm = std::max<int64_t>(1, std::min<int64_t>((int64_t)(ix->smem_per_sm - 1024) / (int64_t)(lp->smem + 1024),
                                                                      2048 / (lp->warps * 32)));
        // finer CTAs balance the tail; with a large k every CTA pays for big candidate sorts, so fewer
        const int64_t waves = ix->opt_waves > 0 ? ix->opt_waves : (k > 256 ? 4 : 10);
        const int64_t want = waves * per_sm * ix->sm_count;  // CTAs in flight x waves
        splits = (int)std::max<int64_t>(1, (want + Q - 1) / std::max<int64_t>(Q, 1));
        // ... but a warp should walk at least ~32 tiles: the per-chunk setup (cursor starts, table
        // ring, candidate compaction, one more list for k_merge) is not free (B: 489 tiles per query)
        if (ix->opt_waves <= 0) splits = (int)std::min<int64_t>(splits, std::max<int64_t>(1, (lp->n_tiles + lp->warps * 16) / (lp->warps * 32)));
    }
    if (k > BM25_SMALL_K && ix->opt_splits <= 0) splits = std::min(splits, 2);  // 2k-key candidate buffers per CTA
    splits = std::max(1, std::min(splits, max_splits));
    lp->tiles_per_chunk = (lp->n_tiles + splits * lp->warps - 1) / (splits * lp->warps);
    lp->n_chunks = (lp->n_tiles + lp->tiles_per_chunk - 1) / lp->tiles_per_chunk;
    lp->splits = (lp->n_chunks + lp->warps - 1) / lp->warps;
    lp->seg_docs = lp->tiles_per_chunk * lp->tile_docs;
    lp->seg_rows = lp->n_chunks;
    plan_mode(ix, lp);
    if (getenv("BM25_B200_DEBUG_PLAN"))
        fprintf(stderr, "plan: tile_docs=%d warps=%d splits=%d tiles_per_chunk=%d n_chunks=%d cap=%d smem=%zu cand_global=%d smem_per_sm=%zu\n",
                lp->tile_docs, lp->warps, lp->splits, lp->tiles_per_chunk, lp->n_chunks, lp->cap, lp->smem, lp->cand_global, ix->smem_per_sm);
    return BM25_OK;
}

template <typename Kern>
int configure_smem(Kern kern, size_t need, size_t smem_optin, size_t* configured) {
    if (*configured >= need) return BM25_OK;
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, kern));
    const size_t max_dyn = smem_optin - fa.sharedSizeBytes;
    if (need > max_dyn) return fail(BM25_ERR_UNSUPPORTED, "kernel needs %zu B of shared memory", need);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn));
    *configured = max_dyn;
    return BM25_OK;
}

int launch_score(bm25_index* ix, const LaunchPlan& lp, const SearchArgs& a, int64_t Q, bool dense, cudaStream_t st) {
    const int64_t grid = Q * lp.splits;
    if (grid > 0x7fffffffLL) return fail(BM25_ERR_UNSUPPORTED, "grid too large");
    int rc;
    if (dense) {
        static thread_local size_t configured[64] = {0};
        if ((rc = configure_smem(k_scores_dense, lp.smem, ix->smem_optin, &configured[ix->device % 64]))) return rc;
        k_scores_dense<<<(unsigned)grid, kThreads, lp.smem, st>>>(a);
    } else {
        static thread_local size_t configured[2][64] = {{0}, {0}};
        static thread_local size_t configured_s[2][64] = {{0}, {0}};
        if (a.T <= 32 && !ix->opt_generic_kernel) {  // lane-per-term specialisation
            static thread_local size_t configured_p[2][64] = {{0}, {0}};
            if (a.pk) {  // compressed handle: 4-byte packed postings
                if (lp.warps <= BM25_LB_T / 32) {
                    if ((rc = configure_smem(k_score_topk_s<BM25_LB_T, true>, lp.smem, ix->smem_optin, &configured_p[0][ix->device % 64]))) return rc;
                    k_score_topk_s<BM25_LB_T, true><<<(unsigned)grid, lp.warps * 32, lp.smem, st>>>(a);
                } else {
                    if ((rc = configure_smem(k_score_topk_s<512, true>, lp.smem, ix->smem_optin, &configured_p[1][ix->device % 64]))) return rc;
                    k_score_topk_s<512, true><<<(unsigned)grid, lp.warps * 32, lp.smem, st>>>(a);
                }
            } else if (lp.warps <= BM25_LB_T / 32) {
                if ((rc = configure_smem(k_score_topk_s<BM25_LB_T, false>, lp.smem, ix->smem_optin, &configured_s[0][ix->device % 64]))) return rc;
                k_score_topk_s<BM25_LB_T, false><<<(unsigned)grid, lp.warps * 32, lp.smem, st>>>(a);
            } else {
                if ((rc = configure_smem(k_score_topk_s<512, false>, lp.smem, ix->smem_optin, &configured_s[1][ix->device % 64]))) return rc;
                k_score_topk_s<512, false><<<(unsigned)grid, lp.warps * 32, lp.smem, st>>>(a);
            }
        } else if (lp.warps <= BM25_LB_T / 32) {
            if ((rc = configure_smem(k_score_topk<BM25_LB_T>, lp.smem, ix->smem_optin, &configured[0][ix->device % 64]))) return rc;
            k_score_topk<BM25_LB_T><<<(unsigned)grid, lp.warps * 32, lp.smem, st>>>(a);
        } else {
            if ((rc = configure_smem(k_score_topk<512>, lp.smem, ix->smem_optin, &configured[1][ix->device % 64]))) return rc;
            k_score_topk<512><<<(unsigned)grid, lp.warps * 32, lp.smem, st>>>(a);
        }
    }
    ++g_launches;
    CU(cudaGetLastError());
    return BM25_OK;
}

// k_segments, or -- when `order` is given -- k_segments_order: the same work plus the batch ordering
// (k_query_order) in the last CTA of the same launch.
int launch_segments(bm25_index* ix, const LaunchPlan& lp, const int32_t* d_queries, int64_t Q, int64_t T, int k,
                    u64* theta_q, bool all_terms, cudaStream_t st, const OrderArgs* order = nullptr) {
    const int64_t n_qt = Q * T;
    int rc = ix->ws_seg.reserve((size_t)n_qt * (lp.seg_rows + 1));
    if (rc) return rc;
    // threshold priming needs positive weights, a shared per-query threshold and k <= 2^(levels-1)
    int level = 0;
    while ((1 << level) < k) ++level;
    const bool prime = theta_q && ix->d_bounds && !lp.general && level < kBoundLevels && !ix->opt_no_priming;
    SegArgs g{};
    g.tptr = ix->d_tptr;
    g.term_row = all_terms ? nullptr : ix->tt.d_term_row;
    g.ids = ix->d_ids;
    g.queries = d_queries;
    g.n_qt = n_qt;
    g.T = (int)T;
    g.n_terms = (int)ix->n_terms;
    g.row_docs = lp.seg_docs;
    g.n_rows = lp.seg_rows;
    g.seg = ix->ws_seg.p;
    g.bounds = prime ? ix->d_bounds : nullptr;
    g.level = level;
    g.theta_q = theta_q;
    if (order) {
        const int64_t blocks = (n_qt * 32 + 1023) / 1024 + 1;  // + the ordering CTA
        const size_t smem = order->P <= 4096 ? (size_t)std::max(order->P, 1024) * 8 : 0;
        k_segments_order<<<(unsigned)blocks, 1024, smem, st>>>(g, *order);
    } else {
        const int64_t blocks = (n_qt * 32 + 255) / 256;
        k_segments<<<(unsigned)blocks, 256, 0, st>>>(g);
    }
    ++g_launches;
    CU(cudaGetLastError());
    return BM25_OK;
}

int launch_merge(const MergeArgs& m, int device, size_t smem_optin, cudaStream_t st) {
    const size_t smem = (size_t)m.P * 8 + (size_t)m.k_out + 16;
    static thread_local size_t configured[64] = {0};
    if (smem > 48 * 1024 && configured[device % 64] < smem) {
        cudaFuncAttributes fa;
        CU(cudaFuncGetAttributes(&fa, k_merge));
        const size_t max_dyn = smem_optin - fa.sharedSizeBytes;
        if (smem > max_dyn) return fail(BM25_ERR_UNSUPPORTED, "merge needs %zu B of shared memory", smem);
        CU(cudaFuncSetAttribute(k_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn));
        configured[device % 64] = max_dyn;
    }
    // CTA size follows the sort size: a 32-key merge (k = 10, two tile ranges) needs no 512 threads
    int threads = kThreads;
    while (threads > 64 && threads >= m.P) threads >>= 1;
    k_merge<<<(unsigned)m.Q, threads, smem, st>>>(m);
    ++g_launches;
    CU(cudaGetLastError());
    return BM25_OK;
}

// k_out above BM25_SMALL_K: global-memory merge.  `scratch` / `present` are stream-ordered allocations
// of this call (cudaMallocAsync) -- a large-k merge is the rare call, not the steady path.
int launch_merge_large(MergeArgs m, cudaStream_t st) {
    const int64_t total = (int64_t)m.n_lists * m.k_in;
    m.P = next_pow2(m.k_out);
    m.n_pad = (total + 1) & ~(int64_t)1;
    u64* scratch = nullptr;
    unsigned char* present = nullptr;
    const size_t bytes = (size_t)m.Q * (size_t)(m.n_pad + m.P) * 8;
    if (cudaMallocAsync(&scratch, bytes, st) != cudaSuccess ||
        cudaMallocAsync(&present, (size_t)m.Q * m.k_out, st) != cudaSuccess) {
        cudaGetLastError();
        if (scratch) cudaFreeAsync(scratch, st);
        return fail(BM25_ERR_OOM, "cudaMallocAsync of %zu bytes of merge scratch failed", bytes);
    }
    m.scratch = scratch;
    m.present_g = present;
    k_merge_large<<<(unsigned)m.Q, kThreads, 0, st>>>(m);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    cudaFreeAsync(present, st);
    if (e != cudaSuccess) return fail(BM25_ERR_CUDA, "k_merge_large launch failed: %s", cudaGetErrorString(e));
    return BM25_OK;
}

int merge_P(int64_t total, int k_out) {
    if (total <= 8192) return std::max(next_pow2(total), next_pow2(k_out));
    return std::max(8192, next_pow2(2 * (int64_t)k_out));
}

int search_locked(bm25_index* ix, const int32_t* d_queries, int64_t Q, int64_t T, int k, int32_t* d_out_ids,
                  float* d_out_scores, cudaStream_t st) {
    LaunchPlan lp;
    int rc = make_plan(ix, Q, T, k, &lp);
    if (rc) return rc;
    if ((rc = ix->ws_partial.reserve((size_t)Q * lp.splits * k))) return rc;
    if ((rc = ix->ws_theta.reserve((size_t)Q))) return rc;
    if (lp.cand_global && (rc = ix->ws_cand.reserve((size_t)Q * lp.splits * lp.cap))) return rc;
    if ((rc = ix->ws_seg.reserve((size_t)Q * T * (lp.seg_rows + 1)))) return rc;
    if ((rc = ensure_table(ix, lp.tile_docs))) return rc;
    if ((rc = ensure_packed(ix, lp.tile_docs))) return rc;
    // the workspace is per handle: a search on another stream waits for the previous one
    if (ix->ws_used && ix->ws_stream != st) CU(cudaStreamWaitEvent(st, ix->ws_done, 0));
    if (ix->opt_poison) {  // debug: uninitialised workspace reads must not pass by luck
        CU(cudaMemsetAsync(ix->ws_seg.p, 0xff, ix->ws_seg.bytes(), st));
        CU(cudaMemsetAsync(ix->ws_partial.p, 0xff, ix->ws_partial.bytes(), st));
        if (ix->ws_cand.p) CU(cudaMemsetAsync(ix->ws_cand.p, 0xff, ix->ws_cand.bytes(), st));
    }
    CU(cudaMemsetAsync(ix->ws_theta.p, 0, (size_t)Q * sizeof(u64), st));
    const bool timing = ix->opt_timing != 0;
    if (timing) {
        for (auto& e : ix->ev)
            if (!e) CU(cudaEventCreate(&e));
        ix->ev_valid = false;
        CU(cudaEventRecord(ix->ev[0], st));
    }
    // cross-query L2 sharing: CTAs of queries with the same heaviest term become neighbours; the
    // ordering runs in the same launch as the segment table
    const bool qsort = !ix->opt_no_query_sort && Q > 1 && Q <= (1 << 20);
    OrderArgs o{};
    if (qsort) {
        const int P = next_pow2(Q);
        if ((rc = ix->ws_qperm.reserve((size_t)Q))) return rc;
        if (P > 4096 && (rc = ix->ws_qkey.reserve((size_t)P))) return rc;
        o.tptr = ix->d_tptr;
        o.queries = d_queries;
        o.Q = (int)Q;
        o.T = (int)T;
        o.n_terms = (int)ix->n_terms;
        o.P = P;
        o.keys_g = ix->ws_qkey.p;
        o.perm = ix->ws_qperm.p;
    }
    if ((rc = launch_segments(ix, lp, d_queries, Q, T, k, ix->opt_no_theta_share ? nullptr : ix->ws_theta.p, false, st,
                              qsort ? &o : nullptr)))
        return rc;
    if (timing) CU(cudaEventRecord(ix->ev[1], st));
    SearchArgs a{};
    a.ids = ix->d_ids;
    a.w = ix->d_w;
    a.term_row = ix->tt.d_term_row;
    a.tab = ix->tt.d_tab;
    a.queries = d_queries;
    a.qperm = qsort ? ix->ws_qperm.p : nullptr;
    a.seg = ix->ws_seg.p;
    a.partial = ix->ws_partial.p;
    a.theta_q = ix->opt_no_theta_share ? nullptr : ix->ws_theta.p;
    a.cand_global = lp.cand_global ? ix->ws_cand.p : nullptr;
    a.dense_out = nullptr;
    a.theta0 = lp.theta0;
    a.Q = (int)Q;
    a.T = (int)T;
    a.k = k;
    a.n_terms = (int)ix->n_terms;
    a.n_docs = (int)ix->n_docs;
    a.tile_docs = lp.tile_docs;
    a.n_tiles = lp.n_tiles;
    a.tiles_per_chunk = lp.tiles_per_chunk;
    a.n_chunks = lp.n_chunks;
    a.splits = lp.splits;
    a.tiles_per_split = 0;
    a.cap = lp.cap;
    a.general = lp.general;
    a.no_hot = ix->opt_no_hot;
    a.poison = ix->opt_poison;
    a.sp_major = ix->opt_q_major ? 0 : 1;
    a.bulk_clear = ix->opt_no_bulk_clear ? 0 : 1;
    a.wrange = ix->d_wrange;
    a.no_epoch = ix->opt_no_epoch;
    a.pk = (ix->weight_format != BM25_WEIGHTS_FP32 && !ix->opt_no_packed) ? ix->tt.d_pk : nullptr;
    if ((rc = launch_score(ix, lp, a, Q, false, st))) return rc;
    if (timing) CU(cudaEventRecord(ix->ev[2], st));
    MergeArgs m{};
    m.keys = ix->ws_partial.p;
    m.out_ids = d_out_ids;
    m.out_scores = d_out_scores;
    m.Q = Q;
    m.n_lists = lp.splits;
    m.k_in = k;
    m.k_out = k;
    m.P = merge_P((int64_t)lp.splits * k, k);
    m.id_offset = ix->doc_id_base;
    m.fill = 1;
    if ((rc = (k > BM25_SMALL_K ? launch_merge_large(m, st) : launch_merge(m, ix->device, ix->smem_optin, st)))) return rc;
    if (timing) {
        CU(cudaEventRecord(ix->ev[3], st));
        ix->ev_valid = true;
    }
    CU(cudaEventRecord(ix->ws_done, st));
    ix->ws_stream = st;
    ix->ws_used = true;
    return BM25_OK;
}

int check_search_args(const bm25_index* ix, int64_t Q, int64_t T, int k) {
    if (!ix) return fail(BM25_ERR_INVALID, "index handle is NULL");
    if (Q < 0 || T < 1) return fail(BM25_ERR_INVALID, "bad query shape [%lld, %lld]", (long long)Q, (long long)T);
    if (Q * T > 0x7fffffffLL) return fail(BM25_ERR_UNSUPPORTED, "query batch too large");
    if (k < 1) return fail(BM25_ERR_INVALID, "k must be >= 1 (got %d)", k);
    if (k > ix->n_docs)
        return fail(BM25_ERR_INVALID, "kth(=-%d) out of bounds (%lld)", k, (long long)ix->n_docs);
    if (k > BM25_MAX_K)
        return fail(BM25_ERR_INVALID, "k=%d exceeds BM25_MAX_K=%d, the largest top-k this build selects on the device", k,
                    BM25_MAX_K);
    return BM25_OK;
}

}  // namespace

extern "C" {

const char* bm25_last_error(void) { return g_err.c_str(); }
const char* bm25_version(void) { return "bm25_b200 0.1.0 (sm_100a)"; }
int64_t bm25_kernel_launches(void) { return g_launches.load(); }

int bm25_index_create(const int32_t* h_indptr, const int32_t* h_indices, const float* h_data, int64_t n_terms,
                      int64_t n_docs, int64_t nnz, int device, int64_t doc_id_base, bm25_index** out) {
    if (!out) return fail(BM25_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!h_indptr || (nnz > 0 && (!h_indices || !h_data))) return fail(BM25_ERR_INVALID, "NULL index array");
    if (n_terms < 0 || n_docs < 0 || nnz < 0 || nnz > 0x7fffffffLL || n_docs > BM25_MAX_DOCS)
        return fail(BM25_ERR_INVALID, "bad index shape (terms=%lld docs=%lld nnz=%lld)", (long long)n_terms,
                    (long long)n_docs, (long long)nnz);
    if (n_docs + doc_id_base > 0x7fffffffLL)
        return fail(BM25_ERR_INVALID, "doc_id_base + n_docs exceeds int32");
    cudaDeviceProp prop;
    int rc = check_device(device, &prop);
    if (rc) return rc;
    std::vector<int32_t> c_ptr, c_idx;
    std::vector<float> c_dat;
    bool sorted = true, positive = true;
    rc = canonicalise_host(h_indptr, h_indices, h_data, n_terms, n_docs, nnz, c_ptr, c_idx, c_dat, &sorted,
                           &positive);
    if (rc) return rc;
    if (!sorted) {
        h_indptr = c_ptr.data();
        h_indices = c_idx.data();
        h_data = c_dat.data();
        nnz = (int64_t)c_idx.size();
    }
    DeviceGuard g(device);
    if (!g.ok) return fail(BM25_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    bm25_index* ix = new (std::nothrow) bm25_index();
    if (!ix) return fail(BM25_ERR_OOM, "out of host memory");
    ix->device = device;
    ix->n_terms = n_terms;
    ix->n_docs = n_docs;
    ix->nnz = nnz;
    ix->doc_id_base = doc_id_base;
    ix->all_positive = positive;
    ix->was_sorted = sorted;
    ix->h_indptr.assign(h_indptr, h_indptr + n_terms + 1);
    auto cleanup = [&](int code) {
        bm25_index_destroy(ix);
        return code;
    };
    // raw CSC arrays go to the device only as the source of the re-bucketing copy
    int32_t *raw_ptr = nullptr, *raw_ids = nullptr;
    float* raw_w = nullptr;
    auto free_raw = [&]() {
        if (raw_ptr) cudaFree(raw_ptr);
        if (raw_ids) cudaFree(raw_ids);
        if (raw_w) cudaFree(raw_w);
    };
    if (cudaMalloc(&raw_ptr, (size_t)(n_terms + 1) * 4) != cudaSuccess ||
        cudaMalloc(&raw_ids, (size_t)std::max<int64_t>(nnz, 1) * 4) != cudaSuccess ||
        cudaMalloc(&raw_w, (size_t)std::max<int64_t>(nnz, 1) * 4) != cudaSuccess) {
        cudaGetLastError();
        free_raw();
        return cleanup(fail(BM25_ERR_OOM, "cudaMalloc of the index (%lld postings) failed", (long long)nnz));
    }
    if (cudaMemcpy(raw_ptr, h_indptr, (size_t)(n_terms + 1) * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        (nnz > 0 && (cudaMemcpy(raw_ids, h_indices, (size_t)nnz * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
                     cudaMemcpy(raw_w, h_data, (size_t)nnz * 4, cudaMemcpyHostToDevice) != cudaSuccess))) {
        cudaError_t e = cudaGetLastError();
        free_raw();
        return cleanup(fail(BM25_ERR_CUDA, "copying the index to device %d failed: %s", device,
                            cudaGetErrorString(e)));
    }
    rc = finish_create(ix, prop, raw_ptr, raw_ids, raw_w);
    free_raw();
    if (rc) return cleanup(rc);
    *out = ix;
    return BM25_OK;
}

int bm25_index_create_device(const int32_t* d_indptr, const int32_t* d_indices, const float* d_data,
                             int64_t n_terms, int64_t n_docs, int64_t nnz, int device, int64_t doc_id_base,
                             int borrow, bm25_index** out) {
    if (!out) return fail(BM25_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!d_indptr || (nnz > 0 && (!d_indices || !d_data))) return fail(BM25_ERR_INVALID, "NULL index array");
    if (n_terms < 0 || n_docs < 0 || nnz < 0 || nnz > 0x7fffffffLL || n_docs > BM25_MAX_DOCS)
        return fail(BM25_ERR_INVALID, "bad index shape (terms=%lld docs=%lld nnz=%lld)", (long long)n_terms,
                    (long long)n_docs, (long long)nnz);
    if (n_docs + doc_id_base > 0x7fffffffLL)
        return fail(BM25_ERR_INVALID, "doc_id_base + n_docs exceeds int32");
    cudaDeviceProp prop;
    int rc = check_device(device, &prop);
    if (rc) return rc;
    DeviceGuard g(device);
    if (!g.ok) return fail(BM25_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    // validate on the device
    unsigned long long* d_flags = nullptr;
    unsigned long long h_flags[8] = {0};
    CU(cudaMalloc(&d_flags, sizeof h_flags));
    cudaMemset(d_flags, 0, sizeof h_flags);
    if (nnz > 0) {
        k_validate_postings<<<prop.multiProcessorCount * 8, 256>>>(d_indices, d_data, nnz, n_docs, d_flags);
        ++g_launches;
    }
    if (n_terms > 0) {
        k_validate_indptr<<<prop.multiProcessorCount * 4, 256>>>(d_indptr, d_indices, n_terms, nnz, d_flags);
        ++g_launches;
    }
    cudaError_t e = cudaMemcpy(h_flags, d_flags, sizeof h_flags, cudaMemcpyDeviceToHost);
    cudaFree(d_flags);
    if (e != cudaSuccess) return fail(BM25_ERR_CUDA, "index validation failed: %s", cudaGetErrorString(e));
    if (h_flags[5]) return fail(BM25_ERR_INVALID, "indptr is malformed (%llu defects)", h_flags[5]);
    if (h_flags[0]) return fail(BM25_ERR_INVALID, "%llu doc ids are outside [0, %lld)", h_flags[0], (long long)n_docs);
    if (h_flags[1]) return fail(BM25_ERR_INVALID, "%llu weights are not finite", h_flags[1]);
    if (h_flags[3] != h_flags[4])
        return fail(BM25_ERR_INVALID,
                    "device-resident columns must be strictly increasing in doc id (%llu inversions); "
                    "use bm25_index_create for non-canonical input",
                    h_flags[3] - h_flags[4]);
    bm25_index* ix = new (std::nothrow) bm25_index();
    if (!ix) return fail(BM25_ERR_OOM, "out of host memory");
    ix->device = device;
    ix->n_terms = n_terms;
    ix->n_docs = n_docs;
    ix->nnz = nnz;
    ix->doc_id_base = doc_id_base;
    ix->all_positive = (h_flags[2] == 0);
    ix->was_sorted = true;
    ix->h_indptr.resize(n_terms + 1);
    auto cleanup = [&](int code) {
        bm25_index_destroy(ix);
        return code;
    };
    if (cudaMemcpy(ix->h_indptr.data(), d_indptr, (n_terms + 1) * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return cleanup(fail(BM25_ERR_CUDA, "reading indptr back failed"));
    // `borrow` is accepted for source compatibility: the index is always re-bucketed into
    // library-owned memory, the caller's arrays are only read during this call
    (void)borrow;
    if ((rc = finish_create(ix, prop, d_indptr, d_indices, d_data))) return cleanup(rc);
    *out = ix;
    return BM25_OK;
}

int bm25_index_compress(bm25_index* ix, int weight_format) {
    if (!ix) return fail(BM25_ERR_INVALID, "index handle is NULL");
    if (weight_format != BM25_WEIGHTS_BF16)
        return fail(BM25_ERR_INVALID, "weight_format must be BM25_WEIGHTS_BF16 (%d), got %d", BM25_WEIGHTS_BF16, weight_format);
    DeviceGuard g(ix->device);
    if (!g.ok) return fail(BM25_ERR_CUDA, "cudaSetDevice(%d) failed", ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    if (ix->weight_format == weight_format) return BM25_OK;
    CU(cudaDeviceSynchronize());  // no search may still be reading the weights
    if (ix->n_terms > 0 && ix->nnz > 0) {
        // nothing is modified unless every weight stays finite in bf16 (largest finite bf16: 0x7f7f0000)
        std::vector<float2> wr((size_t)ix->n_terms);
        CU(cudaMemcpy(wr.data(), ix->d_wrange, wr.size() * sizeof(float2), cudaMemcpyDeviceToHost));
        const float lim = 3.3895313892515355e38f;
        for (int64_t t = 0; t < ix->n_terms; ++t)
            if (wr[t].y >= wr[t].x && (wr[t].y > lim || wr[t].x < -lim))
                return fail(BM25_ERR_INVALID, "term %lld holds a weight that overflows bf16; the handle is unchanged", (long long)t);
        unsigned long long* d_flags = nullptr;
        unsigned long long h_flags[2] = {0, 0};
        CU(cudaMalloc(&d_flags, sizeof h_flags));
        cudaMemset(d_flags, 0, sizeof h_flags);
        const int grid = (int)std::min<int64_t>(ix->n_terms, (int64_t)ix->sm_count * 16);
        k_quantize<<<grid, 128>>>(ix->d_tptr, (int)ix->n_terms, ix->d_w, ix->d_wrange, d_flags);
        ++g_launches;
        cudaError_t e = cudaMemcpy(h_flags, d_flags, sizeof h_flags, cudaMemcpyDeviceToHost);
        cudaFree(d_flags);
        if (e != cudaSuccess) return fail(BM25_ERR_CUDA, "k_quantize failed: %s", cudaGetErrorString(e));
        ix->weight_format = weight_format;  // the weights are rounded from here on, whatever follows
        if (h_flags[1]) return fail(BM25_ERR_INVALID, "%llu weights overflow bf16", h_flags[1]);
        if (h_flags[0]) ix->all_positive = false;  // a weight rounded to zero: every document competes
    }
    ix->weight_format = weight_format;
    ix->tt.pk_tile_docs = 0;  // packed from the rounded weights on the next search
    ix->tt_stash.pk_tile_docs = 0;
    int rc = compute_bounds(ix);
    if (rc) return rc;
    CU(cudaDeviceSynchronize());
    return BM25_OK;
}

int bm25_index_destroy(bm25_index* ix) {
    if (!ix) return BM25_OK;
    {
        DeviceGuard g(ix->device);
        if (ix->d_tptr) cudaFree(ix->d_tptr);
        if (ix->d_ids) cudaFree(ix->d_ids);
        if (ix->d_wrange) cudaFree(ix->d_wrange);
        if (ix->d_w) cudaFree(ix->d_w);
        ix->tt.release();
        ix->tt_stash.release();
        if (ix->d_bounds) cudaFree(ix->d_bounds);
        if (ix->ws_done) cudaEventDestroy(ix->ws_done);
        ix->ws_seg.release();
        ix->ws_partial.release();
        ix->ws_theta.release();
        ix->ws_cand.release();
        ix->ws_qkey.release();
        ix->ws_qperm.release();
        ix->ws_queries.release();
        ix->ws_out_ids.release();
        ix->ws_out_scores.release();
        ix->pin_queries.release();
        ix->pin_out_ids.release();
        ix->pin_out_scores.release();
        if (ix->own_stream) cudaStreamDestroy(ix->own_stream);
        for (auto& e : ix->ev)
            if (e) cudaEventDestroy(e);
    }
    delete ix;
    return BM25_OK;
}

int bm25_index_get_info(const bm25_index* ix, bm25_index_info* out) {
    if (!ix || !out) return fail(BM25_ERR_INVALID, "NULL argument");
    out->n_terms = ix->n_terms;
    out->n_docs = ix->n_docs;
    out->nnz = ix->nnz;
    out->doc_id_base = ix->doc_id_base;
    out->device_bytes = ix->device_bytes();
    out->device = ix->device;
    out->tile_docs = ix->tile_docs();
    out->n_tiles = ix->n_tiles();
    out->all_positive = ix->all_positive ? 1 : 0;
    out->was_sorted = ix->was_sorted ? 1 : 0;
    out->sm_count = ix->sm_count;
    out->weight_format = ix->weight_format;
    out->posting_bytes = ix->weight_format == BM25_WEIGHTS_FP32 ? 8 : 4;
    return BM25_OK;
}

int bm25_index_set_option(bm25_index* ix, const char* name, int64_t value) {
    if (!ix || !name) return fail(BM25_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lock(ix->mu);
    if (!strcmp(name, "tile_docs")) {
        if (value < 0 || value > (1 << 16)) return fail(BM25_ERR_INVALID, "tile_docs out of range");
        ix->opt_tile_docs = (int)value;
    } else if (!strcmp(name, "splits")) {
        if (value < 0 || value > 65536) return fail(BM25_ERR_INVALID, "splits out of range");
        ix->opt_splits = (int)value;
    } else if (!strcmp(name, "consumer_warps")) {
        if (value < 0 || value > 16) return fail(BM25_ERR_INVALID, "consumer_warps must be 0..16");
        ix->opt_warps = (int)value;
    } else if (!strcmp(name, "cap")) {
        if (value < 0 || value > (1 << 14)) return fail(BM25_ERR_INVALID, "cap out of range");
        ix->opt_cap = (int)value;
    } else if (!strcmp(name, "waves")) {
        if (value < 0 || value > 64) return fail(BM25_ERR_INVALID, "waves out of range");
        ix->opt_waves = (int)value;
    } else if (!strcmp(name, "no_theta_share")) {
        ix->opt_no_theta_share = value ? 1 : 0;
    } else if (!strcmp(name, "no_priming")) {
        ix->opt_no_priming = value ? 1 : 0;
    } else if (!strcmp(name, "heavy_min")) {
        if (value < 0 || value > (1 << 28)) return fail(BM25_ERR_INVALID, "heavy_min out of range");
        ix->opt_heavy_min = (int)value;
    } else if (!strcmp(name, "no_epoch")) {
        ix->opt_no_epoch = value ? 1 : 0;
    } else if (!strcmp(name, "no_packed")) {
        ix->opt_no_packed = value ? 1 : 0;
    } else if (!strcmp(name, "q_major")) {
        ix->opt_q_major = value ? 1 : 0;
    } else if (!strcmp(name, "generic_kernel")) {
        ix->opt_generic_kernel = value ? 1 : 0;
    } else if (!strcmp(name, "no_query_sort")) {
        ix->opt_no_query_sort = value ? 1 : 0;
    } else if (!strcmp(name, "no_bulk_clear")) {
        ix->opt_no_bulk_clear = value ? 1 : 0;
    } else if (!strcmp(name, "poison")) {
        ix->opt_poison = value ? 1 : 0;
    } else if (!strcmp(name, "cand_smem")) {
        ix->opt_cand_smem = value ? 1 : 0;
    } else if (!strcmp(name, "no_hot")) {
        ix->opt_no_hot = value ? 1 : 0;
    } else if (!strcmp(name, "force_general")) {
        ix->opt_force_general = value ? 1 : 0;
    } else if (!strcmp(name, "timing")) {
        ix->opt_timing = value ? 1 : 0;
    } else {
        return fail(BM25_ERR_INVALID, "unknown option '%s'", name);
    }
    return BM25_OK;
}

int bm25_index_get_timing(bm25_index* ix, float* out_ms3) {
    if (!ix || !out_ms3) return fail(BM25_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lock(ix->mu);
    if (!ix->ev_valid) return fail(BM25_ERR_INVALID, "no timed search recorded (set option \"timing\" to 1 first)");
    DeviceGuard g(ix->device);
    CU(cudaEventSynchronize(ix->ev[3]));
    for (int i = 0; i < 3; ++i) CU(cudaEventElapsedTime(out_ms3 + i, ix->ev[i], ix->ev[i + 1]));
    return BM25_OK;
}

int bm25_search(bm25_index* ix, const int32_t* d_queries, int64_t Q, int64_t T, int k, int32_t* d_out_ids,
                float* d_out_scores, void* cuda_stream) {
    int rc = check_search_args(ix, Q, T, k);
    if (rc) return rc;
    if (Q == 0) return BM25_OK;
    if (!d_queries || !d_out_ids || !d_out_scores) return fail(BM25_ERR_INVALID, "NULL buffer");
    DeviceGuard g(ix->device);
    if (!g.ok) return fail(BM25_ERR_CUDA, "cudaSetDevice(%d) failed", ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    return search_locked(ix, d_queries, Q, T, k, d_out_ids, d_out_scores, (cudaStream_t)cuda_stream);
}

int bm25_search_host(bm25_index* ix, const int32_t* h_queries, int64_t Q, int64_t T, int k, int32_t* h_out_ids,
                     float* h_out_scores) {
    int rc = check_search_args(ix, Q, T, k);
    if (rc) return rc;
    if (Q == 0) return BM25_OK;
    if (!h_queries || !h_out_ids || !h_out_scores) return fail(BM25_ERR_INVALID, "NULL buffer");
    for (int64_t i = 0; i < Q * T; ++i)
        if (h_queries[i] >= ix->n_terms)
            return fail(BM25_ERR_INVALID,
                        "The maximum token ID in the query (%d) is higher than the number of tokens in the index.",
                        h_queries[i]);
    DeviceGuard g(ix->device);
    if (!g.ok) return fail(BM25_ERR_CUDA, "cudaSetDevice(%d) failed", ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    const size_t nq = (size_t)Q * T, no = (size_t)Q * k;
    // ids and scores share one device buffer and one pinned buffer: [no] int32 ids | [no] fp32 scores,
    // so the results come back in ONE device-to-host copy
    if ((rc = ix->pin_queries.reserve(nq)) || (rc = ix->pin_out_ids.reserve(2 * no)) ||
        (rc = ix->ws_queries.reserve(nq)) || (rc = ix->ws_out_ids.reserve(2 * no)))
        return rc;
    cudaStream_t st = ix->own_stream;
    memcpy(ix->pin_queries.p, h_queries, nq * 4);
    CU(cudaMemcpyAsync(ix->ws_queries.p, ix->pin_queries.p, nq * 4, cudaMemcpyHostToDevice, st));
    float* d_scores = reinterpret_cast<float*>(ix->ws_out_ids.p + no);
    if ((rc = search_locked(ix, ix->ws_queries.p, Q, T, k, ix->ws_out_ids.p, d_scores, st))) return rc;
    // caller buffers in page-locked memory (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory)
    // receive the results by DMA directly; pageable buffers go through the handle's pinned staging area
    cudaPointerAttributes pa_i, pa_s;
    const bool direct = cudaPointerGetAttributes(&pa_i, h_out_ids) == cudaSuccess && pa_i.type == cudaMemoryTypeHost &&
                        cudaPointerGetAttributes(&pa_s, h_out_scores) == cudaSuccess && pa_s.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (direct) {
        if (reinterpret_cast<const int32_t*>(h_out_scores) == h_out_ids + no) {  // one contiguous [2][Q][k] block
            CU(cudaMemcpyAsync(h_out_ids, ix->ws_out_ids.p, 2 * no * 4, cudaMemcpyDeviceToHost, st));
        } else {
            CU(cudaMemcpyAsync(h_out_ids, ix->ws_out_ids.p, no * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(h_out_scores, d_scores, no * 4, cudaMemcpyDeviceToHost, st));
        }
        CU(cudaStreamSynchronize(st));
        return BM25_OK;
    }
    CU(cudaMemcpyAsync(ix->pin_out_ids.p, ix->ws_out_ids.p, 2 * no * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(h_out_ids, ix->pin_out_ids.p, no * 4);
    memcpy(h_out_scores, ix->pin_out_ids.p + no, no * 4);
    return BM25_OK;
}

int bm25_scores_dense(bm25_index* ix, const int32_t* d_queries, int64_t Q, int64_t T, float* d_out,
                      void* cuda_stream) {
    if (!ix) return fail(BM25_ERR_INVALID, "index handle is NULL");
    if (Q < 0 || T < 1) return fail(BM25_ERR_INVALID, "bad query shape");
    if (Q == 0 || ix->n_docs == 0) return BM25_OK;
    if (!d_queries || !d_out) return fail(BM25_ERR_INVALID, "NULL buffer");
    DeviceGuard g(ix->device);
    if (!g.ok) return fail(BM25_ERR_CUDA, "cudaSetDevice(%d) failed", ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    LaunchPlan lp;
    int rc = make_plan_dense(ix, Q, T, &lp);
    if (rc) return rc;
    if (ix->ws_used && ix->ws_stream != st) CU(cudaStreamWaitEvent(st, ix->ws_done, 0));
    if ((rc = launch_segments(ix, lp, d_queries, Q, T, 1, nullptr, true, st))) return rc;
    SearchArgs a{};
    a.ids = ix->d_ids;
    a.w = ix->d_w;
    a.queries = d_queries;
    a.seg = ix->ws_seg.p;
    a.dense_out = d_out;
    a.Q = (int)Q;
    a.T = (int)T;
    a.k = 1;
    a.n_docs = (int)ix->n_docs;
    a.tile_docs = lp.tile_docs;
    a.n_tiles = lp.n_tiles;
    a.splits = lp.splits;
    a.tiles_per_split = lp.tiles_per_split;
    a.cap = 0;
    if ((rc = launch_score(ix, lp, a, Q, true, st))) return rc;
    CU(cudaEventRecord(ix->ws_done, st));
    ix->ws_stream = st;
    ix->ws_used = true;
    return BM25_OK;
}

int bm25_scores_dense_host(bm25_index* ix, const int32_t* h_queries, int64_t Q, int64_t T, float* h_out) {
    if (!ix) return fail(BM25_ERR_INVALID, "index handle is NULL");
    if (Q < 0 || T < 1) return fail(BM25_ERR_INVALID, "bad query shape");
    if (Q == 0 || ix->n_docs == 0) return BM25_OK;
    if (!h_queries || !h_out) return fail(BM25_ERR_INVALID, "NULL buffer");
    for (int64_t i = 0; i < Q * T; ++i)
        if (h_queries[i] >= ix->n_terms) return fail(BM25_ERR_INVALID, "token id %d out of range", h_queries[i]);
    DeviceGuard g(ix->device);
    if (!g.ok) return fail(BM25_ERR_CUDA, "cudaSetDevice(%d) failed", ix->device);
    int32_t* d_q = nullptr;
    float* d_o = nullptr;
    CU(cudaMalloc(&d_q, Q * T * 4));
    if (cudaMalloc(&d_o, (size_t)Q * ix->n_docs * 4) != cudaSuccess) {
        cudaFree(d_q);
        cudaGetLastError();
        return fail(BM25_ERR_OOM, "cudaMalloc of the dense score slab failed");
    }
    int rc = BM25_OK;
    if (cudaMemcpy(d_q, h_queries, Q * T * 4, cudaMemcpyHostToDevice) != cudaSuccess)
        rc = fail(BM25_ERR_CUDA, "H2D of queries failed");
    if (!rc) rc = bm25_scores_dense(ix, d_q, Q, T, d_o, ix->own_stream);
    if (!rc && cudaStreamSynchronize(ix->own_stream) != cudaSuccess)
        rc = fail(BM25_ERR_CUDA, "dense scoring failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (!rc && cudaMemcpy(h_out, d_o, (size_t)Q * ix->n_docs * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = fail(BM25_ERR_CUDA, "D2H of dense scores failed");
    cudaFree(d_q);
    cudaFree(d_o);
    return rc;
}

int bm25_merge_topk(const int32_t* d_ids, const float* d_scores, int n_lists, int64_t list_stride, int64_t Q,
                    int k_in, int k_out, int32_t* d_out_ids, float* d_out_scores, int device, void* cuda_stream) {
    if (list_stride == 0) list_stride = Q * k_in;
    if (list_stride < Q * k_in) return fail(BM25_ERR_INVALID, "list_stride smaller than one list");
    if (n_lists < 1 || k_in < 1 || k_out < 1 || Q < 0) return fail(BM25_ERR_INVALID, "bad merge shape");
    if ((int64_t)n_lists * k_in < k_out)
        return fail(BM25_ERR_INVALID, "merge needs n_lists*k_in >= k_out (%d*%d < %d)", n_lists, k_in, k_out);
    if (k_out > BM25_MAX_K) return fail(BM25_ERR_UNSUPPORTED, "k_out=%d exceeds BM25_MAX_K", k_out);
    if (Q == 0) return BM25_OK;
    if (!d_ids || !d_scores || !d_out_ids || !d_out_scores) return fail(BM25_ERR_INVALID, "NULL buffer");
    cudaDeviceProp prop;
    int rc = check_device(device, &prop);
    if (rc) return rc;
    DeviceGuard g(device);
    if (!g.ok) return fail(BM25_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    MergeArgs m{};
    m.keys = nullptr;
    m.in_ids = d_ids;
    m.in_scores = d_scores;
    m.out_ids = d_out_ids;
    m.out_scores = d_out_scores;
    m.Q = Q;
    m.list_stride = list_stride;
    m.n_lists = n_lists;
    m.k_in = k_in;
    m.k_out = k_out;
    m.P = merge_P((int64_t)n_lists * k_in, k_out);
    m.id_offset = 0;
    m.fill = 0;
    if (k_out > BM25_SMALL_K) return launch_merge_large(m, (cudaStream_t)cuda_stream);
    return launch_merge(m, device, prop.sharedMemPerBlockOptin, (cudaStream_t)cuda_stream);
}

int bm25_posting_bytes(const bm25_index* ix, const int32_t* h_queries, int64_t Q, int64_t T, int k,
                       int64_t* out_bytes) {
    if (!ix || !h_queries || !out_bytes) return fail(BM25_ERR_INVALID, "NULL argument");
    int64_t postings = 0;
    for (int64_t i = 0; i < Q * T; ++i) {
        const int32_t t = h_queries[i];
        if (t < 0) continue;
        if (t >= ix->n_terms) return fail(BM25_ERR_INVALID, "token id %d out of range", t);
        postings += ix->h_indptr[t + 1] - ix->h_indptr[t];
    }
    *out_bytes = (ix->weight_format == BM25_WEIGHTS_FP32 ? 8 : 4) * postings + 8 * (int64_t)k * Q;
    return BM25_OK;
}

}  // extern "C"
