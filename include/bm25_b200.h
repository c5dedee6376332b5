/* bm25_b200.h -- C ABI of libbm25_b200.so: the B200-native (sm_100a) BM25 query hot path.
 *
 * Path: CSC posting gather -> per-document score accumulation -> top-k
 * (reference hot loop: bm25_native.py:129-158 `_compute_relevance_from_scores`, and the MAX graph
 * gather/sum/top_k of gpu_bm25/common.py:64-78 that it replaces).
 *
 * Conventions (mirroring the reference's custom-op ABI, operations/graph_operation.mojo:27-45:
 * caller-allocated outputs, stream-ordered, errors surfaced to the host as exceptions):
 *   - plain C, no C++/torch types; every entry point returns an int status (0 = BM25_OK) and never
 *     throws; the message of the last failure on the calling thread is bm25_last_error().
 *   - "d_" pointers are device pointers on the index's device, "h_" pointers are host pointers.
 *   - outputs are allocated by the caller; the library never frees caller memory.
 *   - device entry points only enqueue work on `cuda_stream` (a cudaStream_t, NULL = legacy
 *     default stream) and return; *_host entry points synchronise before returning.
 *   - there is NO CPU fallback: without a CUDA device every entry point fails with
 *     BM25_ERR_NO_DEVICE.
 *   - threading: a handle has ONE workspace.  Calls may come from any thread and use any stream:
 *     enqueueing is serialised by a host mutex and a search that runs on a different stream than
 *     the previous search of the same handle first waits (on the device) for that search, so the
 *     searches of one handle never overlap.  Use one handle per stream for concurrent execution.
 *
 * Result order: score descending, ties by ascending document id.  Documents with no matching
 * posting have score +0.0 and fill the tail when fewer than k documents match (the reference
 * returns k ids in that case too, bm25_native.py:147-158).
 */
#ifndef BM25_B200_H
#define BM25_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BM25_OK 0
#define BM25_ERR_INVALID 1     /* bad argument / malformed index or query            */
#define BM25_ERR_CUDA 2        /* a CUDA runtime call or kernel launch failed        */
#define BM25_ERR_NO_DEVICE 3   /* no usable CUDA device (no CPU fallback exists)     */
#define BM25_ERR_UNSUPPORTED 4 /* valid request outside this build's limits (e.g. k) */
#define BM25_ERR_OOM 5         /* host or device allocation failed                   */

#define BM25_MAX_K 65536 /* largest supported top-k */
#define BM25_SMALL_K 6144 /* up to here the final merge sorts in shared memory; above, in global memory */
#define BM25_MAX_DOCS (0x7fffffffLL - 65536) /* largest n_docs of one handle (int32 tile arithmetic) */

#define BM25_WEIGHTS_FP32 0 /* int32 doc id + fp32 weight, 8 bytes per posting (the bm25s on-disk dtypes) */
#define BM25_WEIGHTS_BF16 1 /* compressed: uint16 tile-local doc id + bf16 weight, 4 bytes per posting    */

typedef struct bm25_index bm25_index; /* opaque; owns the HBM-resident CSC arrays + workspace */

typedef struct bm25_index_info {
    int64_t n_terms;      /* V: number of CSC columns                          */
    int64_t n_docs;       /* N: number of CSC rows (documents)                 */
    int64_t nnz;          /* postings                                          */
    int64_t doc_id_base;  /* added to every returned doc id (document shards)  */
    int64_t device_bytes; /* bytes of HBM owned by the handle (index+workspace)*/
    int32_t device;       /* CUDA device ordinal                               */
    int32_t tile_docs;    /* documents per shared-memory score tile            */
    int32_t n_tiles;      /* ceil(N / tile_docs)                               */
    int32_t all_positive; /* 1 if every weight is > 0 (enables the pruned path)*/
    int32_t was_sorted;   /* 1 if every column arrived sorted by doc id        */
    int32_t sm_count;     /* SMs of the device                                 */
    int32_t weight_format;/* BM25_WEIGHTS_FP32 or BM25_WEIGHTS_BF16            */
    int32_t posting_bytes;/* bytes per posting the score kernel streams: 8 or 4*/
} bm25_index_info;

/* Index loader (replaces: nothing in the reference loads the on-disk bm25s CSC index
 * animal_index_bm25/{indptr,indices,data}.csc.index.npy -- bm25_test.py:35-42 only round-trips it
 * through third-party bm25s; BM25v.index, bm25_native.py:59-74, takes the same three arrays as a
 * scipy csc_matrix).  Canonicalises the columns (sorted by doc id, duplicates summed) and pins them
 * in HBM re-bucketed for document-range tiles: int32 doc ids + fp32 weights in 16-byte aligned,
 * padded posting lists, {start, end} per term, and -- built on the first search -- a per-(heavy
 * term, document tile) first-posting table.
 *   indptr  [n_terms+1] int32, indices [nnz] int32 in [0,n_docs), data [nnz] fp32 (finite). */
int bm25_index_create(const int32_t* h_indptr, const int32_t* h_indices, const float* h_data,
                      int64_t n_terms, int64_t n_docs, int64_t nnz, int device,
                      int64_t doc_id_base, bm25_index** out);

/* Same, from arrays already resident on `device` (e.g. a synthetic index generated in HBM).
 * The arrays must already be canonical (each column strictly increasing in doc id).  They are only
 * read during this call: the index is always re-bucketed into library-owned memory (`borrow` is
 * accepted for source compatibility and ignored). */
int bm25_index_create_device(const int32_t* d_indptr, const int32_t* d_indices, const float* d_data,
                             int64_t n_terms, int64_t n_docs, int64_t nnz, int device,
                             int64_t doc_id_base, int borrow, bm25_index** out);

/* Index compression (new; SURVEY.md 8f row 4 -- the dtypes the reference's index declares are
 * params.index.json:1-12 "dtype": "float32", "int_dtype": "int32").  Converts the handle IN PLACE to
 * the compressed posting format: every weight is rounded to bf16 (round to nearest even) and the
 * score kernel for queries of <= 32 term slots streams 4-byte postings {uint16 byte offset of the
 * document's slot inside its document tile, bf16 weight} instead of 8-byte {int32, fp32}.  From
 * then on the handle IS the quantised index: searches and dense scores are bit-identical to the
 * reference run on the CSC matrix with bf16-rounded weights (fp32 accumulation, query-term order).
 * The 8-byte arrays are kept for light terms, wider queries and bm25_scores_dense.  Irreversible;
 * requires tile_docs <= 8192.  Not thread-safe against concurrent searches of the same handle
 * beyond the handle mutex (it waits for the device). */
int bm25_index_compress(bm25_index* index, int weight_format);

int bm25_index_destroy(bm25_index* index);
int bm25_index_get_info(const bm25_index* index, bm25_index_info* out);

/* Tuning knobs; value 0 restores the default.  They change how the work is cut, never the result:
 *   "tile_docs"      documents per warp score tile (multiple of 128; default 2048)
 *   "consumer_warps" warps (= document chunks) per CTA, 1..16 (default 8)
 *   "splits"         CTAs per query (default: enough for "waves" waves of resident CTAs)
 *   "waves"          target number of CTA waves when "splits" is automatic (default 10; 4 for k > 256)
 *   "cap"            candidate-buffer keys per CTA (default max(4k, 512) up to k = 1024, else 2k; power of two)
 *   "force_general"  1: treat the index as if it held non-positive weights (every doc competes)
 *   "cand_smem"      1: keep the candidate buffer in shared memory also for k > 256
 *   "heavy_min"      a term gets a row in the tile table when df*16 >= heavy_min * n_tiles (default 16,
 *                    i.e. one posting per document tile on average); lighter terms are walked by cursors
 *   "generic_kernel" 1: use the any-T kernel also for queries of <= 32 term slots (A/B switch)
 *   "no_query_sort" / "q_major" / "no_bulk_clear" / "no_epoch" / "no_packed"   1: keep the batch order /
 *                    query-major CTA order / vector-store tile clear / zero the tile after every tile /
 *                    read the 8-byte arrays of a compressed handle (A/B switches)
 *   "poison"         1 (debug): fill workspace and shared memory with 0xff before every search
 *   "no_hot" / "no_priming" / "no_theta_share"   1: disable the hot-list epilogue / the load-time
 *                    threshold priming / the per-query threshold shared between CTAs (A/B switches)
 *   "timing"         1: record CUDA events around the three kernels of every search */
int bm25_index_set_option(bm25_index* index, const char* name, int64_t value);

/* With option "timing" = 1 every bm25_search records CUDA events on its stream around its three
 * kernels; this waits for the last recorded search and returns their device durations in ms:
 * out_ms3 = {segment table, score accumulation + per-range top-k, merge}. */
int bm25_index_get_timing(bm25_index* index, float* out_ms3);

/* The hot path (replaces BM25v.search / _compute_relevance_from_scores, bm25_native.py:76-158,
 * and gpu_execute_query's gather->sum->top_k graph, gpu_bm25/common.py:64-85).
 *   d_queries   [Q,T] int32 row-major term ids, negative = padding (bm25_native.py:151); ids
 *               >= n_terms are rejected by the *_host entry point and ignored by this one.
 *   d_out_ids   [Q,k] int32   (doc id + doc_id_base)
 *   d_out_scores[Q,k] fp32
 * Requires 1 <= k <= min(n_docs, BM25_MAX_K) (the reference raises for k > n_docs).  k above
 * BM25_SMALL_K takes a slower, global-memory final merge. */
int bm25_search(bm25_index* index, const int32_t* d_queries, int64_t Q, int64_t T, int k,
                int32_t* d_out_ids, float* d_out_scores, void* cuda_stream);

/* Same with host buffers: H2D of the queries, the kernels, D2H of the results, then a
 * synchronise -- the end-to-end call a Python caller makes (BM25v.search drop-in). */
int bm25_search_host(bm25_index* index, const int32_t* h_queries, int64_t Q, int64_t T, int k,
                     int32_t* h_out_ids, float* h_out_scores);

/* Debug / parity: the dense per-query score vectors (what `doc_toks[:, q].sum(axis=1)` returns,
 * bm25_native.py:152).  d_out [Q, n_docs] fp32. */
int bm25_scores_dense(bm25_index* index, const int32_t* d_queries, int64_t Q, int64_t T,
                      float* d_out, void* cuda_stream);
int bm25_scores_dense_host(bm25_index* index, const int32_t* h_queries, int64_t Q, int64_t T,
                           float* h_out);

/* Multi-GPU final merge (new; the reference is single-device): merges n_lists candidate lists
 * laid out [n_lists, Q, k_in] (the layout an all-gather of per-shard results produces) into the
 * global top k_out per query.  `list_stride` = elements between consecutive lists in both arrays
 * (0 = Q*k_in, i.e. dense), so ids and scores may live interleaved in one all-gather buffer
 * [n_lists][2][Q][k_in].  Device pointers on `device`. */
int bm25_merge_topk(const int32_t* d_ids, const float* d_scores, int n_lists, int64_t list_stride,
                    int64_t Q, int k_in, int k_out, int32_t* d_out_ids, float* d_out_scores,
                    int device, void* cuda_stream);

/* Algorithmic bytes of a batch: sum_q 8*sum_t df(t) + 8*k (4*sum_t df(t) for a
 * compressed handle).  Host queries. */
int bm25_posting_bytes(const bm25_index* index, const int32_t* h_queries, int64_t Q, int64_t T,
                       int k, int64_t* out_bytes);

/* Number of CUDA kernels this library has launched so far in this process. */
int64_t bm25_kernel_launches(void);

const char* bm25_last_error(void);
const char* bm25_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BM25_B200_H */
