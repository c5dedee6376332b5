/* CPU oracle, C restatement -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restates the reference's CPU hot loop (reference bm25_native.py:129-158: per query, gather the
 * query terms' CSC columns, scatter-add them in query-term order into a dense fp32 score vector,
 * then top-k; bm25_native.py:204-214: partial selection + descending sort of the k).
 * Tie order: (score descending, doc id ascending) -- the reference leaves it unspecified.
 * Used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs; validated bit-for-bit against oracle/bm25_oracle.py, which is itself pinned against
 * golden vectors produced by the reference (tests/golden/).
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC -o oracle/libbm25_oracle.so oracle/bm25_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float s; int32_t d; } cand_t;

/* a is "worse" than b: lower score, or equal score and higher doc id */
static inline int worse(cand_t a, cand_t b) { return a.s < b.s || (a.s == b.s && a.d > b.d); }

static void sift_down(cand_t* h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && worse(h[l], h[m])) m = l;
        if (r < n && worse(h[r], h[m])) m = r;
        if (m == i) return;
        cand_t t = h[i]; h[i] = h[m]; h[m] = t; i = m;
    }
}

static int cmp_best_first(const void* pa, const void* pb) {
    cand_t a = *(const cand_t*)pa, b = *(const cand_t*)pb;
    if (worse(b, a)) return -1;
    if (worse(a, b)) return 1;
    return 0;
}

/* bm25_native.py:150-152 : dense fp32 accumulation in query-term order */
void oracle_scores_dense(const int32_t* indptr, const int32_t* indices, const float* data,
                         int64_t n_docs, const int32_t* query, int64_t T, float* out) {
    memset(out, 0, (size_t)n_docs * sizeof(float));
    for (int64_t j = 0; j < T; ++j) {
        int32_t t = query[j];
        if (t < 0) continue; /* padding, bm25_native.py:151 */
        for (int32_t p = indptr[t]; p < indptr[t + 1]; ++p) out[indices[p]] += data[p];
    }
}

/* bm25_native.py:204-214 : k best of a dense vector, best first */
static void topk_dense(const float* sc, int64_t n_docs, int k, cand_t* heap, int32_t* ids, float* vals) {
    int n = 0;
    for (int64_t d = 0; d < n_docs; ++d) {
        cand_t c = { sc[d], (int32_t)d };
        if (n < k) {
            heap[n++] = c;
            if (n == k) for (int i = k / 2 - 1; i >= 0; --i) sift_down(heap, k, i);
        } else if (worse(heap[0], c)) {
            heap[0] = c; sift_down(heap, k, 0);
        }
    }
    qsort(heap, (size_t)n, sizeof(cand_t), cmp_best_first);
    for (int i = 0; i < n; ++i) { ids[i] = heap[i].d; vals[i] = heap[i].s; }
}

/* BM25v.search on raw CSC arrays.  Returns 0, or -1 when k > n_docs (the reference raises). */
int oracle_search(const int32_t* indptr, const int32_t* indices, const float* data, int64_t n_docs,
                  const int32_t* queries, int64_t Q, int64_t T, int k,
                  int32_t* out_ids, float* out_scores, int n_threads) {
    if (k > n_docs || k < 0) return -1;
    if (k == 0 || Q == 0) return 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    #pragma omp parallel
    {
        float* sc = (float*)malloc((size_t)n_docs * sizeof(float));
        cand_t* heap = (cand_t*)malloc((size_t)k * sizeof(cand_t));
        #pragma omp for schedule(dynamic, 1)
        for (int64_t q = 0; q < Q; ++q) {
            oracle_scores_dense(indptr, indices, data, n_docs, queries + q * T, T, sc);
            topk_dense(sc, n_docs, k, heap, out_ids + q * k, out_scores + q * k);
        }
        free(sc); free(heap);
    }
    return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
