/*
 * oracle.c -- CPU restatement of archi's retrieval hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  Nothing under archi_b200/ may import, link or call it.
 *
 * PARITY UNPINNED.  The arithmetic of the reference's path does not live in /root/reference: it
 * runs inside third-party engines that are neither vendored nor installable here:
 *   - pgvector (Docker tag pgvector/pgvector:pg17, floating)  -> distance operators <=>, <->, <#>
 *   - pg_textsearch 0.4.2                                     -> BM25 operator <@>
 *   - sentence-transformers 5.1.2 (Pooling(mean) + Normalize) -> pool + L2 normalise
 * This file restates their *published* algorithms and anchors on the reference's own call sites
 * and score conventions; the reference's tests hold no numeric golden vectors for this path, only
 * score-convention known answers (tests/unit/test_postgres_vectorstore.py:196,259-261,304-306,
 * 352-366), which tests/test_oracle.py checks.
 *
 * Reference call sites restated:
 *   src/data_manager/vectorstore/postgres_vectorstore.py:74-78    metric -> operator map
 *   src/data_manager/vectorstore/postgres_vectorstore.py:317-332  SELECT emb <op> q AS distance ... ORDER BY distance ASC LIMIT k
 *   src/data_manager/vectorstore/postgres_vectorstore.py:361      score = 1 - distance (cosine) | distance (l2, inner_product)
 *   src/data_manager/vectorstore/postgres_vectorstore.py:441-456  semantic = 1.0 - (emb <op> q); combined = semantic*ws + COALESCE(bm25,0)*wb; ORDER BY combined DESC LIMIT k
 *
 * Precision contract of the "reference CPU path" restated here: float4 storage, float
 * accumulators walked in index order, result widened to double (pgvector's documented
 * behaviour [external]).  The fp64 *truth* ranking lives in oracle.py (numpy).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_COSINE = 0, ORC_L2 = 1, ORC_IP = 2 };

static inline float bf16_to_f32(uint16_t h)
{
    uint32_t u = ((uint32_t)h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

/* `<#>` = negative inner product; `<->` = sqrt(sum (a-b)^2); `<=>` = 1 - a.b/sqrt(|a|^2 |b|^2),
 * similarity clamped to [-1,1].  Float accumulators, double result. */
double orc_distance_f32(int metric, int dim, const float *a, const float *b)
{
    if (metric == ORC_IP) {
        float dot = 0.0f;
        for (int i = 0; i < dim; i++) dot += a[i] * b[i];
        return (double)-dot;
    }
    if (metric == ORC_L2) {
        float acc = 0.0f;
        for (int i = 0; i < dim; i++) {
            float diff = a[i] - b[i];
            acc += diff * diff;
        }
        return sqrt((double)acc);
    }
    float dot = 0.0f, na = 0.0f, nb = 0.0f;
    for (int i = 0; i < dim; i++) {
        dot += a[i] * b[i];
        na += a[i] * a[i];
        nb += b[i] * b[i];
    }
    double sim = (double)dot / sqrt((double)na * (double)nb);
    if (sim > 1.0) sim = 1.0;
    else if (sim < -1.0) sim = -1.0;
    return 1.0 - sim;
}

/* Same, corpus row stored as bf16 (our bf16 storage mode: the stored value is what is searched). */
static double distance_row(int metric, int dim, const void *row, int row_is_bf16, const float *q,
                           float *scratch)
{
    if (!row_is_bf16) return orc_distance_f32(metric, dim, (const float *)row, q);
    const uint16_t *r = (const uint16_t *)row;
    for (int i = 0; i < dim; i++) scratch[i] = bf16_to_f32(r[i]);
    return orc_distance_f32(metric, dim, scratch, q);
}

/* Bounded max-heap on (key, id): keeps the k smallest keys; among equal keys the lower id wins. */
typedef struct { double key; int64_t id; } ent_t;

static inline int ent_worse(ent_t a, ent_t b) /* a sorts after b */
{
    return a.key > b.key || (a.key == b.key && a.id > b.id);
}
static void heap_sift_down(ent_t *h, int n, int i)
{
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && ent_worse(h[l], h[m])) m = l;
        if (r < n && ent_worse(h[r], h[m])) m = r;
        if (m == i) return;
        ent_t t = h[i]; h[i] = h[m]; h[m] = t;
        i = m;
    }
}
static void heap_offer(ent_t *h, int *n, int k, ent_t e)
{
    if (*n < k) {
        int i = (*n)++;
        h[i] = e;
        while (i > 0) {
            int p = (i - 1) / 2;
            if (!ent_worse(h[i], h[p])) break;
            ent_t t = h[i]; h[i] = h[p]; h[p] = t;
            i = p;
        }
    } else if (k > 0 && ent_worse(h[0], e)) {
        h[0] = e;
        heap_sift_down(h, k, 0);
    }
}
static int ent_cmp(const void *a, const void *b)
{
    const ent_t *x = a, *y = b;
    if (ent_worse(*x, *y)) return 1;
    if (ent_worse(*y, *x)) return -1;
    return 0;
}

/*
 * ORDER BY distance ASC LIMIT k over a sequential scan (postgres_vectorstore.py:317-332).
 * mask: optional bitmask, bit i set = row i passes the WHERE clause (collection / metadata
 * filter / not-deleted, :296-310); NULL = all rows.
 * Outputs: out_dist[k] (double, ascending), out_ids[k]; unfilled tail is id -1, dist +inf.
 * Returns the number of rows produced = min(k, rows passing).
 */
int orc_scan_topk(int metric, const void *corpus, int corpus_is_bf16, int64_t n, int dim,
                  const float *query, int k, const uint32_t *mask, double *out_dist,
                  int64_t *out_ids)
{
    if (k <= 0) return 0;
    ent_t *heap = malloc(sizeof(ent_t) * (size_t)k);
    float *scratch = malloc(sizeof(float) * (size_t)(dim > 0 ? dim : 1));
    int cnt = 0;
    size_t row_bytes = (size_t)dim * (corpus_is_bf16 ? 2 : 4);
    for (int64_t i = 0; i < n; i++) {
        if (mask && !((mask[i >> 5] >> (i & 31)) & 1u)) continue;
        ent_t e;
        e.key = distance_row(metric, dim, (const char *)corpus + row_bytes * (size_t)i,
                             corpus_is_bf16, query, scratch);
        e.id = i;
        if (e.key != e.key) e.key = INFINITY; /* NaN distances sort last */
        heap_offer(heap, &cnt, k, e);
    }
    qsort(heap, (size_t)cnt, sizeof(ent_t), ent_cmp);
    for (int i = 0; i < k; i++) {
        out_dist[i] = i < cnt ? heap[i].key : INFINITY;
        out_ids[i] = i < cnt ? heap[i].id : -1;
    }
    free(heap);
    free(scratch);
    return cnt;
}

/* score = 1 - distance for cosine, the raw distance otherwise (postgres_vectorstore.py:361). */
double orc_score_from_distance(int metric, double distance)
{
    return metric == ORC_COSINE ? 1.0 - distance : distance;
}

/*
 * hybrid_search's SQL (postgres_vectorstore.py:435-457): for every row passing WHERE,
 *   semantic = 1.0 - (emb <op> q)          (the same expression for all three metrics)
 *   combined = semantic * ws + COALESCE(bm25, 0) * wb        ORDER BY combined DESC LIMIT k
 * bm25: dense per-row array, NaN encodes SQL NULL; may be NULL (all NULL).
 * Outputs: out_combined[k] descending, out_ids[k].
 */
int orc_hybrid_topk(int metric, const void *corpus, int corpus_is_bf16, int64_t n, int dim,
                    const float *query, const double *bm25, double ws, double wb, int k,
                    const uint32_t *mask, double *out_combined, int64_t *out_ids)
{
    if (k <= 0) return 0;
    ent_t *heap = malloc(sizeof(ent_t) * (size_t)k);
    float *scratch = malloc(sizeof(float) * (size_t)(dim > 0 ? dim : 1));
    int cnt = 0;
    size_t row_bytes = (size_t)dim * (corpus_is_bf16 ? 2 : 4);
    for (int64_t i = 0; i < n; i++) {
        if (mask && !((mask[i >> 5] >> (i & 31)) & 1u)) continue;
        double d = distance_row(metric, dim, (const char *)corpus + row_bytes * (size_t)i,
                                corpus_is_bf16, query, scratch);
        double sem = 1.0 - d;
        double b = (bm25 && bm25[i] == bm25[i]) ? bm25[i] : 0.0;
        ent_t e;
        e.key = -(sem * ws + b * wb); /* DESC order == ascending negated key */
        e.id = i;
        if (e.key != e.key) e.key = INFINITY;
        heap_offer(heap, &cnt, k, e);
    }
    qsort(heap, (size_t)cnt, sizeof(ent_t), ent_cmp);
    for (int i = 0; i < k; i++) {
        out_combined[i] = i < cnt ? -heap[i].key : -INFINITY;
        out_ids[i] = i < cnt ? heap[i].id : -1;
    }
    free(heap);
    free(scratch);
    return cnt;
}

/*
 * Batch of independent queries, one per "backend": Postgres runs one ORDER BY ... LIMIT query on
 * one core, concurrent clients run on different cores.  nthreads workers pull queries.
 */
typedef struct {
    int metric, is_bf16, dim, k, nq, nthreads, tid;
    const void *corpus;
    int64_t n;
    const float *queries;
    const uint32_t *mask;
    double *out_dist;
    int64_t *out_ids;
} batch_arg_t;

static void *batch_worker(void *p)
{
    batch_arg_t *a = p;
    for (int q = a->tid; q < a->nq; q += a->nthreads)
        orc_scan_topk(a->metric, a->corpus, a->is_bf16, a->n, a->dim,
                      a->queries + (size_t)q * a->dim, a->k, a->mask,
                      a->out_dist + (size_t)q * a->k, a->out_ids + (size_t)q * a->k);
    return NULL;
}

int orc_scan_topk_batch(int metric, const void *corpus, int corpus_is_bf16, int64_t n, int dim,
                        const float *queries, int nq, int k, const uint32_t *mask,
                        double *out_dist, int64_t *out_ids, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nq) nthreads = nq > 0 ? nq : 1;
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
    batch_arg_t *args = malloc(sizeof(batch_arg_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        batch_arg_t a = { metric, corpus_is_bf16, dim, k, nq, nthreads, t, corpus, n,
                          queries, mask, out_dist, out_ids };
        args[t] = a;
        pthread_create(&th[t], NULL, batch_worker, &args[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    free(args);
    return 0;
}

/*
 * sentence-transformers Pooling(mean) + Normalize [external], the tail of
 * Embeddings.embed_documents (manager.py:373, postgres_vectorstore.py:143,245,390):
 *   pooled = sum_t h[t]*m[t] / max(sum_t m[t], 1e-9);   out = pooled / max(|pooled|_2, 1e-12)
 * hidden [B, L, H] fp32 row-major, mask [B, L] int64, out [B, H] fp32.  Float arithmetic, tokens
 * walked in order.
 */
void orc_pool_normalize(const float *hidden, const int64_t *mask, int B, int L, int H, float *out)
{
    for (int b = 0; b < B; b++) {
        float *o = out + (size_t)b * H;
        for (int h = 0; h < H; h++) o[h] = 0.0f;
        float msum = 0.0f;
        for (int t = 0; t < L; t++) {
            float m = (float)mask[(size_t)b * L + t];
            msum += m;
            const float *row = hidden + ((size_t)b * L + t) * H;
            for (int h = 0; h < H; h++) o[h] += row[h] * m;
        }
        if (msum < 1e-9f) msum = 1e-9f;
        float nrm = 0.0f;
        for (int h = 0; h < H; h++) {
            o[h] = o[h] / msum;
            nrm += o[h] * o[h];
        }
        nrm = sqrtf(nrm);
        if (nrm < 1e-12f) nrm = 1e-12f;
        for (int h = 0; h < H; h++) o[h] = o[h] / nrm;
    }
}
