/*
 * hnsw.c -- CPU restatement of the reference store's DEFAULT index, for the recall report only.
 * TEST INFRASTRUCTURE (same rules as oracle.c): never linked into the product.
 *
 * The reference builds `CREATE INDEX ... USING hnsw (embedding vector_cosine_ops) WITH (m = 16,
 * ef_construction = 64)` (src/cli/templates/init.sql:280-284; knobs in
 * src/cli/managers/templates_manager.py:427-429) and never sets hnsw.ef_search, so pgvector's default
 * of 40 applies [external].  `ORDER BY embedding <=> q LIMIT k` on the semantic path may therefore be
 * answered approximately.  pgvector is not available here; this file restates the published HNSW
 * algorithm (Malkov & Yashunin, Alg. 1-5: level ~ floor(-ln U / ln m), greedy descent, ef-bounded
 * best-first search per layer, heuristic neighbour selection, 2m links on layer 0) -- PARITY
 * UNPINNED against pgvector's implementation details (tie handling, deletion, vacuum).
 * Distance: cosine distance on unit-length rows = 1 - dot.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n_cap, n, dim, m, m0, efc, max_level, entry;
    const float *data;      /* [n_cap, dim], owned by the caller, unit-length rows */
    int *level;             /* [n_cap] */
    int **links;            /* links[i] = per level: [count, ids...] blocks of (cap_l + 1) ints */
    uint32_t *visited;      /* epoch stamps */
    uint32_t epoch;
    uint64_t rng;
} hnsw_t;

typedef struct { float d; int id; } cand_t;

static float dist(const hnsw_t *h, const float *q, int id)
{
    const float *x = h->data + (size_t)id * h->dim;
    float dot = 0.f;
    for (int i = 0; i < h->dim; i++) dot += q[i] * x[i];
    return 1.0f - dot;
}

/* binary heaps on cand_t: min-heap (closest first) and max-heap (furthest first) */
static void heap_push(cand_t *a, int *n, cand_t c, int maxheap)
{
    int i = (*n)++;
    a[i] = c;
    while (i > 0) {
        int p = (i - 1) / 2;
        int up = maxheap ? (a[i].d > a[p].d) : (a[i].d < a[p].d);
        if (!up) break;
        cand_t t = a[i]; a[i] = a[p]; a[p] = t;
        i = p;
    }
}
static cand_t heap_pop(cand_t *a, int *n, int maxheap)
{
    cand_t top = a[0];
    a[0] = a[--(*n)];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, b = i;
        if (l < *n && (maxheap ? a[l].d > a[b].d : a[l].d < a[b].d)) b = l;
        if (r < *n && (maxheap ? a[r].d > a[b].d : a[r].d < a[b].d)) b = r;
        if (b == i) break;
        cand_t t = a[i]; a[i] = a[b]; a[b] = t;
        i = b;
    }
    return top;
}

static int level_cap(const hnsw_t *h, int lc) { return lc == 0 ? h->m0 : h->m; }
static int *link_block(const hnsw_t *h, int id, int lc)
{
    int off = 0;
    for (int l = 0; l < lc; l++) off += level_cap(h, l) + 1;
    return h->links[id] + off;
}

/* Alg. 2: best-first search of one layer, returns up to ef closest in `out` (unsorted), count in *nout */
static void search_layer(hnsw_t *h, const float *q, const cand_t *eps, int neps, int ef, int lc, cand_t *out, int *nout)
{
    cand_t *cand = malloc(sizeof(cand_t) * (size_t)(h->n + 1));
    cand_t *res = malloc(sizeof(cand_t) * (size_t)(ef + 1));
    int nc = 0, nr = 0;
    h->epoch++;
    for (int i = 0; i < neps; i++) {
        h->visited[eps[i].id] = h->epoch;
        heap_push(cand, &nc, eps[i], 0);
        heap_push(res, &nr, eps[i], 1);
        if (nr > ef) heap_pop(res, &nr, 1);
    }
    while (nc > 0) {
        cand_t c = heap_pop(cand, &nc, 0);
        if (nr >= ef && c.d > res[0].d) break;
        int *blk = link_block(h, c.id, lc);
        for (int j = 1; j <= blk[0]; j++) {
            int e = blk[j];
            if (h->visited[e] == h->epoch) continue;
            h->visited[e] = h->epoch;
            cand_t ce = { dist(h, q, e), e };
            if (nr < ef || ce.d < res[0].d) {
                heap_push(cand, &nc, ce, 0);
                heap_push(res, &nr, ce, 1);
                if (nr > ef) heap_pop(res, &nr, 1);
            }
        }
    }
    memcpy(out, res, sizeof(cand_t) * (size_t)nr);
    *nout = nr;
    free(cand);
    free(res);
}

static int cmp_cand(const void *a, const void *b)
{
    const cand_t *x = a, *y = b;
    return x->d < y->d ? -1 : x->d > y->d ? 1 : (x->id - y->id);
}

/* Alg. 4: heuristic selection of at most M neighbours from candidates W (sorted ascending here) */
static int select_neighbors(const hnsw_t *h, cand_t *w, int nw, int M, int *out)
{
    qsort(w, (size_t)nw, sizeof(cand_t), cmp_cand);
    int n = 0;
    for (int i = 0; i < nw && n < M; i++) {
        int good = 1;
        const float *xi = h->data + (size_t)w[i].id * h->dim;
        for (int j = 0; j < n; j++) {
            if (dist(h, xi, out[j]) < w[i].d) { good = 0; break; }
        }
        if (good) out[n++] = w[i].id;
    }
    return n;
}

hnsw_t *hnsw_create(const float *data, int n_cap, int dim, int m, int ef_construction, uint64_t seed)
{
    hnsw_t *h = calloc(1, sizeof(hnsw_t));
    h->n_cap = n_cap; h->dim = dim; h->m = m; h->m0 = 2 * m; h->efc = ef_construction;
    h->data = data; h->entry = -1; h->max_level = -1;
    h->level = calloc((size_t)n_cap, sizeof(int));
    h->links = calloc((size_t)n_cap, sizeof(int *));
    h->visited = calloc((size_t)n_cap, sizeof(uint32_t));
    h->rng = seed ? seed : 0x9E3779B97F4A7C15ull;
    return h;
}

void hnsw_destroy(hnsw_t *h)
{
    for (int i = 0; i < h->n; i++) free(h->links[i]);
    free(h->links); free(h->level); free(h->visited); free(h);
}

static double next_uniform(hnsw_t *h)
{
    h->rng = h->rng * 6364136223846793005ull + 1442695040888963407ull;
    return ((h->rng >> 11) + 1.0) / 9007199254740993.0;
}

/* Alg. 1: insert row `id` (rows must be inserted in order 0, 1, 2, ...) */
void hnsw_insert(hnsw_t *h, int id)
{
    const float *q = h->data + (size_t)id * h->dim;
    int lvl = (int)floor(-log(next_uniform(h)) / log((double)h->m));
    h->level[id] = lvl;
    int ints = 0;
    for (int l = 0; l <= lvl; l++) ints += level_cap(h, l) + 1;
    h->links[id] = calloc((size_t)ints, sizeof(int));
    h->n = id + 1;
    if (h->entry < 0) { h->entry = id; h->max_level = lvl; return; }

    cand_t ep = { dist(h, q, h->entry), h->entry };
    cand_t *w = malloc(sizeof(cand_t) * (size_t)(h->efc + 1));
    int nw;
    for (int lc = h->max_level; lc > lvl; lc--) {
        search_layer(h, q, &ep, 1, 1, lc, w, &nw);
        for (int i = 0; i < nw; i++) if (w[i].d < ep.d) ep = w[i];
    }
    cand_t *eps = malloc(sizeof(cand_t) * (size_t)(h->efc + 1));
    int neps = 1;
    eps[0] = ep;
    int *sel = malloc(sizeof(int) * (size_t)(h->m0 + 1));
    cand_t *tmp = malloc(sizeof(cand_t) * (size_t)(h->m0 + 2));
    for (int lc = (lvl < h->max_level ? lvl : h->max_level); lc >= 0; lc--) {
        search_layer(h, q, eps, neps, h->efc, lc, w, &nw);
        memcpy(eps, w, sizeof(cand_t) * (size_t)nw);
        neps = nw;
        int ns = select_neighbors(h, w, nw, h->m, sel);
        int *blk = link_block(h, id, lc);
        blk[0] = ns;
        memcpy(blk + 1, sel, sizeof(int) * (size_t)ns);
        const int cap = level_cap(h, lc);
        for (int i = 0; i < ns; i++) {
            int e = sel[i];
            int *eb = link_block(h, e, lc);
            if (eb[0] < cap) {
                eb[++eb[0]] = id;
            } else {
                /* shrink: re-select among the old links plus the new one */
                const float *xe = h->data + (size_t)e * h->dim;
                int nt = 0;
                for (int j = 1; j <= eb[0]; j++) { tmp[nt].id = eb[j]; tmp[nt].d = dist(h, xe, eb[j]); nt++; }
                tmp[nt].id = id; tmp[nt].d = dist(h, xe, id); nt++;
                int keep[64];
                int nk = select_neighbors(h, tmp, nt, cap, keep);
                eb[0] = nk;
                memcpy(eb + 1, keep, sizeof(int) * (size_t)nk);
            }
        }
    }
    if (lvl > h->max_level) { h->max_level = lvl; h->entry = id; }
    free(w); free(eps); free(sel); free(tmp);
}

/* Alg. 5: k nearest with beam ef; ids ascending by distance, -1 padded.  Returns rows found. */
int hnsw_search(hnsw_t *h, const float *q, int k, int ef, int *out_ids, float *out_dist)
{
    for (int i = 0; i < k; i++) { out_ids[i] = -1; out_dist[i] = INFINITY; }
    if (h->entry < 0) return 0;
    cand_t ep = { dist(h, q, h->entry), h->entry };
    cand_t *w = malloc(sizeof(cand_t) * (size_t)((ef > 1 ? ef : 1) + 1));
    int nw;
    for (int lc = h->max_level; lc > 0; lc--) {
        search_layer(h, q, &ep, 1, 1, lc, w, &nw);
        for (int i = 0; i < nw; i++) if (w[i].d < ep.d) ep = w[i];
    }
    search_layer(h, q, &ep, 1, ef, 0, w, &nw);
    qsort(w, (size_t)nw, sizeof(cand_t), cmp_cand);
    int n = nw < k ? nw : k;
    for (int i = 0; i < n; i++) { out_ids[i] = w[i].id; out_dist[i] = w[i].d; }
    free(w);
    return n;
}
