/* TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product path (ebsd_vae_b200/).
 *
 * CPU restatement of the APPROXIMATE index behind the reference's default dictionary,
 *   ChromaLatentVectorDatabase  (latice/index/chroma_db.py:124-130: collection metadata {"hnsw:space": "cosine"},
 *   query at chroma_db.py:254-258),
 * i.e. chroma-hnswlib 0.7.6 (uv.lock:528-529) as chromadb 0.6.3 (uv.lock:553-554) drives it: space "cosine"
 * (rows and queries L2-normalised, distance = 1 - inner product), M = 16, ef_construction = 100, ef_search = 10
 * (chromadb's HnswParams defaults as far as they can be recalled without the wheel; the caller passes them).
 *
 * Neither package is installable in this image (no network) and the reference's own tests mock the collection
 * (tests/index/test_chroma_db.py:267-291), so there is NO golden vector for this file: PARITY UNPINNED.  It restates
 * the published algorithm of hnswlib's HierarchicalNSW (Malkov & Yashunin, "Efficient and robust approximate nearest
 * neighbor search using Hierarchical Navigable Small World graphs"; hnswalg.h: getRandomLevel, searchBaseLayer,
 * getNeighborsByHeuristic2, mutuallyConnectNewElement, addPoint, searchKnn) so that bench.py can report the RECALL of
 * the reference's approximate search next to the exact search of this repository (north star: "with the reference's
 * Chroma/HNSW recall reported alongside").  What it is checked against (tests/test_oracle_hnsw.py): the exact oracle
 * (recall -> 1 as ef grows, every returned distance is the true distance of the returned row, lists ascending) and the
 * structural invariants of the graph (degree bounds, no self links, level distribution).
 *
 * Differences that do not matter for a recall figure: Chroma inserts in batches from several threads (its graphs are
 * not reproducible run to run either); hnswlib draws levels from std::default_random_engine(100) -- restated below as
 * minstd_rand0 + libstdc++'s two-draw generate_canonical<double, 53>.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float d;
    int id;
} Pair;

/* binary max-heap on (d, id), the order of std::priority_queue<std::pair<float, int>> */
typedef struct {
    Pair *a;
    int n, cap;
} Heap;

static int pair_less(Pair x, Pair y) { return x.d < y.d || (x.d == y.d && x.id < y.id); }

static void heap_init(Heap *h, int cap) {
    h->a = (Pair *)malloc(sizeof(Pair) * (size_t)(cap > 16 ? cap : 16));
    h->n = 0;
    h->cap = cap > 16 ? cap : 16;
}
static void heap_free(Heap *h) { free(h->a); }
static void heap_push(Heap *h, float d, int id) {
    if (h->n == h->cap) {
        h->cap *= 2;
        h->a = (Pair *)realloc(h->a, sizeof(Pair) * (size_t)h->cap);
    }
    int i = h->n++;
    Pair v = {d, id};
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!pair_less(h->a[p], v)) break;
        h->a[i] = h->a[p];
        i = p;
    }
    h->a[i] = v;
}
static Pair heap_pop(Heap *h) {
    Pair top = h->a[0], v = h->a[--h->n];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && pair_less(h->a[c], h->a[c + 1])) ++c;
        if (!pair_less(v, h->a[c])) break;
        h->a[i] = h->a[c];
        i = c;
    }
    if (h->n > 0) h->a[i] = v;
    return top;
}

typedef struct {
    int d, M, maxM, maxM0, efc;
    int64_t n;            /* rows inserted */
    const float *rows;    /* [N, d] normalised, owned by the caller */
    int *level;           /* [N] */
    int *link0;           /* [N][maxM0 + 1]: count, neighbours */
    int **linkup;         /* [N] -> [level][maxM + 1] */
    int enter, maxlevel;
    double mult;
    uint32_t rng;         /* minstd_rand0 state */
} Hnsw;

static inline float dist_ip(const Hnsw *h, const float *a, const float *b) {
    float s = 0.f;
    for (int j = 0; j < h->d; ++j) s += a[j] * b[j];
    return 1.0f - s;   /* hnswlib InnerProductDistance */
}
static inline const float *row(const Hnsw *h, int i) { return h->rows + (size_t)i * h->d; }
static inline int *links(const Hnsw *h, int i, int lvl) {
    return lvl == 0 ? h->link0 + (size_t)i * (h->maxM0 + 1) : h->linkup[i] + (size_t)(lvl - 1) * (h->maxM + 1);
}

static uint32_t minstd_next(Hnsw *h) {
    h->rng = (uint32_t)(((uint64_t)h->rng * 16807u) % 2147483647u);
    return h->rng;
}
/* libstdc++ generate_canonical<double, 53>(minstd_rand0): two draws, range R = max - min + 1 = 2147483646 */
static double canonical(Hnsw *h) {
    const double R = 2147483646.0;
    double sum = (double)(minstd_next(h) - 1u);
    sum += (double)(minstd_next(h) - 1u) * R;
    double r = sum / (R * R);
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}
static int random_level(Hnsw *h) {
    double r = -log(canonical(h)) * h->mult;
    return (int)r;
}

typedef struct {
    uint32_t *tag;
    uint32_t epoch;
} Visited;

/* hnswalg.h searchBaseLayer (construction) / searchBaseLayerST (query, no deletions, no filter): `ef` closest rows found
 * from entry point ep at `lvl`, as a max-heap `top` (caller-initialised, emptied here).  The query variant stops as soon
 * as the nearest open candidate is farther than the current bound even when fewer than ef results are held. */
static void search_layer(const Hnsw *h, const float *q, int ep, int lvl, int ef, int query_variant, Visited *vis, Heap *top,
                         Heap *cand) {
    top->n = 0;
    cand->n = 0;
    if (++vis->epoch == 0) {
        memset(vis->tag, 0, sizeof(uint32_t) * (size_t)h->n);
        vis->epoch = 1;
    }
    float lower = dist_ip(h, q, row(h, ep));
    heap_push(top, lower, ep);
    heap_push(cand, -lower, ep);
    vis->tag[ep] = vis->epoch;
    while (cand->n > 0) {
        Pair c = cand->a[0];
        if (-c.d > lower && (top->n == ef || query_variant)) break;
        heap_pop(cand);
        const int *ll = links(h, c.id, lvl);
        for (int j = 1; j <= ll[0]; ++j) {
            const int nb = ll[j];
            if (vis->tag[nb] == vis->epoch) continue;
            vis->tag[nb] = vis->epoch;
            const float dn = dist_ip(h, q, row(h, nb));
            if (top->n < ef || lower > dn) {
                heap_push(cand, -dn, nb);
                heap_push(top, dn, nb);
                if (top->n > ef) heap_pop(top);
                if (top->n > 0) lower = top->a[0].d;
            }
        }
    }
}

/* hnswalg.h getNeighborsByHeuristic2: keep at most M candidates, nearest first, dropping one that is closer to an
 * already kept neighbour than to the query point.  In/out: max-heap `top`. */
static void select_heuristic(const Hnsw *h, Heap *top, int M, Heap *scratch, Pair *keep) {
    if (top->n < M) return;
    scratch->n = 0;
    while (top->n > 0) {
        Pair p = heap_pop(top);
        heap_push(scratch, -p.d, p.id);
    }
    int nk = 0;
    while (scratch->n > 0) {
        if (nk >= M) break;
        Pair cur = heap_pop(scratch);
        const float dq = -cur.d;
        int good = 1;
        for (int j = 0; j < nk; ++j) {
            const float dd = dist_ip(h, row(h, keep[j].id), row(h, cur.id));
            if (dd < dq) {
                good = 0;
                break;
            }
        }
        if (good) keep[nk++] = cur;
    }
    for (int j = 0; j < nk; ++j) heap_push(top, -keep[j].d, keep[j].id);
}

/* hnswalg.h mutuallyConnectNewElement; returns the closest selected neighbour (next entry point) */
static int connect_new(Hnsw *h, int cur, Heap *top, int lvl, Heap *scratch, Heap *other, Pair *keep) {
    const int mmax = lvl ? h->maxM : h->maxM0;
    select_heuristic(h, top, h->M, scratch, keep);
    int sel[256], ns = 0;
    while (top->n > 0) sel[ns++] = heap_pop(top).id;   /* farthest first */
    const int next_ep = sel[ns - 1];
    int *ll = links(h, cur, lvl);
    ll[0] = ns;
    for (int j = 0; j < ns; ++j) ll[1 + j] = sel[j];
    for (int j = 0; j < ns; ++j) {
        int *lo = links(h, sel[j], lvl);
        if (lo[0] < mmax) {
            lo[1 + lo[0]] = cur;
            lo[0] += 1;
        } else {
            other->n = 0;
            heap_push(other, dist_ip(h, row(h, cur), row(h, sel[j])), cur);
            for (int t = 1; t <= lo[0]; ++t) heap_push(other, dist_ip(h, row(h, lo[t]), row(h, sel[j])), lo[t]);
            select_heuristic(h, other, mmax, scratch, keep);
            int c = 0;
            while (other->n > 0) lo[1 + c++] = heap_pop(other).id;
            lo[0] = c;
        }
    }
    return next_ep;
}

static int greedy_descend(const Hnsw *h, const float *q, int ep, int from_level, int to_level_exclusive) {
    float cd = dist_ip(h, q, row(h, ep));
    for (int lvl = from_level; lvl > to_level_exclusive; --lvl) {
        int changed = 1;
        while (changed) {
            changed = 0;
            const int *ll = links(h, ep, lvl);
            for (int j = 1; j <= ll[0]; ++j) {
                const float dn = dist_ip(h, q, row(h, ll[j]));
                if (dn < cd) {
                    cd = dn;
                    ep = ll[j];
                    changed = 1;
                }
            }
        }
    }
    return ep;
}

void *ebsd_oracle_hnsw_build(const float *rows_hat, int64_t N, int d, int M, int ef_construction, unsigned seed) {
    if (N <= 0 || N > 0x7fffffff || M < 2 || M > 64 || d < 1) return NULL;
    Hnsw *h = (Hnsw *)calloc(1, sizeof(Hnsw));
    h->d = d;
    h->M = M;
    h->maxM = M;
    h->maxM0 = 2 * M;
    h->efc = ef_construction > M ? ef_construction : M;
    h->rows = rows_hat;
    h->level = (int *)calloc((size_t)N, sizeof(int));
    h->link0 = (int *)calloc((size_t)N * (size_t)(h->maxM0 + 1), sizeof(int));
    h->linkup = (int **)calloc((size_t)N, sizeof(int *));
    h->enter = -1;
    h->maxlevel = -1;
    h->mult = 1.0 / log((double)M);
    h->rng = seed % 2147483647u;
    if (h->rng == 0) h->rng = 1;
    Visited vis = {(uint32_t *)calloc((size_t)N, sizeof(uint32_t)), 0};
    Heap top, cand, scratch, other;
    heap_init(&top, h->efc + 2);
    heap_init(&cand, 4 * h->efc);
    heap_init(&scratch, h->efc + 2);
    heap_init(&other, h->maxM0 + 2);
    Pair *keep = (Pair *)malloc(sizeof(Pair) * (size_t)(h->maxM0 + 2));
    for (int64_t i = 0; i < N; ++i) {
        const int cur = (int)i;
        const int lvl = random_level(h);
        h->level[cur] = lvl;
        if (lvl > 0) h->linkup[cur] = (int *)calloc((size_t)lvl * (size_t)(h->maxM + 1), sizeof(int));
        h->n = i + 1;
        const int maxl = h->maxlevel;
        int ep = h->enter;
        if (ep != -1) {
            const float *q = row(h, cur);
            if (lvl < maxl) ep = greedy_descend(h, q, ep, maxl, lvl);
            for (int l = lvl < maxl ? lvl : maxl; l >= 0; --l) {
                search_layer(h, q, ep, l, h->efc, 0, &vis, &top, &cand);
                ep = connect_new(h, cur, &top, l, &scratch, &other, keep);
            }
        } else {
            h->enter = 0;
            h->maxlevel = lvl;
        }
        if (lvl > maxl) {
            h->enter = cur;
            h->maxlevel = lvl;
        }
    }
    free(keep);
    heap_free(&top);
    heap_free(&cand);
    heap_free(&scratch);
    heap_free(&other);
    free(vis.tag);
    return h;
}

void ebsd_oracle_hnsw_free(void *hp) {
    Hnsw *h = (Hnsw *)hp;
    if (!h) return;
    for (int64_t i = 0; i < h->n; ++i) free(h->linkup[i]);
    free(h->linkup);
    free(h->link0);
    free(h->level);
    free(h);
}

typedef struct {
    const Hnsw *h;
    const float *queries;
    int64_t q0, q1;
    int k, ef;
    float *out_dist;
    int64_t *out_idx;
} SearchJob;

/* hnswalg.h searchKnn: greedy descent to level 0, then the ef = max(ef, k) beam; the k nearest, ascending distance,
 * rows that were not found are reported as index -1 / distance +inf */
static void *search_worker(void *arg) {
    SearchJob *job = (SearchJob *)arg;
    const Hnsw *h = job->h;
    const int ef = job->ef > job->k ? job->ef : job->k;
    Visited vis = {(uint32_t *)calloc((size_t)h->n, sizeof(uint32_t)), 0};
    Heap top, cand;
    heap_init(&top, ef + 2);
    heap_init(&cand, 4 * ef);
    for (int64_t qi = job->q0; qi < job->q1; ++qi) {
        const float *q = job->queries + (size_t)qi * h->d;
        const int ep = greedy_descend(h, q, h->enter, h->maxlevel, 0);
        search_layer(h, q, ep, 0, ef, 1, &vis, &top, &cand);
        while (top.n > job->k) heap_pop(&top);
        float *od = job->out_dist + (size_t)qi * job->k;
        int64_t *oi = job->out_idx + (size_t)qi * job->k;
        for (int j = 0; j < job->k; ++j) {
            od[j] = INFINITY;
            oi[j] = -1;
        }
        for (int j = top.n - 1; j >= 0; --j) {
            Pair p = heap_pop(&top);
            od[j] = p.d;
            oi[j] = p.id;
        }
    }
    heap_free(&top);
    heap_free(&cand);
    free(vis.tag);
    return NULL;
}

void ebsd_oracle_hnsw_search(void *hp, const float *queries_hat, int64_t Q, int k, int ef, float *out_dist, int64_t *out_idx,
                             int nthreads) {
    const Hnsw *h = (const Hnsw *)hp;
    if (!h || Q <= 0 || k <= 0) return;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if ((int64_t)nthreads > Q) nthreads = (int)Q;
    pthread_t th[64];
    SearchJob jobs[64];
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = (SearchJob){h, queries_hat, Q * t / nthreads, Q * (t + 1) / nthreads, k, ef, out_dist, out_idx};
        pthread_create(&th[t], NULL, search_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* graph statistics for the structural tests: out[0] = max level, out[1] = entry point, out[2] = max degree at level 0,
 * out[3] = max degree above level 0, out[4] = self links, out[5] = out-of-range links, out[6] = rows with level > 0 */
void ebsd_oracle_hnsw_stats(void *hp, int64_t *out) {
    const Hnsw *h = (const Hnsw *)hp;
    memset(out, 0, 7 * sizeof(int64_t));
    out[0] = h->maxlevel;
    out[1] = h->enter;
    for (int64_t i = 0; i < h->n; ++i) {
        if (h->level[i] > 0) out[6] += 1;
        for (int l = 0; l <= h->level[i]; ++l) {
            const int *ll = links(h, (int)i, l);
            if (l == 0 && ll[0] > out[2]) out[2] = ll[0];
            if (l > 0 && ll[0] > out[3]) out[3] = ll[0];
            for (int j = 1; j <= ll[0]; ++j) {
                if (ll[j] == (int)i) out[4] += 1;
                if (ll[j] < 0 || ll[j] >= h->n || h->level[ll[j]] < l) out[5] += 1;
            }
        }
    }
}
