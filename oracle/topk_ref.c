/* Oracle: exact cosine top-k in canonical fp32 arithmetic (plain C, CPU).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never linked into the product.
 *
 * What it restates.  The reference's exact search is FaissLatentVectorDatabase:
 *   - _l2_normalize (latice/index/faiss_db.py:109-113): row / ||row||, zero norm -> divide by 1;
 *   - add_vectors   (faiss_db.py:173-189): fp32 rows, normalised, appended in order;
 *   - query_similar (faiss_db.py:216-256): normalise the query, IndexFlatIP.search -> the k
 *     largest inner products, descending; k clamped to the row count.
 * The Chroma twin (latice/index/chroma_db.py:127-130, 231-259) asks hnswlib for the same thing
 * approximately, in the "cosine" space, and reports distance = 1 - inner product, ascending.
 * faiss-cpu 1.10.0 and chroma-hnswlib 0.7.6 (uv.lock) are third-party and not installable here,
 * and neither defines a summation order or a tie order, so the bit-exact contract is stated here:
 *
 *   n2    = numpy's float32 pairwise sum of the ROUNDED squares s_j = x_j*x_j (D = 16): r_j = s_j + s_{j+8},
 *           n2 = ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7))  -- what np.linalg.norm(axis=1) computes inside
 *           _l2_normalize; pinned BIT FOR BIT by tests/golden/l2_normalize.npz (the reference's own function);
 *           other D: plain ascending sum of rounded squares
 *   norm  = sqrtf(n2) (IEEE, round-to-nearest);  norm == 0  ->  1
 *   xhat_j = x_j / norm                                  (IEEE division)
 *   dot   = fma-chain  sum_j qhat_j*dhat_j, j ascending, starting from 0.0f
 *   order : dot descending, ties -> smaller GLOBAL row index first
 *   reported distance = 1.0f - dot   (monotone in dot, so the order is preserved)
 *
 * For unit vectors ||q||^2 - 2 q.d + ||d||^2 = 2 (1 - q.d): the squared-L2 form named in the
 * north star ranks identically; the kernel and this oracle rank on q.d directly.
 * NaN dots compare false against everything and are never selected.
 *
 * The normalisation leg is pinned to the reference (golden above).  The inner-product search leg stays
 * "parity unpinned" with respect to faiss/hnswlib themselves; it is anchored by the float64 brute-force and the
 * scikit-learn NearestNeighbors(metric="cosine", algorithm="brute") cross-checks in tests/test_oracle_topk.py.
 *
 * Build: gcc -O2 -mfma -ffp-contract=off -pthread -shared -fPIC   (oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <pthread.h>

static inline float dot_chain(const float *a, const float *b, int d) {
    float acc = 0.0f;
    for (int j = 0; j < d; ++j) acc = __builtin_fmaf(a[j], b[j], acc);
    return acc;
}

/* sum of squares exactly as numpy's float32 add.reduce over a contiguous row does it (pairwise_sum, n = 16 < 128:
 * eight running sums, combined as a balanced tree); volatile keeps every operation a separate fp32 rounding */
static float sumsq_numpy(const float *row, int d) {
    if (d == 16) {
        volatile float r[8];
        for (int j = 0; j < 8; ++j) {
            volatile float a = row[j] * row[j], b = row[j + 8] * row[j + 8];
            r[j] = a + b;
        }
        volatile float p0 = r[0] + r[1], p1 = r[2] + r[3], p2 = r[4] + r[5], p3 = r[6] + r[7];
        volatile float q0 = p0 + p1, q1 = p2 + p3;
        return q0 + q1;
    }
    volatile float acc = 0.0f;
    for (int j = 0; j < d; ++j) {
        volatile float s = row[j] * row[j];
        acc = acc + s;
    }
    return acc;
}

void ebsd_oracle_normalize_rows(float *x, int64_t n, int d) {
    for (int64_t i = 0; i < n; ++i) {
        float *row = x + i * d;
        float norm = sqrtf(sumsq_numpy(row, d));
        if (norm == 0.0f) norm = 1.0f;
        for (int j = 0; j < d; ++j) row[j] = row[j] / norm;
    }
}

/* candidate a beats b ? */
static inline int beats(float da, int64_t ia, float db, int64_t ib) {
    return (da > db) || (da == db && ia < ib);
}

/* one query against all rows */
static void topk_one(const float *dict, int64_t N, int64_t index_base, const float *q, int d, int k, float *bd,
                     int64_t *bi) {
    int filled = 0;
    for (int s = 0; s < k; ++s) {
        bd[s] = -INFINITY;
        bi[s] = -1;
    }
    for (int64_t r = 0; r < N; ++r) {
        float dot = dot_chain(q, dict + r * d, d);
        int64_t gi = index_base + r;
        if (filled == k && !beats(dot, gi, bd[k - 1], bi[k - 1])) continue;
        if (dot != dot) continue; /* NaN */
        int pos = filled < k ? filled : k - 1;
        while (pos > 0 && beats(dot, gi, bd[pos - 1], bi[pos - 1])) {
            bd[pos] = bd[pos - 1];
            bi[pos] = bi[pos - 1];
            --pos;
        }
        bd[pos] = dot;
        bi[pos] = gi;
        if (filled < k) ++filled;
    }
}

typedef struct {
    const float *dict, *queries;
    int64_t N, index_base, q_begin, q_end;
    int d, k;
    float *out_dot;
    int64_t *out_idx;
} topk_job;

static void *topk_worker(void *arg) {
    topk_job *j = (topk_job *)arg;
    for (int64_t qi = j->q_begin; qi < j->q_end; ++qi)
        topk_one(j->dict, j->N, j->index_base, j->queries + qi * j->d, j->d, j->k, j->out_dot + qi * j->k,
                 j->out_idx + qi * j->k);
    return 0;
}

/* dict: [N,d] normalised rows; queries: [Q,d] normalised; out_dot/out_idx: [Q,k].
 * Slots that cannot be filled (N < k) hold dot = -INFINITY, idx = -1.
 * nthreads <= 1: scalar, one core; otherwise queries are split over that many pthreads. */
void ebsd_oracle_topk(const float *dict, int64_t N, int64_t index_base, const float *queries, int64_t Q, int d,
                      int k, float *out_dot, int64_t *out_idx, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > Q) nthreads = Q > 0 ? (int)Q : 1;
    topk_job jobs[256];
    pthread_t tids[256];
    for (int t = 0; t < nthreads; ++t) {
        topk_job j = {dict, queries, N, index_base, Q * t / nthreads, Q * (t + 1) / nthreads, d, k, out_dot, out_idx};
        jobs[t] = j;
    }
    if (nthreads == 1) {
        topk_worker(&jobs[0]);
        return;
    }
    for (int t = 0; t < nthreads; ++t) pthread_create(&tids[t], 0, topk_worker, &jobs[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(tids[t], 0);
}

/* k-way merge of R partial lists [R,Q,k] (each sorted by the order above; empty slots idx = -1). */
void ebsd_oracle_topk_merge(const float *dots, const int64_t *idx, int R, int64_t Q, int k, float *out_dot,
                            int64_t *out_idx) {
    for (int64_t qi = 0; qi < Q; ++qi) {
        float *bd = out_dot + qi * k;
        int64_t *bi = out_idx + qi * k;
        int filled = 0;
        for (int s = 0; s < k; ++s) {
            bd[s] = -INFINITY;
            bi[s] = -1;
        }
        for (int r = 0; r < R; ++r) {
            for (int s = 0; s < k; ++s) {
                float dot = dots[((int64_t)r * Q + qi) * k + s];
                int64_t gi = idx[((int64_t)r * Q + qi) * k + s];
                if (gi < 0) continue;
                if (filled == k && !beats(dot, gi, bd[k - 1], bi[k - 1])) continue;
                int pos = filled < k ? filled : k - 1;
                while (pos > 0 && beats(dot, gi, bd[pos - 1], bi[pos - 1])) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = dot;
                bi[pos] = gi;
                if (filled < k) ++filled;
            }
        }
    }
}
