/* CPU oracle for hot path 2 (exhaustive L1 top-k).  TEST INFRASTRUCTURE - see oracle/__init__.py.
 *
 * Restates the algorithm the reference reaches through
 *     index = faiss.read_index(...); index.metric_type = faiss.METRIC_L1
 *     dm, im = index.search(que_arr, k)           (reference src/query_db.py:75-76,87;
 *                                                  bench/cathdb/run_dct.py:51-60)
 * on an index built by faiss.IndexFlatL2(d).add(int8 -> float32)
 *                                                 (reference src/database.py:240-243).
 * The arithmetic lives in the third-party dependency faiss 1.7.4 (reference env.yml:11,15,16),
 * which is not vendored under the reference tree and not installed here; this file restates
 * its published algorithm (IndexFlat::search -> knn_extra_metrics<VectorDistance<METRIC_L1>>):
 *   - the database is stored as float32;
 *   - parallel over queries (faiss: "#pragma omp parallel for"; here a pthread pool that hands
 *     out one query at a time, so the file builds with any C compiler in the image);
 *   - per query a size-k max-heap ordered by (distance, id), initialised to (FLT_MAX, -1);
 *   - database vectors are visited in id order; a vector replaces the heap top only when
 *     its distance is strictly smaller than the top's distance;
 *   - the heap is finally reordered ascending by (distance, id); unfilled slots keep
 *     (FLT_MAX, -1).
 * Net effect: the k smallest entries by (distance, id), ascending.  Pinned by the reference's
 * own fixture test/test/example-search.txt (tests/test_search_oracle.py).
 *
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -pthread -shared -fPIC).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

static inline int worse(float d1, int64_t i1, float d2, int64_t i2) {
    /* (d1,i1) orders after (d2,i2) */
    return d1 > d2 || (d1 == d2 && i1 > i2);
}

static void sift_down(float *hd, int64_t *hi, int k, int pos) {
    float d = hd[pos];
    int64_t id = hi[pos];
    for (;;) {
        int l = 2 * pos + 1, r = l + 1, big = pos;
        float bd = d;
        int64_t bi = id;
        if (l < k && worse(hd[l], hi[l], bd, bi)) { big = l; bd = hd[l]; bi = hi[l]; }
        if (r < k && worse(hd[r], hi[r], bd, bi)) { big = r; bd = hd[r]; bi = hi[r]; }
        if (big == pos) break;
        hd[pos] = hd[big];
        hi[pos] = hi[big];
        pos = big;
    }
    hd[pos] = d;
    hi[pos] = id;
}

/* sum |a - b| in float32 with 16 independent partial sums, the shape faiss' SIMD fvec_L1 has (the compiler turns the
 * inner loop into AVX2 code); for the integer-valued inputs of this path (|sum| <= 480 * 255 < 2^24) every partial sum
 * is exact, so the result does not depend on the summation order */
static float l1_f32(const float *a, const float *b, int d) {
    float acc[16] = {0};
    int i = 0;
    for (; i + 16 <= d; i += 16)
        for (int l = 0; l < 16; ++l) acc[l] += fabsf(a[i + l] - b[i + l]);
    float s = 0.f;
    for (int l = 0; l < 16; ++l) s += acc[l];
    for (; i < d; ++i) s += fabsf(a[i] - b[i]);
    return s;
}

typedef struct {
    const float *q, *db;
    int64_t nq, nb;
    int d, k;
    float *out_d;
    int64_t *out_i;
    int64_t next; /* next query to hand out */
    pthread_mutex_t mu;
} job_t;

static void one_query(const job_t *jb, int64_t qi, float *td, int64_t *ti) {
    const int k = jb->k, d = jb->d;
    float *hd = jb->out_d + qi * k;
    int64_t *hi = jb->out_i + qi * k;
    for (int j = 0; j < k; ++j) { hd[j] = FLT_MAX; hi[j] = -1; }
    const float *x = jb->q + qi * d;
    for (int64_t j = 0; j < jb->nb; ++j) {
        float dis = l1_f32(x, jb->db + j * d, d);
        if (dis < hd[0]) {
            hd[0] = dis;
            hi[0] = j;
            sift_down(hd, hi, k, 0);
        }
    }
    /* heap -> ascending (distance, id); unfilled slots (FLT_MAX, -1) go to the tail */
    int n = 0;
    for (int j = 0; j < k; ++j)
        if (hi[j] != -1) { td[n] = hd[j]; ti[n] = hi[j]; ++n; }
    for (int a = 1; a < n; ++a) { /* insertion sort, n <= k small */
        float dd = td[a];
        int64_t ii = ti[a];
        int b = a - 1;
        while (b >= 0 && worse(td[b], ti[b], dd, ii)) { td[b + 1] = td[b]; ti[b + 1] = ti[b]; --b; }
        td[b + 1] = dd;
        ti[b + 1] = ii;
    }
    for (int j = 0; j < n; ++j) { hd[j] = td[j]; hi[j] = ti[j]; }
    for (int j = n; j < k; ++j) { hd[j] = FLT_MAX; hi[j] = -1; }
}

static void *worker(void *arg) {
    job_t *jb = (job_t *)arg;
    float *td = (float *)malloc(sizeof(float) * (size_t)jb->k);
    int64_t *ti = (int64_t *)malloc(sizeof(int64_t) * (size_t)jb->k);
    for (;;) {
        pthread_mutex_lock(&jb->mu);
        int64_t qi = jb->next++;
        pthread_mutex_unlock(&jb->mu);
        if (qi >= jb->nq) break;
        one_query(jb, qi, td, ti);
    }
    free(td);
    free(ti);
    return NULL;
}

/* float32 queries [nq,d], float32 database [nb,d]  ->  dist float32 [nq,k], ids int64 [nq,k] */
int oracle_l1_topk_f32(const float *q, int64_t nq, const float *db, int64_t nb, int d, int k,
                       float *out_d, int64_t *out_i, int threads) {
    if (nq < 0 || nb < 0 || d <= 0 || k <= 0) return -1;
    job_t jb = {q, db, nq, nb, d, k, out_d, out_i, 0, PTHREAD_MUTEX_INITIALIZER};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((int64_t)threads > nq) threads = nq > 0 ? (int)nq : 1;
    pthread_t tid[256];
    for (int t = 1; t < threads; ++t) pthread_create(&tid[t], NULL, worker, &jb);
    worker(&jb);
    for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
    return 0;
}

/* convenience: int8 inputs converted to float32 exactly as faiss' python wrapper does
 * (np.ascontiguousarray(x, dtype='float32')) */
int oracle_l1_topk_i8(const int8_t *q, int64_t nq, const int8_t *db, int64_t nb, int d, int k,
                      float *out_d, int64_t *out_i, int threads) {
    float *qf = (float *)malloc(sizeof(float) * (size_t)(nq * d + 1));
    float *df = (float *)malloc(sizeof(float) * (size_t)(nb * d + 1));
    if (!qf || !df) { free(qf); free(df); return -2; }
    for (int64_t i = 0; i < nq * d; ++i) qf[i] = (float)q[i];
    for (int64_t i = 0; i < nb * d; ++i) df[i] = (float)db[i];
    int rc = oracle_l1_topk_f32(qf, nq, df, nb, d, k, out_d, out_i, threads);
    free(qf);
    free(df);
    return rc;
}
