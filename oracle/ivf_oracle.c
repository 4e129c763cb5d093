/*
 * TEST INFRASTRUCTURE ONLY -- the CPU oracle for the IVF-Flat hot path.
 *
 * A plain-C restatement of the reference's single-threaded CPU algorithm
 * (wedevxer/CUDA-AcceleratedVectorDatabaseEngine, engine/ivf_flat_index.cpp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's library; the product library
 * (libvdb_b200.so) never links, calls or falls back to it.
 *
 * Parity status: PINNED.  The reference's own tests hold no numeric golden
 * vectors, so this restatement is pinned against outputs of the
 * reference itself: oracle/_ref/libvdbref.so is the reference's unmodified
 * ivf_flat_index.cpp compiled in place (oracle/Makefile), and
 * tests/test_oracle.py requires bit-identical centroids, assignments, probe
 * lists, neighbour ids and distances between the two, plus against the
 * committed fixtures in tests/golden/ (generated from libvdbref.so by
 * tests/golden/make_golden.py).
 *
 * All arithmetic is IEEE fp32, strictly left-to-right, no FMA contraction
 * (build with -ffp-contract=off), exactly as the reference's g++ -O3 x86-64
 * object behaves.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { METRIC_L2 = 0, METRIC_IP = 1, METRIC_COSINE = 2 };

typedef struct {
    float* vec;
    uint64_t* ids;
    size_t count, cap;
} olist;

typedef struct {
    uint32_t dim, nlist;
    int metric;
    float* centroids; /* [nlist][dim] */
    olist* lists;
    uint64_t total;
} oindex;

/* ---- libstdc++ random machinery the reference's train() depends on ------ */

/* std::mt19937 (ISO C++ [rand.eng.mers]; MT19937 of Matsumoto & Nishimura) */
typedef struct {
    uint32_t mt[624];
    int idx;
} mt19937;

static void mt_seed(mt19937* g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

static uint32_t mt_next(mt19937* g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* std::uniform_int_distribution<uint64_t>(0, n-1)(mt19937) as libstdc++ 13
 * implements it for a 32-bit engine and a range below 2^32: Lemire's nearly
 * divisionless method (bits/uniform_int_dist.h, _S_nd).  Reference call site:
 * ivf_flat_index.cpp:54,57. */
static uint64_t uniform_u64_below(mt19937* g, uint64_t n) {
    if (n - 1 >= 0xffffffffull) abort(); /* not restated: > 2^32 training vectors */
    uint32_t range = (uint32_t)n;
    uint64_t product = (uint64_t)mt_next(g) * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
        uint32_t threshold = (uint32_t)(-range) % range;
        while (low < threshold) {
            product = (uint64_t)mt_next(g) * (uint64_t)range;
            low = (uint32_t)product;
        }
    }
    return product >> 32;
}

/* std::uniform_real_distribution<float>(0.0f, b)(mt19937) in libstdc++ 13:
 * generate_canonical<float,24> draws one 32-bit word, converts it to float
 * (round-to-nearest), divides by 2^32, clamps 1.0 to nextafter(1,0); the
 * distribution returns canonical * (b - a) + a.  Call site: :91-92. */
static float uniform_real_0_b(mt19937* g, float b) {
    volatile float sum = (float)mt_next(g);
    volatile float ret = sum / 4294967296.0f;
    if (ret >= 1.0f) ret = nextafterf(1.0f, 0.0f);
    volatile float r = ret * (b - 0.0f);
    return r + 0.0f;
}

/* ---- distance definitions (kernels.cuh:36-80, CPU loops :275-285) ------- */

static float dist_l2(const float* a, const float* b, uint32_t dim) {
    float dist = 0.0f;
    for (uint32_t d = 0; d < dim; ++d) {
        float diff = a[d] - b[d];
        dist += diff * diff;
    }
    return dist;
}

static float dist_ip(const float* a, const float* b, uint32_t dim) {
    float dist = 0.0f;
    for (uint32_t d = 0; d < dim; ++d) dist += a[d] * b[d];
    return -dist;
}

/* The CPU path has branches for L2 and InnerProduct only; any other metric
 * leaves dist at 0.0f (ivf_flat_index.cpp:275-285,308-318,352-362). */
static float dist_metric(int metric, const float* a, const float* b, uint32_t dim) {
    if (metric == METRIC_L2) return dist_l2(a, b, dim);
    if (metric == METRIC_IP) return dist_ip(a, b, dim);
    return 0.0f;
}

/* ---- index object -------------------------------------------------------- */

oindex* oracle_create(uint32_t dim, uint32_t nlist, int metric) {
    if (dim == 0 || nlist == 0) return NULL; /* ctor throws, :17-19 */
    oindex* ix = (oindex*)calloc(1, sizeof(oindex));
    ix->dim = dim;
    ix->nlist = nlist;
    ix->metric = metric;
    ix->centroids = (float*)calloc((size_t)nlist * dim, sizeof(float));
    ix->lists = (olist*)calloc(nlist, sizeof(olist));
    return ix;
}

void oracle_destroy(oindex* ix) {
    if (!ix) return;
    for (uint32_t l = 0; l < ix->nlist; ++l) {
        free(ix->lists[l].vec);
        free(ix->lists[l].ids);
    }
    free(ix->lists);
    free(ix->centroids);
    free(ix);
}

/* assign_to_lists, ivf_flat_index.cpp:259-295: argmin with strict '<', so the
 * lowest centroid index wins ties; honours the index metric. */
void oracle_assign(const oindex* ix, const float* x, uint64_t n, uint32_t* out) {
    for (uint64_t v = 0; v < n; ++v) {
        const float* vec = x + v * ix->dim;
        float min_dist = FLT_MAX;
        uint32_t best = 0;
        for (uint32_t c = 0; c < ix->nlist; ++c) {
            float d = dist_metric(ix->metric, vec, ix->centroids + (size_t)c * ix->dim, ix->dim);
            if (d < min_dist) {
                min_dist = d;
                best = c;
            }
        }
        out[v] = best;
    }
}

/* train, ivf_flat_index.cpp:49-145.  The reference recomputes, for every new
 * seed, the minimum over ALL earlier centroids (:73-84, O(nlist^2 n D)); this
 * restatement keeps the running minimum instead.  Each distance is produced by
 * the same expression and min() is exact, so the result is bit-identical
 * (checked against libvdbref.so in tests/test_oracle.py). */
void oracle_train(oindex* ix, const float* x, uint64_t n) {
    const uint32_t D = ix->dim, K = ix->nlist;
    mt19937 gen;
    mt_seed(&gen, 42); /* :53 */
    uint64_t first = uniform_u64_below(&gen, n); /* :54-57 */
    memcpy(ix->centroids, x + first * D, D * sizeof(float));

    float* mind = (float*)malloc(n * sizeof(float));
    for (uint64_t v = 0; v < n; ++v) mind[v] = FLT_MAX;
    for (uint32_t c = 1; c < K; ++c) { /* :63 */
        const float* newest = ix->centroids + (size_t)(c - 1) * D;
        float total = 0.0f;
        for (uint64_t v = 0; v < n; ++v) {
            float d = dist_l2(x + v * D, newest, D); /* seeding is always L2, :77-81 */
            if (d < mind[v]) mind[v] = d;            /* std::min(min_dist, dist), :83 */
            total += mind[v];                        /* :87 */
        }
        float target = uniform_real_0_b(&gen, total); /* :91-92 */
        float cumsum = 0.0f;
        for (uint64_t v = 0; v < n; ++v) { /* :95-103 */
            cumsum += mind[v];
            if (cumsum >= target) {
                memcpy(ix->centroids + (size_t)c * D, x + v * D, D * sizeof(float));
                break;
            }
        }
        /* if no v satisfies cumsum >= target the slot keeps its zero fill,
         * exactly like the reference (centroids_ is value-initialised, :22). */
    }
    free(mind);

    uint32_t* assign = (uint32_t*)malloc(n * sizeof(uint32_t));
    float* sums = (float*)malloc((size_t)K * D * sizeof(float));
    uint32_t* counts = (uint32_t*)malloc(K * sizeof(uint32_t));
    for (int iter = 0; iter < 10; ++iter) { /* :109 */
        oracle_assign(ix, x, n, assign);
        memset(sums, 0, (size_t)K * D * sizeof(float));
        memset(counts, 0, K * sizeof(uint32_t));
        for (uint64_t v = 0; v < n; ++v) { /* :123-131, input order */
            float* s = sums + (size_t)assign[v] * D;
            const float* vec = x + v * D;
            for (uint32_t d = 0; d < D; ++d) s[d] += vec[d];
            counts[assign[v]]++;
        }
        for (uint32_t c = 0; c < K; ++c) /* :134-141: empty cluster keeps its centroid */
            if (counts[c] > 0)
                for (uint32_t d = 0; d < D; ++d)
                    ix->centroids[(size_t)c * D + d] = sums[(size_t)c * D + d] / counts[c];
    }
    free(assign);
    free(sums);
    free(counts);
}

/* add, ivf_flat_index.cpp:148-202: assign, then append rows and ids to their
 * list in input order. */
void oracle_add(oindex* ix, const float* x, const uint64_t* ids, uint64_t n) {
    uint32_t* assign = (uint32_t*)malloc(n * sizeof(uint32_t));
    oracle_assign(ix, x, n, assign);
    for (uint64_t v = 0; v < n; ++v) {
        olist* l = &ix->lists[assign[v]];
        if (l->count == l->cap) {
            l->cap = l->cap ? l->cap * 2 : 16;
            l->vec = (float*)realloc(l->vec, l->cap * ix->dim * sizeof(float));
            l->ids = (uint64_t*)realloc(l->ids, l->cap * sizeof(uint64_t));
        }
        memcpy(l->vec + l->count * ix->dim, x + v * ix->dim, ix->dim * sizeof(float));
        l->ids[l->count++] = ids[v];
    }
    ix->total += n;
    free(assign);
}

/* bench set-up helper: add() with the assignment step supplied by the caller */
void oracle_load_assigned(oindex* ix, const float* x, const uint64_t* ids, const uint32_t* assign, uint64_t n) {
    for (uint64_t v = 0; v < n; ++v) {
        olist* l = &ix->lists[assign[v]];
        if (l->count == l->cap) {
            l->cap = l->cap ? l->cap * 2 : 16;
            l->vec = (float*)realloc(l->vec, l->cap * ix->dim * sizeof(float));
            l->ids = (uint64_t*)realloc(l->ids, l->cap * sizeof(uint64_t));
        }
        memcpy(l->vec + l->count * ix->dim, x + v * ix->dim, ix->dim * sizeof(float));
        l->ids[l->count++] = ids[v];
    }
    ix->total += n;
}

typedef struct {
    float d;
    uint64_t id;
} cand;

/* std::pair<float,uint64_t>::operator< : (dist, id) ascending */
static int cand_cmp(const void* a, const void* b) {
    const cand* x = (const cand*)a;
    const cand* y = (const cand*)b;
    if (x->d < y->d) return -1;
    if (y->d < x->d) return 1;
    if (x->id < y->id) return -1;
    if (y->id < x->id) return 1;
    return 0;
}

/* select_nprobe_lists, :298-336: first min(nprobe,nlist) of (dist, list id). */
uint32_t oracle_select_nprobe(const oindex* ix, const float* q, uint32_t nprobe, uint32_t* out) {
    cand* cd = (cand*)malloc(ix->nlist * sizeof(cand));
    for (uint32_t c = 0; c < ix->nlist; ++c) {
        cd[c].d = dist_metric(ix->metric, q, ix->centroids + (size_t)c * ix->dim, ix->dim);
        cd[c].id = c;
    }
    qsort(cd, ix->nlist, sizeof(cand), cand_cmp);
    uint32_t np = nprobe < ix->nlist ? nprobe : ix->nlist;
    for (uint32_t p = 0; p < np; ++p) out[p] = (uint32_t)cd[p].id;
    free(cd);
    return np;
}

/* One query: select_nprobe_lists -> search_list_cpu per non-empty probed list
 * (:339-384, top min(k,count) by (dist,id)) -> merge_results (:474-518: sort
 * all, drop UINT64_MAX ids, keep the first occurrence of each id, pad with
 * FLT_MAX / UINT64_MAX).  nprobe is clamped to nlist (the reference reads past
 * probe_lists otherwise, :221-222); buffers are per query, so the stale-buffer
 * quirk of the batched reference loop (:210-211,225) is not reproduced. */
static void search_one(const oindex* ix, const float* q, uint32_t nprobe, uint32_t k,
                       float* D, uint64_t* I) {
    uint32_t* probes = (uint32_t*)malloc((nprobe ? nprobe : 1) * sizeof(uint32_t));
    uint32_t np = oracle_select_nprobe(ix, q, nprobe, probes);
    cand* all = (cand*)malloc(((size_t)np * k + 1) * sizeof(cand));
    size_t nall = 0;
    for (uint32_t p = 0; p < np; ++p) {
        const olist* l = &ix->lists[probes[p]];
        if (l->count == 0) continue;
        cand* c = (cand*)malloc(l->count * sizeof(cand));
        for (size_t i = 0; i < l->count; ++i) {
            c[i].d = dist_metric(ix->metric, q, l->vec + i * ix->dim, ix->dim);
            c[i].id = l->ids[i];
        }
        qsort(c, l->count, sizeof(cand), cand_cmp);
        size_t take = l->count < k ? l->count : k;
        for (size_t i = 0; i < take; ++i)
            if (c[i].id != UINT64_MAX) all[nall++] = c[i];
        free(c);
    }
    qsort(all, nall, sizeof(cand), cand_cmp);
    uint32_t out = 0;
    for (size_t i = 0; i < nall && out < k; ++i) {
        int dup = 0;
        for (uint32_t j = 0; j < out; ++j)
            if (I[j] == all[i].id) {
                dup = 1;
                break;
            }
        if (dup) continue;
        D[out] = all[i].d;
        I[out] = all[i].id;
        ++out;
    }
    for (; out < k; ++out) {
        D[out] = FLT_MAX;
        I[out] = UINT64_MAX;
    }
    free(all);
    free(probes);
}

typedef struct {
    const oindex* ix;
    const float* q;
    uint32_t nq, nprobe, k;
    float* D;
    uint64_t* I;
    int tid, nthreads;
} sjob;

static void* search_worker(void* arg) {
    sjob* j = (sjob*)arg;
    for (uint32_t i = (uint32_t)j->tid; i < j->nq; i += (uint32_t)j->nthreads)
        search_one(j->ix, j->q + (size_t)i * j->ix->dim, j->nprobe, j->k,
                   j->D + (size_t)i * j->k, j->I + (size_t)i * j->k);
    return NULL;
}

/* search, :205-256.  The reference is single-threaded; nthreads > 1 runs
 * independent queries on several host threads for the all-cores baseline. */
void oracle_search(const oindex* ix, const float* q, uint32_t nq, uint32_t nprobe, uint32_t k,
                   float* D, uint64_t* I, int nthreads) {
    if (nthreads <= 1) {
        sjob j = {ix, q, nq, nprobe, k, D, I, 0, 1};
        search_worker(&j);
        return;
    }
    pthread_t* th = (pthread_t*)malloc(nthreads * sizeof(pthread_t));
    sjob* jobs = (sjob*)malloc(nthreads * sizeof(sjob));
    for (int t = 0; t < nthreads; ++t) {
        sjob j = {ix, q, nq, nprobe, k, D, I, t, nthreads};
        jobs[t] = j;
        pthread_create(&th[t], NULL, search_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

void oracle_get_centroids(const oindex* ix, float* out) {
    memcpy(out, ix->centroids, (size_t)ix->nlist * ix->dim * sizeof(float));
}

void oracle_set_centroids(oindex* ix, const float* in) {
    memcpy(ix->centroids, in, (size_t)ix->nlist * ix->dim * sizeof(float));
}

void oracle_list_sizes(const oindex* ix, uint64_t* out) {
    for (uint32_t l = 0; l < ix->nlist; ++l) out[l] = ix->lists[l].count;
}

void oracle_list_ids(const oindex* ix, uint32_t list, uint64_t* out) {
    memcpy(out, ix->lists[list].ids, ix->lists[list].count * sizeof(uint64_t));
}

uint64_t oracle_total_vectors(const oindex* ix) { return ix->total; }

/* Exact flat search (one list holding everything): the restatement of the
 * reference run with nlist=1, nprobe=1, used for the brute-force config. */
void oracle_flat_search(const float* db, const uint64_t* ids, uint64_t n, uint32_t dim, int metric,
                        const float* q, uint32_t nq, uint32_t k, float* D, uint64_t* I) {
    cand* c = (cand*)malloc(n * sizeof(cand));
    for (uint32_t qi = 0; qi < nq; ++qi) {
        for (uint64_t i = 0; i < n; ++i) {
            c[i].d = dist_metric(metric, q + (size_t)qi * dim, db + i * dim, dim);
            c[i].id = ids ? ids[i] : i;
        }
        qsort(c, n, sizeof(cand), cand_cmp);
        for (uint32_t j = 0; j < k; ++j) {
            D[(size_t)qi * k + j] = j < n ? c[j].d : FLT_MAX;
            I[(size_t)qi * k + j] = j < n ? c[j].id : UINT64_MAX;
        }
    }
    free(c);
}
