/*
 * gbdr_oracle.c — CPU restatement of the reference's search path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (libgbdr.so) never links or calls it.
 *
 * Every function restates one reference function in plain, strictly-IEEE C (compile with
 * -ffp-contract=off, no -ffast-math) and cites the reference file:line it follows (paths relative
 * to the reference repository root).  The restatement is pinned against the reference's own C++
 * compiled from /root/reference (oracle/ref_harness.cpp -> oracle/_ref/libgbdr_ref_strict.so) by
 * tests/test_oracle_vs_reference.py and against committed golden vectors in tests/golden/.
 *
 * Floating-point contract ("canonical arithmetic"): the source-level operation order of the
 * reference with every operation individually rounded to fp32 (no FMA contraction, no
 * reassociation).  That is what the reference compiles to with
 * `-O2 -fno-fast-math -ffp-contract=off`; the as-shipped `-Ofast` build may reassociate
 * and is compared with a tolerance instead.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PAD 0xFFFFFFFFu

/* ---------------------------------------------------------------- metrics */

/* L2Metric::Dist, search/support_func.h:107-128.
 * Four lane-strided partial sums over floor(d/4)*4 dims (the d%4 tail is ignored, :111-112),
 * each lane: sum = sum + (a-b)*(a-b) (:122-123), result ((T0+T1)+T2)+T3 (:126). */
float orc_l2(const float *a, const float *b, size_t d) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    size_t d4 = d >> 2;
    for (size_t c = 0; c < d4; ++c) {
        float e0 = a[4 * c + 0] - b[4 * c + 0];
        float e1 = a[4 * c + 1] - b[4 * c + 1];
        float e2 = a[4 * c + 2] - b[4 * c + 2];
        float e3 = a[4 * c + 3] - b[4 * c + 3];
        float p0 = e0 * e0, p1 = e1 * e1, p2 = e2 * e2, p3 = e3 * e3;
        s0 = s0 + p0;
        s1 = s1 + p1;
        s2 = s2 + p2;
        s3 = s3 + p3;
    }
    return ((s0 + s1) + s2) + s3;
}

/* Angular::Dist, search/support_func.h:131-163: NEGATED dot product.
 * 8 AVX lanes over floor(d/8) chunks (:136-141), fold high half onto low half (:143-144),
 * one optional 4-wide step (:146-151), zero-padded tail (:153-157), two hadds (:159-160). */
float orc_angular(const float *x, const float *y, size_t d) {
    float m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    while (d >= 8) {
        for (int i = 0; i < 8; ++i) {
            float p = x[i] * y[i];
            m[i] = m[i] + p;
        }
        x += 8;
        y += 8;
        d -= 8;
    }
    float s[4];
    for (int i = 0; i < 4; ++i) s[i] = m[4 + i] + m[i];
    if (d >= 4) {
        for (int i = 0; i < 4; ++i) {
            float p = x[i] * y[i];
            s[i] = s[i] + p;
        }
        x += 4;
        y += 4;
        d -= 4;
    }
    if (d > 0) {
        for (size_t i = 0; i < d; ++i) {
            float p = x[i] * y[i];
            s[i] = s[i] + p;
        }
        /* lanes >= d multiply 0*0 and add 0: value unchanged */
    }
    float h0 = s[0] + s[1];
    float h1 = s[2] + s[3];
    return -(h0 + h1);
}

/* ------------------------------------------------------------- projection */

/* computeNetLayer, search/support_func.h:624-633.  `out` must hold zeros on entry (:627). */
static void orc_net_layer(const float *layer, const float *in, float *out, int activation,
                          size_t step, size_t d_in, size_t d_out) {
    for (size_t i = 0; i < d_out; ++i) {
        out[i] = out[i] - orc_angular(layer + i * step, in, d_in);
        out[i] = out[i] + layer[i * step + step - 1];
        if (activation && out[i] < 0) out[i] = 0;
    }
}

/* GetLowQueryFromNet + normalizeVector, search/support_func.h:636-658.
 * l1 [dh x (d+1)], l2 [dh2 x (dh+1)], l3 [d_low x (dh2+1)]. */
void orc_project_one(const float *l1, const float *l2, const float *l3, const float *q, size_t d,
                     size_t dh, size_t dh2, size_t d_low, float *out, float *scratch) {
    float *h1 = scratch, *h2 = scratch + dh;
    memset(h1, 0, dh * sizeof(float));
    memset(h2, 0, dh2 * sizeof(float));
    memset(out, 0, d_low * sizeof(float));
    orc_net_layer(l1, q, h1, 1, d + 1, d, dh);
    orc_net_layer(l2, h1, h2, 1, dh + 1, dh, dh2);
    orc_net_layer(l3, h2, out, 0, dh2 + 1, dh2, d_low);
    /* normalizeVector :636-642 — L2Metric against a zero vector, so the d_low%4 tail does not
     * contribute to the norm */
    float norm = 0.f;
    {
        float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (size_t c = 0; c < (d_low >> 2); ++c) {
            float e0 = out[4 * c] - 0.f, e1 = out[4 * c + 1] - 0.f, e2 = out[4 * c + 2] - 0.f,
                  e3 = out[4 * c + 3] - 0.f;
            s0 = s0 + e0 * e0;
            s1 = s1 + e1 * e1;
            s2 = s2 + e2 * e2;
            s3 = s3 + e3 * e3;
        }
        norm = ((s0 + s1) + s2) + s3;
    }
    norm = sqrtf(norm);
    for (size_t i = 0; i < d_low; ++i) out[i] = out[i] / norm;
}

void orc_project(const float *l1, const float *l2, const float *l3, const float *queries,
                 size_t n_q, size_t d, size_t dh, size_t dh2, size_t d_low, float *out) {
    float *scratch = (float *)malloc((dh + dh2) * sizeof(float));
    for (size_t i = 0; i < n_q; ++i)
        orc_project_one(l1, l2, l3, queries + i * d, d, dh, dh2, d_low, out + i * d_low, scratch);
    free(scratch);
}

/* ------------------------------------------------------------ pair heaps */
/* std::priority_queue<std::pair<float,int>> (search/search_function.h:50,55): max-heap under
 * lexicographic (first, second) order.  Elements of one query are distinct pairs, so any
 * correct heap pops the same sequence as libstdc++'s. */
typedef struct {
    float f;
    int32_t i;
} orc_pair;
typedef struct {
    orc_pair *a;
    size_t n, cap;
} orc_heap;

static int pair_less(orc_pair x, orc_pair y) { return x.f < y.f || (!(y.f < x.f) && x.i < y.i); }
static void heap_init(orc_heap *h) {
    h->n = 0;
    h->cap = 64;
    h->a = (orc_pair *)malloc(h->cap * sizeof(orc_pair));
}
static void heap_push(orc_heap *h, float f, int32_t i) {
    if (h->n == h->cap) {
        h->cap *= 2;
        h->a = (orc_pair *)realloc(h->a, h->cap * sizeof(orc_pair));
    }
    size_t c = h->n++;
    orc_pair v = {f, i};
    while (c > 0) {
        size_t p = (c - 1) / 2;
        if (!pair_less(h->a[p], v)) break;
        h->a[c] = h->a[p];
        c = p;
    }
    h->a[c] = v;
}
static void heap_pop(orc_heap *h) {
    orc_pair v = h->a[--h->n];
    size_t c = 0;
    for (;;) {
        size_t l = 2 * c + 1, r = l + 1, m;
        if (l >= h->n) break;
        m = (r < h->n && pair_less(h->a[l], h->a[r])) ? r : l;
        if (!pair_less(v, h->a[m])) break;
        h->a[c] = h->a[m];
        c = m;
    }
    if (h->n) h->a[c] = v;
}

/* ------------------------------------------------------------ beam search */

/* getOneSearchResults + makeStep, search/search_function.h:15-102, single entry point.
 *
 *   adjacency: offsets[n+1] / edges (flattened vector<vector<uint32_t>>)
 *   aux_offsets/aux_edges: the auxiliary graph of use_second_graph == true (:73-80), or NULL
 *                      (use_second_graph == false, the only mode final_test.cpp uses, :85,:88);
 *                      llf and hops_bound as in :47,:73,:82
 *   out_ids/out_dists: the final heap (:96-100) sorted ascending by (dist, id), length k,
 *                      padded with ORC_PAD / +inf
 *   returns hops (:90); *dist_calc as in :52,:29; *scanned = adjacency ids looked at.
 */
typedef struct {
    const float *query, *db;
    uint32_t d;
    int ef;
    uint8_t *visited;
    uint32_t *touched;
    size_t n_touched, cap_touched;
    orc_heap top, cand;
    int dc, scan;
} orc_walk;

/* makeStep, :15-40; returns `found` */
static int orc_make_step(orc_walk *w, const uint32_t *nbrs, uint64_t deg) {
    int found = 0;
    for (uint64_t e = 0; e < deg; ++e) { /* :23 */
        uint32_t nb = nbrs[e];
        ++w->scan;
        if (w->visited[nb]) continue; /* :25 */
        w->visited[nb] = 1;           /* :26 */
        if (w->n_touched == w->cap_touched) {
            w->cap_touched *= 2;
            w->touched = (uint32_t *)realloc(w->touched, w->cap_touched * sizeof(uint32_t));
        }
        w->touched[w->n_touched++] = nb;
        float dn = orc_l2(w->query, w->db + (size_t)nb * w->d, w->d); /* :27-28 */
        ++w->dc;                                                       /* :29 */
        if (w->top.a[0].f > dn || (int)w->top.n < w->ef) {             /* :31 */
            heap_push(&w->cand, -dn, (int32_t)nb);                     /* :32 */
            found = 1;                                                 /* :33 */
            heap_push(&w->top, dn, (int32_t)nb);                       /* :34 */
            if ((int)w->top.n > w->ef) heap_pop(&w->top);              /* :35-36 */
        }
    }
    return found;
}

int orc_search_one_aux(const float *query, const float *db, uint64_t n, uint32_t d,
                       const uint64_t *offsets, const uint32_t *edges, const uint64_t *aux_offsets,
                       const uint32_t *aux_edges, int llf, uint32_t hops_bound, int ef, int k,
                       uint32_t entry, uint8_t *visited /* n bytes, zero on entry, zeroed on exit */,
                       uint32_t *out_ids, float *out_dists, int *dist_calc, int *scanned) {
    orc_walk w;
    w.query = query;
    w.db = db;
    w.d = d;
    w.ef = ef;
    w.visited = visited;
    heap_init(&w.top);
    heap_init(&w.cand);
    w.touched = (uint32_t *)malloc(1024 * sizeof(uint32_t));
    w.n_touched = 0;
    w.cap_touched = 1024;
    w.dc = 1; /* :52 */
    w.scan = 0;
    int hops = 0;
    (void)n;

    float dist = orc_l2(query, db + (size_t)entry * d, d); /* :56-57 */
    heap_push(&w.top, dist, (int32_t)entry);                /* :59 */
    heap_push(&w.cand, -dist, (int32_t)entry);              /* :60 */
    visited[entry] = 1;                                     /* :64 */
    w.touched[w.n_touched++] = entry;

    while (w.cand.n) {                                   /* :65 */
        orc_pair cur = w.cand.a[0];                      /* :66 */
        if (-cur.f > w.top.a[0].f) break;                /* :67 */
        heap_pop(&w.cand);                               /* :69 */
        uint32_t node = (uint32_t)cur.i;
        int aux_found = 0;
        if (aux_offsets && (uint32_t)hops < hops_bound) /* :73 */
            aux_found = orc_make_step(&w, aux_edges + aux_offsets[node],
                                      aux_offsets[node + 1] - aux_offsets[node]);
        if (!(aux_found && llf) || !aux_offsets) /* :82 */
            orc_make_step(&w, edges + offsets[node], offsets[node + 1] - offsets[node]);
        ++hops; /* :90 */
    }
    while ((int)w.top.n > k) heap_pop(&w.top); /* :96-98 */

    int m = (int)w.top.n;
    for (int j = 0; j < k; ++j) {
        out_ids[j] = ORC_PAD;
        if (out_dists) out_dists[j] = INFINITY;
    }
    for (int j = m - 1; j >= 0; --j) { /* pops come out worst first */
        out_ids[j] = (uint32_t)w.top.a[0].i;
        if (out_dists) out_dists[j] = w.top.a[0].f;
        heap_pop(&w.top);
    }
    for (size_t t = 0; t < w.n_touched; ++t) visited[w.touched[t]] = 0;
    free(w.touched);
    free(w.top.a);
    free(w.cand.a);
    if (dist_calc) *dist_calc = w.dc;
    if (scanned) *scanned = w.scan;
    return hops;
}

int orc_search_one(const float *query, const float *db, uint64_t n, uint32_t d,
                   const uint64_t *offsets, const uint32_t *edges, int ef, int k, uint32_t entry,
                   uint8_t *visited, uint32_t *out_ids, float *out_dists, int *dist_calc,
                   int *scanned) {
    return orc_search_one_aux(query, db, n, d, offsets, edges, NULL, NULL, 0, 0, ef, k, entry, visited,
                              out_ids, out_dists, dist_calc, scanned);
}

/* getRealNearest, search/search_function.h:105-125, extended from arg-min to top-k.
 * The reference walks the low-dim heap from its worst element (largest (dist,id)) to its best
 * and keeps the strictly smaller exact distance (:117), so on exact ties the candidate with the
 * WORSE low-dim rank wins.  cand_ids are the low-dim survivors in ascending (dist,id) order as
 * returned by orc_search_one (m valid entries).  Output: top-k by (exact dist asc, low-dim
 * rank desc).  out_ids[0] is the reference's return value. */
void orc_rerank_one(const float *query, const float *db, uint32_t d, const uint32_t *cand_ids,
                    int m, int k, uint32_t *out_ids, float *out_dists) {
    float *dist = (float *)malloc((size_t)(m > 0 ? m : 1) * sizeof(float));
    int *ord = (int *)malloc((size_t)(m > 0 ? m : 1) * sizeof(int));
    int mm = 0;
    for (int j = 0; j < m; ++j) {
        if (cand_ids[j] == ORC_PAD) break;
        dist[j] = orc_l2(db + (size_t)d * cand_ids[j], query, d); /* :110,:116 */
        ord[mm++] = j;
    }
    /* insertion sort by (dist asc, rank desc) */
    for (int a = 1; a < mm; ++a) {
        int v = ord[a], b = a - 1;
        while (b >= 0 && (dist[ord[b]] > dist[v] || (dist[ord[b]] == dist[v] && ord[b] < v))) {
            ord[b + 1] = ord[b];
            --b;
        }
        ord[b + 1] = v;
    }
    for (int j = 0; j < k; ++j) {
        if (j < mm) {
            out_ids[j] = cand_ids[ord[j]];
            if (out_dists) out_dists[j] = dist[ord[j]];
        } else {
            out_ids[j] = ORC_PAD;
            if (out_dists) out_dists[j] = INFINITY;
        }
    }
    free(dist);
    free(ord);
}

/* The query loop of performTest, search/search_function.h:153-186 (three branches), sequential.
 * mode 0: low-dim search beam=ef then re-rank to k in original dim (:158-164, recheck_size = ef)
 * mode 1: low-dim search only (:165-173)
 * mode 2: plain search in the original dim (:174-182)
 * dist_calc gets +ef in mode 0 (:164). */
void orc_search_batch_aux(const float *queries, const float *q_low, const float *db,
                      const float *db_low, uint64_t n, uint32_t d, uint32_t d_low,
                      const uint64_t *offsets, const uint32_t *edges, const uint64_t *aux_offsets,
                      const uint32_t *aux_edges, int llf, uint32_t hops_bound, uint32_t n_q, int ef, int k,
                      int mode, const uint32_t *entry, uint32_t *out_ids, float *out_dists,
                      int32_t *hops, int32_t *dist_calc, int32_t *scanned) {
    uint8_t *visited = (uint8_t *)calloc(n, 1);
    uint32_t *tmp_ids = (uint32_t *)malloc((size_t)ef * sizeof(uint32_t));
    float *tmp_d = (float *)malloc((size_t)ef * sizeof(float));
    for (uint32_t i = 0; i < n_q; ++i) {
        int dc = 0, sc = 0, h;
        if (mode == 0) {
            h = orc_search_one_aux(q_low + (size_t)i * d_low, db_low, n, d_low, offsets, edges, aux_offsets, aux_edges, llf, hops_bound, ef, ef,
                               entry[i], visited, tmp_ids, tmp_d, &dc, &sc);
            orc_rerank_one(queries + (size_t)i * d, db, d, tmp_ids, ef, k,
                           out_ids + (size_t)i * k, out_dists ? out_dists + (size_t)i * k : NULL);
            dc += ef;
        } else if (mode == 1) {
            h = orc_search_one_aux(q_low + (size_t)i * d_low, db_low, n, d_low, offsets, edges, aux_offsets, aux_edges, llf, hops_bound, ef, k,
                               entry[i], visited, out_ids + (size_t)i * k,
                               out_dists ? out_dists + (size_t)i * k : NULL, &dc, &sc);
        } else {
            h = orc_search_one_aux(queries + (size_t)i * d, db, n, d, offsets, edges, aux_offsets, aux_edges, llf, hops_bound, ef, k, entry[i],
                               visited, out_ids + (size_t)i * k,
                               out_dists ? out_dists + (size_t)i * k : NULL, &dc, &sc);
        }
        if (hops) hops[i] = h;
        if (dist_calc) dist_calc[i] = dc;
        if (scanned) scanned[i] = sc;
    }
    free(visited);
    free(tmp_ids);
    free(tmp_d);
}

void orc_search_batch(const float *queries, const float *q_low, const float *db,
                      const float *db_low, uint64_t n, uint32_t d, uint32_t d_low,
                      const uint64_t *offsets, const uint32_t *edges, uint32_t n_q, int ef, int k,
                      int mode, const uint32_t *entry, uint32_t *out_ids, float *out_dists,
                      int32_t *hops, int32_t *dist_calc, int32_t *scanned) {
    orc_search_batch_aux(queries, q_low, db, db_low, n, d, d_low, offsets, edges, NULL, NULL, 0, 0, n_q,
                         ef, k, mode, entry, out_ids, out_dists, hops, dist_calc, scanned);
}

/* --------------------------------------------------------------- kNN build */

typedef struct {
    float d;
    uint32_t id;
} orc_nd;
static int nd_cmp(const void *a, const void *b) {
    const orc_nd *x = (const orc_nd *)a, *y = (const orc_nd *)b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}

/* get_nearestneighbors(xq, xb, k) — dim_red/support_func.py:20-74 (faiss IndexFlatL2 or the torch
 * cdist2+topk fallback): the k smallest squared-L2 neighbours per query row, ascending, the row
 * itself included when xq is xb.  PARITY UNPINNED for tie order and last-ulp order: the reference
 * computes ||a||^2-2ab+||b||^2 in a BLAS-dependent order (support_func.py:44-47); this oracle
 * defines the result as canonical direct-difference distances (orc_l2) sorted by (dist, id).
 * tests/golden/knn_torch_*.npz pins SET agreement with the reference's torch path. */
void orc_knn(const float *Q, uint64_t n_q, const float *B, uint64_t n, uint32_t d, uint32_t k,
             uint32_t *out_ids, float *out_dists) {
    orc_nd *row = (orc_nd *)malloc((size_t)n * sizeof(orc_nd));
    for (uint64_t i = 0; i < n_q; ++i) {
        for (uint64_t j = 0; j < n; ++j) {
            row[j].d = orc_l2(Q + i * d, B + j * d, d);
            row[j].id = (uint32_t)j;
        }
        qsort(row, (size_t)n, sizeof(orc_nd), nd_cmp);
        for (uint32_t j = 0; j < k; ++j) {
            if (j < n) {
                out_ids[i * k + j] = row[j].id;
                if (out_dists) out_dists[i * k + j] = row[j].d;
            } else {
                out_ids[i * k + j] = ORC_PAD;
                if (out_dists) out_dists[i * k + j] = INFINITY;
            }
        }
    }
    free(row);
}

/* cutKNNbyK, search/support_func.h:309-340: the knn_size nearest entries of every list, ordered by distance to the
 * list's own vertex.  Nothing is dropped (the vertex itself stays, at distance 0); the reference std::sorts by dist
 * only, this restatement orders exact ties by id (as orc_gd_prune does).  out_edges capacity n*knn_size. */
void orc_knn_cut(const uint64_t *knn_offsets, const uint32_t *knn_edges, const float *ds, uint64_t n, uint32_t d,
                 uint32_t knn_size, uint64_t *out_offsets, uint32_t *out_edges) {
    uint64_t cap = 1;
    for (uint64_t i = 0; i < n; ++i)
        if (knn_offsets[i + 1] - knn_offsets[i] > cap) cap = knn_offsets[i + 1] - knn_offsets[i];
    orc_nd *row = (orc_nd *)malloc((size_t)cap * sizeof(orc_nd));
    out_offsets[0] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t len = knn_offsets[i + 1] - knn_offsets[i];
        for (uint64_t j = 0; j < len; ++j) {
            row[j].id = knn_edges[knn_offsets[i] + j];
            row[j].d = orc_l2(ds + i * d, ds + (uint64_t)row[j].id * d, d);
        }
        qsort(row, (size_t)len, sizeof(orc_nd), nd_cmp);
        const uint64_t keep = len < knn_size ? len : knn_size;
        for (uint64_t j = 0; j < keep; ++j) out_edges[out_offsets[i] + j] = row[j].id;
        out_offsets[i + 1] = out_offsets[i] + keep;
    }
    free(row);
}

/* ------------------------------------------------------------ GD pruning */

/* hnswlikeGD, search/support_func.h:521-575, + addReverseEdgesForGD :402-445,
 * + getConstantDegreeForGD :466-485.
 * Candidate order: the reference std::sorts Neighbor by dist only (:63-66,:540), which is
 * unstable at exact ties; this restatement orders by (dist, id).
 * out_offsets[n+1]; out_edges capacity n*2*M (+M/2 slack is never needed: forward lists are at
 * most M + M/2 <= 2M long, the reverse pass caps at 2M, :431).
 * getEps() = 1e-10 as float (:41-43). */
void orc_gd_prune(const uint64_t *knn_offsets, const uint32_t *knn_edges, const float *ds,
                  uint64_t n, uint32_t d, int M, int reverse, int need_const_degree,
                  uint64_t *out_offsets, uint32_t *out_edges) {
    const float eps = 1e-10f;
    const int edge = M / 2;
    const size_t cap = (size_t)2 * M;
    uint32_t *g = (uint32_t *)malloc((size_t)n * cap * sizeof(uint32_t));
    uint32_t *deg = (uint32_t *)calloc(n, sizeof(uint32_t));

    size_t maxdeg = 0;
    for (uint64_t i = 0; i < n; ++i) {
        size_t dgr = (size_t)(knn_offsets[i + 1] - knn_offsets[i]);
        if (dgr > maxdeg) maxdeg = dgr;
    }
    orc_nd *nb = (orc_nd *)malloc((maxdeg ? maxdeg : 1) * sizeof(orc_nd));
    for (uint64_t i = 0; i < n; ++i) {
        const float *pi = ds + i * d;
        size_t m = 0;
        for (uint64_t e = knn_offsets[i]; e < knn_offsets[i + 1]; ++e) { /* :532-539 */
            float di = orc_l2(pi, ds + (size_t)knn_edges[e] * d, d);
            if (di > eps) {
                nb[m].id = knn_edges[e];
                nb[m].d = di;
                ++m;
            }
        }
        qsort(nb, m, sizeof(orc_nd), nd_cmp); /* :540 */
        uint32_t *gi = g + i * cap;
        uint32_t dg = 0;
        if (m == 0) { /* reference reads neighbors[0] unchecked (:541): undefined; we emit nothing */
            deg[i] = 0;
            continue;
        }
        gi[dg++] = nb[0].id;                  /* :541 */
        for (size_t j = 1; j < m; ++j) {      /* :542 */
            const float *pp = ds + (size_t)nb[j].id * d;
            int good = 1;
            for (uint32_t l = 0; l < dg; ++l) { /* :545-551 */
                const float *pa = ds + (size_t)gi[l] * d;
                float lhs = orc_l2(pp, pi, d) + eps;
                if (lhs > orc_l2(pp, pa, d)) {
                    good = 0;
                    break;
                }
            }
            if (good) gi[dg++] = nb[j].id; /* :552-554 */
            if ((int)dg == M) break;        /* :555-557 */
        }
        for (int j = 0; j < edge && (size_t)j < m; ++j) { /* :559-563 (bounded by m here) */
            int found = 0;
            for (uint32_t l = 0; l < dg; ++l)
                if (gi[l] == nb[j].id) {
                    found = 1;
                    break;
                }
            if (!found) gi[dg++] = nb[j].id;
        }
        deg[i] = dg;
    }
    free(nb);

    if (reverse) { /* addReverseEdgesForGD :402-445, sequential, i ascending */
        uint32_t *indeg = (uint32_t *)calloc(n, sizeof(uint32_t));
        for (uint64_t i = 0; i < n; ++i) /* :418-422 (only sizes are used afterwards) */
            for (uint32_t j = 0; j < deg[i]; ++j) indeg[g[i * cap + j]]++;
        for (uint64_t i = 0; i < n; ++i) { /* :423-442 */
            int upper = M - (int)indeg[i];
            int thr = upper < M / 2 ? upper : M / 2;
            if (thr > 0) {
                for (uint32_t j = 0; j < deg[i]; ++j) { /* deg[i] re-read every iteration (:429) */
                    uint32_t c = g[i * cap + j];
                    if (deg[c] < (uint32_t)(2 * M)) {
                        int found = 0;
                        for (uint32_t l = 0; l < deg[c]; ++l)
                            if (g[(size_t)c * cap + l] == (uint32_t)i) {
                                found = 1;
                                break;
                            }
                        if (!found) {
                            g[(size_t)c * cap + deg[c]++] = (uint32_t)i;
                            if (--thr <= 0) break;
                        }
                    }
                }
            }
        }
        free(indeg);
    }

    if (need_const_degree) { /* getConstantDegreeForGD :466-485 */
        for (uint64_t i = 0; i < n; ++i) {
            if (deg[i] < (uint32_t)(2 * M)) {
                uint64_t b = knn_offsets[i], e = knn_offsets[i + 1];
                for (uint64_t j = b + 1; j < e; ++j) { /* j starts at 1 (:473) */
                    int found = 0;
                    for (uint32_t l = 0; l < deg[i]; ++l)
                        if (g[i * cap + l] == knn_edges[j]) {
                            found = 1;
                            break;
                        }
                    if (!found) {
                        g[i * cap + deg[i]++] = knn_edges[j];
                        if (deg[i] == (uint32_t)(2 * M)) break;
                    }
                }
            }
        }
    }

    out_offsets[0] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        memcpy(out_edges + out_offsets[i], g + i * cap, deg[i] * sizeof(uint32_t));
        out_offsets[i + 1] = out_offsets[i] + deg[i];
    }
    free(g);
    free(deg);
}
