/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's per-target topological feature path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this file's shared object.  The product (tlc-gnn_b200/) never links, imports or calls it.
 *
 * Parity pinning: validated in this container against the *real* reference imported in place from
 * /root/reference (tests/test_oracle_vs_reference.py, skipped where the tree is absent) and against
 * the committed fixtures in tests/golden/ which were produced by the real reference
 * (oracle/make_golden.py).  The only golden vector in the reference's own tree,
 * Knowledge_Distillation/pimg.py:450-505, is checked in tests/test_oracle_golden.py.
 *
 * Every function cites the reference lines it restates (paths relative to pkuyzy/TLC-GNN).
 *
 * Canonical order: the reference's simplex tie order is whatever networkx's sub-graph
 * view iterates; value multisets are invariant under it.  This restatement fixes ONE order: vicinity
 * vertices ascending by graph id (local id = rank), edges lexicographic (lo,hi) in local ids and
 * oriented (lo,hi); python's stable sort then makes that the tie-break.  It equals what the reference
 * itself produces when handed a graph whose dict order is that order (ref_harness.run_one_stages).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define TLO_MODE_EDGE 0 /* riccidist2dgm.py:311-316 : ball(u) & ball(v)            */
#define TLO_MODE_NODE 1 /* Knowledge_Distillation/data_utils_NC.py:97-100 : ball(u) */
#define TLO_MODE_EDGE_FORCED 2 /* Knowledge_Distillation/data_utils_LP.py:107-112 : (ball(u) & ball(v)) + [u] + [v] */
#define TLO_MODE_EDGE_UNION 3 /* sg2pimg range == 'union' riccidist2dgm.py:242-247 : nodes_u + nodes_v */
#define TLO_MODE_EDGE_REMOVEINTER 4 /* range == 'removeinter' :289-296 : list(set(nodes_union) - nodes_intersec) + [u, v] */

#define TLO_DESC_MIN 0
#define TLO_DESC_MAX 1
#define TLO_DESC_SUM 2

#define TLO_F_NORM 1u         /* build_fv(norm=True)                      riccidist2dgm.py:50-56        */
#define TLO_F_EXTENDED 2u     /* Accelerate_PD                            riccidist2dgm.py:323-326      */
#define TLO_F_KEEP_ZERO 4u    /* KD copy emits zero-persistence pairs     KD/accelerated_PD.py:68-69... */
#define TLO_F_NORM_EPS 8u     /* KD: /(max + 1e-10)                       data_utils_NC.py:54           */
#define TLO_F_SUM_PLAIN 16u   /* python<=3.11 sum(): plain left-to-right; default = 3.12 Neumaier (F5)  */
#define TLO_F_ASIS_FV 32u     /* build_fv as written: one Dijkstra per (vertex, root)  :27-37           */
#define TLO_F_FILT_DEGREE 512u      /* KD filt='degree'    : induced degree / (max + 1e-10)   data_utils_NC.py:126-128 */
#define TLO_F_FILT_CENTRALITY 1024u /* KD filt='centrality': nx.degree_centrality / (max + 1e-10)     :118-121 */
#define TLO_F_FILT_CLUSTERING 2048u /* KD filt='clustering': nx.clustering / (max + 1e-10)            :122-125 */

/* pair kinds, in the order the reference concatenates them (accelerated_PD.py:110, riccidist2dgm.py:328) */
#define TLO_K_UP 0
#define TLO_K_ESS 1
#define TLO_K_DOWN 2
#define TLO_K_ESS_REV 3
#define TLO_K_ONE 4

/* per-target outcome classes, SURVEY.md A.8 (riccidist2dgm.py:318,352-357,54-56; accelerated_PD.py:122) */
#define TLO_ST_OK 0
#define TLO_ST_TRIVIAL 1
#define TLO_ST_EMPTY 2
#define TLO_ST_DISCONNECTED 3
#define TLO_ST_DEGENERATE 4
#define TLO_ST_UNKNOWN_NODE 5
#define TLO_ST_BAD_DESCRIPTOR 6
#define TLO_ST_NO_TREE_EDGES 7

typedef struct {
  int32_t N;
  const int32_t *rowptr; /* [N+1] */
  const int32_t *col;    /* [2M] ascending inside each row */
  const double *kappa;   /* [2M] curvature per directed edge; weight = kappa + 1 (riccidist2dgm.py:225) */
} tlo_graph;

typedef struct {
  int32_t hop, mode, descriptor, resolution;
  uint32_t flags;
  uint32_t img_mask; /* bit k set: pairs of kind k are rasterised */
} tlo_params;

/* optional per-target detail; every pointer may be NULL.  Capacities are the caller's business
 * (n <= N, m <= M, pairs <= n + m + 2). */
typedef struct {
  int32_t n, m, npairs, npos, nneg, lu, lv;
  int32_t *vert;      /* [n] graph ids, ascending */
  int32_t *elo, *ehi; /* [m] local ids, lexicographic */
  double *ew;         /* [m] kappa+1 */
  double *d1, *d2;    /* [n] raw distances (before min/max/sum, before norm) */
  double *fval;       /* [n] chosen descriptor after normalisation */
  int32_t *ord_asc, *ord_desc; /* [m] edge index in sweep order */
  int32_t *pkind, *pbv, *pdv;  /* [npairs] kind, birth vertex, death vertex (local ids) */
  double *pbirth, *pdeath;     /* [npairs] */
  int32_t *pos, *neg;          /* edge indices, sweep order (accelerated_PD.py:99,109) */
} tlo_detail;

/* ------------------------------------------------------------------------------------------ */
/* scratch per worker thread                                                                    */
typedef struct {
  int32_t N;
  int32_t *mark_u, *mark_v; /* [N] BFS depth+1 or 0 */
  int32_t *queue;           /* [N] */
  int32_t *lid;             /* [N] local id or -1 */
  /* grown on demand */
  int32_t ncap, mcap;
  int32_t *vert, *elo, *ehi, *ord_asc, *ord_desc, *uf, *heap_pos, *heap, *par, *rowl, *adj, *adje;
  int32_t *pos, *neg, *tpar, *tpe, *stamp, *bq;
  double *ew, *d1, *d2, *dtmp, *fval, *kasc, *kdesc, *pw;
  /* pairs */
  int32_t pcap;
  int32_t *pkind, *pbv, *pdv;
  double *pbirth, *pdeath;
} tlo_ws;

static void *xrealloc(void *p, size_t sz) {
  void *q = realloc(p, sz ? sz : 1);
  if (!q) abort();
  return q;
}

static tlo_ws *ws_new(int32_t N) {
  tlo_ws *w = (tlo_ws *)calloc(1, sizeof(tlo_ws));
  w->N = N;
  w->mark_u = (int32_t *)calloc((size_t)N + 1, 4);
  w->mark_v = (int32_t *)calloc((size_t)N + 1, 4);
  w->queue = (int32_t *)malloc(((size_t)N + 1) * 4);
  w->lid = (int32_t *)malloc(((size_t)N + 1) * 4);
  for (int32_t i = 0; i < N; i++) w->lid[i] = -1;
  return w;
}
static void ws_reserve(tlo_ws *w, int32_t n, int32_t m) {
  if (n > w->ncap) {
    int32_t c = n * 2 + 16;
    w->ncap = c;
#define GROW(f, T, k) w->f = (T *)xrealloc(w->f, (size_t)(k) * sizeof(T))
    GROW(vert, int32_t, c); GROW(uf, int32_t, c); GROW(heap_pos, int32_t, c); GROW(heap, int32_t, c);
    GROW(par, int32_t, c); GROW(rowl, int32_t, c + 1); GROW(tpar, int32_t, c); GROW(tpe, int32_t, c);
    GROW(stamp, int32_t, c); GROW(bq, int32_t, c);
    GROW(d1, double, c); GROW(d2, double, c); GROW(dtmp, double, c); GROW(fval, double, c); GROW(pw, double, c);
  }
  if (m > w->mcap) {
    int32_t c = m * 2 + 16;
    w->mcap = c;
    GROW(elo, int32_t, c); GROW(ehi, int32_t, c); GROW(ord_asc, int32_t, c); GROW(ord_desc, int32_t, c);
    GROW(adj, int32_t, 2 * (size_t)c); GROW(adje, int32_t, 2 * (size_t)c); GROW(pos, int32_t, c); GROW(neg, int32_t, c);
    GROW(ew, double, c); GROW(kasc, double, c); GROW(kdesc, double, c);
  }
  if (n + m + 4 > w->pcap) {
    int32_t c = (n + m) * 2 + 16;
    w->pcap = c;
    GROW(pkind, int32_t, c); GROW(pbv, int32_t, c); GROW(pdv, int32_t, c);
    GROW(pbirth, double, c); GROW(pdeath, double, c);
  }
}
static void ws_free(tlo_ws *w) {
  if (!w) return;
  free(w->mark_u); free(w->mark_v); free(w->queue); free(w->lid);
  free(w->vert); free(w->elo); free(w->ehi); free(w->ord_asc); free(w->ord_desc); free(w->uf);
  free(w->heap_pos); free(w->heap); free(w->par); free(w->rowl); free(w->adj); free(w->adje);
  free(w->pos); free(w->neg); free(w->tpar); free(w->tpe); free(w->stamp); free(w->bq);
  free(w->ew); free(w->d1); free(w->d2); free(w->dtmp); free(w->fval); free(w->kasc); free(w->kdesc); free(w->pw);
  free(w->pkind); free(w->pbv); free(w->pdv); free(w->pbirth); free(w->pdeath);
  free(w);
}

/* ------------------------------------------------------------------------------------------ */
/* closed ball of radius hop: [root] + targets of nx.bfs_edges(G, root, depth_limit=hop)
 * riccidist2dgm.py:311-314.  mark[x] = depth+1.  returns count, members in w->queue[0..cnt). */
static int32_t ball(const tlo_graph *g, int32_t root, int32_t hop, int32_t *mark, int32_t *queue) {
  int32_t head = 0, tail = 0;
  mark[root] = 1;
  queue[tail++] = root;
  while (head < tail) {
    int32_t x = queue[head++];
    int32_t d = mark[x];
    if (d - 1 >= hop) continue;
    for (int32_t e = g->rowptr[x]; e < g->rowptr[x + 1]; e++) {
      int32_t y = g->col[e];
      if (!mark[y]) { mark[y] = d + 1; queue[tail++] = y; }
    }
  }
  return tail;
}

static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

/* python's sum() over a list of floats, reference build_fv `sum([...])` riccidist2dgm.py:30,35.
 * CPython >= 3.12 (this container: 3.12.3) uses Neumaier compensation after the first float;
 * CPython <= 3.11 (the reference pins 3.7) adds left to right.  SURVEY.md F5. */
typedef struct { double s, c; int k; } pysum;
static inline void pysum_init(pysum *p) { p->s = 0.0; p->c = 0.0; p->k = 0; }
static inline void pysum_add(pysum *p, double x, int plain) {
  if (p->k == 0) { p->s = x; p->k = 1; return; } /* int 0 + float -> float, exact */
  if (plain) { p->s += x; return; }
  double t = p->s + x;
  if (fabs(p->s) >= fabs(x)) p->c += (p->s - t) + x; else p->c += (x - t) + p->s;
  p->s = t;
}
static inline double pysum_get(const pysum *p, int plain) {
  if (p->k == 0) return 0.0;
  if (!plain && p->c != 0.0 && isfinite(p->c)) return p->s + p->c;
  return p->s;
}

/* binary heap keyed by (dist, insertion counter) == networkx's (vu_dist, next(c), u) tuples */
typedef struct { double d; int64_t c; int32_t v; } hitem;
static int hless(const hitem *a, const hitem *b) { return a->d < b->d || (a->d == b->d && a->c < b->c); }
typedef struct { hitem *a; int32_t n, cap; } heap_t;
static void hpush(heap_t *h, hitem it) {
  if (h->n == h->cap) { h->cap = h->cap * 2 + 64; h->a = (hitem *)xrealloc(h->a, (size_t)h->cap * sizeof(hitem)); }
  int32_t i = h->n++;
  while (i > 0) { int32_t p = (i - 1) >> 1; if (!hless(&it, &h->a[p])) break; h->a[i] = h->a[p]; i = p; }
  h->a[i] = it;
}
static hitem hpop(heap_t *h) {
  hitem top = h->a[0], last = h->a[--h->n];
  int32_t i = 0;
  for (;;) {
    int32_t l = 2 * i + 1, r = l + 1, s;
    if (l >= h->n) break;
    s = (r < h->n && hless(&h->a[r], &h->a[l])) ? r : l;
    if (!hless(&h->a[s], &last)) break;
    h->a[i] = h->a[s]; i = s;
  }
  if (h->n > 0) h->a[i] = last;
  return top;
}

/* nx.dijkstra_path(S, src, dst, weight='weight') restated after networkx _dijkstra_multisource
 * (weighted.py): lazy-deletion heap of (dist, counter, node), strict-improvement relaxations,
 * stop when dst is popped.  pred[] gives the path; neighbours are visited in ascending local id.
 * Then `sum([kappa+1 ...])` along src -> dst in that order.  riccidist2dgm.py:29-30. */
static double dijkstra_path_sum(tlo_ws *w, int32_t n, int32_t src, int32_t dst, int plain, heap_t *h) {
  double *seen = w->dtmp;
  int32_t *pred = w->par, *done = w->stamp;
  for (int32_t i = 0; i < n; i++) { seen[i] = INFINITY; done[i] = 0; pred[i] = -1; }
  h->n = 0;
  int64_t cnt = 0;
  seen[src] = 0.0;
  hitem it = {0.0, cnt++, src};
  hpush(h, it);
  while (h->n) {
    hitem t = hpop(h);
    if (done[t.v]) continue;
    done[t.v] = 1;
    if (t.v == dst) break;
    for (int32_t e = w->rowl[t.v]; e < w->rowl[t.v + 1]; e++) {
      int32_t y = w->adj[e];
      double nd = t.d + w->ew[w->adje[e]];
      if (done[y]) continue;
      if (nd < seen[y]) { seen[y] = nd; pred[y] = t.v; hitem q = {nd, cnt++, y}; hpush(h, q); }
    }
  }
  if (!done[dst]) return 100.0; /* except BaseException: dist = 100   riccidist2dgm.py:31-32 */
  /* path = src ... dst; pred chain runs dst -> src, so collect then sum in src -> dst order */
  int32_t len = 0;
  for (int32_t x = dst; x != src; x = pred[x]) w->bq[len++] = x;
  pysum ps; pysum_init(&ps);
  int32_t prev = src;
  for (int32_t i = len - 1; i >= 0; i--) {
    int32_t x = w->bq[i];
    /* weight of (prev,x) */
    double wt = 0.0;
    for (int32_t e = w->rowl[prev]; e < w->rowl[prev + 1]; e++) if (w->adj[e] == x) { wt = w->ew[w->adje[e]]; break; }
    pysum_add(&ps, wt, plain);
    prev = x;
  }
  return pysum_get(&ps, plain);
}

/* "fast oracle" (SURVEY.md 8c): one SSSP per root instead of one Dijkstra per vertex.  The path
 * x -> root is the reverse of the root's shortest-path tree branch; the weights are re-summed in
 * x -> root order exactly as build_fv does.  Tree rule (shared with the CUDA kernel, so the two
 * agree bit for bit even on exact float ties): dist = least fixpoint of d[x] = min_y fl(d[y]+w),
 * parent[x] = smallest local id y with fl(d[y]+w(y,x)) == d[x]. */
static void sssp_root_sums(tlo_ws *w, int32_t n, int32_t root, int plain, heap_t *h, double *out) {
  double *dist = w->dtmp;
  int32_t *done = w->stamp, *par = w->par;
  for (int32_t i = 0; i < n; i++) { dist[i] = INFINITY; done[i] = 0; par[i] = -1; }
  h->n = 0;
  int64_t cnt = 0;
  dist[root] = 0.0;
  hitem it = {0.0, cnt++, root};
  hpush(h, it);
  while (h->n) {
    hitem t = hpop(h);
    if (done[t.v]) continue;
    done[t.v] = 1;
    for (int32_t e = w->rowl[t.v]; e < w->rowl[t.v + 1]; e++) {
      int32_t y = w->adj[e];
      double nd = t.d + w->ew[w->adje[e]];
      if (!done[y] && nd < dist[y]) { dist[y] = nd; hitem q = {nd, cnt++, y}; hpush(h, q); }
    }
  }
  for (int32_t x = 0; x < n; x++) {
    if (x == root || !done[x]) continue;
    for (int32_t e = w->rowl[x]; e < w->rowl[x + 1]; e++) { /* adj ascending -> first hit = smallest id */
      int32_t y = w->adj[e];
      if (done[y] && dist[y] + w->ew[w->adje[e]] == dist[x]) { par[x] = y; w->pw[x] = w->ew[w->adje[e]]; break; }
    }
  }
  for (int32_t x = 0; x < n; x++) {
    if (x == root) { out[x] = 0.0; continue; }
    if (!done[x]) { out[x] = 100.0; continue; }
    pysum ps; pysum_init(&ps);
    for (int32_t y = x; y != root; y = par[y]) pysum_add(&ps, w->pw[y], plain);
    out[x] = pysum_get(&ps, plain);
  }
}

static __thread const double *g_ext_fval; /* per worker thread: filtration values handed in by tlo_run_one_fval */

/* sort contexts (qsort has no closure) */
static __thread const double *g_key; /* per worker thread */
static int cmp_asc(const void *a, const void *b) { /* stable sort by (value) == sort by (value, original index) */
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  if (g_key[x] < g_key[y]) return -1;
  if (g_key[x] > g_key[y]) return 1;
  return (x > y) - (x < y);
}
static int cmp_desc(const void *a, const void *b) { /* sort(reverse=True) keeps original order among equals */
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  if (g_key[x] > g_key[y]) return -1;
  if (g_key[x] < g_key[y]) return 1;
  return (x > y) - (x < y);
}

static inline int32_t uf_find(int32_t *p, int32_t x) { /* path halving, accelerated_PD.py:53-58 */
  while (x != p[x]) { p[x] = p[p[x]]; x = p[x]; }
  return x;
}

static inline double norm_cdf(double x) { return erfc(-x / sqrt(2.0)) / 2.0; } /* PersistenceImager.pyx:60 */

/* PersistenceImager(resolution).transform(dgm) -- isotropic sigma = 1 branch, PersistenceImager.pyx:361-388
 * mesh: _create_mesh :311-314 = linspace(0, 1 + 1/res, res+1, endpoint=False); weight: linear_ramp :22-28 */
static void pimg_accumulate(int32_t res, double birth, double death, double *img) {
  double pers = death - birth;                      /* :366 */
  double wt = pers < 0.0 ? 0.0 : (pers > 1.0 ? 1.0 : (pers - 0.0) * (1.0 - 0.0) / (1.0 - 0.0) + 0.0); /* :22-28 */
  double cb[65], cp[65];
  double pix = 1.0 / res, step = ((1.0 + pix) - 0.0) / (res + 1); /* np.linspace(..., endpoint=False) */
  for (int32_t i = 0; i <= res; i++) {
    double pt = 0.0 + i * step;
    cb[i] = norm_cdf((pt - birth) / 1.0); /* :385 */
    cp[i] = norm_cdf((pt - pers) / 1.0);  /* :386 */
  }
  for (int32_t i = 0; i < res; i++)
    for (int32_t j = 0; j < res; j++) {
      /* curr_img[a,b] = ncdf_p[b]*ncdf_b[a]; 4-term inclusion-exclusion :387-388 */
      double v = cp[j + 1] * cb[i + 1] - cp[j + 1] * cb[i] - cp[j] * cb[i + 1] + cp[j] * cb[i];
      img[i * res + j] += wt * v;
    }
}

/* the transform on a caller-supplied diagram (for the pimg.py:450-505 golden vector) */
void tlo_pimg_transform(const double *dgm, int32_t npts, int32_t res, double *img) {
  for (int32_t i = 0; i < res * res; i++) img[i] = 0.0;
  if (res > 64) return;
  for (int32_t k = 0; k < npts; k++) pimg_accumulate(res, dgm[2 * k], dgm[2 * k + 1], img);
}

static void det_pairs(tlo_detail *det, const tlo_ws *w, int32_t np) {
  if (!det) return;
  det->npairs = np;
  if (det->pkind) memcpy(det->pkind, w->pkind, (size_t)np * 4);
  if (det->pbv) memcpy(det->pbv, w->pbv, (size_t)np * 4);
  if (det->pdv) memcpy(det->pdv, w->pdv, (size_t)np * 4);
  if (det->pbirth) memcpy(det->pbirth, w->pbirth, (size_t)np * 8);
  if (det->pdeath) memcpy(det->pdeath, w->pdeath, (size_t)np * 8);
}

static void add_pair(tlo_ws *w, int32_t *np, int32_t kind, int32_t bv, int32_t dv, double b, double d) {
  int32_t k = (*np)++;
  w->pkind[k] = kind; w->pbv[k] = bv; w->pdv[k] = dv; w->pbirth[k] = b; w->pdeath[k] = d;
}

/* ------------------------------------------------------------------------------------------ */
/* one target through the whole path.  returns the status; img[res*res] is zero unless status<=1 */
static int run_target(const tlo_graph *g, int32_t u, int32_t v, const tlo_params *p, tlo_ws *w, heap_t *h,
                      double *img, tlo_detail *det) {
  const int32_t res = p->resolution, N = g->N;
  const int plain = (p->flags & TLO_F_SUM_PLAIN) != 0, keep0 = (p->flags & TLO_F_KEEP_ZERO) != 0;
  const int node_mode = p->mode == TLO_MODE_NODE, forced = p->mode == TLO_MODE_EDGE_FORCED;
  for (int32_t i = 0; i < res * res; i++) img[i] = 0.0;
  if (det) { det->n = det->m = det->npairs = det->npos = det->nneg = 0; det->lu = det->lv = -1; }
  /* dict_node[u] KeyError -> zeros   riccidist2dgm.py:353 ; graph lacks isolated nodes loaddatas.py:88-92 */
  if (u < 0 || u >= N || (!node_mode && (v < 0 || v >= N))) return TLO_ST_UNKNOWN_NODE;
  if (g->rowptr[u + 1] == g->rowptr[u] || (!node_mode && g->rowptr[v + 1] == g->rowptr[v])) return TLO_ST_UNKNOWN_NODE;

  /* ---- A2 vicinity: riccidist2dgm.py:311-316 / data_utils_NC.py:97-100 ---- */
  int32_t cu = ball(g, u, p->hop, w->mark_u, w->queue);
  int32_t n = 0;
  if (node_mode) {
    ws_reserve(w, cu, 0);
    for (int32_t i = 0; i < cu; i++) w->vert[n++] = w->queue[i];
    for (int32_t i = 0; i < cu; i++) w->mark_u[w->queue[i]] = 0;
  } else {
    /* keep ball(u) members aside: the queue is reused for ball(v) */
    int32_t *bu = (int32_t *)malloc((size_t)cu * 4);
    memcpy(bu, w->queue, (size_t)cu * 4);
    int32_t cv = ball(g, v, p->hop, w->mark_v, w->queue);
    ws_reserve(w, (cu < cv ? cu : cv) + 2, 0);
    if (p->mode == TLO_MODE_EDGE_UNION || p->mode == TLO_MODE_EDGE_REMOVEINTER) {
      const int rem = p->mode == TLO_MODE_EDGE_REMOVEINTER;
      ws_reserve(w, cu + cv + 2, 0);
      for (int32_t i = 0; i < cu; i++) if (!(rem && w->mark_v[bu[i]])) w->vert[n++] = bu[i];             /* nodes_u (minus the intersection) */
      for (int32_t i = 0; i < cv; i++) if (!w->mark_u[w->queue[i]]) w->vert[n++] = w->queue[i];           /* nodes_v not already taken */
      if (rem) {                                              /* + [u, v]   :293 (g.subgraph dedups: only where the removal took them) */
        if (w->mark_v[u]) w->vert[n++] = u;
        if (v != u && w->mark_u[v]) w->vert[n++] = v;
      }
    } else
    for (int32_t i = 0; i < cu; i++) if (w->mark_v[bu[i]]) w->vert[n++] = bu[i];
    if (forced) { /* nodes = list(set(nodes_u) & set(nodes_v)) + [u] + [v]   data_utils_LP.py:111 (g.subgraph dedups) */
      if (!w->mark_v[u]) w->vert[n++] = u;
      if (v != u && !w->mark_u[v]) w->vert[n++] = v;
    }
    for (int32_t i = 0; i < cu; i++) w->mark_u[bu[i]] = 0;
    for (int32_t i = 0; i < cv; i++) w->mark_v[w->queue[i]] = 0;
    free(bu);
  }
  qsort(w->vert, (size_t)n, 4, cmp_i32); /* canonical vertex order */
  if (det) det->n = n;
  if (n == 0) return TLO_ST_EMPTY; /* assert len(components)==1 with 0 components  :318 */
  for (int32_t i = 0; i < n; i++) w->lid[w->vert[i]] = i;
  int32_t m = 0;
  for (int32_t i = 0; i < n; i++) {
    int32_t x = w->vert[i];
    for (int32_t e = g->rowptr[x]; e < g->rowptr[x + 1]; e++) { int32_t y = g->col[e]; if (y > x && w->lid[y] >= 0) m++; }
  }
  ws_reserve(w, n, m);
  m = 0;
  for (int32_t i = 0; i < n; i++) { /* induced subgraph, G.subgraph(nodes) :316 */
    int32_t x = w->vert[i];
    for (int32_t e = g->rowptr[x]; e < g->rowptr[x + 1]; e++) {
      int32_t y = g->col[e];
      if (y > x && w->lid[y] >= 0) { w->elo[m] = i; w->ehi[m] = w->lid[y]; w->ew[m] = g->kappa[e] + 1; m++; }
    }
  }
  int32_t lu = w->lid[u], lv = node_mode ? lu : w->lid[v];
  for (int32_t i = 0; i < n; i++) w->lid[w->vert[i]] = -1;
  if (det) {
    det->m = m; det->lu = lu; det->lv = lv;
    if (det->vert) memcpy(det->vert, w->vert, (size_t)n * 4);
    if (det->elo) memcpy(det->elo, w->elo, (size_t)m * 4);
    if (det->ehi) memcpy(det->ehi, w->ehi, (size_t)m * 4);
    if (det->ew) memcpy(det->ew, w->ew, (size_t)m * 8);
  }
  if ((node_mode || forced) && m == 0) return TLO_ST_EMPTY; /* `return None, None` data_utils_NC.py:103-104, data_utils_LP.py:117-118 */

  /* symmetric local adjacency (ascending neighbour ids) */
  for (int32_t i = 0; i <= n; i++) w->rowl[i] = 0;
  for (int32_t e = 0; e < m; e++) { w->rowl[w->elo[e] + 1]++; w->rowl[w->ehi[e] + 1]++; }
  for (int32_t i = 0; i < n; i++) w->rowl[i + 1] += w->rowl[i];
  {
    int32_t *fill = w->heap_pos;
    for (int32_t i = 0; i < n; i++) fill[i] = w->rowl[i];
    /* lower neighbours first (ascending because edges are lexicographic on hi within... not lo) */
    for (int32_t e = 0; e < m; e++) { int32_t b = w->ehi[e]; w->adj[fill[b]] = w->elo[e]; w->adje[fill[b]] = e; fill[b]++; }
    for (int32_t e = 0; e < m; e++) { int32_t a = w->elo[e]; w->adj[fill[a]] = w->ehi[e]; w->adje[fill[a]] = e; fill[a]++; }
  }

  /* ---- A3 connectivity: assert len(connected_components)==1   :318 ---- */
  {
    int32_t *seen = w->stamp, head = 0, tail = 0, cntc = 0;
    for (int32_t i = 0; i < n; i++) seen[i] = 0;
    seen[0] = 1; w->bq[tail++] = 0;
    while (head < tail) {
      int32_t x = w->bq[head++]; cntc++;
      for (int32_t e = w->rowl[x]; e < w->rowl[x + 1]; e++) { int32_t y = w->adj[e]; if (!seen[y]) { seen[y] = 1; w->bq[tail++] = y; } }
    }
    if (cntc != n) return TLO_ST_DISCONNECTED;
  }

  /* ---- A4 filtration: filtration.build_fv(weight_graph=True, norm)  :20-61 ; KD data_utils_NC.py:34-55 ---- */
  const int roots_in = lu >= 0 && lv >= 0;
  if (!roots_in) {
    for (int32_t i = 0; i < n; i++) { w->d1[i] = 100.0; w->d2[i] = 100.0; } /* NodeNotFound -> 100  :31-32,36-37 */
  } else if (p->flags & TLO_F_ASIS_FV) {
    for (int32_t x = 0; x < n; x++) {
      if (x == lu || x == lv) { w->d1[x] = 0.0; w->d2[x] = 0.0; continue; } /* :22-25 */
      w->d1[x] = dijkstra_path_sum(w, n, x, lu, plain, h);
      w->d2[x] = node_mode ? w->d1[x] : dijkstra_path_sum(w, n, x, lv, plain, h);
    }
  } else {
    sssp_root_sums(w, n, lu, plain, h, w->d1);
    if (node_mode || lv == lu) memcpy(w->d2, w->d1, (size_t)n * 8); else sssp_root_sums(w, n, lv, plain, h, w->d2);
    w->d1[lu] = w->d2[lu] = 0.0; w->d1[lv] = w->d2[lv] = 0.0; /* `if x in [root_1, root_2]` -> all 0 :22-25 */
  }
  if (det) { if (det->d1) memcpy(det->d1, w->d1, (size_t)n * 8); if (det->d2) memcpy(det->d2, w->d2, (size_t)n * 8); }
  {
    double smax = -INFINITY, ssum = -INFINITY;
    for (int32_t x = 0; x < n; x++) {
      double a = w->d1[x], b = w->d2[x];
      double mx = a > b ? a : b, sm = node_mode ? a : a + b;
      if (mx > smax) smax = mx;
      if (sm > ssum) ssum = sm;
    }
    if (p->flags & TLO_F_NORM) {
      if (p->flags & TLO_F_NORM_EPS) { smax = smax + 1e-10; ssum = ssum + 1e-10; } /* data_utils_NC.py:54 */
      else if (smax == 0.0 || ssum == 0.0) return TLO_ST_DEGENERATE;                /* ZeroDivisionError :54-56 */
    }
    for (int32_t x = 0; x < n; x++) {
      double a = w->d1[x], b = w->d2[x], f;
      if (p->descriptor == TLO_DESC_MIN) f = a < b ? a : b;
      else if (p->descriptor == TLO_DESC_MAX) f = a > b ? a : b;
      else f = node_mode ? a : a + b;
      if (p->flags & TLO_F_NORM) f = f / (p->descriptor == TLO_DESC_SUM ? ssum : smax);
      w->fval[x] = f;
    }
  }
  if (p->flags & TLO_F_FILT_CLUSTERING) {
    /* nx.clustering (unweighted): t / (d * (d - 1)) with t = sum over neighbours w of |N(v) & N(w)| (= twice the
       triangles through v), 0 when t == 0; python int / int = the correctly rounded quotient.   networkx cluster.py
       then fv / (max_val + 1e-10)                                                        data_utils_NC.py:122-125 */
    double mx = -INFINITY;
    for (int32_t x = 0; x < n; x++) {
      const int32_t a0 = w->rowl[x], a1 = w->rowl[x + 1];
      long long t = 0;
      for (int32_t e = a0; e < a1; e++) {
        const int32_t y = w->adj[e];
        /* both local rows ascend?  adj rows hold lower neighbours first then higher, each ascending: ascending overall */
        int32_t i = a0, j = w->rowl[y];
        const int32_t j1 = w->rowl[y + 1];
        while (i < a1 && j < j1) {
          const int32_t u1 = w->adj[i], u2 = w->adj[j];
          if (u1 == u2) { t++; i++; j++; } else if (u1 < u2) i++; else j++;
        }
      }
      const long long d = a1 - a0;
      const double cval = t == 0 ? 0.0 : (double)t / (double)(d * (d - 1));
      w->fval[x] = cval;
      if (cval > mx) mx = cval;
    }
    for (int32_t x = 0; x < n; x++) w->fval[x] = w->fval[x] / (mx + 1e-10);
  }
  if (p->flags & (TLO_F_FILT_DEGREE | TLO_F_FILT_CENTRALITY)) {
    /* the PDGNN generators' structural filtrations: no roots, no distances
       degree     : [subgraph.degree()[i]]; fv / (max(filtration_val) + 1e-10)                 data_utils_NC.py:126-128
       centrality : nx.degree_centrality = d * (1.0 / (len(G) - 1.0)) (1 for a lone vertex); fv / (max_val + 1e-10)  :118-121 */
    double mx = -INFINITY;
    const double sc = n > 1 ? 1.0 / ((double)n - 1.0) : 1.0;
    for (int32_t x = 0; x < n; x++) {
      double d = (double)(w->rowl[x + 1] - w->rowl[x]);
      if (p->flags & TLO_F_FILT_CENTRALITY) d = n > 1 ? d * sc : 1.0;
      w->fval[x] = d;
      if (d > mx) mx = d;
    }
    for (int32_t x = 0; x < n; x++) w->fval[x] = w->fval[x] / (mx + 1e-10);
  }
  if (g_ext_fval) {
    /* caller-supplied filtration values in canonical vertex order (tlo_run_one_fval): e.g. the heat kernel signature, which
       the test computes with the reference's own numpy / scipy lines (data_utils_NC.py:87-93,115-117) */
    for (int32_t x = 0; x < n; x++) w->fval[x] = g_ext_fval[x];
  }
  if (det && det->fval) memcpy(det->fval, w->fval, (size_t)n * 8);
  if (p->descriptor < 0 || p->descriptor > 2) return TLO_ST_BAD_DESCRIPTOR; /* KeyError accelerated_PD.py:13 */

  /* ---- A5 keys: perturb_filter_function  accelerated_PD.py:6-23 (f64, this exact parenthesisation) ---- */
  const double ee = 1e-6, max_filter = 101;
  for (int32_t e = 0; e < m; e++) {
    double fa = w->fval[w->elo[e]], fb = w->fval[w->ehi[e]];
    double mx = fa > fb ? fa : fb, mn = fa < fb ? fa : fb; /* python max(a,b)/min(a,b): values only */
    volatile double t1 = (mn + 1) * ee;   /* volatile: forbid FMA contraction (F4) */
    volatile double t2 = (max_filter - mx) * ee;
    w->kasc[e] = mx + t1;
    w->kdesc[e] = mn - t2;
  }
  double min_value = 99999999, max_value = -99999999; /* accelerated_PD.py:28-38 */
  int32_t min_v = -1, max_v = -1;
  for (int32_t x = 0; x < n; x++) {
    if (min_value > w->fval[x]) { min_value = w->fval[x]; min_v = x; }
    if (max_value < w->fval[x]) { max_value = w->fval[x]; max_v = x; }
  }
  int32_t np = 0;
  const double *f = w->fval;

  /* ---- A6a ascending sweep  accelerated_PD.py:40-68.  Vertices sort before any incident edge
   * (asc > max endpoint), so creating every singleton up front is equivalent. ---- */
  for (int32_t e = 0; e < m; e++) w->ord_asc[e] = e;
  g_key = w->kasc;
  qsort(w->ord_asc, (size_t)m, 4, cmp_asc);
  for (int32_t x = 0; x < n; x++) w->uf[x] = x;
  for (int32_t k = 0; k < m; k++) {
    int32_t e = w->ord_asc[k], a = w->elo[e], b = w->ehi[e];
    int32_t pu = uf_find(w->uf, a), pv = uf_find(w->uf, b);
    if (pu != pv) {
      int32_t small = f[pu] <= f[pv] ? pu : pv, large = pu + pv - small; /* :61-63 */
      int32_t max_node = f[a] > f[b] ? a : b;                             /* :64 */
      if (keep0 || f[large] < f[max_node]) add_pair(w, &np, TLO_K_UP, large, max_node, f[large], f[max_node]); /* :65-66 */
      w->uf[large] = small;                                               /* :67 */
    }
  }
  add_pair(w, &np, TLO_K_ESS, min_v, max_v, min_value, max_value); /* :110 */

  /* ---- A6b descending sweep  accelerated_PD.py:70-109 ---- */
  for (int32_t e = 0; e < m; e++) w->ord_desc[e] = e;
  g_key = w->kdesc;
  qsort(w->ord_desc, (size_t)m, 4, cmp_desc);
  for (int32_t x = 0; x < n; x++) w->uf[x] = x;
  int32_t npos = 0, nneg = 0;
  for (int32_t k = 0; k < m; k++) {
    int32_t e = w->ord_desc[k], a = w->elo[e], b = w->ehi[e];
    int32_t pu = uf_find(w->uf, a), pv = uf_find(w->uf, b);
    if (pu != pv) {
      w->neg[nneg++] = e;                                                  /* :99 */
      int32_t small = f[pu] <= f[pv] ? pu : pv, large = pu + pv - small;   /* :100-102 */
      int32_t min_node = f[a] < f[b] ? a : b;                              /* :103-104 */
      if (keep0 || f[small] > f[min_node]) add_pair(w, &np, TLO_K_DOWN, small, min_node, f[small], f[min_node]); /* :105-106 */
      w->uf[small] = large;                                                /* :107 */
    } else {
      w->pos[npos++] = e;                                                  /* :109 */
    }
  }
  add_pair(w, &np, TLO_K_ESS_REV, max_v, min_v, max_value, min_value); /* :110 */
  if (det) {
    det->npos = npos; det->nneg = nneg;
    if (det->ord_asc) memcpy(det->ord_asc, w->ord_asc, (size_t)m * 4);
    if (det->ord_desc) memcpy(det->ord_desc, w->ord_desc, (size_t)m * 4);
    if (det->pos) memcpy(det->pos, w->pos, (size_t)npos * 4);
    if (det->neg) memcpy(det->neg, w->neg, (size_t)nneg * 4);
  }

  /* ---- A6c loops: Accelerate_PD  accelerated_PD.py:115-178 ---- */
  if (p->flags & TLO_F_EXTENDED) {
    if (nneg == 0) { det_pairs(det, w, np); return TLO_ST_NO_TREE_EDGES; } /* list(Nodes)[0] IndexError :122 */
    /* rank of every edge in the ascending sweep = total order used for argmax (value order is the
     * reference's `asc`; among equal asc the reference takes the first in set-iteration order, which
     * is implementation-defined and value-invariant (F3); canonical choice here: latest in the
     * ascending sweep). */
    int32_t *arank = w->adje; /* reuse: [m] */
    for (int32_t k = 0; k < m; k++) arank[w->ord_asc[k]] = k;
    /* BFS tree of Neg edges rooted at the first endpoint of the first Neg edge  :119-125 */
    int32_t *tpar = w->tpar, *tpe = w->tpe; /* parent vertex, edge index to parent */
    {
      /* tree adjacency via rowl/adj rebuilt over neg edges */
      for (int32_t i = 0; i <= n; i++) w->rowl[i] = 0;
      for (int32_t k = 0; k < nneg; k++) { int32_t e = w->neg[k]; w->rowl[w->elo[e] + 1]++; w->rowl[w->ehi[e] + 1]++; }
      for (int32_t i = 0; i < n; i++) w->rowl[i + 1] += w->rowl[i];
      int32_t *fill = w->heap_pos, *tadj = w->adj;
      int32_t *tadje = (int32_t *)malloc((size_t)(2 * nneg + 1) * 4);
      for (int32_t i = 0; i < n; i++) fill[i] = w->rowl[i];
      for (int32_t k = 0; k < nneg; k++) {
        int32_t e = w->neg[k], a = w->elo[e], b = w->ehi[e];
        tadj[fill[a]] = b; tadje[fill[a]++] = e; tadj[fill[b]] = a; tadje[fill[b]++] = e;
      }
      int32_t root = w->elo[w->neg[0]];
      for (int32_t i = 0; i < n; i++) tpar[i] = -1;
      int32_t head = 0, tail = 0;
      tpar[root] = root; tpe[root] = -1; w->bq[tail++] = root;
      while (head < tail) {
        int32_t x = w->bq[head++];
        for (int32_t q = w->rowl[x]; q < w->rowl[x + 1]; q++) {
          int32_t y = tadj[q];
          if (tpar[y] < 0) { tpar[y] = x; tpe[y] = tadje[q]; w->bq[tail++] = y; }
        }
      }
      free(tadje);
    }
    int32_t *stamp = w->stamp;
    for (int32_t i = 0; i < n; i++) stamp[i] = -1;
    for (int32_t k = 0; k < npos; k++) {
      int32_t pe = w->pos[k], p0 = w->elo[pe], p1 = w->ehi[pe];
      /* path_0 = p0 -> root, path_1 = p1 -> root; Loop = symmetric difference  :131-151
       * == (p0 -> lca) + (p1 -> lca).  Mark p0's root path, climb from p1 to the first marked vertex. */
      for (int32_t x = p0;; x = tpar[x]) { stamp[x] = k; if (tpar[x] == x) break; }
      int32_t lca = p1;
      while (stamp[lca] != k) lca = tpar[lca];
      int32_t best = -1, best_child = -1, in_path0 = 0;
      for (int32_t x = p0; x != lca; x = tpar[x]) if (best < 0 || arank[tpe[x]] > arank[best]) { best = tpe[x]; best_child = x; in_path0 = 1; }
      for (int32_t x = p1; x != lca; x = tpar[x]) if (best < 0 || arank[tpe[x]] > arank[best]) { best = tpe[x]; best_child = x; in_path0 = 0; }
      /* large_value / low_value  :160-165 */
      int32_t la = w->elo[best], lb = w->ehi[best];
      int32_t lv_v = f[la] >= f[lb] ? la : lb;
      int32_t lo_v = f[p0] <= f[p1] ? p0 : p1;
      double large_value = f[lv_v], low_value = f[lo_v];
      if (keep0 || large_value > low_value) add_pair(w, &np, TLO_K_ONE, lo_v, lv_v, low_value, large_value);
      /* re-root  :168-176: reverse parent pointers from the positive edge's endpoint on the side of
       * large_edge up to large_edge's child end; the positive edge becomes that endpoint's tree edge */
      int32_t node = in_path0 ? p0 : p1, nodec = in_path0 ? p1 : p0, ec = pe;
      for (;;) {
        int32_t tp = tpar[node], te = tpe[node];
        tpar[node] = nodec; tpe[node] = ec;
        if (node == best_child) break;
        nodec = node; ec = te; node = tp;
      }
    }
  }
  det_pairs(det, w, np);

  /* ---- A7 image: PersistenceImager(resolution).transform(np.array(PD_zero + PD_one))  :327-328 ---- */
  if (res <= 64)
    for (int32_t k = 0; k < np; k++)
      if (p->img_mask & (1u << w->pkind[k])) pimg_accumulate(res, w->pbirth[k], w->pdeath[k], img);
  return roots_in ? TLO_ST_OK : TLO_ST_TRIVIAL;
}

/* ------------------------------------------------------------------------------------------ */
/* public entry points (ctypes)                                                                 */

int tlo_run_one(const tlo_graph *g, int32_t u, int32_t v, const tlo_params *p, double *img, tlo_detail *det) {
  tlo_ws *w = ws_new(g->N);
  heap_t h = {0, 0, 0};
  int st = run_target(g, u, v, p, w, &h, img, det);
  if (st > TLO_ST_TRIVIAL) for (int32_t i = 0; i < p->resolution * p->resolution; i++) img[i] = 0.0;
  free(h.a);
  ws_free(w);
  return st;
}

/* one target with the filtration values supplied by the caller (canonical vertex order, length = the vicinity's n as a
 * previous tlo_run_one reports it): everything after build_fv -- keys, sweeps, loops, image -- as usual */
int tlo_run_one_fval(const tlo_graph *g, int32_t u, int32_t v, const tlo_params *p, const double *fval_in, double *img, tlo_detail *det) {
  g_ext_fval = fval_in;
  int st = tlo_run_one(g, u, v, p, img, det);
  g_ext_fval = NULL;
  return st;
}

/* graph2pi.get_pimg_for_all_edges  riccidist2dgm.py:362-370: rows pre-zeroed, failures stay zero,
 * cnt_compute counts successes.  targets: [E,2] graph ids (-1 = unknown label).
 * The reference maps a thread pool over targets (:367-370); here: pthreads pulling blocks of targets
 * from a shared counter.  nthreads<=0: every online core. */
typedef struct {
  const tlo_graph *g; const int32_t *targets; int64_t E; const tlo_params *p;
  double *pi_out; uint8_t *status; int32_t *n_out, *m_out, *npairs_out;
  int64_t next; int64_t cnt_compute; pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg) {
  batch_job *j = (batch_job *)arg;
  const int32_t r2 = j->p->resolution * j->p->resolution;
  tlo_ws *w = ws_new(j->g->N);
  heap_t h = {0, 0, 0};
  tlo_detail d;
  memset(&d, 0, sizeof d);
  int64_t done = 0;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    int64_t lo = j->next; j->next += 4;
    pthread_mutex_unlock(&j->mu);
    if (lo >= j->E) break;
    int64_t hi = lo + 4 < j->E ? lo + 4 : j->E;
    for (int64_t i = lo; i < hi; i++) {
      double *img = j->pi_out + i * r2;
      int st = run_target(j->g, j->targets[2 * i], j->targets[2 * i + 1], j->p, w, &h, img, &d);
      if (st > TLO_ST_TRIVIAL) for (int32_t k = 0; k < r2; k++) img[k] = 0.0; else done++;
      if (j->status) j->status[i] = (uint8_t)st;
      if (j->n_out) j->n_out[i] = d.n;
      if (j->m_out) j->m_out[i] = d.m;
      if (j->npairs_out) j->npairs_out[i] = d.npairs;
    }
  }
  pthread_mutex_lock(&j->mu);
  j->cnt_compute += done;
  pthread_mutex_unlock(&j->mu);
  free(h.a);
  ws_free(w);
  return NULL;
}

int tlo_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

int64_t tlo_run_batch(const tlo_graph *g, const int32_t *targets, int64_t E, const tlo_params *p, int32_t nthreads,
                      double *pi_out, uint8_t *status, int32_t *n_out, int32_t *m_out, int32_t *npairs_out) {
  batch_job j = {g, targets, E, p, pi_out, status, n_out, m_out, npairs_out, 0, 0, PTHREAD_MUTEX_INITIALIZER};
  if (nthreads <= 0) nthreads = tlo_max_threads();
  if (nthreads > 256) nthreads = 256;
  if (nthreads == 1) { batch_worker(&j); return j.cnt_compute; }
  pthread_t th[256];
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &j);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  return j.cnt_compute;
}
