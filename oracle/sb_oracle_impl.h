/* TEST INFRASTRUCTURE ONLY -- see sb_oracle.c.  This file is included once per type
 * triple with the macros  I (IDType), N (NNZType), V (ValueType; ignored when HAS_V==0),
 * F (FeatureType), HAS_V (0 => ValueType=void), TAG  defined by the includer.
 *
 * Plain-C restatement of the reference's CPU algorithm for the hot path; every function
 * cites the reference file:line (relative to /root/reference/src/sparsebase/) it follows.
 */
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(CAT(sbo, name), TAG)
#define LOCAL(name) CAT(CAT(l, name), TAG)

#if HAS_V
typedef V LOCAL(val_t);
#else
typedef char LOCAL(val_t); /* never dereferenced */
#endif
#define VT LOCAL(val_t)

typedef struct {
  I row, col;
  VT val;
  int64_t idx;
} LOCAL(triple);

/* comparator of format/coo.cc:119-125 / :138-145 -- (row, col), value ignored.  The
 * original index is the last key so that qsort is deterministic; with duplicate
 * (row,col) the reference's unstable std::sort is itself unspecified. */
static int LOCAL(cmp_triple)(const void *a, const void *b) {
  const LOCAL(triple) *x = (const LOCAL(triple) *)a, *y = (const LOCAL(triple) *)b;
  if (x->row != y->row) return x->row < y->row ? -1 : 1;
  if (x->col != y->col) return x->col < y->col ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/* format/coo.cc:96-157 -- COO constructor, ignore_sort=false: serial sortedness check
 * starting from (prev_row, prev_col) = (0, 0); if an inversion is found, sort the
 * caller's arrays IN PLACE by (row, col). */
int FN(coo_ctor_sort)(int64_t n, int64_t m, int64_t nnz, void *row_, void *col_,
                      void *vals_) {
  I *row = (I *)row_, *col = (I *)col_;
  VT *vals = (VT *)vals_;
  (void)vals;
  (void)n;
  (void)m;
  int not_sorted = 0;
  I prev_row = 0, prev_col = 0;
  for (int64_t i = 0; i < nnz; i++) { /* coo.cc:99-107 */
    if (prev_row > row[i] || (prev_row == row[i] && prev_col > col[i])) {
      not_sorted = 1;
      break;
    }
    prev_row = row[i];
    prev_col = col[i];
  }
  if (!not_sorted) return 0;
  LOCAL(triple) *t = (LOCAL(triple) *)malloc((size_t)(nnz ? nnz : 1) * sizeof(*t));
  if (!t) return 1;
  for (int64_t i = 0; i < nnz; i++) { /* coo.cc:116-118 / :134-137 */
    t[i].row = row[i];
    t[i].col = col[i];
    t[i].idx = i;
#if HAS_V
    t[i].val = vals ? vals[i] : (VT)0;
#endif
  }
  qsort(t, (size_t)nnz, sizeof(*t), LOCAL(cmp_triple)); /* coo.cc:119-125 / :138-145 */
  for (int64_t i = 0; i < nnz; i++) {                   /* coo.cc:127-131 / :147-155 */
    row[i] = t[i].row;
    col[i] = t[i].col;
#if HAS_V
    if (vals) vals[i] = t[i].val;
#endif
  }
  free(t);
  return 0;
}

typedef struct {
  I idx;
  VT val;
} LOCAL(pair);

/* std::less<std::pair<IDType,ValueType>> of format/csr.cc:147-148, csc.cc:147-148:
 * lexicographic (index, then value). */
static int LOCAL(cmp_pair)(const void *a, const void *b) {
  const LOCAL(pair) *x = (const LOCAL(pair) *)a, *y = (const LOCAL(pair) *)b;
  if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
#if HAS_V
  if (x->val < y->val) return -1;
  if (y->val < x->val) return 1;
#endif
  return 0;
}

/* format/csr.cc:99-157 (CSR ctor) == format/csc.cc:99-157 (CSC ctor) with the roles of
 * row/col swapped: if ANY segment is not non-decreasing (each segment's scan starts from
 * prev_value = 0, csr.cc:107), sort EVERY segment of length > 1 by (index, value).
 * vals == NULL with a non-void ValueType sorts with value 0 (csr.cc:144). */
int FN(compressed_ctor_sort)(int64_t n, void *ptr_, void *idx_, void *vals_) {
  N *ptr = (N *)ptr_;
  I *idx = (I *)idx_;
  VT *vals = (VT *)vals_;
  (void)vals;
  int not_sorted = 0;
  for (int64_t i = 0; i < n && !not_sorted; i++) { /* csr.cc:104-116 */
    I prev = 0;
    for (N j = ptr[i]; j < ptr[i + 1]; j++) {
      if (idx[j] < prev) {
        not_sorted = 1;
        break;
      }
      prev = idx[j];
    }
  }
  if (!not_sorted) return 0;
  N maxlen = 0;
  for (int64_t i = 0; i < n; i++)
    if (ptr[i + 1] - ptr[i] > maxlen) maxlen = ptr[i + 1] - ptr[i];
  LOCAL(pair) *p = (LOCAL(pair) *)malloc((size_t)(maxlen ? maxlen : 1) * sizeof(*p));
  if (!p) return 1;
  for (int64_t i = 0; i < n; i++) { /* csr.cc:124-156 */
    N start = ptr[i], end = ptr[i + 1];
    if (end - start <= 1) continue;
    for (N j = start; j < end; j++) {
      p[j - start].idx = idx[j];
#if HAS_V
      p[j - start].val = vals ? vals[j] : (VT)0;
#else
      p[j - start].val = 0;
#endif
    }
    qsort(p, (size_t)(end - start), sizeof(*p), LOCAL(cmp_pair));
    for (N j = start; j < end; j++) {
      idx[j] = p[j - start].idx;
#if HAS_V
      if (vals) vals[j] = p[j - start].val;
#endif
    }
  }
  free(p);
  return 0;
}

int FN(csr_ctor_sort)(int64_t n, int64_t m, void *rp, void *col, void *vals) {
  (void)m;
  return FN(compressed_ctor_sort)(n, rp, col, vals);
}

/* converter/converter_order_two.cc:162-212 -- CooCsrFunctionConditional.  The input is a
 * reference COO object, i.e. its constructor (coo.cc:96-157) has already run: this
 * restatement runs it too (on a private copy so that the caller's arrays stay intact --
 * the reference mutates them in place), then: histogram at [row] (:180-183), inclusive
 * scan (:185-187), shift right (:189-192), col/vals copied verbatim (:181, :199-201),
 * CSR ctor (csr.cc:99-157). */
int FN(coo_to_csr)(int64_t n, int64_t m, int64_t nnz, void *row_, void *col_, void *vals_,
                   void *orp_, void *ocol_, void *ovals_) {
  N *row_ptr = (N *)orp_;
  I *ocol = (I *)ocol_;
  VT *ovals = (VT *)ovals_;
  size_t cnt = (size_t)(nnz ? nnz : 1);
  I *row = (I *)malloc(cnt * sizeof(I));
  if (!row) return 1;
  memcpy(row, row_, (size_t)nnz * sizeof(I));
  memcpy(ocol, col_, (size_t)nnz * sizeof(I));
#if HAS_V
  if (vals_) memcpy(ovals, vals_, (size_t)nnz * sizeof(VT));
  FN(coo_ctor_sort)(n, m, nnz, row, ocol, vals_ ? ovals : NULL);
#else
  (void)ovals;
  FN(coo_ctor_sort)(n, m, nnz, row, ocol, NULL);
#endif
  for (int64_t i = 0; i <= n; i++) row_ptr[i] = 0;
  for (int64_t i = 0; i < nnz; i++) row_ptr[row[i]]++;
  for (int64_t i = 1; i <= n; i++) row_ptr[i] += row_ptr[i - 1];
  for (int64_t i = n; i > 0; i--) row_ptr[i] = row_ptr[i - 1];
  row_ptr[0] = 0;
  free(row);
#if HAS_V
  return FN(compressed_ctor_sort)(n, row_ptr, ocol, vals_ ? ovals : NULL);
#else
  return FN(compressed_ctor_sort)(n, row_ptr, ocol, NULL);
#endif
}

/* converter/converter_order_two.cc:71-118 -- CsrCooFunctionConditional: expand row_ptr
 * (:86-95), copy col (:97-99) and vals (:103-113); COO ctor (coo.cc:96-157). */
int FN(csr_to_coo)(int64_t n, int64_t m, void *rp_, void *col_, void *vals_, void *orow_,
                   void *ocol_, void *ovals_) {
  N *rp = (N *)rp_;
  I *orow = (I *)orow_;
  int64_t nnz = (int64_t)rp[n];
  int64_t count = 0;
  for (int64_t i = 0; i < n; i++)
    for (N j = rp[i]; j < rp[i + 1]; j++) orow[count++] = (I)i;
  memcpy(ocol_, col_, (size_t)nnz * sizeof(I));
#if HAS_V
  if (vals_) memcpy(ovals_, vals_, (size_t)nnz * sizeof(VT));
  return FN(coo_ctor_sort)(n, m, nnz, orow, ocol_, vals_ ? ovals_ : NULL);
#else
  (void)ovals_;
  return FN(coo_ctor_sort)(n, m, nnz, orow, ocol_, NULL);
#endif
}

/* converter/converter_order_two.cc:20-70 -- CooCscFunctionConditional on an (already
 * constructed, hence (row,col)-sorted) COO: column histogram at [c+1] (:49-51),
 * inclusive scan (:52-54), stable counting scatter with a per-column cursor (:56-66);
 * col_ptr has n+1 entries -- dims[0], not m (:32).  Then CSC ctor (csc.cc:99-157). */
static int LOCAL(coo_to_csc_sorted)(int64_t n, int64_t nnz, const I *row, const I *col,
                                    const VT *vals, N *col_ptr, I *orow, VT *ovals) {
  N *counter = (N *)calloc((size_t)(n ? n : 1), sizeof(N));
  if (!counter) return 1;
  for (int64_t i = 0; i <= n; i++) col_ptr[i] = 0;
  for (int64_t i = 0; i < nnz; i++) col_ptr[col[i] + 1]++;
  for (int64_t i = 1; i <= n; i++) col_ptr[i] += col_ptr[i - 1];
  for (int64_t i = 0; i < nnz; i++) {
    I c = col[i];
    N pos = col_ptr[c] + counter[c]++;
    orow[pos] = row[i];
#if HAS_V
    if (vals) ovals[pos] = vals[i];
#endif
  }
  free(counter);
  (void)vals;
  (void)ovals;
#if HAS_V
  return FN(compressed_ctor_sort)(n, col_ptr, orow, vals ? ovals : NULL);
#else
  return FN(compressed_ctor_sort)(n, col_ptr, orow, NULL);
#endif
}

int FN(coo_to_csc)(int64_t n, int64_t m, int64_t nnz, void *row_, void *col_, void *vals_,
                   void *ocp_, void *orow_, void *ovals_) {
  size_t cnt = (size_t)(nnz ? nnz : 1);
  I *row = (I *)malloc(cnt * sizeof(I)), *col = (I *)malloc(cnt * sizeof(I));
  VT *vals = NULL;
  if (!row || !col) return 1;
  memcpy(row, row_, (size_t)nnz * sizeof(I));
  memcpy(col, col_, (size_t)nnz * sizeof(I));
#if HAS_V
  if (vals_) {
    vals = (VT *)malloc(cnt * sizeof(VT));
    memcpy(vals, vals_, (size_t)nnz * sizeof(VT));
  }
#endif
  FN(coo_ctor_sort)(n, m, nnz, row, col, vals);
  int rc = LOCAL(coo_to_csc_sorted)(n, nnz, row, col, vals, (N *)ocp_, (I *)orow_,
                                    (VT *)ovals_);
  free(row);
  free(col);
  free(vals);
  return rc;
}

/* converter/converter_order_two.cc:119-128 -- CsrCscFunctionConditional = CsrCoo then
 * CooCsc. */
int FN(csr_to_csc)(int64_t n, int64_t m, void *rp_, void *col_, void *vals_, void *ocp_,
                   void *orow_, void *ovals_) {
  N *rp = (N *)rp_;
  int64_t nnz = (int64_t)rp[n];
  size_t cnt = (size_t)(nnz ? nnz : 1);
  I *row = (I *)malloc(cnt * sizeof(I)), *col = (I *)malloc(cnt * sizeof(I));
  VT *vals = NULL;
  if (!row || !col) return 1;
#if HAS_V
  if (vals_) vals = (VT *)malloc(cnt * sizeof(VT));
#endif
  FN(csr_to_coo)(n, m, rp_, col_, vals_, row, col, vals);
  int rc = LOCAL(coo_to_csc_sorted)(n, nnz, row, col, vals, (N *)ocp_, (I *)orow_,
                                    (VT *)ovals_);
  free(row);
  free(col);
  free(vals);
  return rc;
}

/* reorder/degree_reorder.cc:22-62 -- counting sort of the vertices by degree; each degree
 * bucket is filled from its END backwards (:42-46), whole array reversed if !ascending
 * (:47-53), inverse returned (:54-57).  `mr` is sized n+1 here: the reference allocates n
 * and indexes mr[n] for the maximum-degree bucket (:41-45, heap overflow; same values). */
int FN(degree_reorder)(int64_t n, int64_t m, void *rp_, void *col_, void *vals_,
                       int ascending, void *oinv_) {
  N *row_ptr = (N *)rp_;
  I *inv = (I *)oinv_;
  (void)m;
  (void)col_;
  (void)vals_;
  I *counts = (I *)calloc((size_t)n + 1, sizeof(I));
  I *sorted = (I *)malloc((size_t)(n ? n : 1) * sizeof(I));
  I *mr = (I *)calloc((size_t)n + 1, sizeof(I));
  if (!counts || !sorted || !mr) return 1;
  for (int64_t u = 0; u < n; u++) counts[row_ptr[u + 1] - row_ptr[u]]++;
  for (int64_t u = 1; u < n + 1; u++) counts[u] += counts[u - 1];
  for (int64_t u = 0; u < n; u++) {
    I ec = counts[row_ptr[u + 1] - row_ptr[u]];
    sorted[ec - mr[ec] - 1] = (I)u;
    mr[ec]++;
  }
  if (!ascending)
    for (int64_t i = 0; i < n / 2; i++) {
      I swp = sorted[i];
      sorted[i] = sorted[n - i - 1];
      sorted[n - i - 1] = swp;
    }
  for (int64_t i = 0; i < n; i++) inv[sorted[i]] = (I)i;
  free(mr);
  free(counts);
  free(sorted);
  return 0;
}

/* reorder/rcm_reorder.cc:22-81 -- pseudo-peripheral vertex by repeated FIFO BFS. */
static I LOCAL(peripheral)(const N *xadj, const I *adj, I start, I *distance, I *Q) {
  I r = start;
  I rlevel = -1, qlevel = 0, deg = -1, flag = -1;
  while (rlevel != qlevel) { /* :34 */
    rlevel = qlevel;
    I qrp = 0, qwp = 0;
    distance[r] = 0;
    Q[qwp++] = r;
    while (qrp < qwp) { /* :42-55 */
      I u = Q[qrp++];
      for (N p = xadj[u]; p < xadj[u + 1]; p++) {
        I v = adj[p];
        if (distance[v] == (I)-1) {
          distance[v] = distance[u] + 1;
          Q[qwp++] = v;
          if (distance[v] > qlevel) qlevel = distance[v];
        }
      }
    }
    if (qrp == qlevel + 1) return r; /* :58 (distance[] left set) */
    flag = -1;
    if (rlevel != qlevel) { /* :62-78 */
      for (I i = 0; i < qrp; i++) {
        if (qlevel == distance[Q[i]]) {
          if (flag == -1) {
            deg = (I)(xadj[Q[i] + 1] - xadj[Q[i]] + 1);
            flag = 0;
          }
          if ((I)(xadj[Q[i] + 1] - xadj[Q[i]]) < deg) {
            r = Q[i];
            deg = (I)(xadj[Q[i] + 1] - xadj[Q[i]]);
          }
        }
        distance[Q[i]] = -1;
      }
    }
  }
  return r;
}

typedef struct {
  I deg, id;
} LOCAL(degid);
static int LOCAL(cmp_degid)(const void *a, const void *b) {
  const LOCAL(degid) *x = (const LOCAL(degid) *)a, *y = (const LOCAL(degid) *)b;
  if (x->deg != y->deg) return x->deg < y->deg ? -1 : 1;
  return x->id < y->id ? -1 : (x->id > y->id);
}

/* reorder/rcm_reorder.cc:83-166 -- Cuthill-McKee BFS per connected component from the
 * pseudo-peripheral vertex; the unvisited neighbours of each popped vertex pass through
 * a min-heap keyed (degree, id) (:100, :130-143) -- since all (degree,id) pairs are
 * distinct, draining the heap == appending them sorted by (degree, id); each component's
 * slice is then reversed (:147-153) and the inverse is returned (:158-160).  Isolated
 * vertices are appended in place (:111-116). */
int FN(rcm_reorder)(int64_t n, int64_t m, void *rp_, void *col_, void *vals_, void *oinv_) {
  const N *xadj = (const N *)rp_;
  const I *adj = (const I *)col_;
  I *inv = (I *)oinv_;
  (void)m;
  (void)vals_;
  size_t cnt = (size_t)(n ? n : 1);
  I *Q = (I *)malloc(cnt * sizeof(I)), *Qp = (I *)malloc(cnt * sizeof(I));
  I *Qp2 = (I *)malloc(cnt * sizeof(I)), *distance = (I *)malloc(cnt * sizeof(I));
  unsigned char *Vis = (unsigned char *)calloc(cnt, 1);
  N maxdeg = 0;
  for (int64_t i = 0; i < n; i++)
    if (xadj[i + 1] - xadj[i] > maxdeg) maxdeg = xadj[i + 1] - xadj[i];
  LOCAL(degid) *pq = (LOCAL(degid) *)malloc((size_t)(maxdeg ? maxdeg : 1) * sizeof(*pq));
  if (!Q || !Qp || !Qp2 || !distance || !Vis || !pq) return 1;
  for (int64_t i = 0; i < n; i++) distance[i] = -1;
  int64_t qrp = 0, qwp = 0, qst = 0;
  for (int64_t i = 0; i < n; i++) {
    if (Vis[i]) continue;
    if (xadj[i] == xadj[i + 1]) { /* :111-116 */
      Q[qwp] = (I)i;
      Qp2[qwp++] = (I)i;
      Vis[i] = 1;
      continue;
    }
    I perv = LOCAL(peripheral)(xadj, adj, (I)i, distance, Qp); /* :119 */
    qst = qwp;
    Vis[perv] = 1;
    Q[qwp++] = perv;
    while (qrp < qwp) { /* :125-144 */
      I u = Q[qrp++];
      int64_t k = 0;
      for (N p = xadj[u]; p < xadj[u + 1]; p++) {
        I v = adj[p];
        if (!Vis[v]) {
          pq[k].deg = (I)(xadj[v + 1] - xadj[v]);
          pq[k].id = v;
          k++;
          Vis[v] = 1;
        }
      }
      if (k > 1) qsort(pq, (size_t)k, sizeof(*pq), LOCAL(cmp_degid));
      for (int64_t j = 0; j < k; j++) Q[qwp++] = pq[j].id;
    }
    for (int64_t j = qst; j < qwp; j++) Qp2[j] = Q[qwp - 1 - (j - qst)]; /* :147-153 */
  }
  for (int64_t i = 0; i < n; i++) inv[Qp2[i]] = (I)i; /* :158-160 */
  free(Q);
  free(Qp);
  free(Qp2);
  free(distance);
  free(Vis);
  free(pq);
  return 0;
}

/* permute/permute_order_two.cc:21-79 -- PermuteOrderTwoCSR: invert the row order
 * (:46-48), walk the new rows gathering the old row's entries with columns renumbered
 * through col_order in source order (:64-74), then the CSR constructor of :76-77
 * (ignore_sort=false) sorts the rows (csr.cc:99-157).  NULL order == identity. */
int FN(permute2d)(int64_t n, int64_t m, void *rp_, void *col_, void *vals_, void *ro_,
                  void *co_, void *orp_, void *ocol_, void *ovals_) {
  const N *xadj = (const N *)rp_;
  const I *adj = (const I *)col_;
  const VT *vals = (const VT *)vals_;
  const I *row_order = (const I *)ro_, *col_order = (const I *)co_;
  N *nxadj = (N *)orp_;
  I *nadj = (I *)ocol_;
  VT *nvals = (VT *)ovals_;
  (void)m;
  I *irow = NULL;
  if (row_order) {
    irow = (I *)malloc((size_t)(n ? n : 1) * sizeof(I));
    if (!irow) return 1;
    for (int64_t i = 0; i < n; i++) irow[row_order[i]] = (I)i;
  }
  N c = 0;
  nxadj[0] = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t u = irow ? (int64_t)irow[i] : i;
    nxadj[i + 1] = nxadj[i] + (xadj[u + 1] - xadj[u]);
    for (N v = xadj[u]; v < xadj[u + 1]; v++) {
      nadj[c] = col_order ? col_order[adj[v]] : adj[v];
#if HAS_V
      if (vals) nvals[c] = vals[v];
#endif
      c++;
    }
  }
  free(irow);
  (void)vals;
  (void)nvals;
#if HAS_V
  return FN(compressed_ctor_sort)(n, nxadj, nadj, vals ? nvals : NULL);
#else
  return FN(compressed_ctor_sort)(n, nxadj, nadj, NULL);
#endif
}

/* feature/degrees.cc:93-105 */
int FN(degrees)(int64_t n, int64_t m, void *rp_, void *col_, void *vals_, void *odeg_) {
  const N *rows = (const N *)rp_;
  I *deg = (I *)odeg_;
  (void)m;
  (void)col_;
  (void)vals_;
  for (int64_t i = 0; i < n; i++) deg[i] = (I)(rows[i + 1] - rows[i]);
  return 0;
}

/* feature/degree_distribution.cc:146-162:
 *   dist[i] = (rows[i+1] - rows[i]) / (FeatureType) num_edges   (:158) */
int FN(degree_distribution)(int64_t n, int64_t m, void *rp_, void *col_, void *vals_,
                            void *odist_) {
  const N *rows = (const N *)rp_;
  F *dist = (F *)odist_;
  (void)m;
  (void)col_;
  (void)vals_;
  N num_edges = rows[n];
  for (int64_t i = 0; i < n; i++) dist[i] = (rows[i + 1] - rows[i]) / (F)num_edges;
  return 0;
}

/* feature/degrees_degree_distribution.cc:147-166 (fused degrees + distribution),
 * feature/min_max_avg_degree.cc:168-191 (min / max over rows from rows[1]-rows[0], avg =
 * degree_sum / (FeatureType) num_vertices), feature/bandwidth.cc:92-111 (max |i-j| + 1 over the
 * nonzeros, an int), feature/profile.cc:92-106 (sum over rows of i - min(i, smallest column of
 * the row); the reference accumulates it in IDType, here in int64_t -- identical while it fits).
 * out_scalars = {min_degree, max_degree, bandwidth, profile}; *out_avg in FeatureType. */
int FN(degree_features)(int64_t n, int64_t m, void *rp_, void *col_, void *odeg_, void *odist_,
                        int64_t *out_scalars, void *out_avg_) {
  const N *rows = (const N *)rp_;
  const I *col = (const I *)col_;
  I *deg = (I *)odeg_;
  F *dist = (F *)odist_;
  (void)m;
  N num_edges = rows[n];
  for (int64_t i = 0; i < n; i++) {
    if (deg) deg[i] = (I)(rows[i + 1] - rows[i]);
    if (dist) dist[i] = (rows[i + 1] - rows[i]) / (F)num_edges;
  }
  N mn = n > 0 ? rows[1] - rows[0] : 0, mx = mn;
  for (int64_t i = 1; i < n; i++) {
    N d = rows[i + 1] - rows[i];
    if (d < mn) mn = d;
    if (d > mx) mx = d;
  }
  int64_t bandwidth = 0, profile = 0;
  if (col) {
    for (int64_t i = 0; i < n; i++) {
      int64_t j = i;
      for (N k = rows[i]; k < rows[i + 1]; k++) {
        int64_t c = (int64_t)col[k];
        int64_t w = i >= c ? i - c + 1 : c - i + 1;
        if (w > bandwidth) bandwidth = w;
        if (j > c) j = c;
      }
      profile += i - j;
    }
  }
  out_scalars[0] = (int64_t)mn;
  out_scalars[1] = (int64_t)mx;
  out_scalars[2] = bandwidth;
  out_scalars[3] = profile;
  if (out_avg_) *(F *)out_avg_ = n > 0 ? (rows[n] - rows[0]) / (F)n : (F)0;
  return 0;
}

/* reorder/boba_reorder.cc:35-137 -- BOBAReorder::GetReorderCOO.  The list is sorted by
 * (col, row) (:63-66); the sequential variant (:73-105) places a vertex at its first appearance
 * in the row array of the sorted list, then (vertices never seen there) at its first appearance
 * in the column array, then the vertices without entries by id; the parallel variant (:107-127)
 * states the same order as key[v] = min index of v in rows ++ cols (2 * nnz if absent), ranked
 * by (key, id) -- restated here in that form (its OpenMP loop updates the minimum without
 * atomics; single-threaded it is this).  inv[v] = new position of v, nodes = max(n, m). */
static int FN(boba_cmp)(const void *a_, const void *b_) {
  const I *a = (const I *)a_, *b = (const I *)b_;
  if (a[1] != b[1]) return a[1] < b[1] ? -1 : 1;
  if (a[0] != b[0]) return a[0] < b[0] ? -1 : 1;
  return 0;
}
static int FN(boba_key_cmp)(const void *a_, const void *b_) {
  const int64_t *a = (const int64_t *)a_, *b = (const int64_t *)b_;
  if (a[0] != b[0]) return a[0] < b[0] ? -1 : 1;
  if (a[1] != b[1]) return a[1] < b[1] ? -1 : 1;
  return 0;
}
int FN(boba_reorder)(int64_t n, int64_t m, int64_t nnz, void *row_, void *col_, void *out_inv_) {
  const I *row = (const I *)row_, *col = (const I *)col_;
  I *inv = (I *)out_inv_;
  const int64_t nodes = n >= m ? n : m;
  I *pairs = (I *)malloc(sizeof(I) * 2 * (size_t)(nnz > 0 ? nnz : 1));
  for (int64_t i = 0; i < nnz; i++) {
    pairs[2 * i] = row[i];
    pairs[2 * i + 1] = col[i];
  }
  qsort(pairs, (size_t)nnz, 2 * sizeof(I), FN(boba_cmp));
  int64_t *key = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)(nodes > 0 ? nodes : 1));
  for (int64_t v = 0; v < nodes; v++) {
    key[2 * v] = 2 * nnz;
    key[2 * v + 1] = v;
  }
  for (int64_t i = 0; i < nnz; i++)
    if (i < key[2 * (int64_t)pairs[2 * i]]) key[2 * (int64_t)pairs[2 * i]] = i;
  for (int64_t i = 0; i < nnz; i++)
    if (nnz + i < key[2 * (int64_t)pairs[2 * i + 1]]) key[2 * (int64_t)pairs[2 * i + 1]] = nnz + i;
  qsort(key, (size_t)nodes, 2 * sizeof(int64_t), FN(boba_key_cmp));
  for (int64_t i = 0; i < nodes; i++) inv[key[2 * i + 1]] = (I)i;
  free(key);
  free(pairs);
  return 0;
}

/* reorder/reorder_heatmap.cc:43-120 -- ReorderHeatmapCSRArrayArray: the b x b grid of nonzero
 * densities of the matrix as it would look after the row / column permutations (order[i] = new
 * position of i): bsize = n / b for BOTH dimensions (:62), block index clamped to b - 1
 * (:73-74, :83-84), heat[i * b + j] = density / (row_ptr[n] + .0f) -- a FLOAT division whatever
 * FloatType is (:112), then converted.  Returns 1 where the reference throws (b > n or b > m,
 * :52-56). */
int FN(reorder_heatmap)(int64_t n, int64_t m, void *rp_, void *col_, void *order_r_,
                        void *order_c_, int b, void *out_heat_) {
  const N *rows = (const N *)rp_;
  const I *col = (const I *)col_;
  const I *order_r = (const I *)order_r_, *order_c = (const I *)order_c_;
  F *heat = (F *)out_heat_;
  if (b < 1 || b > n || b > m) return 1;
  const I bsize = (I)(n / b);
  N *density = (N *)calloc((size_t)b * b, sizeof(N));
  for (int64_t i = 0; i < n; i++) {
    I u = order_r[i];
    I bu = u / bsize;
    if (bu >= b) bu = (I)(b - 1);
    for (N k = rows[i]; k < rows[i + 1]; k++) {
      I v = order_c[col[k]];
      I bv = v / bsize;
      if (bv >= b) bv = (I)(b - 1);
      density[(size_t)bu * b + bv]++;
    }
  }
  for (int64_t k = 0; k < (int64_t)b * b; k++) heat[k] = (F)(density[k] / (rows[n] + .0f));
  free(density);
  return 0;
}

/* io/edge_list_reader.cc:28-151 -- EdgeListReader::ReadCOO on an in-memory edge list:
 * self edges dropped when remove_self (:40 / :98), the reverse edge pushed right behind
 * every kept edge when undirected (:43-44 / :101-102), n = max u + 1, m = max v + 1 over the
 * kept edges (:46-47), squared when square || undirected (:51-54), sort by (row, col)
 * (:56-65 / :114-123), unique on (row, col) keeping the first of each run (:67-75 / :125-134).
 * The reference's std::sort is unstable: with duplicate edges of different weights, WHICH
 * weight survives is unspecified there; this restatement (and the CUDA path) keep the first in
 * input order.  Outputs must hold 2 * n_edges entries when undirected.
 * out3 = {n, m, nnz}. */
int FN(edges_to_coo)(int64_t n_edges, void *u_, void *v_, void *w_, int remove_duplicates,
                     int remove_self, int undirected, int square, void *orow_, void *ocol_,
                     void *ovals_, int64_t *out3) {
  const I *u = (const I *)u_, *v = (const I *)v_;
  const VT *w = (const VT *)w_;
  I *orow = (I *)orow_, *ocol = (I *)ocol_;
  VT *ovals = (VT *)ovals_;
  (void)w;
  (void)ovals;
  size_t cap = (size_t)(n_edges > 0 ? n_edges : 1) * (undirected ? 2 : 1);
  LOCAL(triple) *t = (LOCAL(triple) *)malloc(cap * sizeof(*t));
  if (!t) return 1;
  int64_t cnt = 0, n = 0, m = 0;
  for (int64_t i = 0; i < n_edges; i++) {
    if (u[i] != v[i] || !remove_self) {
      t[cnt].row = u[i];
      t[cnt].col = v[i];
      t[cnt].idx = cnt;
#if HAS_V
      t[cnt].val = w ? w[i] : (VT)0;
#endif
      cnt++;
      if (undirected) {
        t[cnt].row = v[i];
        t[cnt].col = u[i];
        t[cnt].idx = cnt;
#if HAS_V
        t[cnt].val = w ? w[i] : (VT)0;
#endif
        cnt++;
      }
      if ((int64_t)u[i] + 1 > n) n = (int64_t)u[i] + 1;
      if ((int64_t)v[i] + 1 > m) m = (int64_t)v[i] + 1;
    }
  }
  if (square || undirected) {
    if (m > n) n = m;
    m = n;
  }
  qsort(t, (size_t)cnt, sizeof(*t), LOCAL(cmp_triple));
  int64_t nnz = 0;
  for (int64_t i = 0; i < cnt; i++) {
    if (remove_duplicates && i > 0 && t[i].row == t[i - 1].row && t[i].col == t[i - 1].col)
      continue;
    orow[nnz] = t[i].row;
    ocol[nnz] = t[i].col;
#if HAS_V
    if (ovals && w) ovals[nnz] = t[i].val;
#endif
    nnz++;
  }
  free(t);
  out3[0] = n;
  out3[1] = m;
  out3[2] = nnz;
  return 0;
}

#undef VT
#undef FN
#undef LOCAL
#undef CAT
#undef CAT_
