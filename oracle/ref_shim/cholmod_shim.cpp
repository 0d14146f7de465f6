// cholmod_shim.cpp -- TEST INFRASTRUCTURE ONLY: the slice of the CHOLMOD C API declared in cholmod.h, implemented from
// the published algorithms (greedy minimum weighted degree on compressed variables; elimination tree + up-looking
// simplicial Cholesky, Davis, "Direct Methods for Sparse Linear Systems", ch. 4).  Lets oracle/_ref run the unmodified
// reference optimiser; not used by, nor linked into, the product.
#include "cholmod.h"
#include "cs.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <vector>

namespace {
struct Symbolic {
  std::vector<int> perm, pinv, parent;
  bool gram = false;   // the analysed matrix was unsymmetric: factor A*A'
};
[[noreturn]] void die(const char* m) { std::fprintf(stderr, "cholmod shim: %s\n", m); std::abort(); }

cholmod_sparse* alloc_sparse(size_t nrow, size_t ncol, size_t nzmax, int stype) {
  cholmod_sparse* A = (cholmod_sparse*)std::calloc(1, sizeof(cholmod_sparse));
  A->nrow = nrow; A->ncol = ncol; A->nzmax = nzmax; A->stype = stype; A->sorted = 1; A->packed = 1; A->xtype = CHOLMOD_REAL;
  A->p = std::calloc(ncol + 1, sizeof(int));
  A->i = std::calloc(std::max<size_t>(nzmax, 1), sizeof(int));
  A->x = std::calloc(std::max<size_t>(nzmax, 1), sizeof(double));
  return A;
}
cholmod_dense* alloc_dense(size_t nrow, size_t ncol) {
  cholmod_dense* X = (cholmod_dense*)std::calloc(1, sizeof(cholmod_dense));
  X->nrow = nrow; X->ncol = ncol; X->nzmax = nrow * ncol; X->d = nrow; X->xtype = CHOLMOD_REAL;
  X->x = std::calloc(std::max<size_t>(nrow * ncol, 1), sizeof(double));
  return X;
}

// C = A * B (all entries, sorted columns)
cholmod_sparse* multiply(const cholmod_sparse* A, const cholmod_sparse* B, bool upper_only) {
  const int m = (int)A->nrow, n = (int)B->ncol;
  const int *Ap = (const int*)A->p, *Ai = (const int*)A->i, *Bp = (const int*)B->p, *Bi = (const int*)B->i;
  const double *Ax = (const double*)A->x, *Bx = (const double*)B->x;
  std::vector<int> Cp(n + 1, 0), Ci, mark(m, -1), rows;
  std::vector<double> Cx, acc(m, 0.0);
  for (int j = 0; j < n; j++) {
    rows.clear();
    for (int q = Bp[j]; q < Bp[j + 1]; q++) {
      const int k = Bi[q];
      const double b = Bx[q];
      for (int t = Ap[k]; t < Ap[k + 1]; t++) {
        const int i = Ai[t];
        if (upper_only && i > j) continue;
        if (mark[i] != j) { mark[i] = j; acc[i] = 0.0; rows.push_back(i); }
        acc[i] += Ax[t] * b;
      }
    }
    std::sort(rows.begin(), rows.end());
    for (int i : rows) { Ci.push_back(i); Cx.push_back(acc[i]); }
    Cp[j + 1] = (int)Ci.size();
  }
  cholmod_sparse* C = alloc_sparse(m, n, Ci.size(), upper_only ? 1 : 0);
  std::memcpy(C->p, Cp.data(), (n + 1) * sizeof(int));
  if (!Ci.empty()) { std::memcpy(C->i, Ci.data(), Ci.size() * sizeof(int)); std::memcpy(C->x, Cx.data(), Cx.size() * sizeof(double)); }
  return C;
}
cholmod_sparse* transpose(const cholmod_sparse* A) {
  const int m = (int)A->nrow, n = (int)A->ncol;
  const int *Ap = (const int*)A->p, *Ai = (const int*)A->i;
  const double* Ax = (const double*)A->x;
  cholmod_sparse* T = alloc_sparse(n, m, Ap[n], 0);
  int *Tp = (int*)T->p, *Ti = (int*)T->i;
  double* Tx = (double*)T->x;
  std::vector<int> cnt(m + 1, 0);
  for (int q = 0; q < Ap[n]; q++) cnt[Ai[q] + 1]++;
  for (int i = 0; i < m; i++) cnt[i + 1] += cnt[i];
  std::memcpy(Tp, cnt.data(), (m + 1) * sizeof(int));
  std::vector<int> fill(cnt.begin(), cnt.end() - 1);
  for (int j = 0; j < n; j++)
    for (int q = Ap[j]; q < Ap[j + 1]; q++) { const int at = fill[Ai[q]]++; Ti[at] = j; Tx[at] = Ax[q]; }
  return T;
}

// fill-reducing ordering of a symmetric pattern given by its upper triangle: variables with identical adjacency (the
// scalar columns of one pose / plane) are merged, then greedy minimum weighted external degree on the explicit
// elimination graph, ties broken by the smallest index (deterministic)
std::vector<int> min_degree(int n, const int* Up, const int* Ui) {
  std::vector<std::vector<int>> adj(n);
  for (int j = 0; j < n; j++)
    for (int q = Up[j]; q < Up[j + 1]; q++) { const int i = Ui[q]; if (i != j) { adj[i].push_back(j); adj[j].push_back(i); } }
  for (int j = 0; j < n; j++) { adj[j].push_back(j); std::sort(adj[j].begin(), adj[j].end()); adj[j].erase(std::unique(adj[j].begin(), adj[j].end()), adj[j].end()); }
  // supervariables: identical closed neighbourhoods
  std::map<std::vector<int>, int> seen;
  std::vector<int> sv_of(n), rep;
  std::vector<std::vector<int>> members;
  for (int j = 0; j < n; j++) {
    auto it = seen.find(adj[j]);
    if (it == seen.end()) { seen[adj[j]] = (int)rep.size(); sv_of[j] = (int)rep.size(); rep.push_back(j); members.push_back({j}); }
    else { sv_of[j] = it->second; members[it->second].push_back(j); }
  }
  const int ns = (int)rep.size();
  std::vector<std::vector<int>> g(ns);
  std::vector<int> w(ns);
  for (int s = 0; s < ns; s++) {
    w[s] = (int)members[s].size();
    for (int v : adj[rep[s]]) if (sv_of[v] != s) g[s].push_back(sv_of[v]);
    std::sort(g[s].begin(), g[s].end());
    g[s].erase(std::unique(g[s].begin(), g[s].end()), g[s].end());
  }
  std::vector<long> deg(ns);
  std::set<std::pair<long, int>> pq;
  for (int s = 0; s < ns; s++) { long d = 0; for (int u : g[s]) d += w[u]; deg[s] = d; pq.insert({d, s}); }
  std::vector<char> done(ns, 0);
  std::vector<int> order, tmp;
  order.reserve(n);
  while (!pq.empty()) {
    const int v = pq.begin()->second;
    pq.erase(pq.begin());
    done[v] = 1;
    for (int x : members[v]) order.push_back(x);
    const std::vector<int> nb = g[v];
    for (int u : nb) {
      pq.erase({deg[u], u});
      tmp.clear();
      std::set_union(g[u].begin(), g[u].end(), nb.begin(), nb.end(), std::back_inserter(tmp));
      g[u].clear();
      long d = 0;
      for (int t : tmp) if (t != u && t != v) { g[u].push_back(t); d += w[t]; }
      deg[u] = d;
      pq.insert({d, u});
    }
    g[v].clear(); g[v].shrink_to_fit();
  }
  return order;
}

// numeric factor of the permuted symmetric matrix C (upper triangle, sorted) -- up-looking Cholesky
struct Numeric { std::vector<int> Lp, Li; std::vector<double> Lx; };
bool chol_up(int n, const std::vector<int>& Cp, const std::vector<int>& Ci, const std::vector<double>& Cx, std::vector<int>& parent, Numeric& N) {
  // elimination tree
  parent.assign(n, -1);
  std::vector<int> anc(n, -1);
  for (int k = 0; k < n; k++)
    for (int q = Cp[k]; q < Cp[k + 1]; q++) {
      int i = Ci[q];
      while (i != -1 && i < k) { const int nx = anc[i]; anc[i] = k; if (nx == -1) parent[i] = k; i = nx; }
    }
  std::vector<std::vector<int>> Lrow(n);     // column j: row indices (diagonal first, increasing)
  std::vector<std::vector<double>> Lval(n);
  std::vector<double> x(n, 0.0);
  std::vector<int> mark(n, -1), stack(n), path(n);
  bool ok = true;
  for (int k = 0; k < n; k++) {
    // nonzero pattern of row k of L = reach of the entries of C(0:k-1, k) in the elimination tree, in topological order
    int top = n;
    mark[k] = k;
    double d = 0.0;
    for (int q = Cp[k]; q < Cp[k + 1]; q++) {
      int i = Ci[q];
      if (i > k) continue;
      if (i == k) { d = Cx[q]; continue; }
      x[i] = Cx[q];
      int len = 0;
      for (; mark[i] != k; i = parent[i]) { path[len++] = i; mark[i] = k; }
      while (len > 0) stack[--top] = path[--len];
    }
    for (; top < n; top++) {
      const int j = stack[top];
      const double lkj = x[j] / Lval[j][0];
      x[j] = 0.0;
      for (size_t t = 1; t < Lrow[j].size(); t++) x[Lrow[j][t]] -= Lval[j][t] * lkj;
      d -= lkj * lkj;
      Lrow[j].push_back(k);
      Lval[j].push_back(lkj);
    }
    if (!(d > 0.0)) { ok = false; d = std::fabs(d) > 0 ? std::fabs(d) : 1e-300; }
    Lrow[k].push_back(k);
    Lval[k].push_back(std::sqrt(d));
  }
  N.Lp.assign(n + 1, 0);
  for (int j = 0; j < n; j++) N.Lp[j + 1] = N.Lp[j] + (int)Lrow[j].size();
  N.Li.resize(N.Lp[n]); N.Lx.resize(N.Lp[n]);
  for (int j = 0; j < n; j++) { std::copy(Lrow[j].begin(), Lrow[j].end(), N.Li.begin() + N.Lp[j]); std::copy(Lval[j].begin(), Lval[j].end(), N.Lx.begin() + N.Lp[j]); }
  return ok;
}
}  // namespace

extern "C" {

int cholmod_start(cholmod_common* c) { std::memset(c, 0, sizeof(*c)); return 1; }
int cholmod_finish(cholmod_common*) { return 1; }
cholmod_sparse* cholmod_allocate_sparse(size_t nrow, size_t ncol, size_t nzmax, int sorted, int packed, int stype, int, cholmod_common*) {
  cholmod_sparse* A = alloc_sparse(nrow, ncol, nzmax, stype);
  A->sorted = sorted; A->packed = packed;
  return A;
}
int cholmod_free_sparse(cholmod_sparse** A, cholmod_common*) {
  if (A && *A) { std::free((*A)->p); std::free((*A)->i); std::free((*A)->x); std::free(*A); *A = nullptr; }
  return 1;
}
cholmod_sparse* cholmod_transpose(cholmod_sparse* A, int, cholmod_common*) { return transpose(A); }
cholmod_sparse* cholmod_ssmult(cholmod_sparse* A, cholmod_sparse* B, int stype, int, int, cholmod_common*) {
  if (stype < 0) die("ssmult: lower-triangular result not provided");
  return multiply(A, B, stype > 0);
}

cholmod_factor* cholmod_analyze(cholmod_sparse* A, cholmod_common*) {
  cholmod_sparse* G = nullptr;
  const cholmod_sparse* S = A;
  Symbolic* sy = new Symbolic();
  if (A->stype == 0) {   // unsymmetric: CHOLMOD factors A*A'
    cholmod_sparse* At = transpose(A);
    G = multiply(A, At, true);
    cholmod_free_sparse(&At, nullptr);
    S = G;
    sy->gram = true;
  } else if (A->stype < 0) {
    die("analyze: lower-triangular storage not provided");
  }
  const int n = (int)S->ncol;
  sy->perm = min_degree(n, (const int*)S->p, (const int*)S->i);
  sy->pinv.assign(n, 0);
  for (int k = 0; k < n; k++) sy->pinv[sy->perm[k]] = k;
  if (G) cholmod_free_sparse(&G, nullptr);
  cholmod_factor* L = (cholmod_factor*)std::calloc(1, sizeof(cholmod_factor));
  L->n = n; L->minor = n; L->is_ll = 1; L->xtype = CHOLMOD_PATTERN;
  L->Perm = std::malloc(std::max(n, 1) * sizeof(int));
  std::memcpy(L->Perm, sy->perm.data(), n * sizeof(int));
  L->impl = sy;
  return L;
}

int cholmod_factorize(cholmod_sparse* A, cholmod_factor* L, cholmod_common* c) {
  Symbolic* sy = (Symbolic*)L->impl;
  cholmod_sparse* G = nullptr;
  const cholmod_sparse* S = A;
  if (A->stype == 0) {
    cholmod_sparse* At = transpose(A);
    G = multiply(A, At, true);
    cholmod_free_sparse(&At, nullptr);
    S = G;
  }
  const int n = (int)S->ncol;
  const int *Sp = (const int*)S->p, *Si = (const int*)S->i;
  const double* Sx = (const double*)S->x;
  // C = P S P' (upper triangle, sorted columns)
  std::vector<std::vector<std::pair<int, double>>> cols(n);
  for (int j = 0; j < n; j++)
    for (int q = Sp[j]; q < Sp[j + 1]; q++) {
      int a = sy->pinv[Si[q]], b = sy->pinv[j];
      if (a > b) std::swap(a, b);
      cols[b].push_back({a, Sx[q]});
    }
  std::vector<int> Cp(n + 1, 0), Ci;
  std::vector<double> Cx;
  for (int j = 0; j < n; j++) {
    std::sort(cols[j].begin(), cols[j].end());
    for (auto& e : cols[j]) { Ci.push_back(e.first); Cx.push_back(e.second); }
    Cp[j + 1] = (int)Ci.size();
  }
  if (G) cholmod_free_sparse(&G, nullptr);
  Numeric N;
  const bool ok = chol_up(n, Cp, Ci, Cx, sy->parent, N);
  std::free(L->p); std::free(L->i); std::free(L->x);
  L->nzmax = N.Li.size();
  L->p = std::malloc((n + 1) * sizeof(int));
  L->i = std::malloc(std::max<size_t>(N.Li.size(), 1) * sizeof(int));
  L->x = std::malloc(std::max<size_t>(N.Lx.size(), 1) * sizeof(double));
  std::memcpy(L->p, N.Lp.data(), (n + 1) * sizeof(int));
  std::memcpy(L->i, N.Li.data(), N.Li.size() * sizeof(int));
  std::memcpy(L->x, N.Lx.data(), N.Lx.size() * sizeof(double));
  L->xtype = CHOLMOD_REAL;
  if (c) c->status = ok ? 0 : 1;
  return 1;
}
int cholmod_change_factor(int, int, int, int, int, cholmod_factor*, cholmod_common*) { return 1; }   // always simplicial LL', packed, monotonic
int cholmod_free_factor(cholmod_factor** L, cholmod_common*) {
  if (L && *L) {
    std::free((*L)->Perm); std::free((*L)->p); std::free((*L)->i); std::free((*L)->x);
    delete (Symbolic*)(*L)->impl;
    std::free(*L);
    *L = nullptr;
  }
  return 1;
}
// (CHOLMOD converts the factor to a sparse matrix and leaves the factor symbolic)
cholmod_sparse* cholmod_factor_to_sparse(cholmod_factor* L, cholmod_common*) {
  if (L->xtype != CHOLMOD_REAL) die("factor_to_sparse: the factor is symbolic");
  const int n = (int)L->n;
  cholmod_sparse* S = (cholmod_sparse*)std::calloc(1, sizeof(cholmod_sparse));
  S->nrow = n; S->ncol = n; S->nzmax = L->nzmax; S->stype = 0; S->sorted = 1; S->packed = 1; S->xtype = CHOLMOD_REAL;
  S->p = L->p; S->i = L->i; S->x = L->x;
  L->p = L->i = L->x = nullptr; L->xtype = CHOLMOD_PATTERN;
  return S;
}
cholmod_dense* cholmod_zeros(size_t nrow, size_t ncol, int, cholmod_common*) { return alloc_dense(nrow, ncol); }
int cholmod_free_dense(cholmod_dense** X, cholmod_common*) {
  if (X && *X) { std::free((*X)->x); std::free(*X); *X = nullptr; }
  return 1;
}
// Y = alpha * op(A) * X + beta * Y
int cholmod_sdmult(cholmod_sparse* A, int transp, double alpha[2], double beta[2], cholmod_dense* X, cholmod_dense* Y, cholmod_common*) {
  const int n = (int)A->ncol;
  const int *Ap = (const int*)A->p, *Ai = (const int*)A->i;
  const double* Ax = (const double*)A->x;
  double *y = (double*)Y->x;
  const double* x = (const double*)X->x;
  if (A->stype != 0) die("sdmult: symmetric storage not provided");
  for (size_t k = 0; k < Y->nrow * Y->ncol; k++) y[k] = (beta[0] == 0.0) ? 0.0 : beta[0] * y[k];
  for (size_t c = 0; c < X->ncol; c++) {
    const double* xc = x + c * X->d;
    double* yc = y + c * Y->d;
    for (int j = 0; j < n; j++)
      for (int q = Ap[j]; q < Ap[j + 1]; q++) {
        if (!transp) yc[Ai[q]] += alpha[0] * Ax[q] * xc[j];
        else yc[j] += alpha[0] * Ax[q] * xc[Ai[q]];
      }
  }
  return 1;
}
cholmod_dense* cholmod_solve(int sys, cholmod_factor* L, cholmod_dense* B, cholmod_common*) {
  const int n = (int)L->n;
  if ((int)B->nrow != n || B->ncol != 1) die("solve: one right-hand side of the factor's order expected");
  cholmod_dense* X = alloc_dense(n, 1);
  double* x = (double*)X->x;
  const double* b = (const double*)B->x;
  const int* perm = (const int*)L->Perm;
  if (sys == CHOLMOD_P) { for (int k = 0; k < n; k++) x[k] = b[perm[k]]; return X; }
  if (sys == CHOLMOD_Pt) { for (int k = 0; k < n; k++) x[perm[k]] = b[k]; return X; }
  if (L->xtype != CHOLMOD_REAL) die("solve: the factor is symbolic");
  const int *Lp = (const int*)L->p, *Li = (const int*)L->i;
  const double* Lx = (const double*)L->x;
  std::memcpy(x, b, n * sizeof(double));
  if (sys == CHOLMOD_L) {
    for (int j = 0; j < n; j++) {
      x[j] /= Lx[Lp[j]];
      for (int q = Lp[j] + 1; q < Lp[j + 1]; q++) x[Li[q]] -= Lx[q] * x[j];
    }
  } else if (sys == CHOLMOD_Lt) {
    for (int j = n - 1; j >= 0; j--) {
      for (int q = Lp[j] + 1; q < Lp[j + 1]; q++) x[j] -= Lx[q] * x[Li[q]];
      x[j] /= Lx[Lp[j]];
    }
  } else {
    die("solve: system not provided by the shim");
  }
  return X;
}

// ---- CSparse: named by the reference's disabled alternative path only ----
#define CS_STUB(name) die("CSparse function " name " reached (USE_CSPARSE is false upstream; the shim does not provide it)")
cs* cs_spalloc(int, int, int, int, int) { CS_STUB("cs_spalloc"); }
cs* cs_spfree(cs*) { CS_STUB("cs_spfree"); }
cs* cs_transpose(const cs*, int) { CS_STUB("cs_transpose"); }
cs* cs_multiply(const cs*, const cs*) { CS_STUB("cs_multiply"); }
css* cs_sqr(int, const cs*, int) { CS_STUB("cs_sqr"); }
csn* cs_qr(const cs*, const css*) { CS_STUB("cs_qr"); }
css* cs_schol(int, const cs*) { CS_STUB("cs_schol"); }
csn* cs_chol(const cs*, const css*) { CS_STUB("cs_chol"); }
css* cs_sfree(css*) { CS_STUB("cs_sfree"); }
csn* cs_nfree(csn*) { CS_STUB("cs_nfree"); }
void* cs_free(void*) { CS_STUB("cs_free"); }
int* cs_pinv(const int*, int) { CS_STUB("cs_pinv"); }
int cs_gaxpy(const cs*, const double*, double*) { CS_STUB("cs_gaxpy"); }
int cs_lsolve(const cs*, double*) { CS_STUB("cs_lsolve"); }
int cs_ltsolve(const cs*, double*) { CS_STUB("cs_ltsolve"); }
}
