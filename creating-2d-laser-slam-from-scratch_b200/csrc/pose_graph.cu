// A dependency-free 2-D pose-graph optimiser behind the ScanSolver plug-in of the mapper (host C++; no device work:
// the back end stays on the CPU in every BASELINE config, SURVEY.md §8(b) "Karto back-end plugin").
// Role of lesson6/src/spa_solver/spa_solver.cc (AddNode :65-70, AddConstraint :72-93, Compute :44-63) without its
// Eigen / SuiteSparse / sba dependencies.  Not a restatement of sba::SysSPA2d — parity unpinned by design; the same
// solver is plugged into the reference Mapper and into ours in tests.
//
// Error of constraint (i -> j, mean d, information L):  e = [ R(th_i)^T (t_j - t_i) - d_xy ;  wrap(th_j - th_i - d_th) ],
// cost = sum e^T L e.  Levenberg-Marquardt on the normal equations (H + lambda diag H) dx = -g, 3x3 blocks, solved
// DIRECTLY by a sparse block L D L^T factorisation (what sba's doSPA does with CHOLMOD / CSparse): a greedy
// minimum-degree elimination order computed once per solve on the graph of the constraints — which also yields every
// column's fill pattern — then a right-looking numeric factorisation per LM trial and two triangular sweeps.  A
// trajectory graph (chain + near links + loop closures) keeps its fill near-linear under that ordering: the
// 5 000-node / 10 747-constraint bench graph solves in tens of milliseconds (the earlier block-Jacobi PCG took 3.5 s).
// The first node is held fixed.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <queue>
#include <new>
#include <vector>

#include "common.cuh"

using namespace b2s;

struct b2s_pose_graph {
  struct Node { int32_t id; double p[3]; };
  struct Con { int a, b; double d[3]; double L[9]; };
  std::vector<Node> nodes;
  std::map<int32_t, int> index;
  std::vector<Con> cons;
  int lm_iterations = 40, cg_iterations = 400;
  double chi_before = 0, chi_after = 0, lm_steps = 0;
};

namespace {

typedef b2s_pose_graph G;

inline double wrap(double a) { return normalize_angle(a); }

bool inv3(const double m[9], double k[9]) {
  k[0] = m[4] * m[8] - m[5] * m[7]; k[1] = m[2] * m[7] - m[1] * m[8]; k[2] = m[1] * m[5] - m[2] * m[4];
  k[3] = m[5] * m[6] - m[3] * m[8]; k[4] = m[0] * m[8] - m[2] * m[6]; k[5] = m[2] * m[3] - m[0] * m[5];
  k[6] = m[3] * m[7] - m[4] * m[6]; k[7] = m[1] * m[6] - m[0] * m[7]; k[8] = m[0] * m[4] - m[1] * m[3];
  const double det = m[0] * k[0] + m[1] * k[3] + m[2] * k[6];
  if (fabs(det) <= 1e-300) return false;
  for (int i = 0; i < 9; i++) k[i] /= det;
  return true;
}

void con_error(const G::Con &c, const std::vector<double> &x, double e[3]) {
  const double *pa = &x[3 * c.a], *pb = &x[3 * c.b];
  const double cs = cos(pa[2]), sn = sin(pa[2]), dx = pb[0] - pa[0], dy = pb[1] - pa[1];
  e[0] = cs * dx + sn * dy - c.d[0];
  e[1] = -sn * dx + cs * dy - c.d[1];
  e[2] = wrap(pb[2] - pa[2] - c.d[2]);
}

double chi2(const G *g, const std::vector<double> &x) {
  double s = 0;
  for (const auto &c : g->cons) {
    double e[3];
    con_error(c, x, e);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) s += e[i] * c.L[3 * i + j] * e[j];
  }
  return s;
}

struct Lin {  // per constraint: Jacobians wrt node a and b, J^T L J blocks
  double Haa[9], Hab[9], Hbb[9];
};

void mat3_tAB(const double A[9], const double B[9], double out[9]) {  // A^T B
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) out[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
void mat3_AB(const double A[9], const double B[9], double out[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) out[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

// ---- sparse block L D L^T ----------------------------------------------------------------------------------------
struct Factor {
  int n = 0;
  std::vector<int> order, pos;               // elimination order / its inverse
  std::vector<std::vector<int>> rows;        // rows[k]: nodes of column order[k] below the diagonal, sorted by pos
  std::vector<std::vector<double>> blk;      // blk[k]: 9 doubles per entry of rows[k] (block (row, order[k]))
  std::vector<double> D, Dinv;               // diagonal blocks by elimination position
};

// greedy minimum degree on the elimination graph; the neighbours a node has when it is eliminated ARE its column's rows
void symbolic(Factor &F, int n, const std::vector<std::pair<int, int>> &edges) {
  F.n = n;
  std::vector<std::vector<int>> adj((size_t)n);
  for (const auto &e : edges)
    if (e.first != e.second) { adj[e.first].push_back(e.second); adj[e.second].push_back(e.first); }
  for (auto &a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
  std::vector<char> dead((size_t)n, 0);
  typedef std::pair<int, int> DI;  // (degree, node), stale entries skipped
  std::priority_queue<DI, std::vector<DI>, std::greater<DI>> pq;
  for (int i = 0; i < n; i++) pq.push(DI((int)adj[i].size(), i));
  F.order.clear(); F.pos.assign((size_t)n, -1); F.rows.assign((size_t)n, {});
  std::vector<int> merged;
  while (!pq.empty()) {
    const DI top = pq.top();
    pq.pop();
    const int v = top.second;
    if (dead[v] || top.first != (int)adj[v].size()) continue;
    dead[v] = 1;
    const int k = (int)F.order.size();
    F.pos[v] = k;
    F.order.push_back(v);
    const std::vector<int> N = adj[v];  // all alive: dead nodes are removed from their neighbours' lists below
    F.rows[k] = N;
    for (int u : N) {  // clique among the neighbours, v removed
      merged.clear();
      std::set_union(adj[u].begin(), adj[u].end(), N.begin(), N.end(), std::back_inserter(merged));
      merged.erase(std::remove_if(merged.begin(), merged.end(), [&](int w) { return w == u || w == v; }), merged.end());
      adj[u].swap(merged);
      pq.push(DI((int)adj[u].size(), u));
    }
    std::vector<int>().swap(adj[v]);
  }
  for (int k = 0; k < n; k++) std::sort(F.rows[k].begin(), F.rows[k].end(), [&](int a, int b) { return F.pos[a] < F.pos[b]; });
  F.blk.assign((size_t)n, {});
  for (int k = 0; k < n; k++) F.blk[k].assign(9 * F.rows[k].size(), 0.0);
  F.D.assign(9 * (size_t)n, 0.0);
  F.Dinv.assign(9 * (size_t)n, 0.0);
}

inline double *find_block(Factor &F, int col_pos, int row_node) {  // block (row_node, order[col_pos]); must exist
  const std::vector<int> &r = F.rows[col_pos];
  const int target = F.pos[row_node];
  size_t lo = 0, hi = r.size();
  while (lo < hi) {
    const size_t mid = (lo + hi) / 2;
    if (F.pos[r[mid]] < target) lo = mid + 1; else hi = mid;
  }
  return (lo < r.size() && r[lo] == row_node) ? &F.blk[col_pos][9 * lo] : nullptr;
}

// numeric right-looking factorisation of the matrix held in F.D / F.blk (lower part); false = not positive definite
bool factorise(Factor &F) {
  const int n = F.n;
  std::vector<double> W;
  for (int k = 0; k < n; k++) {
    double *Dk = &F.D[9 * (size_t)k];
    if (!(Dk[0] > 0.0) || !inv3(Dk, &F.Dinv[9 * (size_t)k])) return false;
    const double *Di = &F.Dinv[9 * (size_t)k];
    const std::vector<int> &R = F.rows[k];
    const size_t m = R.size();
    W.assign(F.blk[k].begin(), F.blk[k].end());  // B_ik before scaling
    for (size_t t = 0; t < m; t++) {             // L_ik = B_ik D^-1
      double *B = &F.blk[k][9 * t], L[9];
      mat3_AB(B, Di, L);
      std::memcpy(B, L, sizeof(L));
    }
    for (size_t tj = 0; tj < m; tj++) {          // A_ij -= L_ik B_jk^T for i >= j among the rows of column k
      const int j = R[tj], pj = F.pos[j];
      const double *Wj = &W[9 * tj];
      {
        const double *Li = &F.blk[k][9 * tj];
        double *T = &F.D[9 * (size_t)pj];
        for (int a = 0; a < 3; a++)
          for (int b = 0; b < 3; b++) T[3 * a + b] -= Li[3 * a] * Wj[3 * b] + Li[3 * a + 1] * Wj[3 * b + 1] + Li[3 * a + 2] * Wj[3 * b + 2];
      }
      size_t cursor = 0;
      const std::vector<int> &Rj = F.rows[pj];
      for (size_t ti = tj + 1; ti < m; ti++) {
        const int i = R[ti];
        while (cursor < Rj.size() && Rj[cursor] != i) cursor++;  // both lists are sorted by pos: one merge walk
        if (cursor == Rj.size()) return false;                   // (cannot happen: fill pattern of the elimination)
        const double *Li = &F.blk[k][9 * ti];
        double *T = &F.blk[pj][9 * cursor];
        for (int a = 0; a < 3; a++)
          for (int b = 0; b < 3; b++) T[3 * a + b] -= Li[3 * a] * Wj[3 * b] + Li[3 * a + 1] * Wj[3 * b + 1] + Li[3 * a + 2] * Wj[3 * b + 2];
      }
    }
  }
  return true;
}

void solve_factored(const Factor &F, std::vector<double> &b) {  // b indexed by node, overwritten with the solution
  const int n = F.n;
  for (int k = 0; k < n; k++) {  // L y = b
    const double *bk = &b[3 * (size_t)F.order[k]];
    for (size_t t = 0; t < F.rows[k].size(); t++) {
      const double *L = &F.blk[k][9 * t];
      double *bi = &b[3 * (size_t)F.rows[k][t]];
      for (int a = 0; a < 3; a++) bi[a] -= L[3 * a] * bk[0] + L[3 * a + 1] * bk[1] + L[3 * a + 2] * bk[2];
    }
  }
  for (int k = 0; k < n; k++) {  // D z = y
    double *bk = &b[3 * (size_t)F.order[k]];
    const double *Di = &F.Dinv[9 * (size_t)k];
    const double t0 = bk[0], t1 = bk[1], t2 = bk[2];
    for (int a = 0; a < 3; a++) bk[a] = Di[3 * a] * t0 + Di[3 * a + 1] * t1 + Di[3 * a + 2] * t2;
  }
  for (int k = n - 1; k >= 0; k--) {  // L^T x = z
    double *bk = &b[3 * (size_t)F.order[k]];
    for (size_t t = 0; t < F.rows[k].size(); t++) {
      const double *L = &F.blk[k][9 * t];
      const double *bi = &b[3 * (size_t)F.rows[k][t]];
      for (int a = 0; a < 3; a++) bk[a] -= L[a] * bi[0] + L[3 + a] * bi[1] + L[6 + a] * bi[2];
    }
  }
}

int solve(G *g) {
  const int n = (int)g->nodes.size();
  if (n < 2 || g->cons.empty()) return 0;
  std::vector<double> x(3 * (size_t)n);
  for (int i = 0; i < n; i++) std::memcpy(&x[3 * i], g->nodes[i].p, sizeof(double) * 3);
  double lambda = 1e-4, cur = chi2(g, x);
  g->chi_before = cur;
  g->lm_steps = 0;
  Factor F;
  {
    std::vector<std::pair<int, int>> edges;
    edges.reserve(g->cons.size());
    for (const auto &c : g->cons) edges.push_back(std::make_pair(c.a, c.b));
    symbolic(F, n, edges);
  }
  std::vector<Lin> lin(g->cons.size());
  std::vector<double> grad(3 * (size_t)n), diag(9 * (size_t)n), dx(3 * (size_t)n);
  for (int it = 0; it < g->lm_iterations; it++) {
    std::fill(grad.begin(), grad.end(), 0.0);
    std::fill(diag.begin(), diag.end(), 0.0);
    for (size_t k = 0; k < g->cons.size(); k++) {
      const G::Con &c = g->cons[k];
      const double *pa = &x[3 * c.a], *pb = &x[3 * c.b];
      const double cs = cos(pa[2]), sn = sin(pa[2]), ddx = pb[0] - pa[0], ddy = pb[1] - pa[1];
      double e[3];
      con_error(c, x, e);
      const double Ja[9] = {-cs, -sn, -sn * ddx + cs * ddy, sn, -cs, -cs * ddx - sn * ddy, 0, 0, -1};
      const double Jb[9] = {cs, sn, 0, -sn, cs, 0, 0, 0, 1};
      double LJa[9], LJb[9], Le[3];
      mat3_AB(c.L, Ja, LJa);
      mat3_AB(c.L, Jb, LJb);
      for (int i = 0; i < 3; i++) Le[i] = c.L[3 * i] * e[0] + c.L[3 * i + 1] * e[1] + c.L[3 * i + 2] * e[2];
      mat3_tAB(Ja, LJa, lin[k].Haa);
      mat3_tAB(Ja, LJb, lin[k].Hab);
      mat3_tAB(Jb, LJb, lin[k].Hbb);
      for (int i = 0; i < 3; i++) {
        grad[3 * c.a + i] += Ja[i] * Le[0] + Ja[3 + i] * Le[1] + Ja[6 + i] * Le[2];
        grad[3 * c.b + i] += Jb[i] * Le[0] + Jb[3 + i] * Le[1] + Jb[6 + i] * Le[2];
      }
      for (int q = 0; q < 9; q++) { diag[9 * c.a + q] += lin[k].Haa[q]; diag[9 * c.b + q] += lin[k].Hbb[q]; }
    }
    bool improved = false;
    for (int tries = 0; tries < 8 && !improved; tries++) {
      // assemble the lower part of (H + lambda diag H) in elimination order; node 0 is fixed: identity row / column
      for (int k = 0; k < n; k++) {
        const int v = F.order[k];
        double *Dk = &F.D[9 * (size_t)k];
        if (v == 0) {
          const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
          std::memcpy(Dk, I, sizeof(I));
        } else {
          std::memcpy(Dk, &diag[9 * (size_t)v], 9 * sizeof(double));
          for (int q = 0; q < 3; q++) Dk[4 * q] = Dk[4 * q] * (1.0 + lambda) + 1e-12;
        }
        std::fill(F.blk[k].begin(), F.blk[k].end(), 0.0);
      }
      for (size_t k = 0; k < g->cons.size(); k++) {
        const int a = g->cons[k].a, b = g->cons[k].b;
        if (a == 0 || b == 0 || a == b) continue;
        // H_ab = Hab (rows a, cols b); the lower-part block sits in the column of whichever node is eliminated first
        if (F.pos[a] < F.pos[b]) {  // block (b, a) = Hab^T
          double *T = find_block(F, F.pos[a], b);
          if (T) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[3 * i + j] += lin[k].Hab[3 * j + i];
        } else {                    // block (a, b) = Hab
          double *T = find_block(F, F.pos[b], a);
          if (T) for (int q = 0; q < 9; q++) T[q] += lin[k].Hab[q];
        }
      }
      if (!factorise(F)) { lambda *= 10.0; continue; }
      for (size_t i = 0; i < grad.size(); i++) dx[i] = i < 3 ? 0.0 : -grad[i];
      solve_factored(F, dx);
      dx[0] = dx[1] = dx[2] = 0.0;
      std::vector<double> xn(x);
      for (int i = 1; i < n; i++) { xn[3 * i] += dx[3 * i]; xn[3 * i + 1] += dx[3 * i + 1]; xn[3 * i + 2] = wrap(xn[3 * i + 2] + dx[3 * i + 2]); }
      const double nxt = chi2(g, xn);
      if (nxt < cur) {
        const double gain = cur - nxt;
        x.swap(xn);
        cur = nxt;
        lambda = std::max(lambda * 0.3, 1e-9);
        improved = true;
        g->lm_steps += 1;
        if (gain <= 1e-12 * (1.0 + cur)) it = g->lm_iterations;  // converged
      } else {
        lambda *= 10.0;
      }
    }
    if (!improved) break;
  }
  g->chi_after = cur;
  for (int i = 0; i < n; i++) std::memcpy(g->nodes[i].p, &x[3 * i], sizeof(double) * 3);
  return n;
}

void cb_add_node(void *u, int32_t id, const double pose[3]) {
  G *g = static_cast<G *>(u);
  if (g->index.count(id)) return;
  g->index[id] = (int)g->nodes.size();
  G::Node nd;
  nd.id = id;
  std::memcpy(nd.p, pose, sizeof(nd.p));
  g->nodes.push_back(nd);
}

void cb_add_constraint(void *u, int32_t a, int32_t b, const double diff[3], const double cov[9]) {
  G *g = static_cast<G *>(u);
  auto ia = g->index.find(a), ib = g->index.find(b);
  if (ia == g->index.end() || ib == g->index.end()) return;
  G::Con c;
  c.a = ia->second; c.b = ib->second;
  std::memcpy(c.d, diff, sizeof(c.d));
  if (!inv3(cov, c.L)) return;  // precisionMatrix = covariance.Inverse() (spa_solver.cc:83)
  g->cons.push_back(c);
}

int32_t cb_compute(void *u, int32_t cap, int32_t *ids, double *poses) {
  G *g = static_cast<G *>(u);
  solve(g);
  const int n = std::min<int>(cap, (int)g->nodes.size());
  for (int i = 0; i < n; i++) {
    ids[i] = g->nodes[i].id;
    std::memcpy(poses + 3 * i, g->nodes[i].p, sizeof(double) * 3);
  }
  return n;
}

void cb_clear(void *) {}  // SpaSolver::Clear only drops the corrections list; nodes and constraints stay

}  // namespace

extern "C" {

b2s_status b2s_pose_graph_create(b2s_pose_graph **out) {
  if (!out) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null argument");
  *out = new (std::nothrow) b2s_pose_graph();
  if (!*out) B2S_FAIL(B2S_ERR_CUDA, "out of host memory");
  return B2S_OK;
}

void b2s_pose_graph_destroy(b2s_pose_graph *g) { delete g; }

b2s_status b2s_pose_graph_as_scan_solver(b2s_pose_graph *g, b2s_scan_solver *out) {
  if (!g || !out) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null argument");
  out->user = g;
  out->add_node = cb_add_node;
  out->add_constraint = cb_add_constraint;
  out->compute = cb_compute;
  out->clear = cb_clear;
  return B2S_OK;
}

b2s_status b2s_pose_graph_set_iterations(b2s_pose_graph *g, int lm_iterations, int cg_iterations) {
  if (!g || lm_iterations < 0 || cg_iterations < 1) B2S_FAIL(B2S_ERR_BAD_PARAMS, "bad iteration counts");
  g->lm_iterations = lm_iterations;
  g->cg_iterations = cg_iterations;
  return B2S_OK;
}

b2s_status b2s_pose_graph_stats(const b2s_pose_graph *g, double out[5]) {
  if (!g || !out) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null argument");
  out[0] = (double)g->nodes.size(); out[1] = (double)g->cons.size();
  out[2] = g->chi_before; out[3] = g->chi_after; out[4] = g->lm_steps;
  return B2S_OK;
}

}  // extern "C"
