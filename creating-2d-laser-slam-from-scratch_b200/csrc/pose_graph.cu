// A dependency-free 2-D pose-graph optimiser behind the ScanSolver plug-in of the mapper (host C++; no device work:
// the back end stays on the CPU in every BASELINE config, SURVEY.md §8(b) "Karto back-end plugin").
// Role of lesson6/src/spa_solver/spa_solver.cc (AddNode :65-70, AddConstraint :72-93, Compute :44-63) without its
// Eigen / SuiteSparse / sba dependencies.  Not a restatement of sba::SysSPA2d — parity unpinned by design; the same
// solver is plugged into the reference Mapper and into ours in tests.
//
// Error of constraint (i -> j, mean d, information L):  e = [ R(th_i)^T (t_j - t_i) - d_xy ;  wrap(th_j - th_i - d_th) ],
// cost = sum e^T L e.  Levenberg-Marquardt on the normal equations H dx = -g (3x3 blocks), solved by conjugate
// gradients with a block-Jacobi preconditioner; the first node is held fixed.
#include <cmath>
#include <cstring>
#include <map>
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

int solve(G *g) {
  const int n = (int)g->nodes.size();
  if (n < 2 || g->cons.empty()) return 0;
  std::vector<double> x(3 * (size_t)n);
  for (int i = 0; i < n; i++) std::memcpy(&x[3 * i], g->nodes[i].p, sizeof(double) * 3);
  double lambda = 1e-4, cur = chi2(g, x);
  g->chi_before = cur;
  g->lm_steps = 0;
  std::vector<Lin> lin(g->cons.size());
  std::vector<double> grad(3 * (size_t)n), diag(9 * (size_t)n), dinv(9 * (size_t)n), dx(3 * (size_t)n), r, z, p, Ap;
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
      // (H + lambda * diag(H)) dx = -grad with node 0 fixed, PCG with the inverse 3x3 diagonal blocks
      for (int i = 0; i < n; i++) {
        double blk[9];
        std::memcpy(blk, &diag[9 * i], sizeof(blk));
        for (int q = 0; q < 3; q++) blk[4 * q] = blk[4 * q] * (1.0 + lambda) + 1e-12;
        if (!inv3(blk, &dinv[9 * i])) std::memset(&dinv[9 * i], 0, sizeof(blk));
      }
      auto apply = [&](const std::vector<double> &v, std::vector<double> &out) {
        out.assign(v.size(), 0.0);
        for (int i = 1; i < n; i++)
          for (int a = 0; a < 3; a++) {
            double s = 0;
            for (int b = 0; b < 3; b++) s += diag[9 * i + 3 * a + b] * v[3 * i + b];
            out[3 * i + a] = s + lambda * diag[9 * i + 4 * a] * v[3 * i + a];
          }
        for (size_t k = 0; k < g->cons.size(); k++) {
          const int a = g->cons[k].a, b = g->cons[k].b;
          for (int i = 0; i < 3; i++) {
            double sab = 0, sba = 0;
            for (int j = 0; j < 3; j++) { sab += lin[k].Hab[3 * i + j] * v[3 * b + j]; sba += lin[k].Hab[3 * j + i] * v[3 * a + j]; }
            if (a != 0) out[3 * a + i] += (b != 0 ? sab : 0.0);
            if (b != 0) out[3 * b + i] += (a != 0 ? sba : 0.0);
          }
        }
      };
      auto precond = [&](const std::vector<double> &v, std::vector<double> &out) {
        out.assign(v.size(), 0.0);
        for (int i = 1; i < n; i++)
          for (int a = 0; a < 3; a++) out[3 * i + a] = dinv[9 * i + 3 * a] * v[3 * i] + dinv[9 * i + 3 * a + 1] * v[3 * i + 1] + dinv[9 * i + 3 * a + 2] * v[3 * i + 2];
      };
      std::fill(dx.begin(), dx.end(), 0.0);
      r.assign(grad.size(), 0.0);
      for (size_t i = 3; i < grad.size(); i++) r[i] = -grad[i];
      precond(r, z);
      p = z;
      double rz = 0, r0 = 0;
      for (size_t i = 0; i < r.size(); i++) { rz += r[i] * z[i]; r0 += r[i] * r[i]; }
      for (int cg = 0; cg < g->cg_iterations && rz > 0; cg++) {
        apply(p, Ap);
        double pAp = 0;
        for (size_t i = 0; i < p.size(); i++) pAp += p[i] * Ap[i];
        if (!(pAp > 0)) break;
        const double alpha = rz / pAp;
        double rr = 0;
        for (size_t i = 0; i < p.size(); i++) { dx[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; rr += r[i] * r[i]; }
        if (rr <= 1e-20 * (r0 + 1e-300)) break;
        precond(r, z);
        double rz2 = 0;
        for (size_t i = 0; i < r.size(); i++) rz2 += r[i] * z[i];
        const double beta = rz2 / rz;
        rz = rz2;
        for (size_t i = 0; i < p.size(); i++) p[i] = z[i] + beta * p[i];
      }
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
