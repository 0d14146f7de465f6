// ORACLE (test infrastructure, NOT product code) -- see oracle_math.hpp header.
//
// Restatement of the iSAM graph container, linearisation, direct solve and the
// Gauss-Newton / Levenberg-Marquardt drivers exactly as pop_planar_slam uses them.
// Parity: pose-graph path pinned against the reference's sphere2500 dataset + ground truth, plane path
// "parity unpinned" (see oracle_math.hpp).
//
// Reference files followed (relative to /root/reference, ISAM = pop_planar_slam/Thirdparty/isam):
//   ISAM/isamlib/Slam.cpp:59-67,91-126,157-210,216-268,395-432
//   ISAM/isamlib/Optimizer.cpp:49-67,114-185,286-366,371-467
//   ISAM/isamlib/numericalDiff.cpp:32-87
//   ISAM/isamlib/Cholesky.cpp:86-132          (damping rule; CHOLMOD itself is a
//       SuiteSparse system library, un-vendored and unpinned -- a direct sparse
//       Cholesky is restated here: fill-reducing order + up-looking factorisation)
//   ISAM/include/isam/Factor.h:67-77,126-139, Node.h:99-154, Graph.h:40-133
//   pop_planar_slam/src/isam_plane3d.h, ISAM/include/isam/slam3d.h (via oracle_math.hpp)
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <queue>
#include <string>
#include <vector>

#include "oracle_math.hpp"

namespace orc {

enum NodeKind { NODE_POSE = 0, NODE_PLANE = 1 };
enum FactorKind { F_POSE_PRIOR = 0, F_ODOMETRY = 1, F_POSE_PLANE = 2, F_PLANE_PRIOR = 3 };
enum Method { GAUSS_NEWTON = 0, LEVENBERG_MARQUARDT = 1 };  // Properties.h (DOG_LEG unused by PPS)
enum Selector { LINPOINT = 0, ESTIMATE = 1 };                // Node.h:40
enum RobustKind { ROBUST_NONE = 0, ROBUST_HUBER = 1, ROBUST_PSEUDO_HUBER = 2 };
enum JacobianMode { JAC_NUMERIC = 0, JAC_ANALYTIC = 1 };

inline int node_dim(int kind) { return kind == NODE_POSE ? 6 : 3; }

struct Node {
  int kind = NODE_POSE;
  bool alive = true;
  bool initialized = false;
  Pose pose, pose0;     // estimate / linearisation point (Node.h:103-104)
  Plane plane, plane0;
  int start = -1;       // column offset, Slam::update_starts Slam.cpp:59-67
};

struct Factor {
  int kind = F_POSE_PLANE;
  bool alive = true;
  int nodes[2] = {-1, -1};
  int n_nodes = 1;
  int dim = 3;
  double meas[6] = {0, 0, 0, 0, 0, 0};  // plane: abcd (normalised) ; pose: x,y,z,yaw,pitch,roll
  double sqrtinf[36];                   // dim x dim row-major, upper triangular
  bool has_rays = false;                // Pose3d_Plane3d_Factor2: measurement re-popped from two ground-edge rays
  double rays[6] = {0, 0, 0, 0, 0, 0};
};

// ISAM/include/isam/Properties.h:86-109 defaults
struct Properties {
  int method = GAUSS_NEWTON;
  double epsilon2 = 1e-2;
  double epsilon_abs = 1e-3;
  double epsilon_rel = 1e-5;
  int max_iterations = 500;
  double lm_lambda0 = 1e-6;
  double lm_lambda_factor = 10.;
  int mod_update = 1;
  int mod_batch = 100;
  int mod_solve = 1;
};

struct TraceEntry {
  double lambda;
  double error_new;
  double error_before;
  double delta_norm;
  int accepted;
};

struct Timers {
  double linearize = 0, solve = 0, chi2 = 0, order = 0, total = 0;
  int n_linearize = 0, n_solve = 0, n_chi2 = 0;
};

static inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// one linearised factor: dense rows (dim x cols) + rhs
struct LinFactor {
  int fid;
  int dim;
  int n_nodes;
  int nodes[2];
  int ndim[2];
  double H[6 * 12];  // row-major, dim x (ndim0+ndim1)
  double rhs[6];     // = -r  (Jacobian.h:98)
};

class Slam {
 public:
  std::vector<Node> nodes;      // index == unique id (insertion order)
  std::vector<Factor> factors;  // index == unique id (insertion order)
  Properties prop;
  int robust_kind = ROBUST_NONE;
  double robust_b = 1.0;
  int jac_mode = JAC_NUMERIC;
  bool reuse_ordering = false;  // reference re-runs cholmod_analyze every call (Cholesky.cpp:98)
  int step = 0;
  std::vector<TraceEntry> trace;
  Timers timers;
  std::string last_error;

  // cached ordering (only when reuse_ordering)
  std::vector<int> cached_order;
  size_t cached_order_sig = 0;

  // ---------------- graph edits ----------------
  int add_pose() { Node n; n.kind = NODE_POSE; nodes.push_back(n); return (int)nodes.size() - 1; }
  int add_plane() { Node n; n.kind = NODE_PLANE; nodes.push_back(n); return (int)nodes.size() - 1; }

  // NodeT::init Node.h:123-126
  void init_pose(int id, const Pose& p) { nodes[id].pose = p; nodes[id].pose0 = p; nodes[id].initialized = true; }
  void init_plane(int id, const Plane& p) { nodes[id].plane = p; nodes[id].plane0 = p; nodes[id].initialized = true; }

  static void fill_sqrtinf(Factor& f, const double* ut) {
    // upper-triangular packed row-major (noise_to_string order, Factor.h:148-155)
    int d = f.dim, k = 0;
    for (int r = 0; r < d; r++)
      for (int c = 0; c < d; c++) f.sqrtinf[r * d + c] = (c >= r) ? ut[k++] : 0.0;
  }

  // Pose3d_Factor  slam3d.h:58-89 ; initialize :75-80
  int add_pose_prior(int pose, const double meas6[6], const double* sqrtinf_ut) {
    Factor f; f.kind = F_POSE_PRIOR; f.n_nodes = 1; f.nodes[0] = pose; f.dim = 6;
    std::memcpy(f.meas, meas6, 6 * sizeof(double));
    fill_sqrtinf(f, sqrtinf_ut);
    if (!nodes[pose].initialized) init_pose(pose, pose_from_xyzypr(meas6));
    factors.push_back(f);
    return (int)factors.size() - 1;
  }

  // Pose3d_Pose3d_Factor  slam3d.h:91-193 ; initialize :123-137
  int add_odometry(int p1, int p2, const double meas6[6], const double* sqrtinf_ut) {
    Factor f; f.kind = F_ODOMETRY; f.n_nodes = 2; f.nodes[0] = p1; f.nodes[1] = p2; f.dim = 6;
    std::memcpy(f.meas, meas6, 6 * sizeof(double));
    fill_sqrtinf(f, sqrtinf_ut);
    if (!nodes[p1].initialized && !nodes[p2].initialized) {
      last_error = "slam3d: Pose3d_Pose3d_Factor requires pose1 or pose2 to be initialized";
      return -1;
    }
    Pose m = pose_from_xyzypr(meas6);
    if (!nodes[p1].initialized && nodes[p2].initialized) {
      Pose z;
      init_pose(p1, pose_oplus(nodes[p2].pose, pose_ominus(z, m)));
    } else if (nodes[p1].initialized && !nodes[p2].initialized) {
      init_pose(p2, pose_oplus(nodes[p1].pose, m));
    }
    factors.push_back(f);
    return (int)factors.size() - 1;
  }

  // Pose3d_Plane3d_Factor  isam_plane3d.h:221-308 ; initialize :252-264
  int add_pose_plane(int pose, int plane, const double meas4[4], const double* sqrtinf_ut) {
    Factor f; f.kind = F_POSE_PLANE; f.n_nodes = 2; f.nodes[0] = pose; f.nodes[1] = plane; f.dim = 3;
    Plane m = plane_from_vec4(meas4);
    std::memcpy(f.meas, m.v, 4 * sizeof(double));
    fill_sqrtinf(f, sqrtinf_ut);
    if (!nodes[pose].initialized) {
      last_error = "Plane3d: Pose3d_Plane3d_Factor requires pose to be initialized";
      return -1;
    }
    if (!nodes[plane].initialized) {
      double T[16];
      pose_oTw(nodes[pose].pose, T);
      init_plane(plane, plane_transform_T(T, m));
    }
    factors.push_back(f);
    return (int)factors.size() - 1;
  }

  // Pose3d_Plane3d_Factor2  isam_plane3d.h:314-424 (precompute_edge_ray :358-370 supplies the two rays)
  int add_pose_plane2(int pose, int plane, const double meas4[4], const double rays6[6], const double* sqrtinf_ut) {
    int fid = add_pose_plane(pose, plane, meas4, sqrtinf_ut);
    if (fid < 0) return fid;
    factors[fid].has_rays = true;
    std::memcpy(factors[fid].rays, rays6, 6 * sizeof(double));
    return fid;
  }

  // Plane3d_Factor  isam_plane3d.h:428-474 ; initialize :443-448
  int add_plane_prior(int plane, const double meas4[4], const double* sqrtinf_ut) {
    Factor f; f.kind = F_PLANE_PRIOR; f.n_nodes = 1; f.nodes[0] = plane; f.dim = 3;
    Plane m = plane_from_vec4(meas4);
    std::memcpy(f.meas, m.v, 4 * sizeof(double));
    fill_sqrtinf(f, sqrtinf_ut);
    if (!nodes[plane].initialized) init_plane(plane, m);
    factors.push_back(f);
    return (int)factors.size() - 1;
  }

  // FactorT::set_measurement Factor.h:206 (Plane3d measure => normalised by the Plane3d ctor at the call site, Mapping.cpp:602)
  void set_measurement(int fid, const double* m) {
    Factor& f = factors[fid];
    if (f.kind == F_POSE_PLANE || f.kind == F_PLANE_PRIOR) {
      Plane p = plane_from_vec4(m);
      std::memcpy(f.meas, p.v, 4 * sizeof(double));
    } else {
      std::memcpy(f.meas, m, 6 * sizeof(double));
    }
  }

  void remove_factor(int fid) { factors[fid].alive = false; }  // Slam.cpp:117-126
  void remove_node(int id) {                                    // Slam.cpp:107-115
    for (size_t i = 0; i < factors.size(); i++) {
      Factor& f = factors[i];
      if (!f.alive) continue;
      for (int k = 0; k < f.n_nodes; k++)
        if (f.nodes[k] == id) { f.alive = false; break; }
    }
    nodes[id].alive = false;
  }

  int num_nodes() const { int c = 0; for (auto& n : nodes) c += n.alive; return c; }
  int num_factors() const { int c = 0; for (auto& f : factors) c += f.alive; return c; }

  // Slam::update_starts Slam.cpp:59-67
  int update_starts() {
    int start = 0;
    for (auto& n : nodes) {
      if (!n.alive) { n.start = -1; continue; }
      n.start = start;
      start += node_dim(n.kind);
    }
    return start;
  }
  int dim_measure() const { int c = 0; for (auto& f : factors) if (f.alive) c += f.dim; return c; }

  // ---------------- error ----------------
  // Factor::basic_error dispatch
  void basic_error(const Factor& f, int s, double* e) const {
    switch (f.kind) {
      case F_POSE_PRIOR: {
        const Node& n = nodes[f.nodes[0]];
        pose_prior_basic_error(s == ESTIMATE ? n.pose : n.pose0, f.meas, e);
      } break;
      case F_ODOMETRY: {
        const Node& a = nodes[f.nodes[0]];
        const Node& b = nodes[f.nodes[1]];
        odometry_basic_error(s == ESTIMATE ? a.pose : a.pose0, s == ESTIMATE ? b.pose : b.pose0, f.meas, e);
      } break;
      case F_POSE_PLANE: {
        const Node& a = nodes[f.nodes[0]];
        const Node& b = nodes[f.nodes[1]];
        Plane m; std::memcpy(m.v, f.meas, 4 * sizeof(double));
        if (f.has_rays) pose_plane2_basic_error(s == ESTIMATE ? a.pose : a.pose0, s == ESTIMATE ? b.plane : b.plane0, f.rays, e);
        else pose_plane_basic_error(s == ESTIMATE ? a.pose : a.pose0, s == ESTIMATE ? b.plane : b.plane0, m, e);
      } break;
      case F_PLANE_PRIOR: {
        const Node& a = nodes[f.nodes[0]];
        Plane m; std::memcpy(m.v, f.meas, 4 * sizeof(double));
        plane_log_error(s == ESTIMATE ? a.plane : a.plane0, m, e);
      } break;
    }
  }

  double cost(double v) const {
    return robust_kind == ROBUST_HUBER ? cost_huber(v, robust_b) : cost_pseudo_huber(v, robust_b);
  }

  // Factor::error  Factor.h:67-77
  void error(const Factor& f, int s, double* err) const {
    double e[6];
    basic_error(f, s, e);
    int d = f.dim;
    for (int r = 0; r < d; r++) {
      double acc = 0;
      for (int c = 0; c < d; c++) acc += f.sqrtinf[r * d + c] * e[c];
      err[r] = acc;
    }
    if (robust_kind != ROBUST_NONE) {
      for (int i = 0; i < d; i++) {
        double val = err[i];
        err[i] = ((val >= 0) ? 1. : (-1.)) * std::sqrt(cost(val));
      }
    }
  }

  // Slam::weighted_errors / chi2  Slam.cpp:254-268
  double chi2(int s) {
    double t0 = now_s();
    double acc = 0;
    double err[6];
    for (auto& f : factors) {
      if (!f.alive) continue;
      error(f, s, err);
      for (int i = 0; i < f.dim; i++) acc += err[i] * err[i];
    }
    timers.chi2 += now_s() - t0;
    timers.n_chi2++;
    return acc;
  }

  // ---------------- node-level ops (Node.h:141-146, Slam.cpp:216-252) ----------------
  void node_self_exmap(Node& n, const double* d) {
    if (n.kind == NODE_POSE) n.pose0 = pose_exmap(n.pose0, d); else n.plane0 = plane_exmap(n.plane0, d);
  }
  void node_apply_exmap(Node& n, const double* d) {
    if (n.kind == NODE_POSE) n.pose = pose_exmap(n.pose0, d); else n.plane = plane_exmap(n.plane0, d);
  }
  void self_exmap(const std::vector<double>& x) {
    int pos = 0;
    for (auto& n : nodes) { if (!n.alive) continue; node_self_exmap(n, &x[pos]); pos += node_dim(n.kind); }
  }
  void apply_exmap(const std::vector<double>& x) {
    int pos = 0;
    for (auto& n : nodes) { if (!n.alive) continue; node_apply_exmap(n, &x[pos]); pos += node_dim(n.kind); }
  }
  void linpoint_to_estimate() { for (auto& n : nodes) { n.pose = n.pose0; n.plane = n.plane0; } }
  void estimate_to_linpoint() { for (auto& n : nodes) { n.pose0 = n.pose; n.plane0 = n.plane; } }

  // ---------------- Jacobians ----------------
  // numericalDiff  ISAM/isamlib/numericalDiff.cpp:41-87 (SYMMETRIC, epsilon = 1e-4),
  // including the restore through update0(vector0()) (Euler round trip for poses,
  // re-normalisation for planes).
  void numerical_jacobian(const Factor& f, double* H /* dim x ncols row-major */, int ncols) {
    const double epsilon = 0.0001;
    int col = 0;
    double yp[6], ym[6];
    for (int k = 0; k < f.n_nodes; k++) {
      Node& n = nodes[f.nodes[k]];
      int dn = node_dim(n.kind);
      for (int j = 0; j < dn; j++, col++) {
        double delta[6] = {0, 0, 0, 0, 0, 0};
        double orig[6];
        if (n.kind == NODE_POSE) pose_vector(n.pose0, orig); else std::memcpy(orig, n.plane0.v, 4 * sizeof(double));
        delta[j] = epsilon;
        node_self_exmap(n, delta);
        error(f, LINPOINT, yp);
        if (n.kind == NODE_POSE) n.pose0 = pose_set_vector(orig); else n.plane0 = plane_from_vec4(orig);
        delta[j] = -epsilon;
        node_self_exmap(n, delta);
        error(f, LINPOINT, ym);
        if (n.kind == NODE_POSE) n.pose0 = pose_set_vector(orig); else n.plane0 = plane_from_vec4(orig);
        for (int r = 0; r < f.dim; r++) H[r * ncols + col] = (yp[r] - ym[r]) / (epsilon + epsilon);
      }
    }
  }

  // Closed-form Jacobians of the same error functions through the same exmaps
  // (SURVEY.md Appendix A).  Not in the reference (it only differentiates
  // numerically); used to pin the restatement and the CUDA kernels' blocks.
  void analytic_jacobian(const Factor& f, double* H, int ncols, double* r_out) const;

  void linearize_factor(int fid, LinFactor& lf) {
    Factor& f = factors[fid];
    lf.fid = fid; lf.dim = f.dim; lf.n_nodes = f.n_nodes;
    int ncols = 0;
    for (int k = 0; k < f.n_nodes; k++) { lf.nodes[k] = f.nodes[k]; lf.ndim[k] = node_dim(nodes[f.nodes[k]].kind); ncols += lf.ndim[k]; }
    double r[6];
    if (jac_mode == JAC_NUMERIC || f.has_rays) {   // (Factor2 has no independent closed form here: numericalDiff as upstream)
      numerical_jacobian(f, lf.H, ncols);
      error(f, LINPOINT, r);  // Factor::jacobian Factor.h:127-128 (after numericalDiff)
    } else {
      analytic_jacobian(f, lf.H, ncols, r);
    }
    for (int i = 0; i < f.dim; i++) lf.rhs[i] = -r[i];
  }

  // Slam::jacobian_partial(-1)  Slam.cpp:395-432
  void jacobian(std::vector<LinFactor>& J) {
    double t0 = now_s();
    update_starts();
    J.clear();
    for (size_t i = 0; i < factors.size(); i++) {
      if (!factors[i].alive) continue;
      LinFactor lf;
      linearize_factor((int)i, lf);
      J.push_back(lf);
    }
    timers.linearize += now_s() - t0;
    timers.n_linearize++;
  }

  // ---------------- direct solve ----------------
  // Optimizer::compute_gauss_newton_step Optimizer.cpp:49-67 -> CholeskyImpl::factorize
  // Cholesky.cpp:68-147: solve (J'J + lambda*diag(J'J)) delta = J' rhs.
  // Output delta in default (insertion) ordering.
  bool solve(const std::vector<LinFactor>& J, double lambda, std::vector<double>& delta);

  // assembled normal equations in insertion ordering (upper triangle CSC), for tests
  void normal_equations(const std::vector<LinFactor>& J, double lambda, std::vector<int>& Ap, std::vector<int>& Ai,
                        std::vector<double>& Ax, std::vector<double>& b);

  // ---------------- drivers ----------------
  // Optimizer::levenberg_marquardt  Optimizer.cpp:371-467
  int levenberg_marquardt() {
    int num_iter = 0;
    double lambda = prop.lm_lambda0;
    estimate_to_linpoint();
    std::vector<LinFactor> J;
    jacobian(J);
    double error = chi2(LINPOINT);
    double error_diff, error_new;
    std::vector<double> delta;
    solve(J, lambda, delta);
    while (((prop.max_iterations <= 0) || (num_iter < prop.max_iterations)) && (norm2(delta) > prop.epsilon2) &&
           (error > prop.epsilon_abs)) {
      num_iter++;
      linpoint_to_estimate();
      self_exmap(delta);
      error_new = chi2(LINPOINT);
      error_diff = error - error_new;
      TraceEntry te{lambda, error_new, error, norm2(delta), error_diff > 0. ? 1 : 0};
      trace.push_back(te);
      if (error_diff > 0.) {
        if (error_diff < prop.epsilon_rel * error) break;
        lambda /= prop.lm_lambda_factor;
        error = error_new;
        jacobian(J);
      } else {
        lambda *= prop.lm_lambda_factor;
        estimate_to_linpoint();
      }
      solve(J, lambda, delta);
    }
    linpoint_to_estimate();
    return num_iter;
  }

  // Optimizer::gauss_newton  Optimizer.cpp:286-366
  int gauss_newton() {
    int num_iter = 0;
    estimate_to_linpoint();
    std::vector<LinFactor> J;
    jacobian(J);
    double error = chi2(LINPOINT);
    double error_new;
    double error_diff = prop.epsilon_rel * error + 1;
    std::vector<double> delta;
    solve(J, 0.0, delta);
    while (((prop.max_iterations <= 0) || (num_iter < prop.max_iterations)) && (norm2(delta) > prop.epsilon2) &&
           (error > prop.epsilon_abs) && (std::fabs(error_diff) > prop.epsilon_rel * error)) {
      num_iter++;
      apply_exmap(delta);
      estimate_to_linpoint();
      jacobian(J);
      error_new = chi2(LINPOINT);
      error_diff = error - error_new;
      TraceEntry te{0.0, error_new, error, norm2(delta), 1};
      trace.push_back(te);
      error = error_new;
      solve(J, 0.0, delta);
    }
    return num_iter;
  }

  // Slam::batch_optimization Slam.cpp:198-210 -> Optimizer::batch_optimize Optimizer.cpp:538-555
  int batch_optimization() {
    double t0 = now_s();
    trace.clear();
    int it = 0;
    if (prop.method == LEVENBERG_MARQUARDT) it = levenberg_marquardt(); else it = gauss_newton();
    timers.total += now_s() - t0;
    return it;
  }

  // Slam::update Slam.cpp:157-196.  PPS sets mod_batch = 1 (Mapping.cpp:35) so
  // every call is batch_optimization_step -> Optimizer::relinearize
  // (Optimizer.cpp:114-185): relinearise, one un-damped GN step, apply_exmap.
  // Other mod_batch values would take the Givens incremental path, which PPS
  // never reaches; it is not restated (returns -1).
  int update() {
    int rc = 0;
    if (step % prop.mod_update == 0) {
      if (step % prop.mod_batch == 0) {
        estimate_to_linpoint();
        std::vector<LinFactor> J;
        jacobian(J);
        std::vector<double> h_gn;
        solve(J, 0.0, h_gn);
        apply_exmap(h_gn);
      } else {
        last_error = "oracle: incremental (Givens) update path not restated; PPS uses mod_batch=1";
        rc = -1;
      }
    }
    step++;
    return rc;
  }

  static double norm2(const std::vector<double>& v) { double s = 0; for (double x : v) s += x * x; return std::sqrt(s); }

 private:
  void min_degree_order(const std::vector<LinFactor>& J, const std::vector<int>& live, std::vector<int>& order);
};

// ---------------------------------------------------------------------------
// analytic Jacobians (derivations in DESIGN.md)
// ---------------------------------------------------------------------------
namespace detail {
inline void skew(const double v[3], double S[9]) {
  S[0] = 0; S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2]; S[4] = 0; S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
// C(m x n) = A(m x k) * B(k x n), row-major
inline void matmul(const double* A, const double* B, double* C, int m, int k, int n) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      double s = 0;
      for (int l = 0; l < k; l++) s += A[i * k + l] * B[l * n + j];
      C[i * n + j] = s;
    }
}
// log-map part: given local plane l (unit 4) and measurement m, e and G4 = de/dl (3x4)
inline void plane_log_and_grad(const double l[4], const double m[4], double e[3], double G4[12]) {
  // M(m) = [[ m_w I + [m_v]x , -m_v ],[ m_v^T , m_w ]]
  double M[16];
  double mv[3] = {m[0], m[1], m[2]}, mw = m[3];
  double S[9]; skew(mv, S);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) M[i * 4 + j] = (i == j ? mw : 0.0) + S[i * 3 + j];
    M[i * 4 + 3] = -mv[i];
    M[12 + i] = mv[i];
  }
  M[15] = mw;
  double dq[4];
  matmul(M, l, dq, 4, 4, 1);
  if (dq[3] < 0) { for (int i = 0; i < 4; i++) dq[i] = -dq[i]; for (int i = 0; i < 16; i++) M[i] = -M[i]; }
  double a[3] = {dq[0], dq[1], dq[2]}, w = dq[3];
  double na = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  double D[12];  // [de/da (3x3) | de/dw (3x1)] as 3x4
  double den = na * na + w * w;
  if (na > 1e-12) {
    double theta = 2.0 * std::atan2(na, w);
    double ah[3] = {a[0] / na, a[1] / na, a[2] / na};
    for (int i = 0; i < 3; i++) e[i] = theta * ah[i];
    double c1 = theta / na, c2 = 2.0 * w / den, c3 = -2.0 * na / den;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) D[i * 4 + j] = c1 * ((i == j ? 1.0 : 0.0) - ah[i] * ah[j]) + c2 * ah[i] * ah[j];
      D[i * 4 + 3] = c3 * ah[i];
    }
  } else {
    // limit na -> 0: e = (2/w) a, de/da = (2/w) I, de/dw = 0
    for (int i = 0; i < 3; i++) e[i] = 2.0 / w * a[i];
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) D[i * 4 + j] = (i == j ? 2.0 * w / den : 0.0);
      D[i * 4 + 3] = 0.0;
    }
  }
  matmul(D, M, G4, 3, 4, 4);
}
// E_b^{-1}(pitch, roll): (yaw,pitch,roll) rates from body rates
inline void euler_rate_inv(double p, double r, double Ei[9]) {
  double cp = std::cos(p), sr = std::sin(r), cr = std::cos(r), tp = std::tan(p);
  Ei[0] = 0; Ei[1] = sr / cp; Ei[2] = cr / cp;
  Ei[3] = 0; Ei[4] = cr;      Ei[5] = -sr;
  Ei[6] = 1; Ei[7] = sr * tp; Ei[8] = cr * tp;
}
}  // namespace detail

inline void Slam::analytic_jacobian(const Factor& f, double* H, int ncols, double* r_out) const {
  using namespace detail;
  double e[6];
  double Jb[6 * 12];  // basic (unweighted) jacobian
  std::memset(Jb, 0, sizeof(Jb));
  if (f.kind == F_POSE_PLANE || f.kind == F_PLANE_PRIOR) {
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
    const Plane* gp;
    if (f.kind == F_POSE_PLANE) {
      const Pose& p = nodes[f.nodes[0]].pose0;
      quat_to_wRo(p.q, R);
      t[0] = p.t[0]; t[1] = p.t[1]; t[2] = p.t[2];
      gp = &nodes[f.nodes[1]].plane0;
    } else {
      gp = &nodes[f.nodes[0]].plane0;
    }
    const double* pi = gp->v;
    double n[3] = {pi[0], pi[1], pi[2]}, d = pi[3];
    // u = [R^T n ; t.n + d]
    double u[4];
    for (int i = 0; i < 3; i++) u[i] = R[0 * 3 + i] * n[0] + R[1 * 3 + i] * n[1] + R[2 * 3 + i] * n[2];
    u[3] = t[0] * n[0] + t[1] * n[1] + t[2] * n[2] + d;
    double s = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + u[3] * u[3]);
    double l[4] = {u[0] / s, u[1] / s, u[2] / s, u[3] / s};
    double G4[12];
    plane_log_and_grad(l, f.meas, e, G4);
    // G = G4 * (I - l l^T)/s   (3x4)
    double Pn[16];
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) Pn[i * 4 + j] = ((i == j ? 1.0 : 0.0) - l[i] * l[j]) / s;
    double G[12];
    matmul(G4, Pn, G, 3, 4, 4);
    // du/dplane = [[R^T,0],[t^T,1]] * 1/2 [[ d I - [n]x ],[ -n^T ]]   (4x3)
    double A[16] = {R[0], R[3], R[6], 0, R[1], R[4], R[7], 0, R[2], R[5], R[8], 0, t[0], t[1], t[2], 1};
    double Sn[9]; skew(n, Sn);
    double B[12];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) B[i * 3 + j] = 0.5 * ((i == j ? d : 0.0) - Sn[i * 3 + j]);
    for (int j = 0; j < 3; j++) B[9 + j] = -0.5 * n[j];
    double AB[12], Jl[9];
    matmul(A, B, AB, 4, 4, 3);
    matmul(G, AB, Jl, 3, 4, 3);
    if (f.kind == F_POSE_PLANE) {
      // du/dpose = [[ 0 , [R^T n]x ],[ n^T , 0 ]]   (4x6)
      double C[24];
      std::memset(C, 0, sizeof(C));
      double Su[9]; skew(u, Su);
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 6 + 3 + j] = Su[i * 3 + j];
      for (int j = 0; j < 3; j++) C[18 + j] = n[j];
      double Jp[18];
      matmul(G, C, Jp, 3, 4, 6);
      for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 6; j++) Jb[i * ncols + j] = Jp[i * 6 + j];
        for (int j = 0; j < 3; j++) Jb[i * ncols + 6 + j] = Jl[i * 3 + j];
      }
    } else {
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Jb[i * ncols + j] = Jl[i * 3 + j];
    }
  } else if (f.kind == F_ODOMETRY) {
    const Pose& p1 = nodes[f.nodes[0]].pose0;
    const Pose& p2 = nodes[f.nodes[1]].pose0;
    odometry_basic_error(p1, p2, f.meas, e);
    double R1[9], R2[9];
    quat_to_wRo(p1.q, R1); quat_to_wRo(p2.q, R2);
    double R1t[9] = {R1[0], R1[3], R1[6], R1[1], R1[4], R1[7], R1[2], R1[5], R1[8]};
    double R12[9];
    matmul(R1t, R2, R12, 3, 3, 3);
    double dt[3] = {p2.t[0] - p1.t[0], p2.t[1] - p1.t[1], p2.t[2] - p1.t[2]};
    double t12[3];
    matmul(R1t, dt, t12, 3, 3, 1);
    double pitch = std::asin(-R12[6]);
    double roll = std::atan2(R12[7], R12[8]);
    double Ei[9]; euler_rate_inv(pitch, roll, Ei);
    double R12t[9] = {R12[0], R12[3], R12[6], R12[1], R12[4], R12[7], R12[2], R12[5], R12[8]};
    double EiR[9]; matmul(Ei, R12t, EiR, 3, 3, 3);
    double St[9]; skew(t12, St);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        Jb[i * ncols + j] = -R1t[i * 3 + j];
        Jb[i * ncols + 3 + j] = St[i * 3 + j];
        Jb[(3 + i) * ncols + 3 + j] = -EiR[i * 3 + j];
        Jb[i * ncols + 6 + j] = R1t[i * 3 + j];
        Jb[(3 + i) * ncols + 9 + j] = Ei[i * 3 + j];
      }
  } else {  // F_POSE_PRIOR
    const Pose& p = nodes[f.nodes[0]].pose0;
    pose_prior_basic_error(p, f.meas, e);
    double v[6]; pose_vector(p, v);
    double Ei[9]; euler_rate_inv(v[4], v[5], Ei);
    for (int i = 0; i < 3; i++) {
      Jb[i * ncols + i] = 1.0;
      for (int j = 0; j < 3; j++) Jb[(3 + i) * ncols + 3 + j] = Ei[i * 3 + j];
    }
  }
  // weight: r = S e ; J = S Jb ; robust row scale (SURVEY Appendix A.1 last lines)
  int d = f.dim;
  double r[6];
  for (int i = 0; i < d; i++) {
    double acc = 0;
    for (int c = 0; c < d; c++) acc += f.sqrtinf[i * d + c] * e[c];
    r[i] = acc;
    for (int j = 0; j < ncols; j++) {
      double a2 = 0;
      for (int c = 0; c < d; c++) a2 += f.sqrtinf[i * d + c] * Jb[c * ncols + j];
      H[i * ncols + j] = a2;
    }
  }
  if (robust_kind != ROBUST_NONE) {
    for (int i = 0; i < d; i++) {
      double val = r[i], av = std::fabs(val), wgt = 1.0, rho = cost(val);
      if (robust_kind == ROBUST_HUBER) {
        if (av >= robust_b) wgt = robust_b / std::sqrt(2 * robust_b * av - robust_b * robust_b);
      } else {
        double sq = std::sqrt(1 + val * val / (robust_b * robust_b));
        wgt = (av > 1e-150) ? av / (sq * std::sqrt(rho)) : 1.0;
      }
      r[i] = ((val >= 0) ? 1. : (-1.)) * std::sqrt(rho);
      for (int j = 0; j < ncols; j++) H[i * ncols + j] *= wgt;
    }
  }
  for (int i = 0; i < d; i++) r_out[i] = r[i];
}

// ---------------------------------------------------------------------------
// direct sparse solve
// ---------------------------------------------------------------------------
// Greedy minimum-degree elimination order on the node (block) graph -- the
// stand-in for cholmod_analyze's fill-reducing ordering (Cholesky.cpp:98).
inline void Slam::min_degree_order(const std::vector<LinFactor>& J, const std::vector<int>& live, std::vector<int>& order) {
  int n = (int)live.size();
  std::vector<int> pos(nodes.size(), -1);
  for (int i = 0; i < n; i++) pos[live[i]] = i;
  std::vector<std::vector<int>> adj(n);
  for (auto& lf : J)
    if (lf.n_nodes == 2) {
      int a = pos[lf.nodes[0]], b = pos[lf.nodes[1]];
      adj[a].push_back(b); adj[b].push_back(a);
    }
  std::vector<int> w(n);
  for (int i = 0; i < n; i++) w[i] = node_dim(nodes[live[i]].kind);
  auto uniq = [](std::vector<int>& v) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); };
  for (auto& a : adj) uniq(a);
  std::vector<char> done(n, 0);
  auto degree = [&](int v) { long d = 0; for (int u : adj[v]) d += w[u]; return d; };
  typedef std::pair<long, int> PI;
  std::priority_queue<PI, std::vector<PI>, std::greater<PI>> pq;
  std::vector<long> curdeg(n);
  for (int i = 0; i < n; i++) { curdeg[i] = degree(i); pq.push(PI(curdeg[i], i)); }
  order.clear(); order.reserve(n);
  std::vector<int> merged;
  while (!pq.empty()) {
    PI top = pq.top(); pq.pop();
    int v = top.second;
    if (done[v] || top.first != curdeg[v]) continue;
    done[v] = 1;
    order.push_back(v);
    const std::vector<int>& N = adj[v];
    for (int u : N) {
      merged.clear();
      merged.reserve(adj[u].size() + N.size());
      std::set_union(adj[u].begin(), adj[u].end(), N.begin(), N.end(), std::back_inserter(merged));
      std::vector<int>& au = adj[u];
      au.clear();
      for (int x : merged) if (x != u && x != v && !done[x]) au.push_back(x);
      curdeg[u] = degree(u);
      pq.push(PI(curdeg[u], u));
    }
    std::vector<int>().swap(adj[v]);
  }
}

inline void Slam::normal_equations(const std::vector<LinFactor>& J, double lambda, std::vector<int>& Ap, std::vector<int>& Ai,
                                   std::vector<double>& Ax, std::vector<double>& b) {
  int n = update_starts();
  std::vector<std::map<int, double>> cols(n);
  b.assign(n, 0.0);
  for (auto& lf : J) {
    int ncols = lf.ndim[0] + (lf.n_nodes == 2 ? lf.ndim[1] : 0);
    int gcol[12];
    int c = 0;
    for (int k = 0; k < lf.n_nodes; k++)
      for (int j = 0; j < lf.ndim[k]; j++) gcol[c++] = nodes[lf.nodes[k]].start + j;
    for (int a = 0; a < ncols; a++) {
      for (int r = 0; r < lf.dim; r++) b[gcol[a]] += lf.H[r * ncols + a] * lf.rhs[r];
      for (int bb = 0; bb < ncols; bb++) {
        int i = gcol[a], j = gcol[bb];
        if (i > j) continue;
        double s = 0;
        for (int r = 0; r < lf.dim; r++) s += lf.H[r * ncols + a] * lf.H[r * ncols + bb];
        cols[j][i] += s;
      }
    }
  }
  Ap.assign(n + 1, 0); Ai.clear(); Ax.clear();
  for (int j = 0; j < n; j++) {
    for (auto& kv : cols[j]) {
      Ai.push_back(kv.first);
      Ax.push_back(kv.first == j ? kv.second * (1 + lambda) : kv.second);
    }
    Ap[j + 1] = (int)Ai.size();
  }
}

inline bool Slam::solve(const std::vector<LinFactor>& J, double lambda, std::vector<double>& delta) {
  double t0 = now_s();
  int n = update_starts();
  // live nodes in insertion order
  std::vector<int> live;
  for (size_t i = 0; i < nodes.size(); i++) if (nodes[i].alive) live.push_back((int)i);
  int nn = (int)live.size();
  // --- ordering ("analyze") ---
  double to = now_s();
  std::vector<int> order;
  size_t sig = J.size() * 1000003u + (size_t)nn;
  if (reuse_ordering && cached_order_sig == sig && (int)cached_order.size() == nn) {
    order = cached_order;
  } else {
    min_degree_order(J, live, order);
    if (reuse_ordering) { cached_order = order; cached_order_sig = sig; }
  }
  timers.order += now_s() - to;
  // new scalar start per node
  std::vector<int> pos(nodes.size(), -1);
  for (int i = 0; i < nn; i++) pos[live[i]] = i;
  std::vector<int> newstart(nn);
  {
    int s = 0;
    for (int k = 0; k < nn; k++) { newstart[order[k]] = s; s += node_dim(nodes[live[order[k]]].kind); }
  }
  // --- assemble permuted upper-triangular blocks ---
  std::map<std::pair<int, int>, std::vector<double>> blocks;  // key (colstart,rowstart)
  std::vector<double> rhs(n, 0.0);
  for (auto& lf : J) {
    int ncols = lf.ndim[0] + (lf.n_nodes == 2 ? lf.ndim[1] : 0);
    int off[2] = {0, lf.ndim[0]};
    for (int ka = 0; ka < lf.n_nodes; ka++) {
      int sa = newstart[pos[lf.nodes[ka]]], da = lf.ndim[ka];
      for (int a = 0; a < da; a++) {
        double s = 0;
        for (int r = 0; r < lf.dim; r++) s += lf.H[r * ncols + off[ka] + a] * lf.rhs[r];
        rhs[sa + a] += s;
      }
      for (int kb = 0; kb < lf.n_nodes; kb++) {
        int sb = newstart[pos[lf.nodes[kb]]], db = lf.ndim[kb];
        if (sa > sb) continue;  // keep upper part: row start <= col start
        std::vector<double>& blk = blocks[std::make_pair(sb, sa)];
        if (blk.empty()) blk.assign(da * db, 0.0);
        for (int a = 0; a < da; a++)
          for (int bb = 0; bb < db; bb++) {
            double s = 0;
            for (int r = 0; r < lf.dim; r++) s += lf.H[r * ncols + off[ka] + a] * lf.H[r * ncols + off[kb] + bb];
            blk[a * db + bb] += s;
          }
      }
    }
  }
  // CSC upper (rows <= col), damping diag *= (1+lambda)  (Cholesky.cpp:91-97)
  std::vector<int> Ap(n + 1, 0), Ai;
  std::vector<double> Ax;
  {
    // blocks are sorted by (colstart,rowstart); emit column by column
    auto it = blocks.begin();
    while (it != blocks.end()) {
      int sb = it->first.first;
      auto jt = it;
      int db = 0;
      // find db from diagonal block
      std::vector<std::pair<int, const std::vector<double>*>> colblks;
      while (jt != blocks.end() && jt->first.first == sb) { colblks.push_back(std::make_pair(jt->first.second, &jt->second)); ++jt; }
      // diagonal block is the last (rowstart == sb) and is square
      const std::vector<double>* diag = colblks.back().second;
      db = (int)std::lround(std::sqrt((double)diag->size()));
      for (int c = 0; c < db; c++) {
        for (auto& cb : colblks) {
          int sa = cb.first;
          int da = (int)cb.second->size() / db;
          for (int a = 0; a < da; a++) {
            int i = sa + a, j = sb + c;
            if (i > j) continue;
            double v = (*cb.second)[a * db + c];
            if (i == j) v *= (1 + lambda);
            Ai.push_back(i); Ax.push_back(v);
          }
        }
        Ap[sb + c + 1] = (int)Ai.size();
      }
      it = jt;
    }
  }
  // --- elimination tree ---
  std::vector<int> parent(n, -1), anc(n, -1);
  for (int k = 0; k < n; k++) {
    for (int p = Ap[k]; p < Ap[k + 1]; p++) {
      int i = Ai[p];
      while (i != -1 && i < k) {
        int inext = anc[i];
        anc[i] = k;
        if (inext == -1) parent[i] = k;
        i = inext;
      }
    }
  }
  // --- symbolic: column counts of L via row-subtree reach ---
  std::vector<int> mark(n, -1), stack(n), cnt(n, 1);
  auto ereach = [&](int k, int& top) {
    top = n;
    mark[k] = k;
    for (int p = Ap[k]; p < Ap[k + 1]; p++) {
      int i = Ai[p];
      if (i >= k) continue;
      int len = 0;  // path goes to the bottom of `stack` first (len <= top always), then to the top
      for (; mark[i] != k; i = parent[i]) { stack[len++] = i; mark[i] = k; }
      while (len > 0) stack[--top] = stack[--len];
    }
  };
  for (int k = 0; k < n; k++) {
    int top;
    ereach(k, top);
    for (int p = top; p < n; p++) cnt[stack[p]]++;
  }
  std::vector<int> Lp(n + 1, 0);
  for (int j = 0; j < n; j++) Lp[j + 1] = Lp[j] + cnt[j];
  std::vector<int> Li(Lp[n]);
  std::vector<double> Lx(Lp[n]);
  std::vector<int> cptr(Lp.begin(), Lp.end() - 1);
  std::vector<double> x(n, 0.0);
  std::fill(mark.begin(), mark.end(), -1);
  // --- numeric up-looking Cholesky ---
  bool ok = true;
  for (int k = 0; k < n; k++) {
    int top;
    ereach(k, top);
    x[k] = 0;
    for (int p = Ap[k]; p < Ap[k + 1]; p++) if (Ai[p] <= k) x[Ai[p]] = Ax[p];
    double d = x[k];
    x[k] = 0;
    for (; top < n; top++) {
      int i = stack[top];
      double lki = x[i] / Lx[Lp[i]];
      x[i] = 0;
      for (int p = Lp[i] + 1; p < cptr[i]; p++) x[Li[p]] -= Lx[p] * lki;
      d -= lki * lki;
      int p = cptr[i]++;
      Li[p] = k; Lx[p] = lki;
    }
    if (!(d > 0)) { ok = false; d = std::fabs(d) + 1e-300; }
    int p = cptr[k]++;
    Li[p] = k; Lx[p] = std::sqrt(d);
  }
  // --- solves: L y = rhs ; L^T z = y ---
  std::vector<double> y(rhs);
  for (int j = 0; j < n; j++) {
    y[j] /= Lx[Lp[j]];
    for (int p = Lp[j] + 1; p < Lp[j + 1]; p++) y[Li[p]] -= Lx[p] * y[j];
  }
  for (int j = n - 1; j >= 0; j--) {
    for (int p = Lp[j] + 1; p < Lp[j + 1]; p++) y[j] -= Lx[p] * y[Li[p]];
    y[j] /= Lx[Lp[j]];
  }
  // un-permute to insertion ordering (Optimizer.cpp:61-64)
  delta.assign(n, 0.0);
  for (int i = 0; i < nn; i++) {
    int d = node_dim(nodes[live[i]].kind);
    int s_old = nodes[live[i]].start, s_new = newstart[i];
    for (int c = 0; c < d; c++) delta[s_old + c] = y[s_new + c];
  }
  timers.solve += now_s() - t0;
  timers.n_solve++;
  if (!ok) last_error = "oracle: normal equations not positive definite";
  return ok;
}

}  // namespace orc
