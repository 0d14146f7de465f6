// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points (ref_*, same shapes as the orc_* / pus_* ones) around the
// UNMODIFIED reference classes: isam::Slam, Pose3d_Node, Pose3d_Factor, Pose3d_Pose3d_Factor (ISAM/include/isam/slam3d.h),
// Plane3d_Node, Pose3d_Plane3d_Factor, Pose3d_Plane3d_Factor2, Plane3d_Factor (pop_planar_slam/src/isam_plane3d.h) and the
// reference optimiser behind them (ISAM/isamlib/{Slam,Optimizer,Cholesky,numericalDiff,...}.cpp).  Built by
// `make -C oracle ref` into oracle/_ref/libisam_ref.so from the reference sources where they lie (/root/reference), against
// the API shims in oracle/ref_shim/ (Eigen3, Boost.Math and SuiteSparse are not installed in the build container).
// Nothing in the product links or loads this; tests use it to pin the oracle restatement and the CUDA path.
#include <cmath>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include <isam/isam.h>
#include <isam/robust.h>
#include "isam_plane3d.h"

using namespace isam;

namespace {
struct Props {   // layout of pus_properties (include/popup_gpu.h)
  int method; double epsilon2, epsilon_abs, epsilon_rel; int max_iterations; double lm_lambda0, lm_lambda_factor;
  int mod_update, mod_batch, mod_solve;
};
enum Kind { K_POSE = 0, K_PLANE = 1 };
enum FKind { F_PRIOR = 0, F_ODO = 1, F_PP = 2, F_LP = 3, F_PP2 = 4 };
struct Handle {
  Slam slam;
  std::vector<Node*> nodes;
  std::vector<int> nkind;
  std::vector<char> nalive;
  std::vector<Factor*> factors;
  std::vector<int> fkind;
  std::vector<char> falive;
  std::vector<double> t_lambda, t_new;
  std::vector<int> t_acc;
  Handle() {
    Properties p = slam.properties();
    p.quiet = true;
    slam.set_properties(p);
  }
  ~Handle() {
    for (Factor* f : factors) delete f;
    for (Node* n : nodes) delete n;
  }
};
thread_local std::string g_err;
double g_robust_b = 1.0;
double huber_cost(double d) { return cost_huber(d, g_robust_b); }               // robust.h:101-108
double pseudo_huber_cost(double d) { return cost_pseudo_huber(d, g_robust_b); } // robust.h:115-118
Handle* H(void* h) { return static_cast<Handle*>(h); }

Pose3d pose_from7(const double* v) { return Pose3d(Point3d(v[0], v[1], v[2]), Rot3d(Eigen::Quaterniond(v[3], v[4], v[5], v[6]))); }
void pose_to7(const Pose3d& p, double* o) {
  o[0] = p.x(); o[1] = p.y(); o[2] = p.z();
  const Eigen::Quaterniond q = p.rot().quaternion();
  o[3] = q.w(); o[4] = q.x(); o[5] = q.y(); o[6] = q.z();
}
Eigen::MatrixXd ut_to_matrix(const double* si, int dim) {
  Eigen::MatrixXd m = Eigen::MatrixXd::Zero(dim, dim);
  int q = 0;
  for (int r = 0; r < dim; r++)
    for (int c = r; c < dim; c++) m(r, c) = si[q++];
  return m;
}
bool ok_node(Handle* s, int id, int kind) { return id >= 0 && id < (int)s->nodes.size() && s->nalive[id] && s->nkind[id] == kind; }
bool ok_factor(Handle* s, int f) { return f >= 0 && f < (int)s->factors.size() && s->falive[f]; }
int push_factor(Handle* s, Factor* f, int kind) {
  s->slam.add_factor(f);
  s->factors.push_back(f); s->fkind.push_back(kind); s->falive.push_back(1);
  return (int)s->factors.size() - 1;
}
}  // namespace

extern "C" {

int ref_create(int, void** out) { *out = new Handle(); return 0; }
int ref_destroy(void* h) { delete H(h); return 0; }
const char* ref_last_error(void) { return g_err.c_str(); }

int ref_add_pose(void* h, const double* v) {
  Pose3d_Node* n = new Pose3d_Node();
  if (v) n->init(pose_from7(v));
  H(h)->slam.add_node(n);
  H(h)->nodes.push_back(n); H(h)->nkind.push_back(K_POSE); H(h)->nalive.push_back(1);
  return (int)H(h)->nodes.size() - 1;
}
int ref_add_plane(void* h, const double* v) {
  Plane3d_Node* n = new Plane3d_Node();
  if (v) n->init(Plane3d(Eigen::Vector4d(v[0], v[1], v[2], v[3])));
  H(h)->slam.add_node(n);
  H(h)->nodes.push_back(n); H(h)->nkind.push_back(K_PLANE); H(h)->nalive.push_back(1);
  return (int)H(h)->nodes.size() - 1;
}
int ref_add_poses(void* h, int n, const double* v, int* out) {
  int first = -1;
  for (int i = 0; i < n; i++) { int id = ref_add_pose(h, v ? v + 7 * i : nullptr); if (!i) first = id; if (out) out[i] = id; }
  return first;
}
int ref_add_planes(void* h, int n, const double* v, int* out) {
  int first = -1;
  for (int i = 0; i < n; i++) { int id = ref_add_plane(h, v ? v + 4 * i : nullptr); if (!i) first = id; if (out) out[i] = id; }
  return first;
}
int ref_init_pose(void* h, int id, const double* v) {
  if (!ok_node(H(h), id, K_POSE)) { g_err = "bad pose id"; return -1; }
  static_cast<Pose3d_Node*>(H(h)->nodes[id])->init(pose_from7(v)); return 0;
}
int ref_init_plane(void* h, int id, const double* v) {
  if (!ok_node(H(h), id, K_PLANE)) { g_err = "bad plane id"; return -1; }
  static_cast<Plane3d_Node*>(H(h)->nodes[id])->init(Plane3d(Eigen::Vector4d(v[0], v[1], v[2], v[3]))); return 0;
}
int ref_init_poses(void* h, int n, const int* ids, const double* v) { for (int i = 0; i < n; i++) if (ref_init_pose(h, ids[i], v + 7 * i) < 0) return -1; return 0; }
int ref_init_planes(void* h, int n, const int* ids, const double* v) { for (int i = 0; i < n; i++) if (ref_init_plane(h, ids[i], v + 4 * i) < 0) return -1; return 0; }
int ref_get_pose(void* h, int id, double* out) {
  if (!ok_node(H(h), id, K_POSE)) { g_err = "bad pose id"; return -1; }
  pose_to7(static_cast<Pose3d_Node*>(H(h)->nodes[id])->value(), out); return 0;
}
int ref_get_plane(void* h, int id, double* out) {
  if (!ok_node(H(h), id, K_PLANE)) { g_err = "bad plane id"; return -1; }
  const Eigen::Vector4d v = static_cast<Plane3d_Node*>(H(h)->nodes[id])->value().vector();
  for (int i = 0; i < 4; i++) out[i] = v(i);
  return 0;
}
int ref_get_poses(void* h, int n, const int* ids, double* out) { for (int i = 0; i < n; i++) if (ref_get_pose(h, ids[i], out + 7 * i) < 0) return -1; return 0; }
int ref_get_planes(void* h, int n, const int* ids, double* out) { for (int i = 0; i < n; i++) if (ref_get_plane(h, ids[i], out + 4 * i) < 0) return -1; return 0; }

int ref_add_pose_prior(void* h, int p, const double* m, const double* si) {
  if (!ok_node(H(h), p, K_POSE)) { g_err = "bad pose id"; return -1; }
  return push_factor(H(h), new Pose3d_Factor(static_cast<Pose3d_Node*>(H(h)->nodes[p]), Pose3d(m[0], m[1], m[2], m[3], m[4], m[5]), SqrtInformation(ut_to_matrix(si, 6))), F_PRIOR);
}
int ref_add_odometry(void* h, int a, int b, const double* m, const double* si) {
  if (!ok_node(H(h), a, K_POSE) || !ok_node(H(h), b, K_POSE)) { g_err = "bad pose id"; return -1; }
  return push_factor(H(h), new Pose3d_Pose3d_Factor(static_cast<Pose3d_Node*>(H(h)->nodes[a]), static_cast<Pose3d_Node*>(H(h)->nodes[b]),
                                                    Pose3d(m[0], m[1], m[2], m[3], m[4], m[5]), SqrtInformation(ut_to_matrix(si, 6))), F_ODO);
}
int ref_add_pose_plane(void* h, int p, int l, const double* m, const double* si) {
  if (!ok_node(H(h), p, K_POSE) || !ok_node(H(h), l, K_PLANE)) { g_err = "bad node id"; return -1; }
  return push_factor(H(h), new Pose3d_Plane3d_Factor(static_cast<Pose3d_Node*>(H(h)->nodes[p]), static_cast<Plane3d_Node*>(H(h)->nodes[l]),
                                                     Plane3d(Eigen::Vector4d(m[0], m[1], m[2], m[3])), SqrtInformation(ut_to_matrix(si, 3)), false), F_PP);
}
// Pose3d_Plane3d_Factor2: the rays are handed to precompute_edge_ray() as an identity invK and one ground segment whose
// end points are (x/z, y/z) of the two rays -- callers pass rays of the form (x, y, 1) with float32-representable x, y
int ref_add_pose_plane2(void* h, int p, int l, const double* m, const double* rays6, const double* si) {
  if (!ok_node(H(h), p, K_POSE) || !ok_node(H(h), l, K_PLANE)) { g_err = "bad node id"; return -1; }
  Pose3d_Plane3d_Factor2* f = new Pose3d_Plane3d_Factor2(static_cast<Pose3d_Node*>(H(h)->nodes[p]), static_cast<Plane3d_Node*>(H(h)->nodes[l]),
                                                         Plane3d(Eigen::Vector4d(m[0], m[1], m[2], m[3])), SqrtInformation(ut_to_matrix(si, 3)), false);
  Eigen::Matrix3f K = Eigen::Matrix3f::Identity();
  Eigen::MatrixXf seg(1, 4);
  seg << (float)(rays6[0] / rays6[2]), (float)(rays6[1] / rays6[2]), (float)(rays6[3] / rays6[5]), (float)(rays6[4] / rays6[5]);
  f->precompute_edge_ray(K, seg);
  return push_factor(H(h), f, F_PP2);
}
int ref_add_plane_prior(void* h, int l, const double* m, const double* si) {
  if (!ok_node(H(h), l, K_PLANE)) { g_err = "bad plane id"; return -1; }
  return push_factor(H(h), new Plane3d_Factor(static_cast<Plane3d_Node*>(H(h)->nodes[l]), Plane3d(Eigen::Vector4d(m[0], m[1], m[2], m[3])), SqrtInformation(ut_to_matrix(si, 3))), F_LP);
}
int ref_add_odometry_bulk(void* h, int n, const int* a, const int* b, const double* m, const double* si, int* out) {
  int first = -1;
  for (int i = 0; i < n; i++) { int f = ref_add_odometry(h, a[i], b[i], m + 6 * i, si + 21 * i); if (f < 0) return f; if (!i) first = f; if (out) out[i] = f; }
  return first;
}
int ref_add_pose_plane_bulk(void* h, int n, const int* a, const int* b, const double* m, const double* si, int* out) {
  int first = -1;
  for (int i = 0; i < n; i++) { int f = ref_add_pose_plane(h, a[i], b[i], m + 4 * i, si + 6 * i); if (f < 0) return f; if (!i) first = f; if (out) out[i] = f; }
  return first;
}
int ref_set_measurement(void* h, int fid, const double* m) {
  if (!ok_factor(H(h), fid)) { g_err = "bad factor id"; return -1; }
  const int k = H(h)->fkind[fid];
  if (k == F_PP || k == F_LP || k == F_PP2) static_cast<FactorT<Plane3d>*>(H(h)->factors[fid])->set_measurement(Plane3d(Eigen::Vector4d(m[0], m[1], m[2], m[3])));
  else static_cast<FactorT<Pose3d>*>(H(h)->factors[fid])->set_measurement(Pose3d(m[0], m[1], m[2], m[3], m[4], m[5]));
  return 0;
}
int ref_get_measurement(void* h, int fid, double* m) {
  if (!ok_factor(H(h), fid)) { g_err = "bad factor id"; return -1; }
  const int k = H(h)->fkind[fid];
  if (k == F_PP || k == F_LP || k == F_PP2) {
    const Eigen::Vector4d v = static_cast<FactorT<Plane3d>*>(H(h)->factors[fid])->measurement().vector();
    for (int i = 0; i < 4; i++) m[i] = v(i);
  } else {
    const Eigen::VectorXd v = static_cast<FactorT<Pose3d>*>(H(h)->factors[fid])->measurement().vector();
    for (int i = 0; i < 6; i++) m[i] = v(i);
  }
  return 0;
}
int ref_remove_factor(void* h, int fid) {
  if (!ok_factor(H(h), fid)) { g_err = "bad factor id"; return -1; }
  H(h)->slam.remove_factor(H(h)->factors[fid]);
  H(h)->falive[fid] = 0;
  return 0;
}
int ref_remove_node(void* h, int id) {
  Handle* s = H(h);
  if (id < 0 || id >= (int)s->nodes.size() || !s->nalive[id]) { g_err = "bad node id"; return -1; }
  // Slam::remove_node removes the adjacent factors too (Slam.cpp:107-115): mirror that in the id tables first
  const std::list<Factor*> adj = s->nodes[id]->factors();
  for (Factor* f : adj)
    for (size_t k = 0; k < s->factors.size(); k++) if (s->factors[k] == f) s->falive[k] = 0;
  s->slam.remove_node(s->nodes[id]);
  s->nalive[id] = 0;
  return 0;
}
int ref_num_nodes(void* h) { return (int)H(h)->slam.get_nodes().size(); }
int ref_num_factors(void* h) { return (int)H(h)->slam.get_factors().size(); }
int ref_node_start(void* h, int id) {
  if (id < 0 || id >= (int)H(h)->nodes.size() || !H(h)->nalive[id]) return -1;
  H(h)->slam.jacobian();   // update_starts() (Slam.cpp:59-67) runs inside jacobian_partial
  return H(h)->nodes[id]->start();
}

int ref_get_properties(void* h, Props* p) {
  const Properties q = H(h)->slam.properties();
  p->method = (q.method == LEVENBERG_MARQUARDT) ? 1 : 0;
  p->epsilon2 = q.epsilon2; p->epsilon_abs = q.epsilon_abs; p->epsilon_rel = q.epsilon_rel; p->max_iterations = q.max_iterations;
  p->lm_lambda0 = q.lm_lambda0; p->lm_lambda_factor = q.lm_lambda_factor; p->mod_update = q.mod_update; p->mod_batch = q.mod_batch;
  p->mod_solve = q.mod_solve;
  return 0;
}
int ref_set_properties(void* h, const Props* p) {
  Properties q = H(h)->slam.properties();
  q.method = p->method == 1 ? LEVENBERG_MARQUARDT : GAUSS_NEWTON;
  q.epsilon2 = p->epsilon2; q.epsilon_abs = p->epsilon_abs; q.epsilon_rel = p->epsilon_rel; q.max_iterations = p->max_iterations;
  q.lm_lambda0 = p->lm_lambda0; q.lm_lambda_factor = p->lm_lambda_factor; q.mod_update = p->mod_update; q.mod_batch = p->mod_batch;
  q.mod_solve = p->mod_solve;
  H(h)->slam.set_properties(q);
  return 0;
}
int ref_set_robust(void* h, int kind, double b) {
  g_robust_b = b;
  H(h)->slam.set_cost_function(kind == 1 ? &huber_cost : (kind == 2 ? &pseudo_huber_cost : (cost_func_t) nullptr));
  return 0;
}

// Slam::batch_optimization (Slam.cpp:198-210).  The LM trace (lambda, accepted, new chi2) is recovered from the
// reference's own progress lines (Optimizer.cpp:420-427), printed at 17 digits because the stream precision is sticky.
int ref_batch_optimize(void* h, int* iters) {
  Handle* s = H(h);
  Properties p = s->slam.properties();
  const bool lm = (p.method == LEVENBERG_MARQUARDT);
  p.quiet = !lm;
  s->slam.set_properties(p);
  std::ostringstream cap;
  std::streambuf* old = std::cout.rdbuf(cap.rdbuf());
  const std::streamsize prec = std::cout.precision(17);
  const int it = s->slam.batch_optimization();
  std::cout.precision(prec);
  std::cout.rdbuf(old);
  p.quiet = true;
  s->slam.set_properties(p);
  s->t_lambda.clear(); s->t_new.clear(); s->t_acc.clear();
  std::istringstream in(cap.str());
  std::string line;
  while (std::getline(in, line)) {
    const size_t a = line.find("(lambda=");
    if (line.compare(0, 12, "LM Iteration") != 0 || a == std::string::npos) continue;
    const double lam = std::atof(line.c_str() + a + 8);
    const size_t r = line.find("residual: ");
    s->t_lambda.push_back(lam);
    s->t_acc.push_back(r != std::string::npos);
    s->t_new.push_back(r != std::string::npos ? std::atof(line.c_str() + r + 10) : NAN);
  }
  if (iters) *iters = it;
  return 0;
}
int ref_update(void* h) { H(h)->slam.update(); return 0; }
int ref_chi2(void* h, double* out) { *out = H(h)->slam.chi2(); return 0; }
int ref_get_trace(void* h, int cap, double* lambda, double* e_new, double* e_before, double* dn, int* acc, int* pcg) {
  Handle* s = H(h);
  const int n = (int)s->t_lambda.size();
  for (int i = 0; i < n && i < cap; i++) {
    if (lambda) lambda[i] = s->t_lambda[i];
    if (e_new) e_new[i] = s->t_new[i];
    if (e_before) e_before[i] = NAN;
    if (dn) dn[i] = NAN;
    if (acc) acc[i] = s->t_acc[i];
    if (pcg) pcg[i] = 0;
  }
  return n;
}

// ---- factor-level access: the reference's own error() / numericalDiff Jacobian of one factor ----
// r_out: Factor::error (sqrtinf * basic_error, robustified when a cost function is set; Factor.h:67-77), returns dim
int ref_factor_error(void* h, int fid, double* r_out) {
  if (!ok_factor(H(h), fid)) { g_err = "bad factor id"; return -1; }
  const Eigen::VectorXd e = H(h)->factors[fid]->error(ESTIMATE);
  for (int i = 0; i < e.size(); i++) r_out[i] = e(i);
  return (int)e.size();
}
// H_out: dim x (sum of node dims), row-major, terms in node order; r_out = residual (rhs = -r upstream, Jacobian.h:98)
int ref_factor_jacobian(void* h, int fid, int /*mode*/, double* H_out, double* r_out) {
  if (!ok_factor(H(h), fid)) { g_err = "bad factor id"; return -1; }
  Factor* f = H(h)->factors[fid];
  for (Node* n : f->nodes()) static_cast<Node*>(n)->estimate_to_linpoint();   // Factor::jacobian differentiates at the linpoint
  Jacobian J = f->jacobian();   // Factor.h:126-139 -> numericalDiff.cpp:41-87
  int ncols = 0;
  for (Terms::const_iterator it = J.terms().begin(); it != J.terms().end(); ++it) ncols += it->term().cols();
  const int dim = f->dim();
  int c0 = 0;
  for (Terms::const_iterator it = J.terms().begin(); it != J.terms().end(); ++it) {
    const Eigen::MatrixXd& T = it->term();
    for (int r = 0; r < dim; r++)
      for (int c = 0; c < T.cols(); c++) H_out[r * ncols + c0 + c] = T(r, c);
    c0 += T.cols();
  }
  const Eigen::VectorXd rhs = J.rhs();
  for (int i = 0; i < dim; i++) r_out[i] = -rhs(i);
  return ncols;
}

// ---- value-type helpers straight from the reference classes ----
void ref_pose_from_xyzypr(const double* v6, double* out7) { pose_to7(Pose3d(v6[0], v6[1], v6[2], v6[3], v6[4], v6[5]), out7); }
void ref_pose_vector(const double* p7, double* out6) { const Eigen::VectorXd v = pose_from7(p7).vector(); for (int i = 0; i < 6; i++) out6[i] = v(i); }
void ref_pose_exmap(const double* p7, const double* d6, double* out7) {
  Eigen::VectorXd d(6);
  for (int i = 0; i < 6; i++) d(i) = d6[i];
  pose_to7(pose_from7(p7).exmap(d), out7);
}
void ref_pose_oplus(const double* a7, const double* b7, double* out7) { pose_to7(pose_from7(a7).oplus(pose_from7(b7)), out7); }
void ref_pose_ominus(const double* a7, const double* b7, double* out7) { pose_to7(pose_from7(a7).ominus(pose_from7(b7)), out7); }
void ref_pose_wTo(const double* p7, double* T16) { const Eigen::Matrix4d T = pose_from7(p7).wTo(); for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T16[r * 4 + c] = T(r, c); }
void ref_pose_oTw(const double* p7, double* T16) { const Eigen::Matrix4d T = pose_from7(p7).oTw(); for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T16[r * 4 + c] = T(r, c); }
void ref_pose_from_mat4(const double* T16, double* out7) {
  Eigen::MatrixXd T(4, 4);
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T(r, c) = T16[r * 4 + c];
  pose_to7(Pose3d(T), out7);
}
void ref_plane_exmap(const double* p4, const double* d3, double* out4) {
  const Eigen::Vector4d v = Plane3d(Eigen::Vector4d(p4[0], p4[1], p4[2], p4[3])).exmap(Eigen::Vector3d(d3[0], d3[1], d3[2])).vector();
  for (int i = 0; i < 4; i++) out4[i] = v(i);
}
// Plane3d(T^T * plane): transform_to(wTo) / transform_from(oTw) (isam_plane3d.h:180-188)
void ref_plane_transform(const double* T16, const double* p4, double* out4) {
  Eigen::Matrix4d T;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T(r, c) = T16[r * 4 + c];
  const Eigen::Vector4d v = Plane3d(Eigen::Vector4d(p4[0], p4[1], p4[2], p4[3])).transform_to(T).vector();
  for (int i = 0; i < 4; i++) out4[i] = v(i);
}
double ref_standard_rad(double t) { return standardRad(t); }

// get_wall_plane_equation (isam_plane3d.cpp:20-55): rays 3 x 2n (column-major pairs), T 4x4 row-major -> n x 4 planes (row-major)
int ref_wall_plane_equation(int n_seg, const double* rays, const double* T16, double* planes_out) {
  Eigen::MatrixXd R(3, 2 * n_seg);
  for (int c = 0; c < 2 * n_seg; c++) for (int r = 0; r < 3; r++) R(r, c) = rays[c * 3 + r];
  Eigen::Matrix4d T;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T(r, c) = T16[r * 4 + c];
  Eigen::MatrixXd out;
  get_wall_plane_equation(R, T, out);
  for (int s = 0; s < out.rows(); s++) for (int c = 0; c < 4; c++) planes_out[s * 4 + c] = out(s, c);
  return (int)out.rows();
}

}  // extern "C"
