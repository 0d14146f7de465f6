// ORACLE (test infrastructure, NOT product code) -- see oracle_math.hpp header.
// C entry points (orc_*) with the same shapes as include/popup_gpu.h (pus_*), so the
// parity tests can drive the oracle and the CUDA library through one Python wrapper.
// Parity: pose-graph path pinned against the reference's sphere2500 dataset + ground truth; plane path
// "parity unpinned" (oracle_math.hpp).
#include <cstring>
#include <string>
#include <vector>

#include "oracle_popup.hpp"
#include "oracle_slam.hpp"

using namespace orc;

namespace {
thread_local std::string g_err;
struct Props {
  int method; double epsilon2, epsilon_abs, epsilon_rel; int max_iterations; double lm_lambda0, lm_lambda_factor;
  int mod_update, mod_batch, mod_solve;
};
inline Pose pose_from7(const double* v) {
  Pose p; p.t[0] = v[0]; p.t[1] = v[1]; p.t[2] = v[2]; p.q = Quat{v[3], v[4], v[5], v[6]}; return p;
}
inline void pose_to7(const Pose& p, double* v) {
  v[0] = p.t[0]; v[1] = p.t[1]; v[2] = p.t[2]; v[3] = p.q.w; v[4] = p.q.x; v[5] = p.q.y; v[6] = p.q.z;
}
inline bool valid_node(Slam* s, int id, int kind) {
  return id >= 0 && id < (int)s->nodes.size() && s->nodes[id].alive && s->nodes[id].kind == kind;
}
}  // namespace

#define S(h) (reinterpret_cast<Slam*>(h))

extern "C" {

int orc_create(int, void** out) { *out = new Slam(); return 0; }
int orc_destroy(void* h) { delete S(h); return 0; }
const char* orc_last_error(void) { return g_err.c_str(); }

int orc_add_pose(void* h, const double* init7) {
  int id = S(h)->add_pose();
  if (init7) S(h)->init_pose(id, pose_from7(init7));
  return id;
}
int orc_add_plane(void* h, const double* abcd) {
  int id = S(h)->add_plane();
  if (abcd) S(h)->init_plane(id, plane_from_vec4(abcd));
  return id;
}
int orc_add_poses(void* h, int n, const double* v, int* out_ids) {
  int first = -1;
  for (int i = 0; i < n; i++) { int id = orc_add_pose(h, v ? v + 7 * i : nullptr); if (i == 0) first = id; if (out_ids) out_ids[i] = id; }
  return first;
}
int orc_add_planes(void* h, int n, const double* v, int* out_ids) {
  int first = -1;
  for (int i = 0; i < n; i++) { int id = orc_add_plane(h, v ? v + 4 * i : nullptr); if (i == 0) first = id; if (out_ids) out_ids[i] = id; }
  return first;
}
int orc_init_pose(void* h, int id, const double* v) {
  if (!valid_node(S(h), id, NODE_POSE)) { g_err = "bad pose id"; return -1; }
  S(h)->init_pose(id, pose_from7(v)); return 0;
}
int orc_init_plane(void* h, int id, const double* v) {
  if (!valid_node(S(h), id, NODE_PLANE)) { g_err = "bad plane id"; return -1; }
  S(h)->init_plane(id, plane_from_vec4(v)); return 0;
}
int orc_init_poses(void* h, int n, const int* ids, const double* v) {
  for (int i = 0; i < n; i++) if (orc_init_pose(h, ids[i], v + 7 * i) < 0) return -1;
  return 0;
}
int orc_init_planes(void* h, int n, const int* ids, const double* v) {
  for (int i = 0; i < n; i++) if (orc_init_plane(h, ids[i], v + 4 * i) < 0) return -1;
  return 0;
}
int orc_get_pose(void* h, int id, double* out7) {
  if (!valid_node(S(h), id, NODE_POSE)) { g_err = "bad pose id"; return -1; }
  pose_to7(S(h)->nodes[id].pose, out7); return 0;
}
int orc_get_plane(void* h, int id, double* out4) {
  if (!valid_node(S(h), id, NODE_PLANE)) { g_err = "bad plane id"; return -1; }
  std::memcpy(out4, S(h)->nodes[id].plane.v, 4 * sizeof(double)); return 0;
}
int orc_get_poses(void* h, int n, const int* ids, double* out) {
  for (int i = 0; i < n; i++) if (orc_get_pose(h, ids[i], out + 7 * i) < 0) return -1;
  return 0;
}
int orc_get_planes(void* h, int n, const int* ids, double* out) {
  for (int i = 0; i < n; i++) if (orc_get_plane(h, ids[i], out + 4 * i) < 0) return -1;
  return 0;
}

int orc_add_pose_prior(void* h, int pose, const double* m, const double* si) {
  if (!valid_node(S(h), pose, NODE_POSE)) { g_err = "bad pose id"; return -1; }
  return S(h)->add_pose_prior(pose, m, si);
}
int orc_add_odometry(void* h, int a, int b, const double* m, const double* si) {
  if (!valid_node(S(h), a, NODE_POSE) || !valid_node(S(h), b, NODE_POSE)) { g_err = "bad pose id"; return -1; }
  int r = S(h)->add_odometry(a, b, m, si);
  if (r < 0) g_err = S(h)->last_error;
  return r;
}
int orc_add_pose_plane(void* h, int pose, int plane, const double* m, const double* si) {
  if (!valid_node(S(h), pose, NODE_POSE) || !valid_node(S(h), plane, NODE_PLANE)) { g_err = "bad node id"; return -1; }
  int r = S(h)->add_pose_plane(pose, plane, m, si);
  if (r < 0) g_err = S(h)->last_error;
  return r;
}
int orc_add_pose_plane2(void* h, int pose, int plane, const double* m, const double* rays, const double* si) {
  if (!valid_node(S(h), pose, NODE_POSE) || !valid_node(S(h), plane, NODE_PLANE)) { g_err = "bad node id"; return -1; }
  int r = S(h)->add_pose_plane2(pose, plane, m, rays, si);
  if (r < 0) g_err = S(h)->last_error;
  return r;
}
int orc_add_plane_prior(void* h, int plane, const double* m, const double* si) {
  if (!valid_node(S(h), plane, NODE_PLANE)) { g_err = "bad plane id"; return -1; }
  return S(h)->add_plane_prior(plane, m, si);
}
int orc_add_odometry_bulk(void* h, int n, const int* a, const int* b, const double* m, const double* si, int* out) {
  int first = -1;
  for (int i = 0; i < n; i++) {
    int f = orc_add_odometry(h, a[i], b[i], m + 6 * i, si + 21 * i);
    if (f < 0) return f;
    if (i == 0) first = f;
    if (out) out[i] = f;
  }
  return first;
}
int orc_add_pose_plane_bulk(void* h, int n, const int* a, const int* b, const double* m, const double* si, int* out) {
  int first = -1;
  for (int i = 0; i < n; i++) {
    int f = orc_add_pose_plane(h, a[i], b[i], m + 4 * i, si + 6 * i);
    if (f < 0) return f;
    if (i == 0) first = f;
    if (out) out[i] = f;
  }
  return first;
}
int orc_set_measurement(void* h, int fid, const double* m) {
  if (fid < 0 || fid >= (int)S(h)->factors.size() || !S(h)->factors[fid].alive) { g_err = "bad factor id"; return -1; }
  S(h)->set_measurement(fid, m); return 0;
}
int orc_get_measurement(void* h, int fid, double* m) {
  if (fid < 0 || fid >= (int)S(h)->factors.size() || !S(h)->factors[fid].alive) { g_err = "bad factor id"; return -1; }
  const Factor& f = S(h)->factors[fid];
  std::memcpy(m, f.meas, (f.dim == 3 ? 4 : 6) * sizeof(double)); return 0;
}
int orc_remove_factor(void* h, int fid) {
  if (fid < 0 || fid >= (int)S(h)->factors.size() || !S(h)->factors[fid].alive) { g_err = "bad factor id"; return -1; }
  S(h)->remove_factor(fid); return 0;
}
int orc_remove_node(void* h, int id) {
  if (id < 0 || id >= (int)S(h)->nodes.size() || !S(h)->nodes[id].alive) { g_err = "bad node id"; return -1; }
  S(h)->remove_node(id); return 0;
}
int orc_num_nodes(void* h) { return S(h)->num_nodes(); }
int orc_num_factors(void* h) { return S(h)->num_factors(); }
int orc_factor_nodes(void* h, int fid, int* out2) {
  if (fid < 0 || fid >= (int)S(h)->factors.size() || !S(h)->factors[fid].alive) { g_err = "bad factor id"; return -1; }
  const Factor& f = S(h)->factors[fid];
  for (int k = 0; k < f.n_nodes; k++) out2[k] = f.nodes[k];
  return f.n_nodes;
}
int orc_node_factors(void* h, int id, int* out, int cap) {
  if (id < 0 || id >= (int)S(h)->nodes.size() || !S(h)->nodes[id].alive) { g_err = "bad node id"; return -1; }
  int c = 0;
  for (size_t i = 0; i < S(h)->factors.size(); i++) {
    const Factor& f = S(h)->factors[i];
    if (!f.alive) continue;
    for (int k = 0; k < f.n_nodes; k++) if (f.nodes[k] == id) { if (c < cap) out[c] = (int)i; c++; break; }
  }
  return c;
}
int orc_node_start(void* h, int id) {
  if (id < 0 || id >= (int)S(h)->nodes.size()) return -1;
  S(h)->update_starts();
  return S(h)->nodes[id].start;
}
int orc_factor_row(void* h, int fid) {
  if (fid < 0 || fid >= (int)S(h)->factors.size() || !S(h)->factors[fid].alive) return -1;
  int row = 0;
  for (int i = 0; i < fid; i++) if (S(h)->factors[i].alive) row += S(h)->factors[i].dim;
  return row;
}

int orc_get_properties(void* h, Props* p) {
  const Properties& q = S(h)->prop;
  p->method = q.method; p->epsilon2 = q.epsilon2; p->epsilon_abs = q.epsilon_abs; p->epsilon_rel = q.epsilon_rel;
  p->max_iterations = q.max_iterations; p->lm_lambda0 = q.lm_lambda0; p->lm_lambda_factor = q.lm_lambda_factor;
  p->mod_update = q.mod_update; p->mod_batch = q.mod_batch; p->mod_solve = q.mod_solve;
  return 0;
}
int orc_set_properties(void* h, const Props* p) {
  Properties& q = S(h)->prop;
  q.method = p->method; q.epsilon2 = p->epsilon2; q.epsilon_abs = p->epsilon_abs; q.epsilon_rel = p->epsilon_rel;
  q.max_iterations = p->max_iterations; q.lm_lambda0 = p->lm_lambda0; q.lm_lambda_factor = p->lm_lambda_factor;
  q.mod_update = p->mod_update; q.mod_batch = p->mod_batch; q.mod_solve = p->mod_solve;
  return 0;
}
int orc_set_robust(void* h, int kind, double b) { S(h)->robust_kind = kind; S(h)->robust_b = b; return 0; }

int orc_batch_optimize(void* h, int* iters) {
  int it = S(h)->batch_optimization();
  if (iters) *iters = it;
  return 0;
}
int orc_update(void* h) {
  int r = S(h)->update();
  if (r < 0) g_err = S(h)->last_error;
  return r;
}
int orc_chi2(void* h, double* out) { *out = S(h)->chi2(ESTIMATE); return 0; }
int orc_get_trace(void* h, int cap, double* lambda, double* e_new, double* e_before, double* dn, int* acc, int* pcg) {
  int n = (int)S(h)->trace.size();
  for (int i = 0; i < n && i < cap; i++) {
    const TraceEntry& t = S(h)->trace[i];
    if (lambda) lambda[i] = t.lambda;
    if (e_new) e_new[i] = t.error_new;
    if (e_before) e_before[i] = t.error_before;
    if (dn) dn[i] = t.delta_norm;
    if (acc) acc[i] = t.accepted;
    if (pcg) pcg[i] = 0;
  }
  return n;
}

// ---- oracle-only hooks ----
int orc_set_jacobian_mode(void* h, int mode) { S(h)->jac_mode = mode; return 0; }
int orc_set_reuse_ordering(void* h, int on) { S(h)->reuse_ordering = on != 0; return 0; }
// timers: linearize, solve, chi2, order, total, n_linearize, n_solve, n_chi2
int orc_get_timers(void* h, double* out8) {
  const Timers& t = S(h)->timers;
  out8[0] = t.linearize; out8[1] = t.solve; out8[2] = t.chi2; out8[3] = t.order; out8[4] = t.total;
  out8[5] = t.n_linearize; out8[6] = t.n_solve; out8[7] = t.n_chi2;
  return 0;
}
int orc_reset_timers(void* h) { S(h)->timers = Timers(); return 0; }
// Jacobian of one factor at the LINPOINT (= estimate copied first). mode 0 numeric (reference), 1 analytic.
// H_out: dim x ncols row-major, r_out: dim (weighted residual). Returns ncols.
int orc_factor_jacobian(void* h, int fid, int mode, double* H_out, double* r_out) {
  Slam* s = S(h);
  if (fid < 0 || fid >= (int)s->factors.size() || !s->factors[fid].alive) { g_err = "bad factor id"; return -1; }
  s->estimate_to_linpoint();
  int saved = s->jac_mode;
  s->jac_mode = mode;
  LinFactor lf;
  s->linearize_factor(fid, lf);
  s->jac_mode = saved;
  s->estimate_to_linpoint();  // undo the Euler round trip of numericalDiff on the linpoint copy
  int ncols = lf.ndim[0] + (lf.n_nodes == 2 ? lf.ndim[1] : 0);
  std::memcpy(H_out, lf.H, sizeof(double) * lf.dim * ncols);
  for (int i = 0; i < lf.dim; i++) r_out[i] = -lf.rhs[i];
  return ncols;
}
// weighted residual of one factor at the estimate
int orc_factor_error(void* h, int fid, double* r_out) {
  Slam* s = S(h);
  if (fid < 0 || fid >= (int)s->factors.size() || !s->factors[fid].alive) { g_err = "bad factor id"; return -1; }
  s->error(s->factors[fid], ESTIMATE, r_out);
  return s->factors[fid].dim;
}
// normal equations (upper CSC, insertion ordering) at the current estimate; returns nnz or needed sizes
long long orc_normal_equations(void* h, double lambda, int* Ap, int* Ai, double* Ax, double* b, long long cap_nnz) {
  Slam* s = S(h);
  s->estimate_to_linpoint();
  std::vector<LinFactor> J;
  s->jacobian(J);
  s->estimate_to_linpoint();
  std::vector<int> ap, ai; std::vector<double> ax, bb;
  s->normal_equations(J, lambda, ap, ai, ax, bb);
  long long nnz = (long long)ai.size();
  if (Ap && cap_nnz >= nnz) {
    std::memcpy(Ap, ap.data(), ap.size() * sizeof(int));
    std::memcpy(Ai, ai.data(), ai.size() * sizeof(int));
    std::memcpy(Ax, ax.data(), ax.size() * sizeof(double));
    std::memcpy(b, bb.data(), bb.size() * sizeof(double));
  }
  return nnz;
}
// one damped solve at the current estimate: delta (insertion ordering). Returns state dimension.
int orc_solve_step(void* h, double lambda, double* delta_out, int cap) {
  Slam* s = S(h);
  s->estimate_to_linpoint();
  std::vector<LinFactor> J;
  s->jacobian(J);
  s->estimate_to_linpoint();
  std::vector<double> d;
  bool ok = s->solve(J, lambda, d);
  if (!ok) { g_err = s->last_error; }
  if ((int)d.size() <= cap) std::memcpy(delta_out, d.data(), d.size() * sizeof(double));
  return (int)d.size();
}
// apply a tangent step to the estimate (NodeT::apply_exmap Node.h:145 on linpoint = estimate)
int orc_apply_delta(void* h, const double* delta, int n) {
  Slam* s = S(h);
  s->estimate_to_linpoint();
  std::vector<double> d(delta, delta + n);
  s->apply_exmap(d);
  return 0;
}

// pose / plane helpers exported for the Python tests & generators' cross-checks
void orc_pose_from_xyzypr(const double* v6, double* out7) { pose_to7(pose_from_xyzypr(v6), out7); }
void orc_pose_vector(const double* p7, double* out6) { pose_vector(pose_from7(p7), out6); }
void orc_pose_exmap(const double* p7, const double* d6, double* out7) { pose_to7(pose_exmap(pose_from7(p7), d6), out7); }
void orc_pose_oplus(const double* a7, const double* b7, double* out7) { pose_to7(pose_oplus(pose_from7(a7), pose_from7(b7)), out7); }
void orc_pose_ominus(const double* a7, const double* b7, double* out7) { pose_to7(pose_ominus(pose_from7(a7), pose_from7(b7)), out7); }
void orc_pose_wTo(const double* p7, double* T16) { pose_wTo(pose_from7(p7), T16); }
void orc_pose_oTw(const double* p7, double* T16) { pose_oTw(pose_from7(p7), T16); }
void orc_pose_from_mat4(const double* T16, double* out7) { pose_to7(pose_from_mat4(T16), out7); }
void orc_plane_exmap(const double* p4, const double* d3, double* out4) {
  Plane p = plane_from_vec4(p4); Plane r = plane_exmap(p, d3); std::memcpy(out4, r.v, 4 * sizeof(double));
}
void orc_plane_transform(const double* T16, const double* p4, double* out4) {
  Plane p = plane_from_vec4(p4); Plane r = plane_transform_T(T16, p); std::memcpy(out4, r.v, 4 * sizeof(double));
}
void orc_plane_log_error(const double* l4, const double* m4, double* e3) {
  Plane l = plane_from_vec4(l4), m = plane_from_vec4(m4); plane_log_error(l, m, e3);
}
double orc_standard_rad(double t) { return standard_rad(t); }

int orc_popup_fit_frames(int, int n_frames, const int* seg_ptr, const float* segs, const float* invK, const float* Ts,
                         float dist_thre, int mode, float* planes_world, float* planes_sensor, float* dist, int* good) {
  for (int f = 0; f < n_frames; f++) {
    int s0 = seg_ptr[f], n = seg_ptr[f + 1] - s0;
    long row0 = s0 + f;
    if (n <= 0) continue;
    popup_fit(segs + 4 * s0, n, invK, Ts + 16 * f, dist_thre, mode, planes_world ? planes_world + 4 * row0 : nullptr,
              planes_sensor ? planes_sensor + 4 * row0 : nullptr, dist ? dist + row0 : nullptr, good ? good + row0 : nullptr,
              nullptr, nullptr);
  }
  return 0;
}

// Mapper_mono::update_plane_measurement  pop_planar_slam/src/Mapping.cpp:590-607
//   for every frame: latest_pose = pose_vertex->value().wTo() cast to float (:598-599);
//   update_plane_equation_from_seg(ground_seg2d_lines, inv_calib, pose, all_planes_sensor_new) (:600);
//   for every kept plane: Plane3d(row.cast<double>()) (:602) -> pose_plane_facs[plane_id]->set_measurement (:603)
int orc_refresh_plane_measurements(void* h, int n_frames, const int* frame_pose, const int* seg_ptr, const float* segs,
                                   const float* invK, int n_map, const int* map_fid, const int* map_frame, const int* map_row,
                                   double* new_meas) {
  Slam* sl = S(h);
  const int n_rows = seg_ptr[n_frames] + n_frames;
  std::vector<float> ps((size_t)n_rows * 4, 0.f);
  for (int f = 0; f < n_frames; f++) {
    const int id = frame_pose[f];
    if (id < 0 || id >= (int)sl->nodes.size() || !sl->nodes[id].alive || sl->nodes[id].kind != NODE_POSE) { g_err = "not a pose node"; return -1; }
    double Td[16];
    pose_wTo(sl->nodes[id].pose, Td);
    float T[16];
    for (int i = 0; i < 16; i++) T[i] = (float)Td[i];
    const int s0 = seg_ptr[f], n = seg_ptr[f + 1] - s0;
    if (n <= 0) continue;
    popup_fit(segs + 4 * s0, n, invK, T, 0.f, 0, nullptr, ps.data() + 4 * (size_t)(s0 + f), nullptr, nullptr, nullptr, nullptr);
  }
  for (int m = 0; m < n_map; m++) {
    const int f = map_frame[m], fid = map_fid[m];
    if (fid < 0 || fid >= (int)sl->factors.size() || !sl->factors[fid].alive || sl->factors[fid].kind != F_POSE_PLANE) { g_err = "not a pose-plane factor"; return -1; }
    const float* r = ps.data() + 4 * (size_t)(seg_ptr[f] + f + map_row[m]);
    double v[4] = {(double)r[0], (double)r[1], (double)r[2], (double)r[3]};
    sl->set_measurement(fid, v);
    if (new_meas) std::memcpy(new_meas + 4 * (size_t)m, sl->factors[fid].meas, 4 * sizeof(double));
  }
  return 0;
}

// Plane3d::project_to_plane  PPS/src/isam_plane3d.h:172-177 (normal() :149-151, d() :154-156) as applied to the
// polygon vertices by Mapper_mono::reproj_to_newplane, Mapping.cpp:609-632 (float -> double -> float)
int orc_project_to_planes(void* h, int n_points, const int* plane_of_point, const float* in, float* out) {
  Slam* sl = S(h);
  for (int i = 0; i < n_points; i++) {
    const int id = plane_of_point[i];
    if (id < 0 || id >= (int)sl->nodes.size() || !sl->nodes[id].alive || sl->nodes[id].kind != NODE_PLANE) { g_err = "not a plane node"; return -1; }
    const double* pl = sl->nodes[id].plane.v;
    const double nn = std::sqrt(pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2]);
    const double nx = pl[0] / nn, ny = pl[1] / nn, nz = pl[2] / nn;
    const double dd = -pl[3] / nn;
    const double px = in[3 * (size_t)i], py = in[3 * (size_t)i + 1], pz = in[3 * (size_t)i + 2];
    const double t = (nx * px + ny * py + nz * pz) - dd;
    out[3 * (size_t)i] = (float)(px - nx * t);
    out[3 * (size_t)i + 1] = (float)(py - ny * t);
    out[3 * (size_t)i + 2] = (float)(pz - nz * t);
  }
  return 0;
}

}  // extern "C"
