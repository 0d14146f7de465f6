// ORACLE (test infrastructure, NOT product code).
//
// CPU restatement of the value types and error functions on the hot path of
// shichaoy/pop_up_slam (iSAM 1.7 + the application's plane vertex / pose-plane
// edge).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, load or call anything under oracle/.
//
// PARITY STATUS: pinned for the pose-graph half, "parity unpinned" for the plane half.
//   The reference ships no tests or golden outputs and cannot be compiled in this image (Eigen3 / CHOLMOD / Boost
//   absent), but it holds one known-answer fixture for this path: ISAM/data/sphere2500.txt with
//   ISAM/data/groundtruth/sphere2500_groundtruth.txt.  tests/test_oracle.py pins Pose3d / Rot3d, the odometry and
//   prior factors, the sqrt-information handling, the Loader conventions and the Gauss-Newton + sparse-Cholesky
//   solve against it: the ground-truth graph (4 949 edges, 2 450 loop closures) is consistent to chi2 = 1.8e-3 when
//   its sequential edges are chained through oplus, and the noisy graph optimises to a normalised chi2 of 0.996 and
//   to within 0.8 % of the sphere radius of the ground truth.
//   Plane3d, the pose-plane factor and the pop-up fit have no reference-held fixture: "parity unpinned", self-pinned
//   only (numeric-vs-analytic Jacobians, exmap/log round trips, independent numpy geometry, scipy splu cross-check of
//   the linear solve, noise-free recovery of ground truth).
//
// All paths cited below are relative to /root/reference.
//   ISAM = pop_planar_slam/Thirdparty/isam ; PPS = pop_planar_slam
#pragma once
#include <cmath>
#include <cstring>
#include <limits>

namespace orc {

static const double kPi = 3.14159265358979323846;
static const double kTwoPi = 2.0 * 3.14159265358979323846;

// ISAM/include/isam/util.h:101-108  standardRad
inline double standard_rad(double t) {
  if (t >= 0.) {
    t = std::fmod(t + kPi, kTwoPi) - kPi;
  } else {
    t = std::fmod(t - kPi, -kTwoPi) + kPi;
  }
  return t;
}

struct Quat {
  double w, x, y, z;
};

// Eigen::Quaterniond operator* (Hamilton product), used at Rot3d.h:232 and
// isam_plane3d.h:108,288.
inline Quat qmul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}

// ISAM/include/isam/Rot3d.h:126-136  Rot3d::delta3_to_quat (note the +theta^2/48
// in the small-angle branch is the reference's own expression).
inline Quat rot_delta3_to_quat(const double d[3]) {
  double theta = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  double S;
  if (theta < 0.0001) {
    S = 0.5 + theta * theta / 48.;
  } else {
    S = std::sin(0.5 * theta) / theta;
  }
  double C = std::cos(0.5 * theta);
  return Quat{C, S * d[0], S * d[1], S * d[2]};
}

// ISAM/include/isam/Rot3d.h:100-112  euler_to_quat
inline Quat euler_to_quat(double yaw, double pitch, double roll) {
  double sy = std::sin(yaw * 0.5), cy = std::cos(yaw * 0.5);
  double sp = std::sin(pitch * 0.5), cp = std::cos(pitch * 0.5);
  double sr = std::sin(roll * 0.5), cr = std::cos(roll * 0.5);
  Quat q;
  q.w = cr * cp * cy + sr * sp * sy;
  q.x = sr * cp * cy - cr * sp * sy;
  q.y = cr * sp * cy + sr * cp * sy;
  q.z = cr * cp * sy - sr * sp * cy;
  return q;
}

// ISAM/include/isam/Rot3d.h:114-124  quat_to_euler
inline void quat_to_euler(const Quat& q, double& yaw, double& pitch, double& roll) {
  const double q0 = q.w, q1 = q.x, q2 = q.y, q3 = q.z;
  roll = std::atan2(2.0 * (q0 * q1 + q2 * q3), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3);
  pitch = std::asin(2.0 * (q0 * q2 - q3 * q1));
  yaw = std::atan2(2.0 * (q0 * q3 + q1 * q2), q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3);
}

// Rot3d.h:96-98 quat_to_wRo = Eigen::Matrix3d(quat) (Eigen's
// QuaternionBase::toRotationMatrix; no normalisation).  R row-major [9].
inline void quat_to_wRo(const Quat& q, double R[9]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Rot3d.h:92-94 wRo_to_quat = Eigen::Quaterniond(Matrix3d) (Eigen's
// quaternionbase_assign_impl<Matrix3>, the trace / largest-diagonal method).
inline Quat wRo_to_quat(const double R[9]) {
  Quat q;
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (R[7] - R[5]) * t;
    q.y = (R[2] - R[6]) * t;
    q.z = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (R[k * 3 + j] - R[j * 3 + k]) * t;
    v[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    v[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}

// ---------------------------------------------------------------------------
// isam::Pose3d  (ISAM/include/isam/Pose3d.h:70-274): translation + quaternion
// ---------------------------------------------------------------------------
struct Pose {
  double t[3];
  Quat q;
  Pose() : t{0, 0, 0}, q{1, 0, 0, 0} {}
};

// Pose3d(x,y,z,yaw,pitch,roll)  Pose3d.h:86, Rot3d::set Rot3d.h:219-226
inline Pose pose_from_xyzypr(const double v[6]) {
  Pose p;
  p.t[0] = v[0]; p.t[1] = v[1]; p.t[2] = v[2];
  p.q = euler_to_quat(v[3], v[4], v[5]);
  return p;
}

// Pose3d::vector()  Pose3d.h:138-145
inline void pose_vector(const Pose& p, double v[6]) {
  v[0] = p.t[0]; v[1] = p.t[1]; v[2] = p.t[2];
  quat_to_euler(p.q, v[3], v[4], v[5]);
}

// Pose3d::set(Vector6d)  Pose3d.h:152-155  (used by NodeT::update0 when
// numericalDiff restores the linearisation point, Node.h:136)
inline Pose pose_set_vector(const double v[6]) {
  double w[6] = {v[0], v[1], v[2], standard_rad(v[3]), standard_rad(v[4]), standard_rad(v[5])};
  return pose_from_xyzypr(w);
}

// Pose3d::exmap  Pose3d.h:131-136 (Point3d::exmap Point3d.h:63-69 additive;
// Rot3d::exmap Rot3d.h:229-233 right-multiplication)
inline Pose pose_exmap(const Pose& p, const double d[6]) {
  Pose r = p;
  r.t[0] += d[0]; r.t[1] += d[1]; r.t[2] += d[2];
  r.q = qmul(p.q, rot_delta3_to_quat(d + 3));
  return r;
}

// 4x4 row-major helpers
inline void mat4_mul(const double A[16], const double B[16], double C[16]) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j];
      C[i * 4 + j] = s;
    }
}

// Pose3d::wTo  Pose3d.h:188-194
inline void pose_wTo(const Pose& p, double T[16]) {
  double R[9];
  quat_to_wRo(p.q, R);
  T[0] = R[0]; T[1] = R[1]; T[2] = R[2];  T[3] = p.t[0];
  T[4] = R[3]; T[5] = R[4]; T[6] = R[5];  T[7] = p.t[1];
  T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = p.t[2];
  T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

// Pose3d::oTw  Pose3d.h:204-213
inline void pose_oTw(const Pose& p, double T[16]) {
  double R[9];
  quat_to_wRo(p.q, R);
  // oRw = R^T ; C = -oRw * t
  double oRw[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]};
  double C[3];
  for (int i = 0; i < 3; i++) C[i] = -(oRw[i * 3] * p.t[0] + oRw[i * 3 + 1] * p.t[1] + oRw[i * 3 + 2] * p.t[2]);
  T[0] = oRw[0]; T[1] = oRw[1]; T[2] = oRw[2];  T[3] = C[0];
  T[4] = oRw[3]; T[5] = oRw[4]; T[6] = oRw[5];  T[7] = C[1];
  T[8] = oRw[6]; T[9] = oRw[7]; T[10] = oRw[8]; T[11] = C[2];
  T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

// Pose3d(const Eigen::MatrixXd& m) 4x4 branch  Pose3d.h:92-98
inline Pose pose_from_mat4(const double m[16]) {
  double T[16];
  for (int i = 0; i < 16; i++) T[i] = m[i] / m[15];
  Pose p;
  p.t[0] = T[3]; p.t[1] = T[7]; p.t[2] = T[11];
  double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  p.q = wRo_to_quat(R);
  // DEVIATION from the reference (the only one in this file): renormalise.  Upstream keeps the raw
  // Eigen::Quaterniond(Matrix3d); the matrix->quaternion->matrix cycle of oplus() amplifies the error of
  // |q|^2 by ~tan^2(theta/2) per cycle, so dead-reckoning a few hundred poses at camera-like attitudes
  // reaches |q| = 1.01 and the Euler residuals stop being functions of a rotation.  Within the reference's
  // own regime (80 key-frames) both versions agree to rounding.
  double nq = std::sqrt(p.q.w * p.q.w + p.q.x * p.q.x + p.q.y * p.q.y + p.q.z * p.q.z);
  p.q.w /= nq; p.q.x /= nq; p.q.y /= nq; p.q.z /= nq;
  return p;
}

// Pose3d::oplus  Pose3d.h:222-224
inline Pose pose_oplus(const Pose& a, const Pose& d) {
  double A[16], D[16], C[16];
  pose_wTo(a, A); pose_wTo(d, D);
  mat4_mul(A, D, C);
  return pose_from_mat4(C);
}

// Pose3d::ominus  Pose3d.h:233-235   (a.ominus(b) = b.oTw() * a.wTo())
inline Pose pose_ominus(const Pose& a, const Pose& b) {
  double B[16], A[16], C[16];
  pose_oTw(b, B); pose_wTo(a, A);
  mat4_mul(B, A, C);
  return pose_from_mat4(C);
}

// ---------------------------------------------------------------------------
// isam::Plane3d  (PPS/src/isam_plane3d.h:27-193): unit homogeneous 4-vector
// ---------------------------------------------------------------------------
struct Plane {
  double v[4];  // a,b,c,d ; ||v||_4 = 1
  Plane() : v{1, 0, 0, 0} {}
};

// Plane3d::_normalize  isam_plane3d.h:36-38 (Eigen normalize())
inline void normalize4(double v[4]) {
  double z = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  if (z > 0) {
    double n = std::sqrt(z);
    v[0] /= n; v[1] /= n; v[2] /= n; v[3] /= n;
  }
}

// Plane3d(const Eigen::Vector4d&)  isam_plane3d.h:59-66
inline Plane plane_from_vec4(const double a[4]) {
  Plane p;
  p.v[0] = a[0]; p.v[1] = a[1]; p.v[2] = a[2]; p.v[3] = a[3];
  normalize4(p.v);
  return p;
}

// boost::math::sinc_pi<double> (Boost.Math special_functions/sinc.hpp; version
// unpinned by the reference, apt libboost of Ubuntu 14.04/16.04): sin(x)/x with
// a Taylor series below eps^(1/4).
inline double boost_sinc_pi(double x) {
  const double taylor_0_bound = std::numeric_limits<double>::epsilon();
  const double taylor_2_bound = std::sqrt(taylor_0_bound);
  const double taylor_n_bound = std::sqrt(taylor_2_bound);
  if (std::fabs(x) >= taylor_n_bound) return std::sin(x) / x;
  double result = 1;
  if (std::fabs(x) >= taylor_0_bound) {
    double x2 = x * x;
    result -= x2 / 6;
    if (std::fabs(x) >= taylor_2_bound) result += (x2 * x2) / 120;
  }
  return result;
}

// Plane3d::delta3_to_quat  isam_plane3d.h:78-92
inline Quat plane_delta3_to_quat(const double d[3]) {
  double theta = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  double S = 0.5 * boost_sinc_pi(0.5 * theta);
  double C = std::cos(0.5 * theta);
  return Quat{C, S * d[0], S * d[1], S * d[2]};
}

// Plane3d::exmap_3dof  isam_plane3d.h:101-127 with plane_type == -1 (always, in
// PPS: Mapping.cpp:498 is commented out): q' = Q(delta) * q_pi, renormalise.
inline Plane plane_exmap(const Plane& p, const double d[3]) {
  Quat qp{p.v[3], p.v[0], p.v[1], p.v[2]};  // toQuaternion :74-76 (w=d; x,y,z=a,b,c)
  Quat q = qmul(plane_delta3_to_quat(d), qp);
  double c[4] = {q.x, q.y, q.z, q.w};       // q.coeffs() is (x,y,z,w)
  return plane_from_vec4(c);
}

// Plane3d::transform_to  isam_plane3d.h:180-182 : Plane3d(wTo^T * pi)
// Plane3d::transform_from :186-188 : Plane3d(oTw^T * pi) -- same function, other matrix.
inline Plane plane_transform_T(const double T[16], const Plane& p) {
  double r[4];
  for (int i = 0; i < 4; i++) {
    double s = 0;
    for (int j = 0; j < 4; j++) s += T[j * 4 + i] * p.v[j];
    r[i] = s;
  }
  return plane_from_vec4(r);
}

// The log map shared by Pose3d_Plane3d_Factor::basic_error (isam_plane3d.h:285-294),
// Plane3d_Factor::basic_error (:456-466): dq = q(l) * conj(q(m)); e = axis*angle
// with the angle wrapped to (-pi,pi].  Eigen::AngleAxisd(Quaterniond) as of
// Eigen >= 3.3 (atan2 form; Eigen 3.2's acos form followed by the reference's own
// wrap at :291-292 yields the same vector).
inline void plane_log_error(const Plane& l, const Plane& m, double e[3]) {
  Quat q{l.v[3], l.v[0], l.v[1], l.v[2]};
  Quat qmc{m.v[3], -m.v[0], -m.v[1], -m.v[2]};
  Quat dq = qmul(q, qmc);
  double n = std::sqrt(dq.x * dq.x + dq.y * dq.y + dq.z * dq.z);
  double angle, ax[3];
  if (n != 0.0) {
    angle = 2.0 * std::atan2(n, std::fabs(dq.w));
    if (dq.w < 0) n = -n;
    ax[0] = dq.x / n; ax[1] = dq.y / n; ax[2] = dq.z / n;
  } else {
    angle = 0; ax[0] = 1; ax[1] = 0; ax[2] = 0;
  }
  if (angle > kPi) angle -= 2. * kPi;
  if (angle < -kPi) angle += 2. * kPi;
  e[0] = ax[0] * angle; e[1] = ax[1] * angle; e[2] = ax[2] * angle;
}

// get_wall_plane_equation  pop_planar_slam/src/isam_plane3d.cpp:20-55 (with ray_plane_interact :14-18) for ONE
// segment (two rays, 3 x 2 column layout of ground_edge_ray): output is not normalised.
inline void wall_plane_from_rays(const double rays[6], const double T[16], double out[4]) {
  // ground_plane_sensor = transToWorld^T * (0,0,-1,0)
  const double gw[4] = {0, 0, -1, 0};
  double gs[4];
  for (int i = 0; i < 4; i++) gs[i] = T[0 * 4 + i] * gw[0] + T[1 * 4 + i] * gw[1] + T[2 * 4 + i] * gw[2] + T[3 * 4 + i] * gw[3];
  double P[2][3];
  for (int k = 0; k < 2; k++) {
    const double* r = rays + 3 * k;
    const double frac = -gs[3] / (gs[0] * r[0] + gs[1] * r[1] + gs[2] * r[2]);
    for (int i = 0; i < 3; i++) P[k][i] = frac * r[i];
  }
  const double t1[3] = {P[1][0] - P[0][0], P[1][1] - P[0][1], P[1][2] - P[0][2]};
  const double t2[3] = {gs[0], gs[1], gs[2]};
  const double nrm[3] = {t1[1] * t2[2] - t1[2] * t2[1], t1[2] * t2[0] - t1[0] * t2[2], t1[0] * t2[1] - t1[1] * t2[0]};
  out[0] = nrm[0]; out[1] = nrm[1]; out[2] = nrm[2];
  out[3] = -(nrm[0] * P[0][0] + nrm[1] * P[0][1] + nrm[2] * P[0][2]);
}

// Pose3d_Plane3d_Factor2::basic_error  isam_plane3d.h:375-419 (_base == NULL): the measured plane is re-popped from
// the precomputed ground-edge rays with the current pose (:384-386: row 0, 4-normalised), then the same log map.
inline void pose_plane2_basic_error(const Pose& pose, const Plane& global_plane, const double rays[6], double e[3]) {
  double T[16];
  pose_wTo(pose, T);
  Plane local = plane_transform_T(T, global_plane);
  double raw[4];
  wall_plane_from_rays(rays, T, raw);
  Plane meas;
  std::memcpy(meas.v, raw, sizeof(raw));
  normalize4(meas.v);
  plane_log_error(local, meas, e);
}

// Pose3d_Plane3d_Factor::basic_error  isam_plane3d.h:271-304 (useRelative=false,
// Mapping.cpp:21 => _base == NULL)
inline void pose_plane_basic_error(const Pose& pose, const Plane& global_plane, const Plane& meas, double e[3]) {
  double T[16];
  pose_wTo(pose, T);
  Plane local = plane_transform_T(T, global_plane);
  plane_log_error(local, meas, e);
}

// Pose3d_Pose3d_Factor::basic_error  ISAM/include/isam/slam3d.h:174-191 (2-node branch)
inline void odometry_basic_error(const Pose& p1, const Pose& p2, const double meas6[6], double e[6]) {
  Pose pred = pose_ominus(p2, p1);
  double v[6];
  pose_vector(pred, v);
  for (int i = 0; i < 6; i++) e[i] = v[i] - meas6[i];
  e[3] = standard_rad(e[3]); e[4] = standard_rad(e[4]); e[5] = standard_rad(e[5]);
}

// Pose3d_Factor::basic_error  ISAM/include/isam/slam3d.h:82-88
inline void pose_prior_basic_error(const Pose& p, const double meas6[6], double e[6]) {
  double v[6];
  pose_vector(p, v);
  for (int i = 0; i < 6; i++) e[i] = v[i] - meas6[i];
  e[3] = standard_rad(e[3]); e[4] = standard_rad(e[4]); e[5] = standard_rad(e[5]);
}

// ISAM/include/isam/robust.h:101-118
inline double cost_huber(double d, double b) {
  double abs_d = std::fabs(d);
  if (abs_d < b) return d * d;
  return 2 * b * abs_d - b * b;
}
inline double cost_pseudo_huber(double d, double b) {
  double b2 = b * b;
  return 2 * b2 * (std::sqrt(1 + d * d / b2) - 1);
}

}  // namespace orc
