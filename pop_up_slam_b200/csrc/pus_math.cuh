// pus_math.cuh -- value types, manifold updates, residuals and closed-form Jacobians of the
// plane-SLAM factors, as __host__ __device__ functions shared by the CUDA kernels and the
// host-side graph container (factor initialisation, value helpers).
//
// What each function reproduces (paths relative to the reference checkout;
// ISAM = pop_planar_slam/Thirdparty/isam, PPS = pop_planar_slam):
//   pose value / exmap ........ ISAM/include/isam/Pose3d.h:131-136, Rot3d.h:126-136,229-233
//   plane value / exmap ....... PPS/src/isam_plane3d.h:27-127
//   pose-plane residual ....... PPS/src/isam_plane3d.h:271-304   (Pose3d_Plane3d_Factor::basic_error)
//   plane prior residual ...... PPS/src/isam_plane3d.h:450-473   (Plane3d_Factor::basic_error)
//   odometry residual ......... ISAM/include/isam/slam3d.h:174-191 (Pose3d_Pose3d_Factor::basic_error)
//   pose prior residual ....... ISAM/include/isam/slam3d.h:82-88   (Pose3d_Factor::basic_error)
//   sqrt-information + robust . ISAM/include/isam/Factor.h:67-77, robust.h:101-118
// The reference differentiates these numerically (ISAM/isamlib/numericalDiff.cpp:41-87, eps=1e-4);
// the kernels use the exact derivatives through the same exmaps ("Jacobians").
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PUS_HD __host__ __device__ __forceinline__
#else
#define PUS_HD inline
#endif

namespace pus {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;

enum RobustKind { ROBUST_NONE = 0, ROBUST_HUBER = 1, ROBUST_PSEUDO_HUBER = 2 };

// ISAM/include/isam/util.h:101-108
PUS_HD double standard_rad(double t) {
  if (t >= 0.) t = fmod(t + kPi, kTwoPi) - kPi;
  else t = fmod(t - kPi, -kTwoPi) + kPi;
  return t;
}

// q = (w,x,y,z); Hamilton product
PUS_HD void quat_mul(const double* a, const double* b, double* r) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
  double z = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}

// rotation matrix of a quaternion (no normalisation, as Eigen's toRotationMatrix), row-major
PUS_HD void quat_to_R(const double* q, double* R) {
  const double tx = 2 * q[1], ty = 2 * q[2], tz = 2 * q[3];
  const double twx = tx * q[0], twy = ty * q[0], twz = tz * q[0];
  const double txx = tx * q[1], txy = ty * q[1], txz = tz * q[1];
  const double tyy = ty * q[2], tyz = tz * q[2], tzz = tz * q[3];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Eigen::Quaterniond(Matrix3d) -- used by Pose3d(Matrix4d) (Pose3d.h:92-98, Rot3d.h:92-94)
PUS_HD void R_to_quat(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (R[7] - R[5]) * t; q[2] = (R[2] - R[6]) * t; q[3] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    v[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    v[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    q[1] = v[0]; q[2] = v[1]; q[3] = v[2];
  }
}

// Rot3d::euler_to_quat  Rot3d.h:100-112
PUS_HD void euler_to_quat(double yaw, double pitch, double roll, double* q) {
  double sy = sin(yaw * 0.5), cy = cos(yaw * 0.5);
  double sp = sin(pitch * 0.5), cp = cos(pitch * 0.5);
  double sr = sin(roll * 0.5), cr = cos(roll * 0.5);
  q[0] = cr * cp * cy + sr * sp * sy;
  q[1] = sr * cp * cy - cr * sp * sy;
  q[2] = cr * sp * cy + sr * cp * sy;
  q[3] = cr * cp * sy - sr * sp * cy;
}

// Rot3d::quat_to_euler  Rot3d.h:114-124
PUS_HD void quat_to_euler(const double* q, double& yaw, double& pitch, double& roll) {
  const double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  roll = atan2(2.0 * (q0 * q1 + q2 * q3), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3);
  pitch = asin(2.0 * (q0 * q2 - q3 * q1));
  yaw = atan2(2.0 * (q0 * q3 + q1 * q2), q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3);
}

// Rot3d::delta3_to_quat  Rot3d.h:126-136 (small-angle branch as written upstream)
PUS_HD void rot_delta_quat(const double* d, double* q) {
  double theta = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  double S = (theta < 0.0001) ? (0.5 + theta * theta / 48.) : (sin(0.5 * theta) / theta);
  q[0] = cos(0.5 * theta); q[1] = S * d[0]; q[2] = S * d[1]; q[3] = S * d[2];
}

// Plane3d::delta3_to_quat  isam_plane3d.h:78-92: S = 0.5*boost::math::sinc_pi(theta/2)
PUS_HD void plane_delta_quat(const double* d, double* q) {
  double theta = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  double x = 0.5 * theta, sinc;
  if (x >= 1.220703125e-4) {  // eps^(1/4): boost's taylor_n_bound for double
    sinc = sin(x) / x;
  } else {
    sinc = 1.0;
    if (x >= 2.220446049250313e-16) {
      double x2 = x * x;
      sinc -= x2 / 6;
      if (x >= 1.4901161193847656e-8) sinc += (x2 * x2) / 120;
    }
  }
  double S = 0.5 * sinc;
  q[0] = cos(x); q[1] = S * d[0]; q[2] = S * d[1]; q[3] = S * d[2];
}

// Pose3d::exmap  Pose3d.h:131-136.  pose = (x,y,z,qw,qx,qy,qz)
PUS_HD void pose_exmap(const double* p, const double* d, double* out) {
  double dq[4], q[4];
  rot_delta_quat(d + 3, dq);
  quat_mul(p + 3, dq, q);
  out[0] = p[0] + d[0]; out[1] = p[1] + d[1]; out[2] = p[2] + d[2];
  out[3] = q[0]; out[4] = q[1]; out[5] = q[2]; out[6] = q[3];
}

PUS_HD void normalize4(double* v) {
  double z = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  if (z > 0) {
    double n = sqrt(z);
    v[0] /= n; v[1] /= n; v[2] /= n; v[3] /= n;
  }
}

// Plane3d::exmap_3dof  isam_plane3d.h:101-127 (plane_type == -1)
PUS_HD void plane_exmap(const double* p, const double* d, double* out) {
  double dq[4], qp[4] = {p[3], p[0], p[1], p[2]}, q[4];
  plane_delta_quat(d, dq);
  quat_mul(dq, qp, q);
  out[0] = q[1]; out[1] = q[2]; out[2] = q[3]; out[3] = q[0];
  normalize4(out);
}

// ---- host-side helpers for factor initialisation (slam3d.h:123-137, isam_plane3d.h:252-264) ----
PUS_HD void pose_to_T(const double* p, double* T) {  // wTo row-major (Pose3d.h:188-194)
  double R[9];
  quat_to_R(p + 3, R);
  T[0] = R[0]; T[1] = R[1]; T[2] = R[2];  T[3] = p[0];
  T[4] = R[3]; T[5] = R[4]; T[6] = R[5];  T[7] = p[1];
  T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = p[2];
  T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}
PUS_HD void pose_to_Tinv(const double* p, double* T) {  // oTw (Pose3d.h:204-213)
  double R[9];
  quat_to_R(p + 3, R);
  double oRw[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]};
  T[0] = oRw[0]; T[1] = oRw[1]; T[2] = oRw[2];
  T[4] = oRw[3]; T[5] = oRw[4]; T[6] = oRw[5];
  T[8] = oRw[6]; T[9] = oRw[7]; T[10] = oRw[8];
  for (int i = 0; i < 3; i++) T[i * 4 + 3] = -(oRw[i * 3] * p[0] + oRw[i * 3 + 1] * p[1] + oRw[i * 3 + 2] * p[2]);
  T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}
PUS_HD void T_to_pose(const double* m, double* p) {  // Pose3d(Matrix4d) Pose3d.h:92-98
  double T[16];
  for (int i = 0; i < 16; i++) T[i] = m[i] / m[15];
  p[0] = T[3]; p[1] = T[7]; p[2] = T[11];
  double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  R_to_quat(R, p + 3);
  // Deviation from the reference (documented in DESIGN.md): the matrix -> quaternion -> matrix cycle of
  // Pose3d::oplus is not norm-preserving in floating point (the error of |q|^2 is amplified by about
  // tan^2(theta/2) per cycle), so a dead-reckoned chain of a few hundred poses at camera-like attitudes
  // (theta > 90 deg) drifts to |q| = 1.01.  Renormalising changes a unit quaternion only at rounding level.
  double* q = p + 3;
  double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= nq; q[1] /= nq; q[2] /= nq; q[3] /= nq;
}
PUS_HD void mat4_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j];
      C[i * 4 + j] = s;
    }
}
PUS_HD void pose_from_xyzypr(const double* v, double* p) {
  p[0] = v[0]; p[1] = v[1]; p[2] = v[2];
  euler_to_quat(v[3], v[4], v[5], p + 3);
}
PUS_HD void pose_oplus(const double* a, const double* d, double* out) {  // Pose3d.h:222-224
  double A[16], D[16], C[16];
  pose_to_T(a, A); pose_to_T(d, D); mat4_mul(A, D, C); T_to_pose(C, out);
}
PUS_HD void pose_ominus(const double* a, const double* b, double* out) {  // Pose3d.h:233-235
  double B[16], A[16], C[16];
  pose_to_Tinv(b, B); pose_to_T(a, A); mat4_mul(B, A, C); T_to_pose(C, out);
}
// Plane3d::transform_to / transform_from: normalize(T^T pi)  isam_plane3d.h:180-188
PUS_HD void plane_transform_T(const double* T, const double* p, double* out) {
  for (int i = 0; i < 4; i++) {
    double s = 0;
    for (int j = 0; j < 4; j++) s += T[j * 4 + i] * p[j];
    out[i] = s;
  }
  normalize4(out);
}

// ---- robust cost per residual component (Factor.h:67-77 + robust.h:101-118) ----
// r <- sign(r) sqrt(rho(r)); returns d/dr of that map (the Jacobian row scale).
PUS_HD double robustify(int kind, double b, double& r) {
  if (kind == ROBUST_NONE) return 1.0;
  double a = fabs(r), w = 1.0, rho;
  if (kind == ROBUST_HUBER) {
    if (a < b) rho = r * r;
    else { rho = 2 * b * a - b * b; w = b / sqrt(rho); }
  } else {
    double b2 = b * b, sq = sqrt(1 + r * r / b2);
    rho = 2 * b2 * (sq - 1);
    w = (a > 1e-150) ? a / (sq * sqrt(rho)) : 1.0;
  }
  r = ((r >= 0) ? 1. : (-1.)) * sqrt(rho);
  return w;
}

// ---- log-map between a (unit) plane l and a measured plane m, with derivative wrt l ----
// e = axis*angle of dq = q(l) * conj(q(m)), angle wrapped to (-pi,pi] (isam_plane3d.h:285-294;
// Eigen::AngleAxisd(Quaterniond), atan2 form).  G4 (3x4 row-major) = de/dl, or nullptr.
PUS_HD void plane_log(const double* l, const double* m, double* e, double* G4, double* G4m = nullptr) {
  const double mw = m[3];
  double a0 = mw * l[0] + (m[1] * l[2] - m[2] * l[1]) - l[3] * m[0];
  double a1 = mw * l[1] + (m[2] * l[0] - m[0] * l[2]) - l[3] * m[1];
  double a2 = mw * l[2] + (m[0] * l[1] - m[1] * l[0]) - l[3] * m[2];
  double w = m[0] * l[0] + m[1] * l[1] + m[2] * l[2] + mw * l[3];
  double sgn = 1.0;
  if (w < 0) { sgn = -1.0; a0 = -a0; a1 = -a1; a2 = -a2; w = -w; }
  double na2 = a0 * a0 + a1 * a1 + a2 * a2;
  double na = sqrt(na2), den = na2 + w * w;
  double k, c2 = 0, c3 = 0, h0 = 0, h1 = 0, h2 = 0;
  if (na > 1e-12) {
    double theta = 2.0 * atan2(na, w);
    k = theta / na;
    h0 = a0 / na; h1 = a1 / na; h2 = a2 / na;
    c2 = 2.0 * w / den - k;   // coefficient of a^ a^T beyond k*I
    c3 = -2.0 * na / den;
  } else {
    k = 2.0 / w;              // limit |a| -> 0 (w > 0 after the flip, ~1 for unit quaternions)
  }
  e[0] = k * a0; e[1] = k * a1; e[2] = k * a2;
  if (!G4) return;
  // D = [k I + c2 h h^T | c3 h] (3x4);  G4 = D * sgn * M(m),
  // M = [[ mw I + [mv]x , -mv ],[ mv^T , mw ]]
  double h[3] = {h0, h1, h2};
  for (int i = 0; i < 3; i++) {
    double Da[3];
    for (int j = 0; j < 3; j++) Da[j] = (i == j ? k : 0.0) + c2 * h[i] * h[j];
    double Dw = c3 * h[i];
    // (Da * (mw I + [mv]x))_j = mw*Da[j] + (Da x mv... ) : row-vector times skew: (v^T [m]x)_j = (m x v)... use explicit
    // [m]x = [[0,-m2,m1],[m2,0,-m0],[-m1,m0,0]] ; (Da^T [m]x) = (Da1*m2 - Da2*m1, Da2*m0 - Da0*m2, Da0*m1 - Da1*m0)
    double g0 = mw * Da[0] + (Da[1] * m[2] - Da[2] * m[1]) + Dw * m[0];
    double g1 = mw * Da[1] + (Da[2] * m[0] - Da[0] * m[2]) + Dw * m[1];
    double g2 = mw * Da[2] + (Da[0] * m[1] - Da[1] * m[0]) + Dw * m[2];
    double g3 = -(Da[0] * m[0] + Da[1] * m[1] + Da[2] * m[2]) + Dw * mw;
    G4[i * 4 + 0] = sgn * g0; G4[i * 4 + 1] = sgn * g1; G4[i * 4 + 2] = sgn * g2; G4[i * 4 + 3] = sgn * g3;
    if (G4m) {
      // de/dm = D * sgn * N(l),  N = [[ -[lv]x - lw I , lv ],[ lv^T , lw ]]   (a = mw lv + mv x lv - lw mv, w = mv.lv + mw lw)
      // (Da^T [lv]x) = (Da1*l2 - Da2*l1, Da2*l0 - Da0*l2, Da0*l1 - Da1*l0)
      const double n0 = -(Da[1] * l[2] - Da[2] * l[1]) - l[3] * Da[0] + Dw * l[0];
      const double n1 = -(Da[2] * l[0] - Da[0] * l[2]) - l[3] * Da[1] + Dw * l[1];
      const double n2 = -(Da[0] * l[1] - Da[1] * l[0]) - l[3] * Da[2] + Dw * l[2];
      const double n3 = (Da[0] * l[0] + Da[1] * l[1] + Da[2] * l[2]) + Dw * l[3];
      G4m[i * 4 + 0] = sgn * n0; G4m[i * 4 + 1] = sgn * n1; G4m[i * 4 + 2] = sgn * n2; G4m[i * 4 + 3] = sgn * n3;
    }
  }
}

// ---- Pose3d_Plane3d_Factor2 (isam_plane3d.h:314-424): the measured wall plane is re-popped from the two precomputed
// ground-edge rays with the CURRENT pose inside the residual: get_wall_plane_equation (isam_plane3d.cpp:20-55, double):
//   gs = wTo^T (0,0,-1,0) = (-R[2,:]^T, -t_z) ;  P_k = -gs_d / (gs_n . r_k) r_k ;  n = (P1 - P0) x gs_n ;  d = -n . P0
// followed by the 4-normalisation of isam_plane3d.h:385.  rays = (r0, r1).  m4 = the normalised plane; when dm (4x6
// row-major) is given it receives d m / d (delta_t, delta_phi) for the pose exmap (t += dt, R <- R Exp(dphi)).
PUS_HD void popup_plane_from_rays(const double* R, const double* t, const double* rays, double* m4, double* dm) {
  const double g[3] = {-R[6], -R[7], -R[8]}, h = t[2];
  const double* r0 = rays;
  const double* r1 = rays + 3;
  const double s0 = g[0] * r0[0] + g[1] * r0[1] + g[2] * r0[2], s1 = g[0] * r1[0] + g[1] * r1[1] + g[2] * r1[2];
  const double f0 = h / s0, f1 = h / s1;
  const double P0[3] = {f0 * r0[0], f0 * r0[1], f0 * r0[2]}, P1[3] = {f1 * r1[0], f1 * r1[1], f1 * r1[2]};
  const double v[3] = {P1[0] - P0[0], P1[1] - P0[1], P1[2] - P0[2]};
  const double n[3] = {v[1] * g[2] - v[2] * g[1], v[2] * g[0] - v[0] * g[2], v[0] * g[1] - v[1] * g[0]};
  const double d = -(n[0] * P0[0] + n[1] * P0[1] + n[2] * P0[2]);
  const double raw[4] = {n[0], n[1], n[2], d};
  const double nn = sqrt(raw[0] * raw[0] + raw[1] * raw[1] + raw[2] * raw[2] + raw[3] * raw[3]);
  for (int i = 0; i < 4; i++) m4[i] = raw[i] / nn;
  if (!dm) return;
  double draw[4][6];
  for (int i = 0; i < 4; i++) for (int k = 0; k < 6; k++) draw[i][k] = 0.0;
  for (int k = 2; k < 6; k++) {   // delta_t_z (k = 2) and delta_phi (k = 3..5); delta_t_x, delta_t_y leave the plane unchanged
    double dh = 0, dg[3] = {0, 0, 0};
    if (k == 2) dh = 1.0;
    else {  // dg = g x e_j
      const int j = k - 3;
      const double ej[3] = {j == 0 ? 1.0 : 0.0, j == 1 ? 1.0 : 0.0, j == 2 ? 1.0 : 0.0};
      dg[0] = g[1] * ej[2] - g[2] * ej[1]; dg[1] = g[2] * ej[0] - g[0] * ej[2]; dg[2] = g[0] * ej[1] - g[1] * ej[0];
    }
    const double c0 = dh / s0 - h * (r0[0] * dg[0] + r0[1] * dg[1] + r0[2] * dg[2]) / (s0 * s0);
    const double c1 = dh / s1 - h * (r1[0] * dg[0] + r1[1] * dg[1] + r1[2] * dg[2]) / (s1 * s1);
    const double dP0[3] = {c0 * r0[0], c0 * r0[1], c0 * r0[2]}, dP1[3] = {c1 * r1[0], c1 * r1[1], c1 * r1[2]};
    const double dv[3] = {dP1[0] - dP0[0], dP1[1] - dP0[1], dP1[2] - dP0[2]};
    const double dn[3] = {dv[1] * g[2] - dv[2] * g[1] + v[1] * dg[2] - v[2] * dg[1],
                          dv[2] * g[0] - dv[0] * g[2] + v[2] * dg[0] - v[0] * dg[2],
                          dv[0] * g[1] - dv[1] * g[0] + v[0] * dg[1] - v[1] * dg[0]};
    const double dd = -(dn[0] * P0[0] + dn[1] * P0[1] + dn[2] * P0[2]) - (n[0] * dP0[0] + n[1] * dP0[1] + n[2] * dP0[2]);
    draw[0][k] = dn[0]; draw[1][k] = dn[1]; draw[2][k] = dn[2]; draw[3][k] = dd;
  }
  for (int k = 0; k < 6; k++) {   // (I - m m^T) / |raw| * draw
    const double mk = m4[0] * draw[0][k] + m4[1] * draw[1][k] + m4[2] * draw[2][k] + m4[3] * draw[3][k];
    for (int i = 0; i < 4; i++) dm[i * 6 + k] = (draw[i][k] - m4[i] * mk) / nn;
  }
}

// weight a 3-row block by the upper-triangular sqrt-information s = (s00,s01,s02,s11,s12,s22)
PUS_HD void weight3(const double* s, double* v0, double* v1, double* v2) {
  double a = *v0, b = *v1, c = *v2;
  *v0 = s[0] * a + s[1] * b + s[2] * c;
  *v1 = s[3] * b + s[4] * c;
  *v2 = s[5] * c;
}

// Pose-plane edge: weighted, robustified residual r[3] and (optionally) Jp[18] (3x6 row-major,
// columns = delta_t, delta_phi) and Jl[9] (3x3, plane tangent).  SURVEY.md Appendix A.1.
// pose == nullptr => plane prior (R = I, t = 0; Jp untouched).
// rays != nullptr => Pose3d_Plane3d_Factor2: `meas` is ignored, the measured plane is re-popped from the rays with the
// current pose (popup_plane_from_rays) and its dependence on the pose enters Jp.
PUS_HD void pose_plane_linearize(const double* pose, const double* plane, const double* meas, const double* sinf,
                                 int robust_kind, double robust_b, double* r, double* Jp, double* Jl, const double* rays = nullptr) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
  if (pose) { quat_to_R(pose + 3, R); t[0] = pose[0]; t[1] = pose[1]; t[2] = pose[2]; }
  double m2[4], dm[24], G4m[12];
  const bool f2 = (rays != nullptr) && (pose != nullptr);
  if (f2) { popup_plane_from_rays(R, t, rays, m2, (Jp != nullptr) ? dm : nullptr); meas = m2; }
  const double n0 = plane[0], n1 = plane[1], n2 = plane[2], d = plane[3];
  double u[4];
  u[0] = R[0] * n0 + R[3] * n1 + R[6] * n2;
  u[1] = R[1] * n0 + R[4] * n1 + R[7] * n2;
  u[2] = R[2] * n0 + R[5] * n1 + R[8] * n2;
  u[3] = t[0] * n0 + t[1] * n1 + t[2] * n2 + d;
  double s = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + u[3] * u[3]);
  double l[4] = {u[0] / s, u[1] / s, u[2] / s, u[3] / s};
  double e[3], G4[12];
  const bool want_J = (Jl != nullptr);
  plane_log(l, meas, e, want_J ? G4 : nullptr, (want_J && f2) ? G4m : nullptr);
  r[0] = e[0]; r[1] = e[1]; r[2] = e[2];
  weight3(sinf, &r[0], &r[1], &r[2]);
  double wr[3];
  for (int i = 0; i < 3; i++) wr[i] = robustify(robust_kind, robust_b, r[i]);
  if (!want_J) return;
  // G = G4 (I - l l^T)/s
  double G[12];
  for (int i = 0; i < 3; i++) {
    double gl = G4[i * 4] * l[0] + G4[i * 4 + 1] * l[1] + G4[i * 4 + 2] * l[2] + G4[i * 4 + 3] * l[3];
    for (int j = 0; j < 4; j++) G[i * 4 + j] = (G4[i * 4 + j] - gl * l[j]) / s;
  }
  double n[3] = {n0, n1, n2};
  double jp[18], jl[9];
  for (int i = 0; i < 3; i++) {
    const double g0 = G[i * 4], g1 = G[i * 4 + 1], g2 = G[i * 4 + 2], g3 = G[i * 4 + 3];
    // d u / d pose = [[0, [u_v]x],[n^T, 0]]
    jp[i * 6 + 0] = g3 * n0; jp[i * 6 + 1] = g3 * n1; jp[i * 6 + 2] = g3 * n2;
    jp[i * 6 + 3] = g1 * u[2] - g2 * u[1];
    jp[i * 6 + 4] = g2 * u[0] - g0 * u[2];
    jp[i * 6 + 5] = g0 * u[1] - g1 * u[0];
    if (f2 && Jp)   // + (de/dm) (dm/dpose)
      for (int k = 2; k < 6; k++)
        jp[i * 6 + k] += G4m[i * 4] * dm[k] + G4m[i * 4 + 1] * dm[6 + k] + G4m[i * 4 + 2] * dm[12 + k] + G4m[i * 4 + 3] * dm[18 + k];
    // GA = G * [[R^T,0],[t^T,1]]
    double ga[4];
    for (int j = 0; j < 3; j++) ga[j] = g0 * R[j * 3 + 0] + g1 * R[j * 3 + 1] + g2 * R[j * 3 + 2] + g3 * t[j];
    ga[3] = g3;
    // d pi / d delta = 1/2 [[ d I - [n]x ],[ -n^T ]]
    jl[i * 3 + 0] = 0.5 * (d * ga[0] - (ga[1] * n[2] - ga[2] * n[1]) - ga[3] * n[0]);
    jl[i * 3 + 1] = 0.5 * (d * ga[1] - (ga[2] * n[0] - ga[0] * n[2]) - ga[3] * n[1]);
    jl[i * 3 + 2] = 0.5 * (d * ga[2] - (ga[0] * n[1] - ga[1] * n[0]) - ga[3] * n[2]);
  }
  // J <- diag(wr) * S * J
  if (pose && Jp) {
    for (int j = 0; j < 6; j++) {
      double a = jp[j], b = jp[6 + j], c = jp[12 + j];
      weight3(sinf, &a, &b, &c);
      Jp[j] = wr[0] * a; Jp[6 + j] = wr[1] * b; Jp[12 + j] = wr[2] * c;
    }
  }
  for (int j = 0; j < 3; j++) {
    double a = jl[j], b = jl[3 + j], c = jl[6 + j];
    weight3(sinf, &a, &b, &c);
    Jl[j] = wr[0] * a; Jl[3 + j] = wr[1] * b; Jl[6 + j] = wr[2] * c;
  }
}

// E_b^{-1}(pitch, roll): (yaw,pitch,roll) rates from body rates
PUS_HD void euler_rate_inv(double p, double rr, double* Ei) {
  double cp = cos(p), sr = sin(rr), cr = cos(rr), tp = tan(p);
  Ei[0] = 0; Ei[1] = sr / cp; Ei[2] = cr / cp;
  Ei[3] = 0; Ei[4] = cr;      Ei[5] = -sr;
  Ei[6] = 1; Ei[7] = sr * tp; Ei[8] = cr * tp;
}

// weight a 6-vector column by the packed upper-triangular 6x6 sqrt-information (21 entries)
PUS_HD void weight6(const double* s, double* v) {
  double out[6];
  int k = 0;
  for (int r = 0; r < 6; r++) {
    double acc = 0;
    for (int c = r; c < 6; c++) acc += s[k++] * v[c];
    out[r] = acc;
  }
  for (int r = 0; r < 6; r++) v[r] = out[r];
}

// Pose factor (odometry: p2 != nullptr; prior: p2 == nullptr): weighted robustified residual r[6]
// and, when J1 != nullptr, J1[36] (wrt p1) and J2[36] (wrt p2; odometry only), row-major 6x6.
PUS_HD void pose_factor_linearize(const double* p1, const double* p2, const double* meas, const double* sinf,
                                  int robust_kind, double robust_b, double* r, double* J1, double* J2) {
  double e[6];
  double jb1[36], jb2[36];
  const bool want_J = (J1 != nullptr);
  if (want_J) for (int i = 0; i < 36; i++) { jb1[i] = 0; jb2[i] = 0; }
  if (p2) {
    double R1[9], R2[9];
    quat_to_R(p1 + 3, R1); quat_to_R(p2 + 3, R2);
    double R12[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) R12[i * 3 + j] = R1[0 * 3 + i] * R2[0 * 3 + j] + R1[1 * 3 + i] * R2[1 * 3 + j] + R1[2 * 3 + i] * R2[2 * 3 + j];
    double dt[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    double t12[3];
    for (int i = 0; i < 3; i++) t12[i] = R1[0 * 3 + i] * dt[0] + R1[1 * 3 + i] * dt[1] + R1[2 * 3 + i] * dt[2];
    double yaw = atan2(R12[3], R12[0]);
    double sp = -R12[6];
    sp = sp > 1.0 ? 1.0 : (sp < -1.0 ? -1.0 : sp);
    double pitch = asin(sp);
    double roll = atan2(R12[7], R12[8]);
    e[0] = t12[0] - meas[0]; e[1] = t12[1] - meas[1]; e[2] = t12[2] - meas[2];
    e[3] = standard_rad(yaw - meas[3]); e[4] = standard_rad(pitch - meas[4]); e[5] = standard_rad(roll - meas[5]);
    if (want_J) {
      double Ei[9];
      euler_rate_inv(pitch, roll, Ei);
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
          jb1[i * 6 + j] = -R1[j * 3 + i];   // -R1^T
          jb2[i * 6 + j] = R1[j * 3 + i];    //  R1^T
          // -Ei * R12^T
          jb1[(3 + i) * 6 + 3 + j] = -(Ei[i * 3 + 0] * R12[j * 3 + 0] + Ei[i * 3 + 1] * R12[j * 3 + 1] + Ei[i * 3 + 2] * R12[j * 3 + 2]);
          jb2[(3 + i) * 6 + 3 + j] = Ei[i * 3 + j];
        }
      // [t12]x
      jb1[0 * 6 + 4] = -t12[2]; jb1[0 * 6 + 5] = t12[1];
      jb1[1 * 6 + 3] = t12[2];  jb1[1 * 6 + 5] = -t12[0];
      jb1[2 * 6 + 3] = -t12[1]; jb1[2 * 6 + 4] = t12[0];
    }
  } else {
    double yaw, pitch, roll;
    quat_to_euler(p1 + 3, yaw, pitch, roll);
    e[0] = p1[0] - meas[0]; e[1] = p1[1] - meas[1]; e[2] = p1[2] - meas[2];
    e[3] = standard_rad(yaw - meas[3]); e[4] = standard_rad(pitch - meas[4]); e[5] = standard_rad(roll - meas[5]);
    if (want_J) {
      double Ei[9];
      euler_rate_inv(pitch, roll, Ei);
      for (int i = 0; i < 3; i++) {
        jb1[i * 6 + i] = 1.0;
        for (int j = 0; j < 3; j++) jb1[(3 + i) * 6 + 3 + j] = Ei[i * 3 + j];
      }
    }
  }
  for (int i = 0; i < 6; i++) r[i] = e[i];
  weight6(sinf, r);
  double wr[6];
  for (int i = 0; i < 6; i++) wr[i] = robustify(robust_kind, robust_b, r[i]);
  if (!want_J) return;
  for (int j = 0; j < 6; j++) {
    double c1[6], c2[6];
    for (int i = 0; i < 6; i++) { c1[i] = jb1[i * 6 + j]; c2[i] = jb2[i * 6 + j]; }
    weight6(sinf, c1);
    for (int i = 0; i < 6; i++) J1[i * 6 + j] = wr[i] * c1[i];
    if (p2 && J2) {
      weight6(sinf, c2);
      for (int i = 0; i < 6; i++) J2[i * 6 + j] = wr[i] * c2[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Reference-Jacobian mode: the blocks the reference itself forms, by central differences through the exmaps
// (ISAM/isamlib/numericalDiff.cpp:41-87, SYMMETRIC, epsilon = 1e-4; Factor::jacobian Factor.h:126-139).  The product
// defaults to the closed forms above; this mode exists so that a solve can follow the reference's own trajectory
// (its Jacobians carry an O(eps^2) truncation error that decides near-tie accept / reject steps).
// As upstream, the perturbed point is restored through update0(vector0()): a pose is rebuilt from its wrapped Euler
// angles (Pose3d::set, Pose3d.h:152-155), a plane re-normalised (Plane3d::set, isam_plane3d.h:139-142); every column is
// (f(x (+) eps e_j) - f(x (+) -eps e_j)) / (eps + eps) of the weighted, robustified residual.
// ---------------------------------------------------------------------------------------------------------------------
PUS_HD void pose_euler_roundtrip(const double* p, double* out) {
  double yaw, pitch, roll;
  quat_to_euler(p + 3, yaw, pitch, roll);
  out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
  euler_to_quat(standard_rad(yaw), standard_rad(pitch), standard_rad(roll), out + 3);
}

PUS_HD void pose_plane_numeric(const double* pose, const double* plane, const double* meas, const double* sinf, int robust_kind,
                               double robust_b, double* r, double* Jp, double* Jl, const double* rays = nullptr) {
  const double eps = 0.0001;
  double p0[7], l0[4] = {plane[0], plane[1], plane[2], plane[3]};
  normalize4(l0);
  if (pose) {
    pose_euler_roundtrip(pose, p0);
    for (int j = 0; j < 6; j++) {
      double d[6] = {0, 0, 0, 0, 0, 0}, pp[7], yp[3], ym[3];
      d[j] = eps;
      pose_exmap(p0, d, pp);
      pose_plane_linearize(pp, l0, meas, sinf, robust_kind, robust_b, yp, nullptr, nullptr, rays);
      d[j] = -eps;
      pose_exmap(p0, d, pp);
      pose_plane_linearize(pp, l0, meas, sinf, robust_kind, robust_b, ym, nullptr, nullptr, rays);
      for (int i = 0; i < 3; i++) Jp[i * 6 + j] = (yp[i] - ym[i]) / (eps + eps);
    }
  }
  for (int j = 0; j < 3; j++) {
    double d[3] = {0, 0, 0}, lp[4], yp[3], ym[3];
    d[j] = eps;
    plane_exmap(l0, d, lp);
    pose_plane_linearize(pose ? p0 : nullptr, lp, meas, sinf, robust_kind, robust_b, yp, nullptr, nullptr, rays);
    d[j] = -eps;
    plane_exmap(l0, d, lp);
    pose_plane_linearize(pose ? p0 : nullptr, lp, meas, sinf, robust_kind, robust_b, ym, nullptr, nullptr, rays);
    for (int i = 0; i < 3; i++) Jl[i * 3 + j] = (yp[i] - ym[i]) / (eps + eps);
  }
  pose_plane_linearize(pose ? p0 : nullptr, l0, meas, sinf, robust_kind, robust_b, r, nullptr, nullptr, rays);
}

PUS_HD void pose_factor_numeric(const double* p1, const double* p2, const double* meas, const double* sinf, int robust_kind,
                                double robust_b, double* r, double* J1, double* J2) {
  const double eps = 0.0001;
  double a0[7], b0[7];
  pose_euler_roundtrip(p1, a0);
  if (p2) pose_euler_roundtrip(p2, b0);
  for (int k = 0; k < (p2 ? 2 : 1); k++) {
    double* J = k ? J2 : J1;
    for (int j = 0; j < 6; j++) {
      double d[6] = {0, 0, 0, 0, 0, 0}, pp[7], yp[6], ym[6];
      d[j] = eps;
      pose_exmap(k ? b0 : a0, d, pp);
      pose_factor_linearize(k ? a0 : pp, p2 ? (k ? pp : b0) : nullptr, meas, sinf, robust_kind, robust_b, yp, nullptr, nullptr);
      d[j] = -eps;
      pose_exmap(k ? b0 : a0, d, pp);
      pose_factor_linearize(k ? a0 : pp, p2 ? (k ? pp : b0) : nullptr, meas, sinf, robust_kind, robust_b, ym, nullptr, nullptr);
      for (int i = 0; i < 6; i++) J[i * 6 + j] = (yp[i] - ym[i]) / (eps + eps);
    }
  }
  pose_factor_linearize(a0, p2 ? b0 : nullptr, meas, sinf, robust_kind, robust_b, r, nullptr, nullptr);
}

// inverse of a symmetric positive definite 3x3 given as full row-major 9; out full 9
PUS_HD void sym3_inverse(const double* A, double* Ai) {
  double a = A[0], b = A[1], c = A[2], d = A[4], e = A[5], f = A[8];
  double C00 = d * f - e * e, C01 = c * e - b * f, C02 = b * e - c * d;
  double det = a * C00 + b * C01 + c * C02;
  double id = 1.0 / det;
  double C11 = a * f - c * c, C12 = b * c - a * e, C22 = a * d - b * b;
  Ai[0] = C00 * id; Ai[1] = C01 * id; Ai[2] = C02 * id;
  Ai[3] = C01 * id; Ai[4] = C11 * id; Ai[5] = C12 * id;
  Ai[6] = C02 * id; Ai[7] = C12 * id; Ai[8] = C22 * id;
}

}  // namespace pus
