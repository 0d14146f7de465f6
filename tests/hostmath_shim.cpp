// Test shim: exposes the __host__ __device__ math of pop_up_slam_b200/csrc/pus_math.cuh to
// ctypes so the CPU test-suite can pin the exact functions the CUDA kernels run.
#include "../pop_up_slam_b200/csrc/pus_math.cuh"
using namespace pus;
extern "C" {
void hm_pose_plane_linearize(const double* pose, const double* plane, const double* meas, const double* sinf, int rk,
                             double rb, double* r, double* Jp, double* Jl) {
  pose_plane_linearize(pose, plane, meas, sinf, rk, rb, r, Jp, Jl);
}
void hm_pose_plane2_linearize(const double* pose, const double* plane, const double* rays, const double* sinf, int rk,
                              double rb, double* r, double* Jp, double* Jl) {
  const double dummy[4] = {1, 0, 0, 0};
  pose_plane_linearize(pose, plane, dummy, sinf, rk, rb, r, Jp, Jl, rays);
}
void hm_plane_prior_linearize(const double* plane, const double* meas, const double* sinf, int rk, double rb, double* r,
                              double* Jl) {
  pose_plane_linearize(nullptr, plane, meas, sinf, rk, rb, r, nullptr, Jl);
}
void hm_pose_plane_residual(const double* pose, const double* plane, const double* meas, const double* sinf, int rk,
                            double rb, double* r) {
  pose_plane_linearize(pose, plane, meas, sinf, rk, rb, r, nullptr, nullptr);
}
void hm_pose_factor_linearize(const double* p1, const double* p2, const double* meas, const double* sinf, int rk, double rb,
                              double* r, double* J1, double* J2) {
  pose_factor_linearize(p1, p2, meas, sinf, rk, rb, r, J1, J2);
}
void hm_pose_factor_residual(const double* p1, const double* p2, const double* meas, const double* sinf, int rk, double rb,
                             double* r) {
  pose_factor_linearize(p1, p2, meas, sinf, rk, rb, r, nullptr, nullptr);
}
void hm_pose_plane_numeric(const double* pose, const double* plane, const double* meas, const double* sinf, int rk,
                           double rb, double* r, double* Jp, double* Jl) {
  pose_plane_numeric(pose, plane, meas, sinf, rk, rb, r, Jp, Jl);
}
void hm_pose_factor_numeric(const double* p1, const double* p2, const double* meas, const double* sinf, int rk, double rb,
                            double* r, double* J1, double* J2) {
  pose_factor_numeric(p1, p2, meas, sinf, rk, rb, r, J1, J2);
}
void hm_pose_exmap(const double* p, const double* d, double* out) { pose_exmap(p, d, out); }
void hm_plane_exmap(const double* p, const double* d, double* out) { plane_exmap(p, d, out); }
void hm_pose_oplus(const double* a, const double* b, double* out) { pose_oplus(a, b, out); }
void hm_pose_ominus(const double* a, const double* b, double* out) { pose_ominus(a, b, out); }
void hm_pose_from_xyzypr(const double* v, double* out) { pose_from_xyzypr(v, out); }
void hm_sym3_inverse(const double* A, double* Ai) { sym3_inverse(A, Ai); }
}
