// pus_popup.cu -- per-frame ground-polyline -> wall-plane "pop-up" fit on the GPU (float32), batched
// over frames: one thread per (frame, plane) output row.  Reproduces
//   popup_plane::get_plane_equation              pop_up_wall/libs/popup_plane.cpp:551-652   (mode 0)
//   popup_plane::update_plane_equation_from_seg  pop_up_wall/libs/popup_plane.cpp:654-705   (mode 0)
//   popup_plane::update_plane_equation_from_seg_fast  popup_plane.cpp:708-749               (mode 1)
//   ray_plane_interact / point_dist_lineseg      pop_up_wall/libs/matrix_utils.cpp:189-193, 290-303
// Compiled with --fmad=false so the float32 arithmetic rounds like the reference's scalar code.
#include <cuda_runtime.h>

#include <string>

#include "../../include/popup_gpu.h"

namespace pus {
extern thread_local std::string g_err;

__device__ __forceinline__ float seg_dist(const float* b, const float* e, const float* q) {
  float dx = e[0] - b[0], dy = e[1] - b[1];
  float length = sqrtf(dx * dx + dy * dy);
  float qx = q[0] - b[0], qy = q[1] - b[1];
  if (length < 0.001f) return sqrtf(qx * qx + qy * qy);
  float t = (qx * dx + qy * dy) / length / length;
  if (t < 0.0f) return sqrtf(qx * qx + qy * qy);
  else if (t > 1.0f) { float ex = q[0] - e[0], ey = q[1] - e[1]; return sqrtf(ex * ex + ey * ey); }
  float px = b[0] + t * dx, py = b[1] + t * dy;
  float rx = q[0] - px, ry = q[1] - py;
  return sqrtf(rx * rx + ry * ry);
}

__global__ void popup_kernel(int n_frames, const int* __restrict__ seg_ptr, const int* __restrict__ row_frame,
                             int n_rows, const float* __restrict__ segs, const float* __restrict__ invK,
                             const float* __restrict__ Ts, float dist_thre, int mode, float* planes_world,
                             float* planes_sensor, float* dist, int* good) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const int f = row_frame[row];
  const int s0 = seg_ptr[f];
  const int n = seg_ptr[f + 1] - s0;
  const int j = row - (s0 + f);  // 0 = ground, 1.. = segments
  if (n <= 0) return;
  float T[16];
#pragma unroll
  for (int i = 0; i < 16; i++) T[i] = Ts[(size_t)f * 16 + i];
  const float gw[4] = {0.f, 0.f, -1.f, 0.f};
  float gs[4];
#pragma unroll
  for (int i = 0; i < 4; i++) gs[i] = T[0 * 4 + i] * gw[0] + T[1 * 4 + i] * gw[1] + T[2 * 4 + i] * gw[2] + T[3 * 4 + i] * gw[3];
  if (j == 0) {
    if (planes_sensor) for (int i = 0; i < 4; i++) planes_sensor[(size_t)row * 4 + i] = gs[i];
    if (mode == 0) {
      if (planes_world) for (int i = 0; i < 4; i++) planes_world[(size_t)row * 4 + i] = gw[i];
      if (dist) dist[row] = T[2 * 4 + 3];
      if (good) good[row] = 1;
    }
    return;
  }
  const int s = s0 + j - 1;
  float K[9];
#pragma unroll
  for (int i = 0; i < 9; i++) K[i] = invK[i];
  float Ps[2][3], Pw[2][3];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    float x = segs[(size_t)s * 4 + 2 * k], y = segs[(size_t)s * 4 + 2 * k + 1];
    float ray[3];
    for (int i = 0; i < 3; i++) ray[i] = K[i * 3 + 0] * x + K[i * 3 + 1] * y + K[i * 3 + 2] * 1.0f;
    float den = gs[0] * ray[0] + gs[1] * ray[1] + gs[2] * ray[2];
    float frac = -gs[3] / den;
    for (int i = 0; i < 3; i++) Ps[k][i] = frac * ray[i];
    if (mode == 0) {
      float h[4];
      for (int i = 0; i < 4; i++) h[i] = T[i * 4 + 0] * Ps[k][0] + T[i * 4 + 1] * Ps[k][1] + T[i * 4 + 2] * Ps[k][2] + T[i * 4 + 3] * 1.0f;
      for (int i = 0; i < 3; i++) Pw[k][i] = h[i] / h[3];
    }
  }
  if (mode == 0) {
    Pw[0][2] = 0.f; Pw[1][2] = 0.f;
    float t1[3] = {Pw[1][0] - Pw[0][0], Pw[1][1] - Pw[0][1], Pw[1][2] - Pw[0][2]};
    float t2[3] = {gw[0], gw[1], gw[2]};
    float nrm[3] = {t1[1] * t2[2] - t1[2] * t2[1], t1[2] * t2[0] - t1[0] * t2[2], t1[0] * t2[1] - t1[1] * t2[0]};
    float d = -(nrm[0] * Pw[0][0] + nrm[1] * Pw[0][1] + nrm[2] * Pw[0][2]);
    float pw[4] = {nrm[0], nrm[1], nrm[2], d};
    if (planes_world) for (int i = 0; i < 4; i++) planes_world[(size_t)row * 4 + i] = pw[i];
    if (planes_sensor)
      for (int i = 0; i < 4; i++)
        planes_sensor[(size_t)row * 4 + i] = T[0 * 4 + i] * pw[0] + T[1 * 4 + i] * pw[1] + T[2 * 4 + i] * pw[2] + T[3 * 4 + i] * pw[3];
    float cam[2] = {T[0 * 4 + 3], T[1 * 4 + 3]};
    float dd = seg_dist(Pw[0], Pw[1], cam);
    if (dist) dist[row] = dd;
    if (good) good[row] = ((Ps[0][2] > 0) && (Ps[1][2] > 0) && (dd < dist_thre)) ? 1 : 0;
  } else {
    float t1[3] = {Ps[1][0] - Ps[0][0], Ps[1][1] - Ps[0][1], Ps[1][2] - Ps[0][2]};
    float t2[3] = {gs[0], gs[1], gs[2]};
    float nrm[3] = {t1[1] * t2[2] - t1[2] * t2[1], t1[2] * t2[0] - t1[0] * t2[2], t1[0] * t2[1] - t1[1] * t2[0]};
    float d = -(nrm[0] * Ps[0][0] + nrm[1] * Ps[0][1] + nrm[2] * Ps[0][2]);
    if (planes_sensor) {
      planes_sensor[(size_t)row * 4 + 0] = nrm[0]; planes_sensor[(size_t)row * 4 + 1] = nrm[1];
      planes_sensor[(size_t)row * 4 + 2] = nrm[2]; planes_sensor[(size_t)row * 4 + 3] = d;
    }
  }
}
// ---- device-resident refresh (pus_refresh_plane_measurements) -------------------------------------------
// Pose3d::wTo (Pose3d.h:188-194) of the frame's pose estimate, cast to float as Mapping.cpp:598-599 does
__global__ void frame_pose_kernel(int n_frames, const int* __restrict__ frame_pose, const double* pose7, float* Ts) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  const double* p = pose7 + (size_t)frame_pose[f] * 7;
  const double qw = p[3], qx = p[4], qy = p[5], qz = p[6];
  // Eigen toRotationMatrix, no normalisation (Rot3d.h:96-98)
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  float* T = Ts + (size_t)f * 16;
  T[0] = (float)(1 - (tyy + tzz)); T[1] = (float)(txy - twz);       T[2] = (float)(txz + twy);        T[3] = (float)p[0];
  T[4] = (float)(txy + twz);       T[5] = (float)(1 - (txx + tzz)); T[6] = (float)(tyz - twx);        T[7] = (float)p[1];
  T[8] = (float)(txz - twy);       T[9] = (float)(tyz + twx);       T[10] = (float)(1 - (txx + tyy)); T[11] = (float)p[2];
  T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
}

// Plane3d(row.cast<double>()) (normalised 4-vector) -> the factor's measurement slot and the host copy
__global__ void store_meas_kernel(int n_map, const int* __restrict__ map_row, const int* __restrict__ map_slot,
                                  const float* __restrict__ planes_sensor, double* pp_meas, double* out) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_map) return;
  const float* r = planes_sensor + (size_t)map_row[m] * 4;
  double v[4] = {(double)r[0], (double)r[1], (double)r[2], (double)r[3]};
  const double z = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  if (z > 0) {
    const double n = sqrt(z);
    v[0] /= n; v[1] /= n; v[2] /= n; v[3] /= n;
  }
  double* dst = pp_meas + (size_t)map_slot[m] * 4;
  for (int i = 0; i < 4; i++) { dst[i] = v[i]; out[(size_t)m * 4 + i] = v[i]; }
}

// Plane3d::project_to_plane (isam_plane3d.h:172-177): normal() = abc/|abc|, d() = -d/|abc|
__global__ void project_kernel(int n, const int* __restrict__ plane_idx, const double* plane4, const float* __restrict__ in, float* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* pl = plane4 + (size_t)plane_idx[i] * 4;
  const double nn = sqrt(pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2]);
  const double nx = pl[0] / nn, ny = pl[1] / nn, nz = pl[2] / nn;
  const double dd = -pl[3] / nn;
  const double px = in[(size_t)i * 3], py = in[(size_t)i * 3 + 1], pz = in[(size_t)i * 3 + 2];
  const double t = (nx * px + ny * py + nz * pz) - dd;
  out[(size_t)i * 3] = (float)(px - nx * t);
  out[(size_t)i * 3 + 1] = (float)(py - ny * t);
  out[(size_t)i * 3 + 2] = (float)(pz - nz * t);
}

// launchers used by pus_engine.cu (all pointers are device pointers)
cudaError_t launch_refresh(cudaStream_t st, int n_frames, const int* d_frame_pose, const double* d_pose7, const int* d_seg_ptr,
                           const int* d_row_frame, int n_rows, const float* d_segs, const float* d_invK, float* d_Ts,
                           float* d_planes_sensor, int n_map, const int* d_map_row, const int* d_map_slot, double* d_pp_meas,
                           double* d_out) {
  frame_pose_kernel<<<(n_frames + 127) / 128, 128, 0, st>>>(n_frames, d_frame_pose, d_pose7, d_Ts);
  popup_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(n_frames, d_seg_ptr, d_row_frame, n_rows, d_segs, d_invK, d_Ts, 0.f, 0, nullptr,
                                                      d_planes_sensor, nullptr, nullptr);
  if (n_map > 0) store_meas_kernel<<<(n_map + 255) / 256, 256, 0, st>>>(n_map, d_map_row, d_map_slot, d_planes_sensor, d_pp_meas, d_out);
  return cudaGetLastError();
}
cudaError_t launch_project(cudaStream_t st, int n, const int* d_plane_idx, const double* d_plane4, const float* d_in, float* d_out) {
  project_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, d_plane_idx, d_plane4, d_in, d_out);
  return cudaGetLastError();
}
}  // namespace pus

using namespace pus;

#define PCHK(call)                                                                  \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e__); rc = -1; goto done; } \
  } while (0)

extern "C" int pus_popup_fit_frames(int device, int n_frames, const int* seg_ptr, const float* segs, const float* invK,
                                    const float* Ts, float dist_thre, int mode, float* planes_world, float* planes_sensor,
                                    float* dist, int* good) {
  int rc = 0;
  if (n_frames <= 0) return 0;
  const int n_seg = seg_ptr[n_frames];
  const int n_rows = n_seg + n_frames;
  int *d_ptr = nullptr, *d_rf = nullptr, *d_good = nullptr;
  float *d_segs = nullptr, *d_K = nullptr, *d_T = nullptr, *d_pw = nullptr, *d_ps = nullptr, *d_dist = nullptr;
  int* rf = new int[n_rows];
  for (int f = 0; f < n_frames; f++)
    for (int r = seg_ptr[f] + f; r < seg_ptr[f + 1] + f + 1; r++) rf[r] = f;
  {
    int ndev = 0;
    cudaError_t de = cudaGetDeviceCount(&ndev);
    if (de != cudaSuccess || ndev <= 0 || device >= ndev) {
      g_err = "pus_popup_fit_frames: no usable CUDA device; libpopup_gpu has no CPU fallback";
      delete[] rf;
      return -1;
    }
  }
  PCHK(cudaSetDevice(device));
  PCHK(cudaMalloc(&d_ptr, (n_frames + 1) * sizeof(int)));
  PCHK(cudaMalloc(&d_rf, n_rows * sizeof(int)));
  PCHK(cudaMalloc(&d_segs, (size_t)(n_seg > 0 ? n_seg : 1) * 4 * sizeof(float)));
  PCHK(cudaMalloc(&d_K, 9 * sizeof(float)));
  PCHK(cudaMalloc(&d_T, (size_t)n_frames * 16 * sizeof(float)));
  PCHK(cudaMalloc(&d_pw, (size_t)n_rows * 4 * sizeof(float)));
  PCHK(cudaMalloc(&d_ps, (size_t)n_rows * 4 * sizeof(float)));
  PCHK(cudaMalloc(&d_dist, (size_t)n_rows * sizeof(float)));
  PCHK(cudaMalloc(&d_good, (size_t)n_rows * sizeof(int)));
  PCHK(cudaMemcpy(d_ptr, seg_ptr, (n_frames + 1) * sizeof(int), cudaMemcpyHostToDevice));
  PCHK(cudaMemcpy(d_rf, rf, n_rows * sizeof(int), cudaMemcpyHostToDevice));
  if (n_seg > 0) PCHK(cudaMemcpy(d_segs, segs, (size_t)n_seg * 4 * sizeof(float), cudaMemcpyHostToDevice));
  PCHK(cudaMemcpy(d_K, invK, 9 * sizeof(float), cudaMemcpyHostToDevice));
  PCHK(cudaMemcpy(d_T, Ts, (size_t)n_frames * 16 * sizeof(float), cudaMemcpyHostToDevice));
  PCHK(cudaMemset(d_pw, 0, (size_t)n_rows * 4 * sizeof(float)));
  PCHK(cudaMemset(d_ps, 0, (size_t)n_rows * 4 * sizeof(float)));
  PCHK(cudaMemset(d_dist, 0, (size_t)n_rows * sizeof(float)));
  PCHK(cudaMemset(d_good, 0, (size_t)n_rows * sizeof(int)));
  popup_kernel<<<(n_rows + 255) / 256, 256>>>(n_frames, d_ptr, d_rf, n_rows, d_segs, d_K, d_T, dist_thre, mode, d_pw, d_ps, d_dist, d_good);
  PCHK(cudaGetLastError());
  PCHK(cudaDeviceSynchronize());
  if (planes_world) PCHK(cudaMemcpy(planes_world, d_pw, (size_t)n_rows * 4 * sizeof(float), cudaMemcpyDeviceToHost));
  if (planes_sensor) PCHK(cudaMemcpy(planes_sensor, d_ps, (size_t)n_rows * 4 * sizeof(float), cudaMemcpyDeviceToHost));
  if (dist) PCHK(cudaMemcpy(dist, d_dist, (size_t)n_rows * sizeof(float), cudaMemcpyDeviceToHost));
  if (good) PCHK(cudaMemcpy(good, d_good, (size_t)n_rows * sizeof(int), cudaMemcpyDeviceToHost));
done:
  delete[] rf;
  cudaFree(d_ptr); cudaFree(d_rf); cudaFree(d_segs); cudaFree(d_K); cudaFree(d_T); cudaFree(d_pw); cudaFree(d_ps); cudaFree(d_dist); cudaFree(d_good);
  return rc;
}
