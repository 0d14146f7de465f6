// ORACLE (test infrastructure, NOT product code) -- see oracle_math.hpp header.
//
// float32 restatement of the per-frame ground-polyline -> wall-plane "pop-up" fit.
// "parity unpinned" (the reference holds no fixtures for it).
//
// Reference (relative to /root/reference):
//   pop_up_wall/libs/popup_plane.cpp:551-652  popup_plane::get_plane_equation
//   pop_up_wall/libs/popup_plane.cpp:654-705  update_plane_equation_from_seg
//   pop_up_wall/libs/popup_plane.cpp:708-749  update_plane_equation_from_seg_fast
//   pop_up_wall/libs/matrix_utils.cpp:189-193 ray_plane_interact
//   pop_up_wall/libs/matrix_utils.cpp:290-303 point_dist_lineseg
//   pop_planar_slam/src/isam_plane3d.cpp:20-55 get_wall_plane_equation (double copy of _fast)
// Build this file with -ffp-contract=off so no FMA contraction changes the f32 rounding.
#pragma once
#include <cmath>

namespace orc {

// matrix_utils.cpp:290-303
inline float point_dist_lineseg(const float b[2], const float e[2], const float q[2]) {
  float dx = e[0] - b[0], dy = e[1] - b[1];
  float length = std::sqrt(dx * dx + dy * dy);
  float qx = q[0] - b[0], qy = q[1] - b[1];
  if (length < 0.001f) return std::sqrt(qx * qx + qy * qy);
  float t = (qx * dx + qy * dy) / length / length;
  if (t < 0.0f) return std::sqrt(qx * qx + qy * qy);
  else if (t > 1.0f) { float ex = q[0] - e[0], ey = q[1] - e[1]; return std::sqrt(ex * ex + ey * ey); }
  float px = b[0] + t * dx, py = b[1] + t * dy;
  float rx = q[0] - px, ry = q[1] - py;
  return std::sqrt(rx * rx + ry * ry);
}

// segs: n x 4 (x1,y1,x2,y2) px ; invK 3x3 row-major ; T 4x4 row-major (sensor->world)
// outputs (any may be null): planes_world (n+1)x4, planes_sensor (n+1)x4,
// dist (n+1), good (n+1) as 0/1, seg3d_world n x 6, seg3d_sensor n x 6.
// mode 0 = get_plane_equation / update_plane_equation_from_seg (via world frame)
// mode 1 = update_plane_equation_from_seg_fast (sensor frame only; planes_world/dist/good untouched)
inline void popup_fit(const float* segs, int n, const float* invK, const float* T, float dist_thre, int mode,
                      float* planes_world, float* planes_sensor, float* dist, int* good, float* seg3d_world,
                      float* seg3d_sensor) {
  if (n <= 0) return;
  const float gw[4] = {0.f, 0.f, -1.f, 0.f};
  float gs[4];
  for (int i = 0; i < 4; i++) gs[i] = T[0 * 4 + i] * gw[0] + T[1 * 4 + i] * gw[1] + T[2 * 4 + i] * gw[2] + T[3 * 4 + i] * gw[3];
  if (planes_sensor) for (int i = 0; i < 4; i++) planes_sensor[i] = gs[i];
  if (mode == 0) {
    if (planes_world) for (int i = 0; i < 4; i++) planes_world[i] = gw[i];
    if (dist) dist[0] = T[2 * 4 + 3];
    if (good) good[0] = 1;
  }
  for (int s = 0; s < n; s++) {
    float Ps[2][3], Pw[2][3];
    for (int k = 0; k < 2; k++) {
      float x = segs[s * 4 + 2 * k], y = segs[s * 4 + 2 * k + 1];
      float ray[3];
      for (int i = 0; i < 3; i++) ray[i] = invK[i * 3 + 0] * x + invK[i * 3 + 1] * y + invK[i * 3 + 2] * 1.0f;
      float den = gs[0] * ray[0] + gs[1] * ray[1] + gs[2] * ray[2];
      float frac = -gs[3] / den;
      for (int i = 0; i < 3; i++) Ps[k][i] = frac * ray[i];
      if (mode == 0) {
        float h[4];
        for (int i = 0; i < 4; i++) h[i] = T[i * 4 + 0] * Ps[k][0] + T[i * 4 + 1] * Ps[k][1] + T[i * 4 + 2] * Ps[k][2] + T[i * 4 + 3] * 1.0f;
        for (int i = 0; i < 3; i++) Pw[k][i] = h[i] / h[3];
      }
    }
    if (seg3d_sensor) for (int k = 0; k < 2; k++) for (int i = 0; i < 3; i++) seg3d_sensor[s * 6 + 3 * k + i] = Ps[k][i];
    if (mode == 0) {
      Pw[0][2] = 0.f; Pw[1][2] = 0.f;  // popup_plane.cpp:576-579 "make it exact zero"
      if (seg3d_world) for (int k = 0; k < 2; k++) for (int i = 0; i < 3; i++) seg3d_world[s * 6 + 3 * k + i] = Pw[k][i];
      float t1[3] = {Pw[1][0] - Pw[0][0], Pw[1][1] - Pw[0][1], Pw[1][2] - Pw[0][2]};
      float t2[3] = {gw[0], gw[1], gw[2]};
      float nrm[3] = {t1[1] * t2[2] - t1[2] * t2[1], t1[2] * t2[0] - t1[0] * t2[2], t1[0] * t2[1] - t1[1] * t2[0]};
      float d = -(nrm[0] * Pw[0][0] + nrm[1] * Pw[0][1] + nrm[2] * Pw[0][2]);
      float pw[4] = {nrm[0], nrm[1], nrm[2], d};
      if (planes_world) for (int i = 0; i < 4; i++) planes_world[(s + 1) * 4 + i] = pw[i];
      if (planes_sensor)
        for (int i = 0; i < 4; i++)
          planes_sensor[(s + 1) * 4 + i] = T[0 * 4 + i] * pw[0] + T[1 * 4 + i] * pw[1] + T[2 * 4 + i] * pw[2] + T[3 * 4 + i] * pw[3];
      float cam[2] = {T[0 * 4 + 3], T[1 * 4 + 3]};
      float dd = point_dist_lineseg(Pw[0], Pw[1], cam);
      if (dist) dist[s + 1] = dd;
      if (good) good[s + 1] = ((Ps[0][2] > 0) && (Ps[1][2] > 0) && (dd < dist_thre)) ? 1 : 0;  // popup_plane.cpp:619-637
    } else {
      float t1[3] = {Ps[1][0] - Ps[0][0], Ps[1][1] - Ps[0][1], Ps[1][2] - Ps[0][2]};
      float t2[3] = {gs[0], gs[1], gs[2]};
      float nrm[3] = {t1[1] * t2[2] - t1[2] * t2[1], t1[2] * t2[0] - t1[0] * t2[2], t1[0] * t2[1] - t1[1] * t2[0]};
      float d = -(nrm[0] * Ps[0][0] + nrm[1] * Ps[0][1] + nrm[2] * Ps[0][2]);
      if (planes_sensor) { float ps[4] = {nrm[0], nrm[1], nrm[2], d}; for (int i = 0; i < 4; i++) planes_sensor[(s + 1) * 4 + i] = ps[i]; }
    }
  }
}

}  // namespace orc
