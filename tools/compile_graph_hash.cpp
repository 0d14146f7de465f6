// Host-only regression harness for compile_graph(): builds a random graph (loop closures, repeated observations,
// removed factors), compiles it twice into the same object and prints an FNV hash of every array of the device image.
// Use it to prove that a host-side refactor leaves the image bit-identical:
//   g++ -O2 -std=c++17 -DHDR='"<old checkout>/pop_up_slam_b200/csrc/pus_graph.hpp"' -x c++ tools/compile_graph_hash.cpp -o /tmp/h_old
//   g++ -O2 -std=c++17 -DHDR='"pop_up_slam_b200/csrc/pus_graph.hpp"'               -x c++ tools/compile_graph_hash.cpp -o /tmp/h_new
//   for a in "40 6 1" "300 8 2" "2000 12 3" "20000 20 5"; do /tmp/h_old $a; /tmp/h_new $a; done      (args: poses, edges/pose, seed [, force_levels])
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <random>
#include HDR
using namespace pus;
static uint64_t H = 1469598103934665603ull;
template <typename T> void hv(const std::vector<T>& v) { const unsigned char* p = (const unsigned char*)v.data(); for (size_t i = 0; i < v.size() * sizeof(T); i++) { H ^= p[i]; H *= 1099511628211ull; } H ^= v.size(); H *= 1099511628211ull; }
void hi(int x) { H ^= (uint64_t)(uint32_t)x; H *= 1099511628211ull; }
static uint64_t image_hash(const Compiled& c) {
  H = 1469598103934665603ull;
  for (int x : {c.N, c.M, c.Epl, c.Epf, c.Elp, c.SP, c.nc_pad, c.ntile, c.nslot, c.nblk, c.nc, c.n_upart, c.n_ypart, c.nce, c.ngrp, c.n_hv, c.n_heavy, c.n_huge, c.n_f2, c.ntile_pl}) hi(x);
  hv(c.pose_node); hv(c.plane_node); hv(c.node_idx); hv(c.pose_val); hv(c.plane_val);
  hv(c.pp_fid); hv(c.pp_pose); hv(c.pp_plane); hv(c.pp_ptr); hv(c.pm2pl); hv(c.pm_part); hv(c.ypart_ptr); hv(c.tile_ptr); hv(c.blk_part_ptr); hv(c.grp_of_slot);
  hv(c.pp_meas); hv(c.pp_sinf); hv(c.pp_rays); hv(c.pp_kind);
  hv(c.pl2pm); hv(c.pl_ptr); hv(c.pl_plane); hv(c.pl_pose); hv(c.pl_part); hv(c.upart_ptr); hv(c.pp_end); hv(c.heavy); hv(c.huge);
  hv(c.pf_fid); hv(c.pf_i); hv(c.pf_j); hv(c.pinc_ptr); hv(c.pinc); hv(c.pnbr); hv(c.pf_meas); hv(c.pf_sinf);
  hv(c.lp_fid); hv(c.lp_plane); hv(c.linc_ptr); hv(c.linc); hv(c.lp_meas); hv(c.lp_sinf);
  hv(c.blk_grp_ptr); hv(c.grp_plane); hv(c.grp_mem_ptr); hv(c.grp_mem); hv(c.blk_simple); hv(c.grp_info);
  hv(c.ce_ptr); hv(c.ce_node); hv(c.ce_plane); hv(c.ce_lo); hv(c.ce_hi); hv(c.n2ce_ptr); hv(c.n2ce);
  hv(c.hv_plane); hv(c.lp_ptr); hv(c.lp_cea); hv(c.lp_ceb); hv(c.fp_ptr); hv(c.fp_f);
  // (members added in round 2: resident-loop records, assembly tasks, the level-2 pairs of the three-level preconditioner)
  for (int x : {c.res_nt, c.res_ng, c.res_np, c.n_atask, c.n_asplit, c.levels, c.nc2, c.nce2, c.ng2}) hi(x);
  hv(c.grp_info2); hv(c.at_plane); hv(c.at_lo); hv(c.at_hi); hv(c.at_ptr); hv(c.as_plane);
  hv(c.ce2_ptr); hv(c.ce2_node); hv(c.ce2_plane); hv(c.ce2_lo); hv(c.ce2_hi); hv(c.g2_ptr); hv(c.g2_ce);
  return H;
}
int main(int argc, char** argv) {
  int N = atoi(argv[1]), per = atoi(argv[2]), seed = atoi(argv[3]);
  int M = std::max(3, N / 10);
  Graph g; std::mt19937 rng(seed);
  auto U = [&](double a, double b) { return std::uniform_real_distribution<double>(a, b)(rng); };
  double s21[21], s6[6] = {1, 0.1, 0, 2, 0, 1};
  int q = 0; for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) s21[q++] = r == c ? 1.0 + r : 0.0;
  std::vector<int> poses, planes;
  for (int i = 0; i < N; i++) {
    double p7[7] = {U(-5, 5), U(-5, 5), U(0.5, 2), 1, 0, 0, 0};
    poses.push_back(g.add_node(NODE_POSE, i == 0 ? p7 : nullptr));
    double m6[6] = {U(-1, 1), U(-1, 1), 0, U(-.2, .2), 0, 0};
    if (i == 0) g.add_pose_prior(poses[0], m6, s21); else g.add_odometry(poses[i - 1], poses[i], m6, s21);
    if (i > 20 && (rng() % 7) == 0) g.add_odometry(poses[rng() % (i - 10)], poses[i], m6, s21);   // loop closures
    int k = 1 + rng() % per;
    for (int j = 0; j < k; j++) {
      int l = (rng() % 3 == 0) ? 0 : (int)(rng() % M);
      while ((int)planes.size() <= l) { planes.push_back(g.add_node(NODE_PLANE, nullptr)); }
      double m4[4] = {U(-1, 1), U(-1, 1), U(-1, 1), U(-3, 3)};
      int f = g.add_pose_plane(poses[i], planes[l], m4, s6);
      if (rng() % 50 == 0) g.add_pose_plane(poses[i], planes[l], m4, s6);   // duplicate observation
      if (f > 0 && rng() % 97 == 0) g.remove_factor(f);
    }
  }
  if (!planes.empty()) { double m4[4] = {0, 0, 1, 0}; g.add_plane_prior(planes[0], m4, s6); }
  if (argc > 4) g.force_levels = atoi(argv[4]);
  Compiled c; std::string err;
  auto t0 = std::chrono::steady_clock::now();
  bool ok = compile_graph(g, c, err);
  double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  const uint64_t h1 = ok ? image_hash(c) : 0;
  auto t1 = std::chrono::steady_clock::now();
  ok = compile_graph(g, c, err) && ok;   // a second compile into the same object must give the same image
  const double ms2 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
  const uint64_t h2 = image_hash(c);
  ok = ok && (h1 == h2);   // the re-compile into the same object gives the same image
  H = h2;
  printf("N=%d per=%d seed=%d ok=%d E=%d hash=%016llx  (%.2f ms, again %.2f ms) %s\n", N, per, seed, ok, c.Epl, (unsigned long long)H, ms, ms2, err.c_str());
}
