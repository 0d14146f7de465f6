// pus_graph.hpp -- host side of the drop-in boundary: the factor-graph container that mirrors
// isam::Slam's graph bookkeeping, and the "compiler" that flattens it into the HBM layout the
// CUDA kernels sweep.  Pure C++ (no CUDA), so it is unit-tested on the CPU.
//
// Reference semantics reproduced (paths relative to the reference checkout;
// ISAM = pop_planar_slam/Thirdparty/isam, PPS = pop_planar_slam):
//   insertion-ordered ids / lists ...... ISAM/isamlib/Slam.cpp:47-48,91-126, ISAM/include/isam/Graph.h:40-133
//   column / row offsets ................ Slam::update_starts Slam.cpp:59-67, jacobian_partial Slam.cpp:395-432
//   factor initialisation of nodes ...... slam3d.h:75-80,123-137 ; PPS/src/isam_plane3d.h:252-264,443-448
//   measurement update .................. ISAM/include/isam/Factor.h:203-206
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "pus_math.cuh"

namespace pus {

enum NodeKind { NODE_POSE = 0, NODE_PLANE = 1 };
enum FactorKind { F_POSE_PRIOR = 0, F_ODOMETRY = 1, F_POSE_PLANE = 2, F_PLANE_PRIOR = 3 };

constexpr int kBlockPoses = 16;            // poses per dense preconditioner block
constexpr int kBlockDim = 6 * kBlockPoses; // 96
// Three-level additive preconditioner of the reduced pose system:
//   level 1  exact inverses of the 16-pose diagonal blocks of S
//   level 2  piecewise-linear ("hat") trajectory modes with one node every kL2Spacing poses, solved by block-Jacobi in
//            groups of kGroupNodes nodes (96 x 96 blocks of P2^T S P2, same size as the level-1 blocks)
//   level 3  hat modes with one node every kCoarseSpacing poses, Galerkin operator inverted densely (<= 1920^2)
// Graphs of up to kL2Spacing * kMaxCoarseNodes = 5120 poses use two levels (1 + an exactly inverted hat level with a node
// every 16 poses: fewest PCG iterations, dense A_c^-1 <= 1920^2); larger graphs use all three, so the hat level with the
// fine spacing never has to be inverted (or applied) as a dense matrix and the iteration count stays flat in N.
constexpr int kL2Spacing = 16;             // = kBlockPoses: every pose block is exactly one level-2 interval
constexpr int kGroupNodes = 16;            // level-2 nodes per block-Jacobi group (6 * 16 = kBlockDim)
constexpr int kCoarseSpacing = 128;        // base spacing (poses) of the level-3 nodes; multiple of kBlockPoses
constexpr int kMaxCoarseNodes = 320;       // the level-3 spacing grows in steps of 128 so that the dense A_c stays <= 1920^2
constexpr int kHeavyCoarse = 16;           // planes touching more coarse nodes than this are dense rank-3 updates of A_c
constexpr int kPivotNodes = 8;             // coarse nodes per pivot block of the blocked Gauss-Jordan inversion (48 scalars)
constexpr int kTile = 32;                  // edges per warp tile
constexpr int kWStride = 18 * kTile;       // doubles per W tile
constexpr int kMaxGrp = 256;               // plane groups per 16-pose block held in shared memory
constexpr int kMaxPart = 64;               // (tile, pose) partial sums per block held in shared memory
constexpr int kAsmChunk = 128;             // plane-major edge slots per Hll / gl assembly task (one warp)

struct HNode {
  int kind = NODE_POSE;
  bool alive = true;
  bool initialized = false;
  double v[7] = {0, 0, 0, 1, 0, 0, 0};  // pose: x y z qw qx qy qz ; plane: a b c d
};

struct HFactor {
  int kind = F_POSE_PLANE;
  bool alive = true;
  int nodes[2] = {-1, -1};
  int n_nodes = 1;
  int dim = 3;
  double meas[6] = {0, 0, 0, 0, 0, 0};
  double sinf[21];  // packed upper-triangular
  bool has_rays = false;   // Pose3d_Plane3d_Factor2: the measured plane is re-popped from two ground-edge rays (sensor frame)
  double rays[6] = {0, 0, 0, 0, 0, 0};
};

struct Graph {
  std::vector<HNode> nodes;
  std::vector<HFactor> factors;
  std::string err;
  uint64_t topo_version = 1;  // bumped on any structural edit (node/factor add/remove)
  int force_levels = 0;       // 0: two or three preconditioner levels by graph size; 2 / 3: forced (tests, studies)

  bool ok_node(int id, int kind) const { return id >= 0 && id < (int)nodes.size() && nodes[id].alive && nodes[id].kind == kind; }
  bool ok_factor(int f) const { return f >= 0 && f < (int)factors.size() && factors[f].alive; }

  int add_node(int kind, const double* v) {
    HNode n;
    n.kind = kind;
    if (v) {
      if (kind == NODE_POSE) std::memcpy(n.v, v, 7 * sizeof(double));
      else { std::memcpy(n.v, v, 4 * sizeof(double)); normalize4(n.v); }
      n.initialized = true;
    } else if (kind == NODE_PLANE) { n.v[0] = 1; n.v[1] = n.v[2] = n.v[3] = 0; }
    nodes.push_back(n);
    topo_version++;
    return (int)nodes.size() - 1;
  }
  void init_node(int id, const double* v) {
    HNode& n = nodes[id];
    if (n.kind == NODE_POSE) std::memcpy(n.v, v, 7 * sizeof(double));
    else { std::memcpy(n.v, v, 4 * sizeof(double)); normalize4(n.v); }
    n.initialized = true;
  }

  int push_factor(HFactor& f) { factors.push_back(f); topo_version++; return (int)factors.size() - 1; }

  int add_pose_prior(int pose, const double* m, const double* si) {
    if (!ok_node(pose, NODE_POSE)) { err = "add_pose_prior: bad pose id"; return -1; }
    HFactor f; f.kind = F_POSE_PRIOR; f.n_nodes = 1; f.nodes[0] = pose; f.dim = 6;
    std::memcpy(f.meas, m, 6 * sizeof(double)); std::memcpy(f.sinf, si, 21 * sizeof(double));
    if (!nodes[pose].initialized) { double p[7]; pose_from_xyzypr(m, p); init_node(pose, p); }  // slam3d.h:75-80
    return push_factor(f);
  }
  int add_odometry(int a, int b, const double* m, const double* si) {
    if (!ok_node(a, NODE_POSE) || !ok_node(b, NODE_POSE)) { err = "add_odometry: bad pose id"; return -1; }
    if (!nodes[a].initialized && !nodes[b].initialized) {
      err = "slam3d: Pose3d_Pose3d_Factor requires pose1 or pose2 to be initialized";  // slam3d.h:124-125
      return -1;
    }
    HFactor f; f.kind = F_ODOMETRY; f.n_nodes = 2; f.nodes[0] = a; f.nodes[1] = b; f.dim = 6;
    std::memcpy(f.meas, m, 6 * sizeof(double)); std::memcpy(f.sinf, si, 21 * sizeof(double));
    double mp[7]; pose_from_xyzypr(m, mp);
    if (!nodes[a].initialized) {            // slam3d.h:127-132: p2.oplus(z.ominus(measure))
      double z[7] = {0, 0, 0, 1, 0, 0, 0}, inv[7], out[7];
      pose_ominus(z, mp, inv); pose_oplus(nodes[b].v, inv, out); init_node(a, out);
    } else if (!nodes[b].initialized) {     // slam3d.h:133-137: p1.oplus(measure)
      double out[7]; pose_oplus(nodes[a].v, mp, out); init_node(b, out);
    }
    return push_factor(f);
  }
  int add_pose_plane(int pose, int plane, const double* m, const double* si) {
    if (!ok_node(pose, NODE_POSE) || !ok_node(plane, NODE_PLANE)) { err = "add_pose_plane: bad node id"; return -1; }
    if (!nodes[pose].initialized) { err = "Plane3d: Pose3d_Plane3d_Factor requires pose to be initialized"; return -1; }
    HFactor f; f.kind = F_POSE_PLANE; f.n_nodes = 2; f.nodes[0] = pose; f.nodes[1] = plane; f.dim = 3;
    std::memcpy(f.meas, m, 4 * sizeof(double)); normalize4(f.meas);
    std::memcpy(f.sinf, si, 6 * sizeof(double));
    if (!nodes[plane].initialized) {        // isam_plane3d.h:256-262: measure.transform_from(p.oTw())
      double T[16], g[4]; pose_to_Tinv(nodes[pose].v, T); plane_transform_T(T, f.meas, g); init_node(plane, g);
    }
    return push_factor(f);
  }
  // Pose3d_Plane3d_Factor2 (isam_plane3d.h:314-424): same nodes / initialisation, plus the precomputed rays (:358-370)
  int add_pose_plane2(int pose, int plane, const double* m, const double* rays6, const double* si) {
    const int f = add_pose_plane(pose, plane, m, si);
    if (f < 0) return f;
    factors[f].has_rays = true;
    std::memcpy(factors[f].rays, rays6, 6 * sizeof(double));
    return f;
  }
  int add_plane_prior(int plane, const double* m, const double* si) {
    if (!ok_node(plane, NODE_PLANE)) { err = "add_plane_prior: bad plane id"; return -1; }
    HFactor f; f.kind = F_PLANE_PRIOR; f.n_nodes = 1; f.nodes[0] = plane; f.dim = 3;
    std::memcpy(f.meas, m, 4 * sizeof(double)); normalize4(f.meas);
    std::memcpy(f.sinf, si, 6 * sizeof(double));
    if (!nodes[plane].initialized) init_node(plane, f.meas);  // isam_plane3d.h:443-448
    return push_factor(f);
  }
  void remove_factor(int f) { factors[f].alive = false; topo_version++; }
  void remove_node(int id) {  // Slam::remove_node Slam.cpp:107-115: adjacent factors go too
    for (auto& f : factors) {
      if (!f.alive) continue;
      for (int k = 0; k < f.n_nodes; k++) if (f.nodes[k] == id) { f.alive = false; break; }
    }
    nodes[id].alive = false;
    topo_version++;
  }
  int num_nodes() const { int c = 0; for (auto& n : nodes) c += n.alive; return c; }
  int num_factors() const { int c = 0; for (auto& f : factors) c += f.alive; return c; }
  int node_start(int id) const {
    if (id < 0 || id >= (int)nodes.size() || !nodes[id].alive) return -1;
    int s = 0;
    for (int i = 0; i < id; i++) if (nodes[i].alive) s += nodes[i].kind == NODE_POSE ? 6 : 3;
    return s;
  }
  int factor_row(int fid) const {
    if (!ok_factor(fid)) return -1;
    int r = 0;
    for (int i = 0; i < fid; i++) if (factors[i].alive) r += factors[i].dim;
    return r;
  }
};

// ------------------------------------------------------------------------------------------------
// Flattened problem (host copy of what goes to HBM).  All index arrays are int32.
// ------------------------------------------------------------------------------------------------
struct Compiled {
  int N = 0, M = 0, Epl = 0, Epf = 0, Elp = 0;
  int SP = kCoarseSpacing;
  int nc_pad = 0;  // nc rounded up to whole pivot blocks; the padding nodes carry identity blocks
  int ntile = 0, nslot = 0, nblk = 0, nc = 0, n_upart = 0, n_ypart = 0, nce = 0, ngrp = 0, n_hv = 0, n_heavy = 0, n_huge = 0;
  std::vector<int> pose_node, plane_node;   // idx -> node id
  std::vector<int> node_idx;                // node id -> idx (pose idx or plane idx), -1 dead
  std::vector<double> pose_val, plane_val;  // [N*7], [M*4]
  // pose-plane edges, pose-major
  // (slot-indexed: every 16-pose block's edges are padded to whole 32-edge tiles; pad slots have pp_pose = -1)
  std::vector<int> pp_fid, pp_pose, pp_plane, pp_ptr, pm2pl, pm_part, ypart_ptr, tile_ptr, blk_part_ptr, grp_of_slot;
  std::vector<double> pp_meas, pp_sinf, pp_rays;   // pp_rays: [slots][6] when any edge is a Factor2 (else one dummy entry)
  std::vector<int> pp_kind;                        // 1 = Factor2 (uses pp_rays), else 0
  int n_f2 = 0;
  // plane-major view
  std::vector<int> pl2pm, pl_ptr, pl_plane, pl_pose, pl_part, upart_ptr, pp_end, heavy, huge;
  int ntile_pl = 0;
  // pose factors (prior / odometry)
  std::vector<int> pf_fid, pf_i, pf_j, pinc_ptr, pinc, pnbr;
  std::vector<double> pf_meas, pf_sinf;
  // plane priors
  std::vector<int> lp_fid, lp_plane, linc_ptr, linc;
  std::vector<double> lp_meas, lp_sinf;
  // dense-block groups: per pose block, its edges grouped by plane
  std::vector<int> blk_grp_ptr, grp_plane, grp_mem_ptr, grp_mem, blk_simple, grp_info;
  // block-resident PCG (small / medium graphs): the plane exchange is per (pose block, plane) instead of per plane-major tile
  // run -- grp_info2 = {plane, first slot of the plane's per-block partial sums, their number, this group's own slot};
  // res_* = largest tile / group / (tile, pose)-run count of any pose block (uniform shared-memory strides)
  std::vector<int> grp_info2;
  int res_nt = 0, res_ng = 0, res_np = 0;
  // Hll / gl assembly tasks: the plane-major slots of every plane cut into chunks of kAsmChunk (a plane seen from thousands of
  // poses is gathered by many warps instead of one); at_* per task, at_ptr per plane, as_plane = planes with more than one task
  std::vector<int> at_plane, at_lo, at_hi, at_ptr, as_plane;
  int n_atask = 0, n_asplit = 0;
  // coarse (hat) space: (plane, coarse node) pairs
  std::vector<int> ce_ptr, ce_node, ce_plane, ce_lo, ce_hi, n2ce_ptr, n2ce;
  std::vector<int> hv_plane, lp_ptr, lp_cea, lp_ceb, fp_ptr, fp_f;   // coarse assembly: heavy planes, per node-pair lists
  // level 2: (plane, node) pairs at spacing kL2Spacing, and per block-Jacobi group its pairs sorted by (plane, node)
  int levels = 2;   // preconditioner levels in use (2: SP = 16 k hats inverted exactly; 3: level-2 block-Jacobi + level 3 at SP = 256 k)
  int nc2 = 0, nce2 = 0, ng2 = 0;
  std::vector<int> ce2_ptr, ce2_node, ce2_plane, ce2_lo, ce2_hi, g2_ptr, g2_ce;
};

inline int coarse_spacing(int N, int levels) {
  const int base = (levels == 3) ? kCoarseSpacing : kL2Spacing;
  return base * std::max(1, (N + base * kMaxCoarseNodes - 1) / (base * kMaxCoarseNodes));
}
inline int coarse_nodes(int N, int sp) { return N <= 1 ? 1 : (N - 1 + sp - 1) / sp + 1; }

inline bool compile_graph(const Graph& g, Compiled& c, std::string& err) {
  c = Compiled();
  const int nn = (int)g.nodes.size();
  c.node_idx.assign(nn, -1);
  for (int i = 0; i < nn; i++) {
    const HNode& n = g.nodes[i];
    if (!n.alive) continue;
    if (!n.initialized) { err = "node " + std::to_string(i) + " is not initialised"; return false; }
    if (n.kind == NODE_POSE) { c.node_idx[i] = (int)c.pose_node.size(); c.pose_node.push_back(i); }
    else { c.node_idx[i] = (int)c.plane_node.size(); c.plane_node.push_back(i); }
  }
  c.N = (int)c.pose_node.size(); c.M = (int)c.plane_node.size();
  const int N = c.N, M = c.M;
  c.pose_val.resize((size_t)N * 7); c.plane_val.resize((size_t)M * 4);
  for (int p = 0; p < N; p++) std::memcpy(&c.pose_val[(size_t)p * 7], g.nodes[c.pose_node[p]].v, 7 * sizeof(double));
  for (int l = 0; l < M; l++) std::memcpy(&c.plane_val[(size_t)l * 4], g.nodes[c.plane_node[l]].v, 4 * sizeof(double));

  // ---- split factors ----
  std::vector<int> ppf;  // pose-plane factor ids in insertion order
  {
    int npp = 0, npf = 0, nlp = 0;
    for (const HFactor& F : g.factors) {
      if (!F.alive) continue;
      if (F.kind == F_POSE_PLANE) npp++; else if (F.kind == F_PLANE_PRIOR) nlp++; else npf++;
    }
    ppf.reserve(npp);
    c.pf_fid.reserve(npf); c.pf_i.reserve(npf); c.pf_j.reserve(npf); c.pf_meas.reserve((size_t)npf * 6); c.pf_sinf.reserve((size_t)npf * 21);
    c.lp_fid.reserve(nlp); c.lp_plane.reserve(nlp); c.lp_meas.reserve((size_t)nlp * 4); c.lp_sinf.reserve((size_t)nlp * 6);
  }
  for (int f = 0; f < (int)g.factors.size(); f++) {
    const HFactor& F = g.factors[f];
    if (!F.alive) continue;
    switch (F.kind) {
      case F_POSE_PLANE: ppf.push_back(f); break;
      case F_POSE_PRIOR:
      case F_ODOMETRY:
        c.pf_fid.push_back(f);
        c.pf_i.push_back(c.node_idx[F.nodes[0]]);
        c.pf_j.push_back(F.kind == F_ODOMETRY ? c.node_idx[F.nodes[1]] : -1);
        c.pf_meas.insert(c.pf_meas.end(), F.meas, F.meas + 6);
        c.pf_sinf.insert(c.pf_sinf.end(), F.sinf, F.sinf + 21);
        break;
      case F_PLANE_PRIOR:
        c.lp_fid.push_back(f);
        c.lp_plane.push_back(c.node_idx[F.nodes[0]]);
        c.lp_meas.insert(c.lp_meas.end(), F.meas, F.meas + 4);
        c.lp_sinf.insert(c.lp_sinf.end(), F.sinf, F.sinf + 6);
        break;
    }
  }
  c.Epf = (int)c.pf_fid.size(); c.Elp = (int)c.lp_fid.size();
  // ---- pose-major ordering of the pose-plane edges (stable: insertion order within a pose), padded so that
  //      every 16-pose block owns whole 32-edge tiles ----
  c.Epl = (int)ppf.size();
  const int E = c.Epl;
  c.nblk = (N + kBlockPoses - 1) / kBlockPoses;
  // stable counting sort by pose index (keys gathered once: the comparison sort chased the factor records)
  std::vector<int> order(E), key(E);
  {
    std::vector<int> start(N + 1, 0);
    for (int i = 0; i < E; i++) { key[i] = c.node_idx[g.factors[ppf[i]].nodes[0]]; start[key[i] + 1]++; }
    for (int p = 0; p < N; p++) start[p + 1] += start[p];
    for (int i = 0; i < E; i++) order[start[key[i]]++] = i;
  }
  c.pp_ptr.assign(N + 1, 0);
  c.tile_ptr.assign(c.nblk + 1, 0);
  {
    // count edges per pose, then lay the slots out block by block
    std::vector<int> cnt(N, 0);
    for (int i = 0; i < E; i++) cnt[key[i]]++;   // (the keys gathered for the counting sort)
    int slot = 0;
    for (int k = 0; k < c.nblk; k++) {
      c.tile_ptr[k] = slot / kTile;
      int p0 = k * kBlockPoses, p1 = std::min(N, p0 + kBlockPoses);
      for (int p = p0; p < p1; p++) { c.pp_ptr[p] = slot; slot += cnt[p]; }
      slot = (slot + kTile - 1) / kTile * kTile;
    }
    c.pp_ptr[N] = slot;  // only used as the end of the last pose when it has no padding after it
    c.tile_ptr[c.nblk] = slot / kTile;
    c.nslot = slot;
    // end of pose p's range = start + cnt (pp_ptr[p+1] may include padding): keep an explicit end array in pp_ptr
    // by storing starts in pp_ptr[0..N) and ends in pose_end
    c.ntile = slot / kTile;
    const int slots = c.nslot;
    c.pp_fid.assign(slots, -1); c.pp_pose.assign(slots, -1); c.pp_plane.assign(slots, 0);
    c.pp_meas.assign((size_t)slots * 4, 0.0); c.pp_sinf.assign((size_t)slots * 6, 0.0);
    for (int s2 = 0; s2 < slots; s2++) c.pp_meas[(size_t)s2 * 4] = 1.0;
    c.pp_rays.assign(1, 0.0); c.pp_kind.assign(1, 0); c.n_f2 = 0;
    std::vector<int> fill(c.pp_ptr.begin(), c.pp_ptr.begin() + N);
    for (int i = 0; i < E; i++) {
      const HFactor& F = g.factors[ppf[order[i]]];
      int p = key[order[i]];
      int e = fill[p]++;
      c.pp_fid[e] = ppf[order[i]];
      c.pp_pose[e] = p;
      c.pp_plane[e] = c.node_idx[F.nodes[1]];
      std::memcpy(&c.pp_meas[(size_t)e * 4], F.meas, 4 * sizeof(double));
      if (F.has_rays) {
        if (c.pp_rays.size() < (size_t)slots * 6) { c.pp_rays.assign((size_t)slots * 6, 0.0); c.pp_kind.assign(slots, 0); }
        std::memcpy(&c.pp_rays[(size_t)e * 6], F.rays, 6 * sizeof(double));
        c.pp_kind[e] = 1; c.n_f2++;
      }
      std::memcpy(&c.pp_sinf[(size_t)e * 6], F.sinf, 6 * sizeof(double));
    }
    c.pp_end.assign(N, 0);
    for (int p = 0; p < N; p++) c.pp_end[p] = c.pp_ptr[p] + cnt[p];
  }
  const int slots = c.nslot;
  // partial slots for the pose-major sweep: one per (tile, pose) run, numbered in slot order (block-local ranges)
  c.pm_part.assign(slots, -1);
  c.ypart_ptr.assign(N + 1, 0);
  c.blk_part_ptr.assign(c.nblk + 1, 0);
  {
    int np = 0, prev = -1;
    for (int e = 0; e < slots; e++) {
      int p = c.pp_pose[e];
      if (p < 0) { prev = -1; continue; }
      bool head = (e % kTile == 0) || (p != prev);
      if (head) { np++; c.ypart_ptr[p + 1]++; }
      c.pm_part[e] = np - 1;
      prev = p;
    }
    c.n_ypart = np;
    for (int p = 0; p < N; p++) c.ypart_ptr[p + 1] += c.ypart_ptr[p];
    for (int k = 0; k <= c.nblk; k++) c.blk_part_ptr[k] = c.ypart_ptr[std::min(N, k * kBlockPoses)];
  }
  // ---- plane-major view (slots of its own: dense, E real entries then padding) ----
  const int pslots = (E + kTile - 1) / kTile * kTile;
  c.ntile_pl = pslots / kTile;
  std::vector<int> pord(E);   // live slots in slot order, stably counting-sorted by plane
  {
    std::vector<int> start(M + 1, 0);
    for (int e = 0; e < slots; e++) if (c.pp_pose[e] >= 0) start[c.pp_plane[e] + 1]++;
    for (int l = 0; l < M; l++) start[l + 1] += start[l];
    for (int e = 0; e < slots; e++) if (c.pp_pose[e] >= 0) pord[start[c.pp_plane[e]]++] = e;
  }
  c.pl2pm.assign(pslots, -1); c.pl_plane.assign(pslots, -1); c.pl_pose.assign(pslots, 0); c.pl_part.assign(pslots, -1);
  c.pm2pl.assign(slots, -1);
  c.pl_ptr.assign(M + 1, 0); c.upart_ptr.assign(M + 1, 0);
  {
    int np = 0;
    for (int s2 = 0; s2 < E; s2++) {
      int e = pord[s2];
      c.pl2pm[s2] = e; c.pm2pl[e] = s2;
      c.pl_plane[s2] = c.pp_plane[e]; c.pl_pose[s2] = c.pp_pose[e];
      c.pl_ptr[c.pp_plane[e] + 1]++;
      bool head = (s2 % kTile == 0) || (c.pl_plane[s2] != c.pl_plane[s2 - 1]);
      if (head) { np++; c.upart_ptr[c.pl_plane[s2] + 1]++; }
      c.pl_part[s2] = np - 1;
    }
    c.n_upart = np;
    for (int l = 0; l < M; l++) { c.pl_ptr[l + 1] += c.pl_ptr[l]; c.upart_ptr[l + 1] += c.upart_ptr[l]; }
    c.heavy.clear(); c.huge.clear();   // planes with more than 8 partial sums are summed by a warp, more than 512 by a CTA
    for (int l = 0; l < M; l++) {
      const int n = c.upart_ptr[l + 1] - c.upart_ptr[l];
      if (n > 512) c.huge.push_back(l);
      else if (n > 8) c.heavy.push_back(l);
    }
    c.n_heavy = (int)c.heavy.size(); c.n_huge = (int)c.huge.size();
    if (c.heavy.empty()) c.heavy.push_back(-1);
    if (c.huge.empty()) c.huge.push_back(-1);
  }
  // ---- Hll / gl assembly tasks ----
  c.at_plane.clear(); c.at_lo.clear(); c.at_hi.clear(); c.as_plane.clear();
  c.at_ptr.assign(M + 1, 0);
  for (int l = 0; l < M; l++) {
    const int s0 = c.pl_ptr[l], s1 = c.pl_ptr[l + 1];
    c.at_ptr[l] = (int)c.at_plane.size();
    int lo = s0;
    do {   // (a plane without edges still gets one empty task: its priors are added there)
      const int hi = std::min(s1, lo + kAsmChunk);
      c.at_plane.push_back(l); c.at_lo.push_back(lo); c.at_hi.push_back(hi);
      lo = hi;
    } while (lo < s1);
    if ((int)c.at_plane.size() - c.at_ptr[l] > 1) c.as_plane.push_back(l);
  }
  c.at_ptr[M] = (int)c.at_plane.size();
  c.n_atask = (int)c.at_plane.size();
  c.n_asplit = (int)c.as_plane.size();
  if (c.at_plane.empty()) { c.at_plane.push_back(0); c.at_lo.push_back(0); c.at_hi.push_back(0); }
  if (c.as_plane.empty()) c.as_plane.push_back(-1);
  // ---- incidence lists ----
  c.pinc_ptr.assign(N + 1, 0);
  for (int f = 0; f < c.Epf; f++) { c.pinc_ptr[c.pf_i[f] + 1]++; if (c.pf_j[f] >= 0) c.pinc_ptr[c.pf_j[f] + 1]++; }
  for (int p = 0; p < N; p++) c.pinc_ptr[p + 1] += c.pinc_ptr[p];
  c.pinc.assign(c.pinc_ptr[N], 0);
  {
    std::vector<int> fill(c.pinc_ptr.begin(), c.pinc_ptr.end() - 1);
    for (int f = 0; f < c.Epf; f++) {
      c.pinc[fill[c.pf_i[f]]++] = (f << 1) | 0;
      if (c.pf_j[f] >= 0) c.pinc[fill[c.pf_j[f]]++] = (f << 1) | 1;
    }
  }
  // per pose: its first two pose-pose neighbours resolved (factor*2+side, other pose) and where the generic
  // incidence loop resumes -- one 32-byte record instead of a four-level index chain in the PCG pose phase
  c.pnbr.assign((size_t)N * 8, -1);
  for (int p = 0; p < N; p++) {
    int nf = 0, kk = c.pinc_ptr[p];
    const int i1 = c.pinc_ptr[p + 1];
    for (; kk < i1 && nf < 2; kk++) {
      const int inc = c.pinc[kk], f = inc >> 1, side = inc & 1;
      if (c.pf_j[f] < 0) continue;
      c.pnbr[(size_t)p * 8 + 2 * nf] = inc;
      c.pnbr[(size_t)p * 8 + 2 * nf + 1] = side ? c.pf_i[f] : c.pf_j[f];
      nf++;
    }
    c.pnbr[(size_t)p * 8 + 4] = kk;
    c.pnbr[(size_t)p * 8 + 5] = i1;
  }
  c.linc_ptr.assign(M + 1, 0);
  for (int f = 0; f < c.Elp; f++) c.linc_ptr[c.lp_plane[f] + 1]++;
  for (int l = 0; l < M; l++) c.linc_ptr[l + 1] += c.linc_ptr[l];
  c.linc.assign(c.linc_ptr[M], 0);
  {
    std::vector<int> fill(c.linc_ptr.begin(), c.linc_ptr.end() - 1);
    for (int f = 0; f < c.Elp; f++) c.linc[fill[c.lp_plane[f]]++] = f;
  }
  // ---- dense-block groups: the edges of every pose block grouped by plane ----
  c.blk_grp_ptr.assign(c.nblk + 1, 0);
  c.blk_simple.assign(c.nblk, 0);
  c.grp_of_slot.assign(slots, 0);
  c.grp_mem_ptr.clear(); c.grp_mem.clear(); c.grp_plane.clear();
  c.grp_mem.reserve(E); c.grp_plane.reserve(E / 2 + 16); c.grp_mem_ptr.reserve(E / 2 + 16);
  std::vector<unsigned long long> ekey;
  std::vector<int> es;
  for (int k = 0; k < c.nblk; k++) {
    int e0 = c.tile_ptr[k] * kTile, e1 = c.tile_ptr[k + 1] * kTile;
    // (plane, slot) packed into one integer: an integer sort gives the order of a stable sort by plane over the slot order
    ekey.clear();
    for (int e = e0; e < e1; e++) if (c.pp_pose[e] >= 0) ekey.push_back(((unsigned long long)(unsigned)c.pp_plane[e] << 32) | (unsigned)e);
    std::sort(ekey.begin(), ekey.end());
    es.resize(ekey.size());
    for (size_t i = 0; i < ekey.size(); i++) es[i] = (int)(unsigned)(ekey[i] & 0xffffffffull);
    int g0 = (int)c.grp_plane.size();
    for (int i = 0; i < (int)es.size(); i++) {
      if (i == 0 || c.pp_plane[es[i]] != c.pp_plane[es[i - 1]]) {
        c.grp_plane.push_back(c.pp_plane[es[i]]);
        c.grp_mem_ptr.push_back((int)c.grp_mem.size());
      }
      c.grp_of_slot[es[i]] = (int)c.grp_plane.size() - 1 - g0;
      c.grp_mem.push_back(es[i]);
    }
    c.blk_grp_ptr[k + 1] = (int)c.grp_plane.size();
    {  // fast dense-block build: at most 16 tiles, 96 planes and no pose observing one plane twice
      bool simple = (c.tile_ptr[k + 1] - c.tile_ptr[k] <= 16) && ((int)c.grp_plane.size() - g0 <= 96);
      for (int i = 1; i < (int)es.size() && simple; i++)
        if (c.pp_plane[es[i]] == c.pp_plane[es[i - 1]] && c.pp_pose[es[i]] == c.pp_pose[es[i - 1]]) simple = false;
      c.blk_simple[k] = simple ? 1 : 0;
    }
    if (c.blk_grp_ptr[k + 1] - g0 > kMaxGrp || c.blk_part_ptr[k + 1] - c.blk_part_ptr[k] > kMaxPart) {
      err = "pose block " + std::to_string(k) + " observes too many planes (limits: " + std::to_string(kMaxGrp) +
            " distinct planes / " + std::to_string(kMaxPart) + " tile runs per 16 poses)";
      return false;
    }
  }
  c.grp_mem_ptr.push_back((int)c.grp_mem.size());
  c.ngrp = (int)c.grp_plane.size();
  c.grp_info.assign((size_t)std::max(1, c.ngrp) * 4, 0);   // {plane, first partial sum, number of partial sums, -}
  for (int g = 0; g < c.ngrp; g++) {
    const int l = c.grp_plane[g];
    c.grp_info[(size_t)g * 4] = l;
    c.grp_info[(size_t)g * 4 + 1] = c.upart_ptr[l];
    c.grp_info[(size_t)g * 4 + 2] = c.upart_ptr[l + 1] - c.upart_ptr[l];
  }
  {
    // per-block partial sums of the block-resident PCG: plane l owns slots [ub_ptr[l], ub_ptr[l+1]), one per observing pose
    // block in ascending block order (groups are numbered block by block, so a counting pass keeps that order)
    std::vector<int> ub_ptr(M + 1, 0);
    for (int g2 = 0; g2 < c.ngrp; g2++) ub_ptr[c.grp_plane[g2] + 1]++;
    for (int l = 0; l < M; l++) ub_ptr[l + 1] += ub_ptr[l];
    std::vector<int> fill(ub_ptr.begin(), ub_ptr.end() - 1);
    c.grp_info2.assign((size_t)std::max(1, c.ngrp) * 4, 0);
    for (int g2 = 0; g2 < c.ngrp; g2++) {
      const int l = c.grp_plane[g2];
      c.grp_info2[(size_t)g2 * 4] = l;
      c.grp_info2[(size_t)g2 * 4 + 1] = ub_ptr[l];
      c.grp_info2[(size_t)g2 * 4 + 2] = ub_ptr[l + 1] - ub_ptr[l];
      c.grp_info2[(size_t)g2 * 4 + 3] = fill[l]++;
    }
    c.res_nt = c.res_ng = c.res_np = 0;
    for (int k = 0; k < c.nblk; k++) {
      c.res_nt = std::max(c.res_nt, c.tile_ptr[k + 1] - c.tile_ptr[k]);
      c.res_ng = std::max(c.res_ng, c.blk_grp_ptr[k + 1] - c.blk_grp_ptr[k]);
      c.res_np = std::max(c.res_np, c.blk_part_ptr[k + 1] - c.blk_part_ptr[k]);
    }
  }
  // ---- (plane, node) pairs of the two hat levels ----
  auto build_pairs = [&](int SPx, std::vector<int>& ce_ptr, std::vector<int>& ce_node, std::vector<int>& ce_plane, std::vector<int>& ce_lo,
                         std::vector<int>& ce_hi) {
    ce_ptr.assign(M + 1, 0);
    ce_node.clear(); ce_plane.clear(); ce_lo.clear(); ce_hi.clear();
    for (int l = 0; l < M; l++) {
      int s0 = c.pl_ptr[l], s1 = c.pl_ptr[l + 1];
      int last = -1;
      for (int s = s0; s < s1; s++) {
        int p = c.pl_pose[s];
        int c0 = p / SPx;
        int cand[2] = {c0, (p % SPx) ? c0 + 1 : -1};
        for (int q = 0; q < 2; q++) {
          int nd = cand[q];
          if (nd < 0 || nd <= last) continue;
          // slots of plane l supporting node nd: poses in ((nd-1)*SP, (nd+1)*SP)
          // slots are pose-sorted within the plane: lo = first slot with pose > (nd-1)*SP, hi = first with pose >= (nd+1)*SP
          const int* pb = c.pl_pose.data();
          const int lo = (int)(std::upper_bound(pb + s0, pb + s1, (nd - 1) * SPx) - pb);
          const int hi = (int)(std::lower_bound(pb + lo, pb + s1, (nd + 1) * SPx) - pb);
          ce_node.push_back(nd); ce_plane.push_back(l); ce_lo.push_back(lo); ce_hi.push_back(hi);
          last = nd;
        }
      }
      ce_ptr[l + 1] = (int)ce_node.size();
    }
  };
  c.levels = g.force_levels ? g.force_levels : (N > kL2Spacing * kMaxCoarseNodes ? 3 : 2);
  c.SP = coarse_spacing(N, c.levels);
  const int SPc = c.SP;
  c.nc = coarse_nodes(N, SPc);
  c.nc_pad = (c.nc + kPivotNodes - 1) / kPivotNodes * kPivotNodes;
  build_pairs(SPc, c.ce_ptr, c.ce_node, c.ce_plane, c.ce_lo, c.ce_hi);
  c.nc2 = coarse_nodes(N, kL2Spacing);
  c.ng2 = (c.nc2 + kGroupNodes - 1) / kGroupNodes;
  if (c.levels == 3) build_pairs(kL2Spacing, c.ce2_ptr, c.ce2_node, c.ce2_plane, c.ce2_lo, c.ce2_hi);
  else { c.ce2_ptr.assign(M + 1, 0); c.ce2_node.assign(1, 0); c.ce2_plane.assign(1, 0); c.ce2_lo.assign(1, 0); c.ce2_hi.assign(1, 0); c.ng2 = 0; }
  c.nce2 = (c.levels == 3) ? (int)c.ce2_node.size() : 0;
  if (c.levels == 3) {
    // per group: its pairs in (plane, node) order = the order they already have, restricted to the group's nodes
    c.g2_ptr.assign(c.ng2 + 1, 0);
    for (int i = 0; i < c.nce2; i++) c.g2_ptr[c.ce2_node[i] / kGroupNodes + 1]++;
    for (int gq = 0; gq < c.ng2; gq++) c.g2_ptr[gq + 1] += c.g2_ptr[gq];
    c.g2_ce.assign(std::max(1, c.nce2), 0);
    std::vector<int> fill(c.g2_ptr.begin(), c.g2_ptr.end() - 1);
    for (int i = 0; i < c.nce2; i++) c.g2_ce[fill[c.ce2_node[i] / kGroupNodes]++] = i;
  } else {
    c.g2_ptr.assign(1, 0); c.g2_ce.assign(1, 0);
  }
  c.nce = (int)c.ce_node.size();
  c.n2ce_ptr.assign(c.nc + 1, 0);
  for (int i = 0; i < c.nce; i++) c.n2ce_ptr[c.ce_node[i] + 1]++;
  for (int a = 0; a < c.nc; a++) c.n2ce_ptr[a + 1] += c.n2ce_ptr[a];
  c.n2ce.assign(c.nce, 0);
  {
    std::vector<int> fill(c.n2ce_ptr.begin(), c.n2ce_ptr.end() - 1);
    for (int i = 0; i < c.nce; i++) c.n2ce[fill[c.ce_node[i]]++] = i;
  }
  // ---- output-stationary assembly of A_c = P^T S P: for every coarse node pair (a, b) the plane products
  // Wc[a,l] Hll^-1 Wc[b,l]^T of the light planes and the pose-pose factors coupling the two supports; planes that
  // touch more than kHeavyCoarse nodes (the ground plane touches all) are applied as dense rank-3 updates instead
  {
    const int nc = c.nc;
    const size_t np = (size_t)nc * nc;
    c.hv_plane.clear();
    std::vector<char> heavy(M, 0);
    for (int l = 0; l < M; l++)
      if (c.ce_ptr[l + 1] - c.ce_ptr[l] > kHeavyCoarse) { heavy[l] = 1; c.hv_plane.push_back(l); }
    c.n_hv = (int)c.hv_plane.size();
    if (c.hv_plane.empty()) c.hv_plane.push_back(-1);
    c.lp_ptr.assign(np + 1, 0);
    for (int l = 0; l < M; l++) {
      if (heavy[l]) continue;
      for (int i = c.ce_ptr[l]; i < c.ce_ptr[l + 1]; i++)
        for (int j = c.ce_ptr[l]; j < c.ce_ptr[l + 1]; j++) c.lp_ptr[(size_t)c.ce_node[i] * nc + c.ce_node[j] + 1]++;
    }
    for (size_t q = 0; q < np; q++) c.lp_ptr[q + 1] += c.lp_ptr[q];
    c.lp_cea.assign(std::max(1, c.lp_ptr[np]), 0);
    c.lp_ceb.assign(std::max(1, c.lp_ptr[np]), 0);
    {
      std::vector<int> fill(c.lp_ptr.begin(), c.lp_ptr.end() - 1);
      for (int l = 0; l < M; l++) {   // plane-ascending order inside every pair: fixed summation order
        if (heavy[l]) continue;
        for (int i = c.ce_ptr[l]; i < c.ce_ptr[l + 1]; i++)
          for (int j = c.ce_ptr[l]; j < c.ce_ptr[l + 1]; j++) {
            int at = fill[(size_t)c.ce_node[i] * nc + c.ce_node[j]]++;
            c.lp_cea[at] = i; c.lp_ceb[at] = j;
          }
      }
    }
    auto nodes_of = [&](int p, int* out) {   // coarse nodes whose hat function is non-zero at pose p
      int n = 0, c0 = p / SPc;
      if (c0 < nc) out[n++] = c0;
      if ((p % SPc) && c0 + 1 < nc) out[n++] = c0 + 1;
      return n;
    };
    c.fp_ptr.assign(np + 1, 0);
    for (int pass = 0; pass < 2; pass++) {
      std::vector<int> fill;
      if (pass == 1) {
        for (size_t q = 0; q < np; q++) c.fp_ptr[q + 1] += c.fp_ptr[q];
        c.fp_f.assign(std::max(1, c.fp_ptr[np]), 0);
        fill.assign(c.fp_ptr.begin(), c.fp_ptr.end() - 1);
      }
      for (int f = 0; f < c.Epf; f++) {
        if (c.pf_j[f] < 0) continue;
        for (int side = 0; side < 2; side++) {
          const int p = side ? c.pf_j[f] : c.pf_i[f], o = side ? c.pf_i[f] : c.pf_j[f];
          int na[2], nb[2];
          const int ka = nodes_of(p, na), kb = nodes_of(o, nb);
          for (int x = 0; x < ka; x++)
            for (int y = 0; y < kb; y++) {
              const size_t q = (size_t)na[x] * nc + nb[y];
              if (pass == 0) c.fp_ptr[q + 1]++;
              else c.fp_f[fill[q]++] = f * 2 + side;
            }
        }
      }
    }
  }
  return true;
}

}  // namespace pus
