// Replays a recorded frame sequence through include/isam_facade.hpp exactly the way
// pop_planar_slam's Mapper_mono::processFrame drives iSAM (pop_planar_slam/src/Mapping.cpp:31-43, 464-554):
// new Pose3d_Node, prior | odometry factor, new Plane3d_Node per unseen plane (+ ground prior),
// Pose3d_Plane3d_Factor per observation, batch_optimization() every 5th frame, update() otherwise.
// Prints the final chi2, poses and planes; tests/test_facade.py compares them with the oracle.
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <vector>

#include "../include/isam_facade.hpp"

using namespace isam;

static int run(int argc, char** argv);
// a diagonal covariance for either flavour of the facade's dense types (built-in, or Eigen with -DISAM_FACADE_USE_EIGEN: what
// Mapping.cpp itself does at :64-67 -- MatrixXd::Zero + diagonal entries)
static MatrixXd diag_matrix(const std::vector<double>& d) {
  MatrixXd m((int)d.size(), (int)d.size());
  for (size_t i = 0; i < d.size(); i++)
    for (size_t j = 0; j < d.size(); j++) m((int)i, (int)j) = (i == j) ? d[i] : 0.0;
  return m;
}

int main(int argc, char** argv) {
  try { return run(argc, argv); }
  catch (const std::exception& e) { fprintf(stderr, "facade_replay: %s\n", e.what()); return 1; }
}
static int run(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: facade_replay <frames.txt>\n"); return 2; }
  FILE* f = fopen(argv[1], "r");
  if (!f) { perror("open"); return 2; }
  int n_frames, n_planes, ground;
  if (fscanf(f, "%d %d %d", &n_frames, &n_planes, &ground) != 3) return 2;
  Slam slam;
  Properties prop = slam.properties();   // Mapping.cpp:31-43
  prop.method = LEVENBERG_MARQUARDT;
  prop.mod_batch = 1;
  prop.quiet = true;
  prop.epsilon2 *= 0.1;
  prop.epsilon_abs *= 0.1;
  prop.epsilon_rel *= 0.1;
  slam.set_properties(prop);
  std::vector<Pose3d_Node*> poses;
  std::vector<Plane3d_Node*> planes(n_planes, nullptr);
  double pose_sig[6], ground_sig;
  for (int i = 0; i < 6; i++) if (fscanf(f, "%lf", &pose_sig[i]) != 1) return 2;
  if (fscanf(f, "%lf", &ground_sig) != 1) return 2;
  std::vector<double> pv(6);
  for (int i = 0; i < 6; i++) pv[i] = pose_sig[i] * pose_sig[i];
  Covariance poseCov(diag_matrix(pv));
  Covariance groundCov(diag_matrix({ground_sig * ground_sig, ground_sig * ground_sig, ground_sig * ground_sig}));
  for (int fr = 0; fr < n_frames; fr++) {
    double o[6];
    for (int i = 0; i < 6; i++) if (fscanf(f, "%lf", &o[i]) != 1) return 2;
    Pose3d temp_pose(o[0], o[1], o[2], o[3], o[4], o[5]);
    Pose3d estimate_pose;
    if (!poses.empty()) estimate_pose = poses.back()->value().oplus(temp_pose);   // Mapping.cpp:414-416
    Pose3d_Node* poseNode = new Pose3d_Node();
    slam.add_node(poseNode);
    if (poses.empty()) {
      slam.add_factor(new Pose3d_Factor(poseNode, temp_pose, poseCov));            // Mapping.cpp:470-473
    } else {
      poseNode->init(estimate_pose);                                               // Mapping.cpp:475
      slam.add_factor(new Pose3d_Pose3d_Factor(poses.back(), poseNode, temp_pose, poseCov));
    }
    poses.push_back(poseNode);
    if (fr == 0) estimate_pose = poseNode->value();
    int nobs;
    if (fscanf(f, "%d", &nobs) != 1) return 2;
    for (int k = 0; k < nobs; k++) {
      int pl; double m[4], sig;
      if (fscanf(f, "%d %lf %lf %lf %lf %lf", &pl, &m[0], &m[1], &m[2], &m[3], &sig) != 6) return 2;
      Plane3d measure(Vector4d(m[0], m[1], m[2], m[3]));
      if (!planes[pl]) {
        planes[pl] = new Plane3d_Node();
        slam.add_node(planes[pl]);
        planes[pl]->init(measure.transform_from(estimate_pose.oTw()));             // Mapping.cpp:497-499
        if (pl == ground) slam.add_factor(new Plane3d_Factor(planes[pl], Plane3d(Vector4d(0, 0, -1, 0)), groundCov));
      }
      Covariance cov(diag_matrix({sig * sig, sig * sig, sig * sig}));
      slam.add_factor(new Pose3d_Plane3d_Factor(poseNode, planes[pl], measure, cov, false));
    }
    if (fr % 5 == 0) slam.batch_optimization(); else slam.update();                 // Mapping.cpp:550-554
  }
  printf("chi2 %.17g\n", slam.chi2());
  printf("nodes %d factors %d\n", (int)slam.get_nodes().size(), (int)slam.get_factors().size());
  for (auto* p : poses) { double v[7]; p->value().to7(v); printf("pose %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", v[0], v[1], v[2], v[3], v[4], v[5], v[6]); }
  for (auto* p : planes) if (p) { Vector4d v = p->value().vector(); printf("plane %.17g %.17g %.17g %.17g\n", v(0), v(1), v(2), v(3)); }
  return 0;
}
