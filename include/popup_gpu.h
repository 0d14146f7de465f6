/*
 * popup_gpu.h -- C-ABI of libpopup_gpu.so, the B200-native plane-SLAM back end.
 *
 * This is the drop-in boundary for the ONE hot path of shichaoy/pop_up_slam that this
 * repository accelerates: the factor-graph optimiser that pop_planar_slam/src/Mapping.cpp
 * drives through iSAM's C++ class API, plus the per-frame pop-up wall fit.  The reference
 * boundary is an in-process C++ API (no FFI exists upstream); every entry point below names
 * the reference interface it replaces (paths relative to the reference checkout;
 * ISAM = pop_planar_slam/Thirdparty/isam, PPS = pop_planar_slam, PUW = pop_up_wall).
 * include/isam_facade.hpp rebuilds the isam:: classes on top of this ABI so Mapping.cpp
 * compiles unchanged; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - opaque handle, single-threaded per handle (the reference is single-threaded);
 *   - every function returns an int status: >= 0 success (ids / counts where documented),
 *     < 0 error, message via pus_last_error();
 *   - all arrays are caller-owned host memory, copied during the call;
 *   - node / factor ids are assigned in insertion order, per handle, starting at 0
 *     (ISAM/isamlib/Slam.cpp:47-48, Node.h:62-64, Factor.h:92);
 *   - a pose value is 7 doubles  x y z qw qx qy qz   (Pose3d = Point3d + Rot3d quaternion,
 *     ISAM/include/isam/Pose3d.h:78-79, Rot3d.h:144);
 *   - a pose measurement is 6 doubles  x y z yaw pitch roll  (Pose3d::vector(), Pose3d.h:138-145);
 *   - a plane is 4 doubles  a b c d , normalised to unit 4-norm on entry
 *     (Plane3d(const Vector4d&), PPS/src/isam_plane3d.h:59-66);
 *   - sqrt-information matrices are upper triangular, packed row-major
 *     (6 doubles for 3x3, 21 for 6x6; Noise.h:36-62, Factor.h:84-91,148-155).
 *   There is NO CPU fallback: if no CUDA device / kernel image is usable, calls fail.
 *
 * Limits (a solve fails with a message instead of producing a wrong answer):
 *   - any 16 consecutive poses (in insertion order) may observe at most 256 distinct planes and produce at most 64
 *     (32-edge tile, pose) runs, i.e. about 96 plane observations per pose -- fixed shared-memory sizes of the fused
 *     pose phase; upstream iSAM has no such limit (the reference's own graphs: <= 10 planes per key-frame);
 *   - Slam::update() only with mod_batch = 1 (what PPS sets); relative-parameterised plane factors, anchor nodes,
 *     dog-leg and covariance recovery are not provided.
 */
#ifndef POPUP_GPU_H
#define POPUP_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pus_handle;

/* isam::Properties (ISAM/include/isam/Properties.h:37-110) -- the fields PPS touches
 * (PPS/src/Mapping.cpp:32-43) plus the LM constants. */
typedef struct pus_properties {
  int method;              /* 0 GAUSS_NEWTON (default), 1 LEVENBERG_MARQUARDT */
  double epsilon2;         /* stop when ||delta|| <= epsilon2        (1e-2) */
  double epsilon_abs;      /* stop when chi2 <= epsilon_abs          (1e-3) */
  double epsilon_rel;      /* stop when d(chi2) < epsilon_rel*chi2   (1e-5) */
  int max_iterations;      /* (500) */
  double lm_lambda0;       /* (1e-6) */
  double lm_lambda_factor; /* (10) */
  int mod_update;          /* (1) */
  int mod_batch;           /* (100); PPS sets 1 */
  int mod_solve;           /* (1) */
} pus_properties;

/* Solver knobs that have no counterpart in the reference (it uses a direct CHOLMOD solve,
 * ISAM/isamlib/Cholesky.cpp:68-147; here the reduced pose system is solved by block-PCG). */
typedef struct pus_solver_options {
  double pcg_rel_tol;  /* ||r||_Minv / ||b||_Minv target, default 1e-8 (four decades under the 1e-4 parity bar).  At 1e-7 the bench
                        * workload is 8 % faster and still follows the reference trajectory, but the estimates of an UNCONVERGED
                        * 4 200-pose Huber corridor after 5 capped iterations drift to 1.15e-4 of the oracle's: too close to the bar */
  int pcg_max_iter;    /* default 2000 */
  int ctas_per_sm;     /* persistent grid = ctas_per_sm * #SM, 0 = auto */
  int team_ctas;       /* CTAs cooperating on one graph; 0 = auto (whole grid for one graph) */
  int reserved[4];     /* diagnostics, all 0 by default:
                        *   [0] = 1  rebuild the preconditioner for every linear solve (no lazy refresh)
                        *   [1] = 1  no PCG warm start after a rejected LM step
                        *   [2]      bit 0: sub-phase timers in pus_stats.phase_ms[13..23]; bit 1: always stage W / Wt
                        *            tiles by bulk async copy (the large-graph data path); bit 2: never; bit 3: force the
                        *            three-level preconditioner (default: graphs > 5120 poses); bit 4: force two levels;
                        *            bit 5: no block-resident PCG loop; bit 6 / 7: other timer sets; bits 8-15: the CTA of the
                        *            team that keeps the timers; bit 16: never run a team as a thread-block cluster (global-
                        *            memory barrier only; the environment variable PUS_CLUSTER=0 does the same process-wide)
                        *   [3] > 0  lazy-refresh threshold in percent of the post-build iteration count (default 130, three
                        *            levels 200) */
} pus_solver_options;

typedef struct pus_stats {
  int lm_iterations;      /* trial steps of the last batch_optimize (Optimizer.cpp num_iter) */
  int accepted;           /* accepted steps */
  int relinearizations;   /* Jacobian sweeps */
  int chi2_evals;         /* trial chi2 sweeps */
  long long pcg_iterations; /* total PCG iterations of the last call */
  double chi2_initial, chi2_final;
  double kernel_ms;       /* device time of the solve kernel(s), CUDA events on the handle's stream */
  double h2d_ms, d2h_ms;
  long long h2d_bytes, d2h_bytes;
  int n_poses, n_planes, n_pose_plane, n_odometry, n_pose_prior, n_plane_prior;
  int gpu_launches;       /* kernels launched by the last call */
  int grid_ctas, block_threads;
  double phase_ms[24];    /* in-kernel %globaltimer split: 0 linearise, 1 Schur set-up, 2 PCG, 3 back-substitution + update,
                             4 chi2; 8..12 set-up parts (Hll^-1, dense blocks, Wc, A_c, A_c^-1); 16..20 PCG parts
                             (plane sweep, plane solve, pose sweep, pose side + dot, preconditioner + dot) */
} pus_stats;

/* ---- lifetime ------------------------------------------------------------------------- */
/* isam::Slam::Slam()  ISAM/isamlib/Slam.cpp:69-78 */
int pus_create(int device, pus_handle* out);
int pus_destroy(pus_handle h);
const char* pus_last_error(void);
/* run the solve kernels on this CUDA stream (cudaStream_t as void*), 0 = library-owned stream */
int pus_set_stream(pus_handle h, void* cuda_stream);

/* ---- vertices --------------------------------------------------------------------------- */
/* new Pose3d_Node() + Slam::add_node  (slam3d.h:39, Slam.cpp:91-94; PPS Mapping.cpp:464-465).
 * init7 may be NULL (uninitialised node, NodeT::NodeT Node.h:107-110). Returns the node id. */
int pus_add_pose(pus_handle h, const double* init7);
/* new Plane3d_Node() + add_node  (isam_plane3d.h:197-210; Mapping.cpp:483-484) */
int pus_add_plane(pus_handle h, const double* abcd);
/* bulk forms: values may be NULL; out_ids may be NULL; returns the first id */
int pus_add_poses(pus_handle h, int n, const double* init7s, int* out_ids);
int pus_add_planes(pus_handle h, int n, const double* abcds, int* out_ids);
/* NodeT::init  Node.h:123-126 (Mapping.cpp:475, 499) */
int pus_init_pose(pus_handle h, int id, const double* init7);
int pus_init_plane(pus_handle h, int id, const double* abcd);
/* bulk NodeT::init over n vertices (values packed like the single-vertex forms) */
int pus_init_poses(pus_handle h, int n, const int* ids, const double* init7s);
int pus_init_planes(pus_handle h, int n, const int* ids, const double* abcds);
/* NodeT::value()  Node.h:130 */
int pus_get_pose(pus_handle h, int id, double* out7);
int pus_get_plane(pus_handle h, int id, double* out4);
int pus_get_poses(pus_handle h, int n, const int* ids, double* out7s);
int pus_get_planes(pus_handle h, int n, const int* ids, double* out4s);

/* ---- edges ------------------------------------------------------------------------------ */
/* Pose3d_Factor(node, prior, noise) + add_factor  (slam3d.h:58-89; Mapping.cpp:472-473) */
int pus_add_pose_prior(pus_handle h, int pose, const double* xyzypr, const double* sqrtinf_ut21);
/* Pose3d_Pose3d_Factor(n1, n2, measure, noise) + add_factor  (slam3d.h:91-193; Mapping.cpp:477-478).
 * Initialises whichever pose is uninitialised (slam3d.h:123-137). */
int pus_add_odometry(pus_handle h, int pose1, int pose2, const double* xyzypr, const double* sqrtinf_ut21);
/* Pose3d_Plane3d_Factor(pose, plane, measure, noise, relative=false) + add_factor
 * (isam_plane3d.h:221-308; Mapping.cpp:513,523). Initialises the plane on first sight (:252-264). */
int pus_add_pose_plane(pus_handle h, int pose, int plane, const double* meas_abcd, const double* sqrtinf_ut6);
/* Plane3d_Factor(plane, prior, noise) + add_factor  (isam_plane3d.h:428-474; Mapping.cpp:502-503) */
/* Pose3d_Plane3d_Factor2 (pop_planar_slam/src/isam_plane3d.h:314-424): like pus_add_pose_plane, but the residual
 * re-pops the measured wall plane from the two precomputed ground-edge rays (sensor frame, rays6 = r0, r1 as
 * precompute_edge_ray :358-370 forms them: invK * (x, y, 1)) with the CURRENT pose on every evaluation
 * (get_wall_plane_equation, isam_plane3d.cpp:20-55); meas4 only initialises an un-initialised plane node. */
int pus_add_pose_plane2(pus_handle h, int pose, int plane, const double* meas4, const double* rays6, const double* sqrtinf_ut6);
int pus_add_plane_prior(pus_handle h, int plane, const double* abcd, const double* sqrtinf_ut6);
/* bulk forms (same semantics, applied in array order); out_fids may be NULL; returns first fid */
int pus_add_odometry_bulk(pus_handle h, int n, const int* pose1, const int* pose2, const double* xyzypr,
                          const double* sqrtinf_ut21, int* out_fids);
int pus_add_pose_plane_bulk(pus_handle h, int n, const int* pose, const int* plane, const double* meas_abcd,
                            const double* sqrtinf_ut6, int* out_fids);
/* FactorT::set_measurement / measurement()  Factor.h:203-206 (Mapping.cpp:603, 668) */
int pus_set_measurement(pus_handle h, int fid, const double* meas);
int pus_get_measurement(pus_handle h, int fid, double* meas);
/* Slam::remove_factor / remove_node  Slam.cpp:107-126 (Mapping.cpp:673, 699) */
int pus_remove_factor(pus_handle h, int fid);
int pus_remove_node(pus_handle h, int id);
/* Graph::num_nodes / num_factors  Graph.h:61-62 (Mapping.cpp:547-548) */
int pus_num_nodes(pus_handle h);
int pus_num_factors(pus_handle h);
/* Factor::nodes() Factor.h:79 / Node::factors() Node.h:88 (Mapping.cpp:664, 667) */
int pus_factor_nodes(pus_handle h, int fid, int* out2);             /* returns #nodes */
int pus_node_factors(pus_handle h, int id, int* out, int capacity); /* returns #factors (insertion order) */
/* column offset of a node / row offset of a factor in the stacked Jacobian
 * (Slam::update_starts Slam.cpp:59-67; jacobian_partial Slam.cpp:395-432) -- the
 * "edge indexing bit-exact" contract. -1 if removed. */
int pus_node_start(pus_handle h, int id);
int pus_factor_row(pus_handle h, int fid);

/* ---- configuration ---------------------------------------------------------------------- */
/* Slam::properties / set_properties  Slam.h:92-101 (Mapping.cpp:32-43) */
int pus_get_properties(pus_handle h, pus_properties* out);
int pus_set_properties(pus_handle h, const pus_properties* in);
/* Slam::set_cost_function  Slam.cpp:212-214 with robust.h:101-118:
 * kind 0 none, 1 cost_huber(d,b), 2 cost_pseudo_huber(d,b); applied per residual component
 * (Factor.h:67-77). */
int pus_set_robust(pus_handle h, int kind, double b);
/* Properties::force_numerical_jacobian (ISAM/include/isam/Properties.h:44-45) / Factor::jacobian -> numericalDiff
 * (Factor.h:126-139, ISAM/isamlib/numericalDiff.cpp:41-87).  numeric = 0 (default): closed-form Jacobian blocks;
 * numeric = 1: the reference's own scheme on the device -- central differences, epsilon = 1e-4, through the exmaps, of the
 * weighted robustified residual, linearisation point restored through the Euler round trip as upstream.  With it a solve
 * follows the reference's trajectory (same accept / reject sequence) where its truncation error decides near-tie steps. */
int pus_set_jacobian_mode(pus_handle h, int numeric);
int pus_get_solver_options(pus_handle h, pus_solver_options* out);
int pus_set_solver_options(pus_handle h, const pus_solver_options* in);

/* ---- optimisation ----------------------------------------------------------------------- */
/* int Slam::batch_optimization()  Slam.cpp:198-210 -> Optimizer::batch_optimize
 * (Optimizer.cpp:538-555): GN (286-366) or LM (371-467) per properties.method.
 * Host graph -> HBM, device-resident loop, estimates -> host mirrors. */
int pus_batch_optimize(pus_handle h, int* iterations);
/* UpdateStats Slam::update()  Slam.cpp:157-196; with mod_batch==1 (PPS) = relinearise +
 * one un-damped Gauss-Newton step (Optimizer::relinearize Optimizer.cpp:114-185). */
int pus_update(pus_handle h);
/* double Slam::chi2()  Slam.cpp:266-268 (ESTIMATE) */
int pus_chi2(pus_handle h, double* out);
/* many independent graphs in one launch (BASELINE config 4); one CTA team per graph */
int pus_batch_optimize_many(pus_handle* hs, int n, int* iterations);

/* split form of pus_batch_optimize for measurement: upload (graph compile + H2D),
 * solve from the uploaded initial estimate with everything resident in HBM (repeatable),
 * download (D2H into the host mirrors). */
int pus_upload(pus_handle h);
int pus_solve_resident(pus_handle h, int* iterations);
int pus_download(pus_handle h);
int pus_upload_many(pus_handle* hs, int n);
int pus_solve_resident_many(pus_handle* hs, int n, int* iterations);
int pus_download_many(pus_handle* hs, int n);

int pus_get_stats(pus_handle h, pus_stats* out);
/* per LM trial step: lambda used, chi2 of the trial, chi2 before, ||delta||, accepted flag,
 * PCG iterations of the solve that produced delta. Returns the number of entries. */
int pus_get_trace(pus_handle h, int capacity, double* lambda, double* chi2_new, double* chi2_before,
                  double* delta_norm, int* accepted, int* pcg_iters);

/* ---- pop-up wall fit (float32) ---------------------------------------------------------- */
/* popup_plane::get_plane_equation (PUW/libs/popup_plane.cpp:551-652, mode 0, also
 * update_plane_equation_from_seg :654-705) and update_plane_equation_from_seg_fast
 * (:708-749, mode 1), batched over frames:
 *   seg_ptr[n_frames+1]  CSR offsets into segs (4 floats each: x1 y1 x2 y2, pixels)
 *   invK[9] row-major, Ts[n_frames*16] row-major camera-to-world
 *   outputs have one row per (frame, plane) with plane 0 = ground, i.e. n_seg(f)+1 rows
 *   per frame at row offset seg_ptr[f]+f:  planes_world/sensor 4 floats, dist 1, good 1 (0/1).
 * Any output may be NULL. */
int pus_popup_fit_frames(int device, int n_frames, const int* seg_ptr, const float* segs, const float* invK,
                         const float* Ts, float dist_thre, int mode, float* planes_world, float* planes_sensor,
                         float* dist, int* good);

/* ---- device-resident measurement refresh and polygon re-projection -------- */
/* Mapper_mono::update_plane_measurement (pop_planar_slam/src/Mapping.cpp:590-607): after a solve every past
 * frame re-pops its wall planes from its ground segments with the frame's LATEST pose estimate
 * (popup_plane::update_plane_equation_from_seg, popup_plane.cpp:654-705, float32) and stores the sensor-frame
 * plane of every kept observation as the new measurement of its pose-plane factor
 * (Plane3d(Vector4d) normalisation, isam_plane3d.h:59-66; FactorT::set_measurement, Factor.h:203-206).
 * Here the poses are read from the device-resident estimates of the last solve / upload (uploaded first if the
 * graph is not resident), the fit and the factor store run on the device, and only the new measurements come
 * back to refresh the host mirrors (pus_get_measurement):
 *   frame_pose[n_frames]   pose node id of every frame
 *   seg_ptr[n_frames+1], segs, invK   as pus_popup_fit_frames
 *   map_fid / map_frame / map_row [n_map]   factor id, frame index, plane row inside the frame
 *                                           (0 = ground, 1 + i = segment i: good_plane_indices)
 *   new_meas[n_map*4]      optional copy of the stored measurements */
int pus_refresh_plane_measurements(pus_handle h, int n_frames, const int* frame_pose, const int* seg_ptr,
                                   const float* segs, const float* invK, int n_map, const int* map_fid,
                                   const int* map_frame, const int* map_row, double* new_meas);
/* Resident form of the same sweep (the reference re-pops EVERY past frame after EVERY solve, Mapping.cpp:590-607, while the
 * frames' segment lists never change): pus_refresh_bind uploads the tables once (again whenever frames / factors are added);
 * pus_refresh_run re-pops every bound frame from the device-resident pose estimates and stores the new measurements straight
 * into the device factor store.  new_meas = NULL: nothing returns to the host -- the host mirrors (pus_get_measurement,
 * FactorT::measurement()) are refreshed lazily on their next read or at the next layout rebuild; new_meas != NULL: [n_map][4]
 * out, mirrors refreshed at once.  A structural edit between bind and run is fine as long as the bound frames / factors still
 * exist (the layout-dependent tables are rebuilt by the run). */
int pus_refresh_bind(pus_handle h, int n_frames, const int* frame_pose, const int* seg_ptr, const float* segs, const float* invK,
                     int n_map, const int* map_fid, const int* map_frame, const int* map_row);
int pus_refresh_run(pus_handle h, double* new_meas);
/* Plane3d::project_to_plane (isam_plane3d.h:172-177) over point lists, as Mapper_mono::reproj_to_newplane
 * applies it to every stored polygon vertex (Mapping.cpp:609-632): pts_out[i] = float(project(double(pts_in[i])))
 * onto the current device-resident estimate of plane node plane_of_point[i]. */
int pus_project_to_planes(pus_handle h, int n_points, const int* plane_of_point, const float* pts_in, float* pts_out);

/* ---- one graph spanning several ranks (second bullet) --------------------------- */
/* Every rank holds the whole graph (build it identically on each) and runs its own persistent kernel; the PCG
 * phases -- the two Schur sweeps, the pose phase, the preconditioner, the dot products -- are split over the CTAs of
 * ALL ranks.  The vectors they exchange (plane partial sums, p, q, r, x, z, the coarse residuals, the reduction
 * partials) sit in one "mirror arena" per rank; owners store their entries into every peer's arena through
 * NVLink peer memory and the ranks meet at a cross-rank barrier (system-scope atomics) where the single-GPU path
 * has its team barrier -- no host round trip and no NCCL call inside the solve.  Linearisation, the Schur
 * set-up and the LM update are replicated, so all ranks take identical decisions and end with identical estimates.
 *   pus_span_export   uploads the graph if needed and writes the 64-byte CUDA IPC handle of this rank's arena
 *   pus_span_connect  opens the peers' handles (handles = world x 64 bytes, in rank order; own entry ignored)
 *   pus_span_optimize batch_optimization() of the spanning graph; call on every rank
 *   pus_span_disconnect closes the peer mappings (also done by pus_destroy)
 * pus_span_emulate_optimize runs the same protocol with `world` handles on ONE device (one CTA team per handle inside
 * a single launch) -- the test vehicle for the protocol on a single GPU. */
int pus_span_export(pus_handle h, void* ipc_handle_64);
int pus_span_connect(pus_handle h, int rank, int world, const void* handles);
int pus_span_optimize(pus_handle h, int* iterations);
int pus_span_disconnect(pus_handle h);
int pus_span_emulate_optimize(pus_handle* handles, int world, int* iterations);

/* ---- graph text I/O (host only) ---------------------------------------------- */
/* Slam::save (ISAM/isamlib/Slam.cpp:84-89 -> Graph::write, Graph.h:120-131): every factor, then every node, one per
 * line: "<Factor name> <node ids> <measure> {sqrtinf upper triangle, row-wise, comma separated}" (Factor.h:148-155,
 * 169-190, 208-211) and "<Type>_Node <id> <value>" (Node.h:148-153); Pose3d prints "(x, y, z; yaw, pitch, roll)"
 * (Pose3d.h:169-172), Plane3d "(a, b, c; d)" (isam_plane3d.h:190-192).  The plane prior keeps the reference's
 * name "Pose3d_Factor" (isam_plane3d.h:438).  precision <= 0 = the stream default of the reference (6 digits). */
int pus_save_graph(pus_handle h, const char* path, int precision);
/* The 3-D part of the iSAM dataset grammar (ISAM/isam/Loader.cpp:316-392): EDGE3 i j x y z roll pitch yaw [21
 * sqrt-information entries, rotational block re-ordered] with the reversal of edges that point backwards and the
 * 100*I prior on the first pose (Loader.cpp:48-64), POSE3D_INIT; EDGE3_INIT / POSE3D_TRUE / EDGE3_TRUE / SOLVE are
 * skipped as upstream; the 2-D and stereo keywords are rejected.  Nodes are left to the factors' initialize().
 * n_poses / n_factors (optional) receive what was added. */
int pus_load_isam_dataset(pus_handle h, const char* path, int* n_poses, int* n_factors);

/* ---- debug / test hooks ----------------------------------------------------------------- */
/* copy a named device buffer of the last upload/solve to the host as doubles
 * ("Hpp","gp","Hll","gl","W","Wt","Hoff","Minv","delta", ...). Returns element count, <0 if unknown. */
long long pus_debug_fetch(pus_handle h, const char* name, double* out, long long capacity);
/* run single stages on the uploaded graph: 0 linearise at the current linpoint (fills Hpp..W),
 * 1 build Schur preconditioner for `lambda`, 2 one solve for `lambda` (fills "delta"),
 * 3 y = S*x for the vector previously stored with pus_debug_store("x") */
int pus_debug_run_stage(pus_handle h, int stage, double lambda);
long long pus_debug_store(pus_handle h, const char* name, const double* in, long long count);

#ifdef __cplusplus
}
#endif
#endif /* POPUP_GPU_H */
