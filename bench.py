#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric   : Levenberg-Marquardt iterations per second (and ms per solve) of the plane-SLAM back end
workload : BASELINE config 3 -- synthetic corridor, 5 000 poses / 500 planes / 50 000 pose-plane + 5 000
           odometry edges, per-component Huber cost, PPS's iSAM properties, 20 LM iterations max
step     : one complete batch_optimization() of that graph from its dead-reckoned initial estimate
value    : whole-job LM iterations/s with the graph resident in HBM (pus_solve_resident), CUDA-event time
e2e      : the same through the reference-facing C-ABI call (pus_batch_optimize) with host buffers: vertex
           values H2D, solve, estimates + trace D2H, host mirrors refreshed -- all inside the timed region
N > 1    : one process per GPU (torchrun), each rank solves its own replica (seed = rank): weak scaling,
           no data-path collective; value = sum of iterations / max-over-ranks time
--impl reference : the CPU restatement of the reference's iSAM path (oracle/, numeric Jacobians, direct
           sparse Cholesky re-analysed every iteration) on the host cores; the reference is single-threaded.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from pop_up_slam_b200 import graphgen as gg  # noqa: E402

WORKLOAD = "config3_corridor_5000p_500pl_50000e_huber"


def ncu_dram_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one lm_kernel launch on this workload, from the committed
    `ncu --set full` capture (profiles/r1_c3_ncu_raw.csv); None if the file is missing."""
    import csv
    path = os.path.join(ROOT, "profiles", "r1_c3_ncu_raw.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            tot += float(vals[i]) * scale[units[i]]
        return tot
    except Exception:
        return None


NCU_DRAM_BYTES_PER_LAUNCH = ncu_dram_bytes()   # the solve is L2-resident: DRAM traffic << algorithmic bytes


def roofline_bytes(dims, relin, chi2_evals, pcg_iters):
    """SURVEY.md 8(d) algorithmic bytes (fp64 values, int32 ids)."""
    N, M, E_pl, E_od = dims["N"], dims["M"], dims["E_pl"], dims["E_od"]
    b_lin = 232 * E_pl + 512 * E_od + 392 * N + 128 * M
    b_chi = 88 * E_pl + 224 * E_od + 56 * N + 32 * M
    b_pcg = 288 * E_pl + 288 * E_od + 576 * N + 120 * M
    return relin * b_lin + chi2_evals * b_chi + pcg_iters * b_pcg, dict(b_lin=b_lin, b_chi=b_chi, b_pcg=b_pcg)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_workload(api, seed):
    g = gg.make_config(3, seed=seed)
    ids = gg.build_bulk(api, g)
    gg.configure(api, g)
    return g, ids


def run_reference(args, rank, world):
    """CPU arm: the oracle in its faithful mode (numeric Jacobians, ordering recomputed per solve)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_api import OracleAPI
    g = gg.make_config(3, seed=0)
    times, iters = [], []
    for step in range(args.warmup + args.steps):
        api = OracleAPI()
        api.set_jacobian_mode(0)
        api.set_reuse_ordering(0)
        gg.build_bulk(api, g)
        gg.configure(api, g)
        t0 = time.perf_counter()
        it = api.batch_optimize()
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt); iters.append(it)
        tm = api.timers()
    total_t, total_it = sum(times), sum(iters)
    value = total_it / total_t
    line = {
        "impl": "reference", "metric": "lm_iterations_per_s", "value": value, "unit": "LM iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "lm_iterations_per_solve": iters[0], "seed": 0, **g.dims()},
        "cpu_baseline": {"value": value, "unit": "LM iterations/s", "cores": 1, "kind": "port",
                         "sample": f"{len(times)} full solve(s) of the bench graph ({iters[0]} LM iterations each), single thread "
                                   "(the reference path has no threading); numeric Jacobians eps=1e-4, sparse Cholesky re-ordered per solve",
                         "phase_s_last_solve": {k: float(v) for k, v in tm.items() if k in ("linearize", "solve", "chi2", "order", "total")},
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": "LM iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def cpu_baseline_sample(g):
    """Bounded CPU sample for the N=1 line: one faithful-mode solve of the bench graph by the oracle."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_api import OracleAPI
    api = OracleAPI()
    api.set_jacobian_mode(0)
    api.set_reuse_ordering(0)
    gg.build_bulk(api, g)
    gg.configure(api, g)
    t0 = time.perf_counter()
    it = api.batch_optimize()
    dt = time.perf_counter() - t0
    tm = api.timers()
    return {"value": it / dt, "unit": "LM iterations/s", "cores": 1, "kind": "port", "ms_per_solve": 1e3 * dt, "lm_iterations": it,
            "sample": "1 full solve of the bench graph, single thread (reference path is single-threaded), numeric Jacobians, "
                      "direct sparse Cholesky with the ordering recomputed every iteration (as cholmod_analyze is upstream)",
            "phase_s": {k: float(v) for k, v in tm.items() if k in ("linearize", "solve", "chi2", "order", "total")},
            "host_cores_available": os.cpu_count(), "chi2_final": api.chi2()}, api


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch64", action="store_true")
    ap.add_argument("--no-stress", action="store_true", help="skip the config-5 (HBM-bound) leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pop_up_slam_b200 import capi
    from pop_up_slam_b200.capi import GpuGraphAPI

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(dev)          # the library launches on this stream; the CUDA events below are recorded on it
    torch.cuda.set_stream(stream)

    api = GpuGraphAPI(device=local_rank)
    api.set_stream(stream.cuda_stream)
    g, ids = build_workload(api, seed=0)     # identical replica on every rank: per-GPU work is fixed (weak scaling)
    pose_ids, plane_ids = ids["pose_ids"], ids["plane_ids"]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- resident arm: graph in HBM, device time by CUDA events ----------------
    api.upload()
    for _ in range(args.warmup):
        flush.fill_(1.0)
        api.solve_resident()
    sampler = ClockSampler(local_rank)
    sync_all()
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    iters_res, stats_res = [], []
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(float(k))           # L2 flush between timed steps (outside the event pair)
        ev[k][0].record(stream)
        iters_res.append(api.solve_resident())
        ev[k][1].record(stream)
        stats_res.append(api.stats())
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    ms_res = [a.elapsed_time(b) for a, b in ev]
    api.download()

    # ---------------- end-to-end arm: host buffers in, host buffers out, through the C-ABI ----------------
    pose_host = torch.from_numpy(g.poses_init.copy()).pin_memory().numpy()
    plane_host = torch.from_numpy(g.planes_init.copy()).pin_memory().numpy()

    def e2e_step():
        # the caller's buffers -> host mirrors (NodeT::init), then Slam::batch_optimization through the C-ABI
        # (H2D of the vertex values, solve, D2H of estimates + trace), then read the result back
        api.init_poses(pose_ids, pose_host)
        api.init_planes(plane_ids, plane_host)
        it = api.batch_optimize()
        out = api.get_poses(pose_ids[-1:])
        return it, out

    for _ in range(args.warmup):
        e2e_step()
    sync_all()
    iters_e2e, t_e2e, st_e2e = [], [], None
    for k in range(args.steps):
        flush.fill_(float(k))
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        it, _ = e2e_step()
        torch.cuda.synchronize(dev)
        t_e2e.append(time.perf_counter() - t0)
        iters_e2e.append(it)
        st_e2e = api.stats()
    sync_all()

    # ---------------- reduce over ranks ----------------
    from pop_up_slam_b200.parallel import reduce_throughput
    tot_it, tot_ms = reduce_throughput(sum(iters_res), sum(ms_res), world, dev)
    tot_e2e_it, tot_e2e_s = reduce_throughput(sum(iters_e2e), sum(t_e2e), world, dev)
    value = tot_it / (tot_ms * 1e-3)
    e2e_value = tot_e2e_it / tot_e2e_s

    # config 4 (64 TUM-scale graphs sharded over the ranks): every rank takes part
    batch64 = None
    if not args.no_batch64:
        try:
            batch64 = bench_batch64(capi, GpuGraphAPI, local_rank, stream, world, rank)
        except Exception as e:  # report, never hide
            batch64 = {"error": str(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    st = stats_res[-1]
    dims = g.dims()
    peak, peak_src = measured_peak_gbs()
    nbytes, per = roofline_bytes(dims, st["relinearizations"], st["chi2_evals"], st["pcg_iterations"])
    kern_ms = statistics.mean(s["kernel_ms"] for s in stats_res)
    achieved = nbytes / (kern_ms * 1e-3) / 1e9
    line = {
        "metric": "lm_iterations_per_s", "value": value, "unit": "LM iterations/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, **dims, "lm_iterations_per_solve": iters_res[-1], "accepted": st["accepted"],
                   "pcg_iterations_per_solve": st["pcg_iterations"], "pcg_rel_tol": api.get_solver_options().pcg_rel_tol, "seed": 0,
                   "properties": g.properties, "robust": {"kind": "huber", "b": g.robust_b},
                   "l2_flush": "256 MiB device write before every timed step (outside the per-step CUDA-event pair)",
                   "replicas": "one graph per rank, no data-path collective"},
        "ms_per_solve": tot_ms / args.steps, "wall_s_timed_region": t_wall,
        "e2e": {"value": e2e_value, "unit": "LM iterations/s", "ms_per_solve": 1e3 * tot_e2e_s / args.steps,
                "h2d_bytes_per_step": int(st_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"]) + 7 * 8,
                "timed": "pus_init_poses/planes of every vertex from pinned host arrays + pus_batch_optimize (H2D, solve, D2H) + "
                         "pus_get_poses; wall clock around the call, device synchronised on both sides"},
        "gpu_launches": args.steps * st["gpu_launches"],
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                     "traffic_source": "profiles/r1_c3_ncu_raw.csv (dram__bytes_read.sum + dram__bytes_write.sum, one --set full capture of this workload)",
                     "kernel": "lm_kernel (one persistent launch per solve)", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": nbytes, "bytes_per_unit": per,
                     "note": "config 3 fits in L2 (W = 7.2 MB): latency/barrier-bound by construction, see profiles/ for the HBM-bound config 5",
                     "phase_ms": st["phase_ms"], "grid_ctas": st["grid_ctas"], "block_threads": st["block_threads"]},
        "solve": {"chi2_initial": st["chi2_initial"], "chi2_final": st["chi2_final"]},
    }
    if not args.no_cpu_baseline and world == 1:
        cb, orc = cpu_baseline_sample(g)
        line["cpu_baseline"] = cb
        line["solve"]["chi2_final_cpu_numeric_jacobians"] = cb["chi2_final"]
        line["speedup_vs_cpu"] = {"resident": value / cb["value"], "e2e": e2e_value / cb["value"]}
        try:
            line["measurement_refresh"] = bench_refresh(api, orc, g, ids)
        except Exception as e:  # report, never hide
            line["measurement_refresh"] = {"error": str(e)}
    if batch64 is not None:
        line["batch64"] = batch64
    if not args.no_stress and world == 1:
        try:
            line["stress_c5"] = bench_stress(GpuGraphAPI, local_rank, stream)
        except Exception as e:  # report, never hide
            line["stress_c5"] = {"error": str(e)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_refresh(api, orc, g, ids):
    """Mapper_mono::update_plane_measurement over every frame of the bench graph (10 ground segments per
    frame, every pose-plane factor re-measured), host segments in -> new measurements out, GPU entry point vs the
    oracle's restatement on one host core."""
    rng = np.random.default_rng(0)
    nf = g.n_poses
    nseg = np.full(nf, 10)
    seg_ptr = np.concatenate([[0], np.cumsum(nseg)]).astype(np.int32)
    n = int(seg_ptr[-1])
    segs = np.stack([rng.uniform(0, 640, n), rng.uniform(300, 480, n), rng.uniform(0, 640, n), rng.uniform(300, 480, n)], axis=1).astype(np.float32)
    invK = np.linalg.inv(np.array([[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1.0]])).astype(np.float32)
    mf = g.pp_pose.astype(np.int32)
    mr = (np.arange(len(mf)) % 11).astype(np.int32)
    t_gpu = []
    for _ in range(5):
        t0 = time.perf_counter()
        api.refresh_plane_measurements(ids["pose_ids"], seg_ptr, segs, invK, ids["pp_fids"], mf, mr)
        t_gpu.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    orc.refresh_plane_measurements(ids["pose_ids"], seg_ptr, segs, invK, ids["pp_fids"], mf, mr)
    t_cpu = time.perf_counter() - t0
    return {"frames": int(nf), "segments": n, "factors_refreshed": int(len(mf)), "gpu_ms": 1e3 * statistics.median(t_gpu),
            "cpu_port_ms": 1e3 * t_cpu, "timed": "wall clock around the C-ABI call incl. H2D of the segments and D2H of the new measurements"}


def bench_stress(GpuGraphAPI, device, stream):
    """BASELINE config 5 (50 k poses / 5 k planes / 1 M edges): the HBM-bound graph.  Three LM iterations, resident,
    CUDA events; algorithmic bytes as for the headline roofline plus the 96x96 preconditioner blocks the PCG streams."""
    import torch
    g = gg.make_config(5, seed=0, max_iterations=3)
    a = GpuGraphAPI(device=device)
    a.set_stream(stream.cuda_stream)
    gg.build_bulk(a, g)
    gg.configure(a, g)
    a.upload()
    a.solve_resident()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    e0.record(stream)
    for _ in range(reps):
        it = a.solve_resident()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = a.stats()
    dims = g.dims()
    nbytes, per = roofline_bytes(dims, st["relinearizations"], st["chi2_evals"], st["pcg_iterations"])
    binv = st["pcg_iterations"] * ((dims["N"] + 15) // 16) * 96 * 96 * 8
    peak, peak_src = measured_peak_gbs()
    return {"workload": "BASELINE config 5: 50k poses / 5k planes / 1M pose-plane + 50k odometry edges, 3 LM iterations",
            "ms_per_solve": ms, "lm_iterations": int(it), "pcg_iterations": st["pcg_iterations"],
            "ms_per_pcg_iteration": ms / max(1, st["pcg_iterations"]),
            "algorithmic_gbs": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / peak,
            "with_preconditioner_blocks_gbs": (nbytes + binv) / (ms * 1e-3) / 1e9,
            "with_preconditioner_blocks_frac": (nbytes + binv) / (ms * 1e-3) / 1e9 / peak,
            "peak": peak, "peak_source": peak_src, "bytes_per_unit": per, "grid_ctas": st["grid_ctas"]}


def bench_batch64(capi, GpuGraphAPI, device, stream, world, rank):
    """BASELINE config 4: 64 independent TUM-scale graphs, round-robined over the ranks (parallel.shard), each rank
    solving its share in one persistent launch (one CTA team per graph); no data-path collective.  Whole-job numbers:
    units summed over ranks / max-over-ranks device time."""
    import torch
    from pop_up_slam_b200 import parallel
    mine = parallel.shard(64, rank, world)
    apis, graphs = [], []
    for s in mine:
        g = gg.make_config(2, seed=s)
        a = GpuGraphAPI(device=device)
        a.set_stream(stream.cuda_stream)
        gg.build_bulk(a, g)
        gg.configure(a, g)
        apis.append(a); graphs.append(g)
    capi.upload_many(apis)
    capi.solve_resident_many(apis)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        its = capi.solve_resident_many(apis)
    e1.record(stream)
    torch.cuda.synchronize()
    ms_local = e0.elapsed_time(e1) / reps
    nbytes = 0
    for a, g in zip(apis, graphs):
        st = a.stats()
        nbytes += roofline_bytes(g.dims(), st["relinearizations"], st["chi2_evals"], st["pcg_iterations"])[0]
    dev = torch.device("cuda", device)
    tot_its, ms = parallel.reduce_throughput(float(its.sum()), ms_local, world, dev)
    tot_bytes, _ = parallel.reduce_throughput(float(nbytes), ms_local, world, dev)
    peak, _ = measured_peak_gbs()
    gbs = tot_bytes / (ms * 1e-3) / 1e9
    return {"graphs": 64, "n_gpus": world, "graphs_per_rank": len(mine), "ms_per_batch": ms, "graphs_per_s": 64 / (ms * 1e-3),
            "lm_iterations_per_s": tot_its / (ms * 1e-3), "lm_iterations_total": int(tot_its),
            "algorithmic_gbs": gbs, "frac_of_hbm_roofline": gbs / (peak * world),
            "workload": "64 x config 2 (300 poses, 60 planes, 2100 edges), seeds 0..63, round-robin over ranks",
            "grid_ctas": apis[0].stats()["grid_ctas"]}


if __name__ == "__main__":
    main()
