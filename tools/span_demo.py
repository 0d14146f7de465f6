"""One graph spanning the ranks of a torchrun job (one process per GPU): every rank builds the same graph, the PCG
phases are split over all GPUs through CUDA-IPC peer memory.  Prints parity against the single-GPU solve and timings.
usage: torchrun --nproc-per-node N tools/span_demo.py [config] [max_iterations] [n_poses n_planes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from pop_up_slam_b200 import graphgen as gg, parallel
from pop_up_slam_b200.capi import GpuGraphAPI

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
kw = {}
if len(sys.argv) > 2: kw["max_iterations"] = int(sys.argv[2])
if len(sys.argv) > 4: kw["n_poses"], kw["n_planes"] = int(sys.argv[3]), int(sys.argv[4])
g = gg.make_config(cfg, seed=0, **kw)

ref = GpuGraphAPI(device=local)
ir = gg.build_bulk(ref, g); gg.configure(ref, g)
it_ref = ref.batch_optimize()
it_ref = ref.batch_optimize() if False else it_ref
st_ref = ref.stats()
P_ref = ref.get_poses(ir["pose_ids"]); c_ref = ref.chi2()

api = GpuGraphAPI(device=local)
ia = gg.build_bulk(api, g); gg.configure(api, g)
t0 = time.perf_counter()
it = parallel.span_optimize(api, rank, world)
dt = time.perf_counter() - t0
st = api.stats()
P = api.get_poses(ia["pose_ids"]); c = api.chi2()
diff = float(np.abs(P - P_ref).max())
allP = [None] * world
if world > 1:
    dist.all_gather_object(allP, P.tobytes())
    same = all(b == allP[0] for b in allP)
else:
    same = True
if rank == 0:
    print("graph", g.dims(), "world", world)
    print("single GPU : iters %d chi2 %.9g kernel %.2f ms pcg %d grid %d" % (it_ref, c_ref, st_ref["kernel_ms"], st_ref["pcg_iterations"], st_ref["grid_ctas"]))
    print("spanning   : iters %d chi2 %.9g kernel %.2f ms pcg %d grid/rank %d  (call incl. IPC exchange %.1f ms)" %
          (it, c, st["kernel_ms"], st["pcg_iterations"], st["grid_ctas"], dt * 1e3))
    print("max |pose - single| = %.3e ; all ranks bit-identical: %s ; speed-up of the kernel %.2fx" % (diff, same, st_ref["kernel_ms"] / st["kernel_ms"]))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
