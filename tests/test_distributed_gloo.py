"""world_size-2 test (gloo, CPU) of the only multi-GPU path this hot loop has: independent graphs sharded across
ranks with no data-path collective; the gather / throughput reductions are the plumbing bench.py uses.  The solver
injected here is the CPU oracle (tests may use it); on GPUs the same functions run with GpuGraphAPI + NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_api import OracleAPI
    from pop_up_slam_b200 import graphgen as gg, parallel

    graphs = [gg.make_config(1, seed=s) for s in range(5)]

    def make():
        a = OracleAPI()
        a.set_jacobian_mode(1)
        return a

    local = parallel.solve_sharded(graphs, make, gg.build_bulk, gg.configure, rank, world)
    assert sorted(local) == parallel.shard(len(graphs), rank, world)
    merged = parallel.gather_solutions(local, world)
    units, secs = parallel.reduce_throughput(sum(v["iterations"] for v in local.values()), 1.0 + rank, world)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), keys=sorted(merged), units=units, secs=secs,
             **{f"poses{k}": v["poses"] for k, v in merged.items()}, **{f"chi2_{k}": v["chi2"] for k, v in merged.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_api import OracleAPI
    from pop_up_slam_b200 import graphgen as gg, parallel
    graphs = [gg.make_config(1, seed=s) for s in range(5)]

    def make():
        a = OracleAPI()
        a.set_jacobian_mode(1)
        return a

    ref = parallel.solve_sharded(graphs, make, gg.build_bulk, gg.configure, 0, 1)
    total_it = sum(v["iterations"] for v in ref.values())
    for r in range(world):
        d = np.load(os.path.join(tmp_path, f"rank{r}.npz"))
        assert d["keys"].tolist() == [0, 1, 2, 3, 4]                 # every rank holds every solution after the gather
        assert float(d["units"]) == total_it and float(d["secs"]) == 2.0   # sum of units, max of times
        for k in range(5):
            assert np.array_equal(d[f"poses{k}"], ref[k]["poses"])
            assert float(d[f"chi2_{k}"]) == ref[k]["chi2"]
    assert parallel.shard(7, 1, 3) == [1, 4]
