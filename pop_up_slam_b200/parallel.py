"""Multi-GPU plumbing for the one way this path shards: independent graphs (sequences, disjoint sub-graphs)
round-robined over ranks, one process per GPU, no data-path collective.  torch.distributed is
only used to gather the solved estimates / reduce the timing (NCCL on GPUs, gloo in the CPU test-suite).

One graph spanning several ranks (`span_optimize`): every rank builds the same graph; the PCG phases are split
over the CTAs of all ranks inside the persistent kernels, which exchange their vectors through CUDA-IPC-mapped peer
memory and meet at a cross-rank barrier -- torch.distributed only carries the 64-byte IPC handles (DESIGN.md
"Multi-GPU").
"""
import numpy as np


def shard(n_items, rank, world):
    """Indices of the graphs rank `rank` of `world` owns (round-robin, deterministic)."""
    return list(range(rank, n_items, world))


def solve_sharded(graphs, make_api, build, configure, rank=0, world=1, solve_many=None):
    """Build and solve this rank's share of `graphs`.  Returns {graph index: dict(poses, planes, iterations, chi2)}.

    make_api()                 -> a GraphAPI (pus_* on a GPU rank; the tests inject the CPU oracle)
    build(api, g) / configure  -> graphgen.build_bulk / graphgen.configure (or equivalents)
    solve_many(apis)           -> optional batched launch (capi.batch_optimize_many); default: one call per graph
    """
    mine = shard(len(graphs), rank, world)
    apis, ids = [], []
    for i in mine:
        api = make_api()
        ids.append(build(api, graphs[i]))
        configure(api, graphs[i])
        apis.append(api)
    if solve_many is not None and len(apis) > 1:
        its = list(solve_many(apis))
    else:
        its = [a.batch_optimize() for a in apis]
    out = {}
    for i, api, idd, it in zip(mine, apis, ids, its):
        out[i] = dict(poses=api.get_poses(idd["pose_ids"]), planes=api.get_planes(idd["plane_ids"]), iterations=int(it),
                      chi2=api.chi2())
    return out


def gather_solutions(local, world=1):
    """All ranks end up with every graph's solution (all_gather_object; a few hundred KB per graph)."""
    if world == 1:
        return dict(local)
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return merged


def reduce_throughput(units, seconds, world=1, device=None):
    """Whole-job throughput bookkeeping: sum of the units every rank processed and the max over ranks of its time."""
    if world == 1:
        return float(units), float(seconds)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()), float(t.item())


def span_optimize(api, rank=0, world=1):
    """batch_optimization() of ONE graph held identically by every rank (GpuGraphAPI `api` on this rank's GPU).
    Exchanges the arenas' CUDA IPC handles over torch.distributed, connects, solves; returns the LM iteration count.
    All ranks end with the same estimates."""
    if world == 1:
        return api.batch_optimize()
    import torch.distributed as dist
    mine = api.span_export()
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    api.span_connect(rank, world, handles)
    dist.barrier()
    try:
        it = api.span_optimize()
    finally:
        dist.barrier()
        api.span_disconnect()
    return it
