"""One process per GPU (torchrun): plumbing for the parts of the path that shard -- independent
chains and nsig candidates (R/Optimal_sigs.R:3-8,27 runs them as separate R processes).  There is
no data-path collective: ranks only meet at barriers, to take the max of their device times and to
gather the small per-task BIC vectors.  torch.distributed is plumbing here, nothing more."""
import os

import numpy as np


def env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


import contextlib
import sys


@contextlib.contextmanager
def stdout_to_stderr():
    """NCCL prints its version banner on stdout when the first communicator is created; callers
    that promise ONE line of JSON on stdout wrap communicator creation in this."""
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (nccl on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    rank, world, local = env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        with stdout_to_stderr():
            dist.init_process_group(backend, rank=rank, world_size=world, **kw)
            dist.barrier()                              # creates the communicator (and its banner) now
    return rank, world, local


def barrier():
    import torch
    import torch.distributed as dist
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(x):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def deal_tasks(costs, n_bins):
    """Longest-processing-time-first dealing of tasks onto n_bins devices/ranks (the same rule
    signer_gibbs_batch applies inside one process).  Returns a list of task-index lists."""
    order = sorted(range(len(costs)), key=lambda t: (-costs[t], t))
    load = [0.0] * n_bins
    bins = [[] for _ in range(n_bins)]
    for t in order:
        b = min(range(n_bins), key=lambda i: (load[i], i))
        bins[b].append(t)
        load[b] += costs[t]
    return bins


def task_cost(tk):
    """Measured on B200: a sweep costs about (12 + K) units -- the step-1 part barely depends on K."""
    return (12 + tk["K"]) * (tk["burn"] + tk["eval"])


def run_tasks_distributed(tasks, runner, cost=task_cost):
    """Every rank runs its share of `tasks` with runner(list_of_tasks) -> list_of_results and all
    ranks receive the results in task order (all_gather_object of small BIC vectors)."""
    import torch.distributed as dist
    rank, world, _ = env()
    bins = deal_tasks([cost(t) for t in tasks], world)
    mine = bins[rank]
    res = runner([tasks[t] for t in mine]) if mine else []
    pairs = list(zip(mine, res))
    if world > 1 and dist.is_initialized():
        gathered = [None] * world
        dist.all_gather_object(gathered, pairs)
        pairs = [p for g in gathered for p in g]
    out = [None] * len(tasks)
    for t, r in pairs:
        out[t] = r
    return out
