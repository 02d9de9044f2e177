"""Replica bookkeeping for multi-GPU runs (replicas only -- one process per GPU, disjoint sequences,
no collective on the data path).  torch.distributed is used for exactly two things: a start barrier and the reduction
of per-rank timings / counters into the whole-job figure (max time over ranks, units summed over ranks)."""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass
class ReplicaEnv:
    rank: int
    world: int
    local_rank: int

    @staticmethod
    def from_env() -> "ReplicaEnv":
        return ReplicaEnv(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def sequence_seed(base_seed: int, rank: int, index: int = 0, world: int = 1) -> int:
    """Sequence `index` of replica `rank`: sequence s of the job goes to GPU s mod N."""
    return base_seed + index * world + rank


def sequences_of_rank(n_sequences: int, rank: int, world: int = 1):
    """Indices of the job's sequences that replica `rank` serves (BASELINE.json configs[4]: 64 independent sequences over
    1/2/4/8 GPUs): sequence s goes to GPU s mod N, so the shares differ by at most one."""
    return list(range(rank, n_sequences, world))


def init(env: ReplicaEnv, backend: str = "nccl", device_id=None):
    if env.world <= 1:
        return None
    import torch.distributed as dist
    if not dist.is_initialized():
        kw = {"device_id": device_id} if device_id is not None else {}
        dist.init_process_group(backend, rank=env.rank, world_size=env.world, **kw)
    return dist


def barrier(env: ReplicaEnv):
    if env.world > 1:
        import torch.distributed as dist
        dist.barrier()


def aggregate(env: ReplicaEnv, ms_local: float, units_local: int, extra_max=(), extra_sum=(), device="cpu"):
    """Whole-job figures: (max over ranks of ms_local, sum over ranks of units_local, maxes of extra_max, sums of extra_sum)."""
    if env.world <= 1:
        return ms_local, units_local, list(extra_max), list(extra_sum)
    import torch
    import torch.distributed as dist
    mx = torch.tensor([ms_local, *extra_max], dtype=torch.float64, device=device)
    sm = torch.tensor([float(units_local), *[float(x) for x in extra_sum]], dtype=torch.float64, device=device)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return float(mx[0]), int(round(float(sm[0]))), [float(x) for x in mx[1:]], [float(x) for x in sm[1:]]


def throughput(units_total: int, ms_max: float) -> float:
    """units all ranks processed / the slowest rank's time (the job is done when the last replica is)."""
    return units_total / (ms_max * 1e-3)
