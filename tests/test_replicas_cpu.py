"""Multi-GPU = independent replicas.  The only cross-rank logic is the start barrier and the reduction of
timings/counters; it is exercised here with the gloo backend, world size 2, on CPU."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch.multiprocessing as mp

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))     # spawned workers re-import this module without conftest
import gtb  # noqa: E402,F401
from tinyllama_cpp_b200 import replicas as R  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    env = R.ReplicaEnv.from_env()
    dist = R.init(env, backend="gloo")
    R.barrier(env)
    ms_local = 100.0 + 50.0 * rank             # rank 1 is the slow replica
    ms, units, mx, sm = R.aggregate(env, ms_local, 256, extra_max=[ms_local * 2], extra_sum=[1])
    q.put((rank, ms, units, mx, sm, R.sequence_seed(7, rank, index=3, world=world)))
    dist.barrier()
    dist.destroy_process_group()


def test_replica_aggregation_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, units, mx, sm, seed in out:
        assert ms == 150.0 and units == 512 and mx == [300.0] and sm == [2.0]
        assert seed == 7 + 3 * 2 + rank          # disjoint sequences per replica
    assert R.throughput(512, 150.0) == pytest.approx(512 / 0.150)


def test_sequence_partition_is_disjoint_and_complete():
    for world in (1, 2, 4, 8, 3):
        parts = [R.sequences_of_rank(64, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(64))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert R.sequences_of_rank(3, 5, 8) == []            # more GPUs than sequences: the extra replicas idle


def test_single_replica_needs_no_process_group():
    env = R.ReplicaEnv(0, 1, 0)
    assert R.init(env) is None
    assert R.aggregate(env, 12.5, 3) == (12.5, 3, [], [])
