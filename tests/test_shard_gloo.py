"""N > 1 host logic on CPU: sequence sharding and the max-over-ranks / sum-over-ranks reduction bench.py reports,
exercised with a world of 2 over gloo (127.0.0.1 rendezvous)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vins_mono_b200 import shard  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seqs = shard.sequences_of_rank(rank, world, 3)
        # rank 0: 30 frames in 100 ms, rank 1: 30 frames in 150 ms -> 60 frames / 0.150 s
        rate, sec, cnt = shard.aggregate_rate(100.0 + 50.0 * rank, 30)
        dist.barrier()
        out.put((rank, seqs, rate, sec, cnt))
    finally:
        dist.destroy_process_group()


def test_sequences_of_rank_partition():
    world, spg = 8, 64
    owned = [shard.sequences_of_rank(r, world, spg) for r in range(world)]
    flat = [s for o in owned for s in o]
    assert flat == list(range(world * spg))            # disjoint, complete, ordered
    with pytest.raises(ValueError):
        shard.sequences_of_rank(8, 8)


def test_aggregate_without_process_group():
    rate, sec, cnt = shard.aggregate_rate(200.0, 40)
    assert rate == pytest.approx(200.0) and sec == pytest.approx(0.2) and cnt == 40


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=90) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4, 5]
    for _, _, rate, sec, cnt in res:                   # every rank sees the same whole-job figure
        assert cnt == 60 and sec == pytest.approx(0.150) and rate == pytest.approx(400.0)
