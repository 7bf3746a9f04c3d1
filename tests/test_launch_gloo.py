"""CPU tests (world_size 2, gloo) of the host-side launch logic used by bench.py --gpus N: the NCCL-id
hand-off, the max/sum over ranks, and the process-grid / row-partition arithmetic that must agree with the
reference generator (par_laplace.c:66-101, ij.c:7785-7787)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hypre_ve_b200 import launch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        payload = bytes(range(128)) if rank == 0 else b""
        got = launch.broadcast_bytes(payload, 0, 128, "cpu")
        mx = launch.reduce_over_ranks([10.0 + rank, 5.0 - rank], "max", "cpu")
        sm = launch.reduce_over_ranks([float(rank + 1)], "sum", "cpu")
        grid = launch.process_grid(world)
        box, first = launch.local_box(rank, (7, 5, 3), grid)
        dist.barrier()
        q.put((rank, got == bytes(range(128)), mx, sm, box, first))
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, mx, sm, box, first in res:
        assert ok                                    # every rank holds rank 0's 128 bytes
        assert mx == [11.0, 5.0] and sm == [3.0]     # max / sum over the two ranks
    # -P 2 1 1 on a 7 x 5 x 3 grid: x is split 4 + 3, rows are numbered rank by rank
    assert res[0][4] == (4, 5, 3) and res[0][5] == 0
    assert res[1][4] == (3, 5, 3) and res[1][5] == 60


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_process_grid_covers_the_domain(world):
    P, Q, R = launch.process_grid(world)
    assert P * Q * R == world
    dims = (13, 9, 6)
    seen = set()
    total = 0
    firsts = []
    for rank in range(world):
        p, q, r = launch.rank_coords(rank, P, Q, R)
        assert 0 <= p < P and 0 <= q < Q and 0 <= r < R and (p, q, r) not in seen
        seen.add((p, q, r))
        box, first = launch.local_box(rank, dims, (P, Q, R))
        firsts.append(first)
        total += box[0] * box[1] * box[2]
    assert total == dims[0] * dims[1] * dims[2]
    assert firsts == sorted(firsts) and firsts[0] == 0


def test_partitioning_matches_the_reference_rule():
    assert launch.partitioning(10, 3) == [0, 4, 7, 10]      # 10 = 4 + 3 + 3: the first n % parts pieces are longer
    assert launch.partitioning(8, 2) == [0, 4, 8]
    assert launch.partitioning(2, 4) == [0, 1, 2, 2, 2]
