"""Host-side helpers for the one-process-per-GPU launch (bench.py, scripts/dist_run.py).

Nothing here touches the data path: the halo exchanges and Krylov reductions live inside libhypre_b200.so
(NCCL over NVLink).  torch.distributed is used for exactly three things -- carrying the library's 128-byte
NCCL unique id from rank 0 to the other ranks, the barrier around the timed region, and the max / sum of the
per-rank timings -- and these work the same over `gloo` on CPUs, which is how tests/test_launch_gloo.py covers
them with world_size 2.  The process-grid arithmetic mirrors the reference driver and generator."""
import numpy as np

# ranks -> P x Q x R process grid, the `-P` arguments the bench uses (1x1x1, 2x1x1, 2x2x1, 2x2x2)
GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def process_grid(world):
    if world not in GRIDS:
        raise ValueError("supported GPU counts: 1, 2, 4, 8")
    return GRIDS[world]


def rank_coords(rank, P, Q, R):
    """(p, q, r) of a rank on the process grid (test/ij.c:7785-7787)"""
    p = rank % P
    q = ((rank - p) // P) % Q
    r = (rank - p - P * q) // (P * Q)
    return p, q, r


def partitioning(n, parts):
    """hypre_GeneratePartitioning (parcsr_ls/par_laplace.c via utilities): first n % parts pieces get one extra"""
    size, rest = n // parts, n % parts
    starts = [0]
    for k in range(parts):
        starts.append(starts[-1] + size + (1 if k < rest else 0))
    return starts


def local_box(rank, dims, grid):
    """the (nx, ny, nz) box a rank owns and its first global row: ranks own contiguous row blocks, numbered
    rank by rank, each box in x-fastest order (par_laplace.c:66-101 + hypre_map)"""
    (nx, ny, nz), (P, Q, R) = dims, grid
    px, py, pz = partitioning(nx, P), partitioning(ny, Q), partitioning(nz, R)
    sizes = []
    for k in range(P * Q * R):
        p, q, r = rank_coords(k, P, Q, R)
        sizes.append((px[p + 1] - px[p]) * (py[q + 1] - py[q]) * (pz[r + 1] - pz[r]))
    p, q, r = rank_coords(rank, P, Q, R)
    box = (px[p + 1] - px[p], py[q + 1] - py[q], pz[r + 1] - pz[r])
    return box, int(np.sum(sizes[:rank]))


def broadcast_bytes(payload, src, nbytes, device):
    """carry `nbytes` bytes (the NCCL unique id of the library's communicator) from rank `src` to every rank"""
    import torch
    import torch.distributed as dist
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def reduce_over_ranks(values, op, device):
    """elementwise max ('max') or sum ('sum') of a list of floats over all ranks (device timings: max)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return [float(x) for x in t.cpu()]
