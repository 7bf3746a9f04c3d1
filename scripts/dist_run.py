"""Multi-GPU runs of the other BASELINE.json configs (parity-test cases, not bench lines), under torchrun:
  config 3: --stencil 27 --global-edge 256 --relax 13 --gs-blocks T     (strong scaling: 256^3 split over N GPUs)
  config 4: --global-edge 384 --agg-nl 1 --aniso 0.001                  (anisotropic diffusion, aggressive coarsening), or as written
            --global-edge 384 --agg-nl 1 --difconv --solver 3          (ij -difconv: nonsymmetric, AMG-GMRES)
  config 5: --stencil 7 --edge-per-gpu 256 --spmv-sweep                 (weak scaling + SpMV bandwidth sweep)
Prints one JSON line on rank 0."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import hypre_ve_b200 as hb

GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
ap = argparse.ArgumentParser()
ap.add_argument("--stencil", type=int, default=7)
ap.add_argument("--global-edge", type=int, default=0)
ap.add_argument("--edge-per-gpu", type=int, default=0)
ap.add_argument("--relax", type=int, default=18)
ap.add_argument("--gs-blocks", type=int, default=1)
ap.add_argument("--spmv-sweep", action="store_true")
ap.add_argument("--agg-nl", type=int, default=0)
ap.add_argument("--aniso", type=float, default=1.0, help="cz of -c 1 1 cz (config 4: 0.001)")
ap.add_argument("--difconv", action="store_true", help="ij -difconv: GenerateDifConv with -c 1 1 aniso, --conv a a a, --atype")
ap.add_argument("--conv", type=float, default=1.0)
ap.add_argument("--atype", type=int, default=0)
ap.add_argument("--solver", type=int, default=1, help="1 AMG-PCG, 3 AMG-GMRES(5), 9 AMG-BiCGSTAB (ij -solver)")
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
P, Q, R = GRIDS[world]
h = hb.Handle(lr)
if world > 1:
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(hb.Comm.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    comm = hb.Comm.nccl(h, world, rank, bytes(idt.cpu().numpy().tobytes()))
else:
    comm = hb.Comm.single(h)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def allmax(v):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def allsum(v):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.item()


out = {"n_gpus": world, "grid": [P, Q, R], "stencil": a.stencil}
if a.global_edge:
    dims = (a.global_edge,) * 3
else:
    e = a.edge_per_gpu or 256
    dims = (e * P, e * Q, e * R)
out["global_dims"] = list(dims)
if a.difconv:
    A = hb.DistMatrix.difconv(h, comm, *dims, P, Q, R, c=(1.0, 1.0, a.aniso), a=(a.conv,) * 3, atype=a.atype)
    out.update(problem="difconv", a=[a.conv] * 3, atype=a.atype)
else:
    A = hb.DistMatrix.laplacian(h, comm, *dims, P, Q, R, a.stencil, c=(1.0, 1.0, a.aniso))
inf = A.info
n, nnz = inf["local_rows"], inf["local_nnz"]
if a.spmv_sweep:
    x, y = A.vector(1.0), h.zeros(n)
    for _ in range(5):
        A.matvec(1.0, x, 0.0, None, y)
    barrier(); h.timer_start()
    for _ in range(100):
        A.matvec(1.0, x, 0.0, None, y)
    ms = allmax(h.timer_stop_ms() / 100)
    gn, gnnz = allsum(float(n)), allsum(float(nnz))
    byts = 12.0 * gnnz + 20.0 * gn
    out.update(spmv_ms=ms, spmv_gbs=byts / ms / 1e6, algorithmic_gb=byts / 1e9)
else:
    prm = hb.Amg(h, RelaxType=a.relax, RelaxTypeUp=(14 if a.relax == 13 else a.relax), GSBlocks=a.gs_blocks, AggNumLevels=a.agg_nl,
                 ModuleRAP2=0)
    b, x = A.vector(1.0), A.vector(0.0)
    res = []
    for k in range(a.steps + 1):
        barrier()
        amg = hb.DistAmg(h, comm, prm, A)
        s_ms = amg.setup_ms
        h.fill(x, 0.0)
        h.timer_start()
        if a.solver == 3:
            its, rel, _, _ = hb.dist_gmres(h, comm, A, amg, b, x, tol=1e-8, max_iter=200)
        elif a.solver == 9:
            its, rel, _, _ = hb.dist_bicgstab(h, comm, A, amg, b, x, tol=1e-8, max_iter=200)
        else:
            its, rel, _ = hb.dist_pcg(h, comm, A, amg, b, x, tol=1e-8, max_iter=200)
        v_ms = h.timer_stop_ms()
        nl = amg.num_levels
        amg.destroy()
        if k:
            res.append((allmax(s_ms), allmax(v_ms)))
    out.update(solver=a.solver, agg_nl=a.agg_nl, c=[1.0, 1.0, a.aniso], relax=a.relax, gs_blocks_per_gpu=a.gs_blocks, levels=nl, iterations=its, final_rel_res=rel,
               setup_s=sum(r[0] for r in res) / len(res) / 1e3, solve_s=sum(r[1] for r in res) / len(res) / 1e3)
if rank == 0:
    print(json.dumps(out))
A.destroy()
if world > 1:
    dist.destroy_process_group()
