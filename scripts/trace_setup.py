"""One traced (B200_TRACE=1) distributed setup under torchrun: per-phase, per-level wall-clock on stderr."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import torch, torch.distributed as dist
import hypre_ve_b200 as hb
from hypre_ve_b200 import launch
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
h = hb.Handle(lr)
uid = launch.broadcast_bytes(hb.Comm.nccl_unique_id() if rank == 0 else b"", 0, 128, "cuda")
comm = hb.Comm.nccl(h, world, rank, uid)
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fx = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[world]
A = hb.DistMatrix.laplacian(h, comm, n1 * fx[0], n1 * fx[1], n1 * fx[2], 1, 1, world, 7)
prm = hb.Amg(h, ModuleRAP2=0)
for k in range(3):
    if k == 2:
        os.environ["B200_TRACE"] = "1"
    dist.barrier(); torch.cuda.synchronize()
    amg = hb.DistAmg(h, comm, prm, A)
    os.environ.pop("B200_TRACE", None)
    if rank == 0:
        print("setup_ms", amg.setup_ms, file=sys.stderr)
    amg.destroy()
dist.destroy_process_group()
