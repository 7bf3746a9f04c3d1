#!/usr/bin/env python3
"""bench.py -- BoomerAMG-PCG setup+solve seconds and SpMV HBM GB/s on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU build

A step = one full pass of the hot path on BASELINE config 2 (`ij -laplacian -n 256 256 256
-solver 1 -pmis -interptype 6 -Pmx 4 -rlx 18 -mod_rap2 1`, rhs = 1, x0 = 0, tol 1e-8):
BoomerAMG setup + PCG solve.  Prints ONE JSON line (rank 0).

  value        device-timed setup+solve seconds per step, matrix/rhs resident in HBM (CUDA events)
  e2e          the same through the C-ABI with HOST CSR/rhs buffers: H2D copy of A and b, setup,
               solve, D2H copy of x inside the timed region
  roofline     dominant kernel = the streaming CSR SpMV family (SpMV / residual / l1-Jacobi / P / R):
               timed live as a standalone y=A0*x loop (the `ij -solver -1` analogue),
               algorithmic bytes 12*nnz + 4*(N+1) + 16*N / CUDA-event time
  cpu_baseline oracle/_ref (the reference compiled in place) timed on the host cores

Multi-GPU (N>1): one process per GPU under torchrun; rows partitioned over a P x Q x R process grid
like `ij -P` (1x1x1, 2x1x1, 2x2x1, 2x2x2), weak scaling with 256^3 unknowns per GPU (config 5 at
N=8: 512^3).  Halo exchange and Krylov reductions go over NCCL/NVLink inside libhypre_b200.so; torch
only launches the ranks and broadcasts the NCCL unique id.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N1 = 256                       # config 2 grid edge
# the driver's own flags for config 2, `ij -n 256 256 256 -solver 1 -pmis -rlx 18`: Galerkin product in the library's
# default fused order (ModuleRAP2 0; 22 iterations at 256^3), on one GPU and across ranks alike
REF_ARGS = ["-pmis", "-rlx", "18", "-keepT", "1", "-nodump"]
REF_ARGS_DIST = REF_ARGS
WORKLOAD = ("ij 3D 7-pt Laplacian 256^3 BoomerAMG-PCG, PMIS + ext+i(Pmx 4) interp + l1-Jacobi, tol 1e-8 "
            "(ij -n 256 256 256 -solver 1 -pmis -rlx 18)")


def spmv_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (None if absent)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "spmv_traffic.json")))["traffic_bytes_per_launch"]
    except Exception:
        return None


_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line.  Native libraries print there too (NCCL's version banner, for one),
    so file descriptor 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k] == "Active"})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(n1, threads, dist_flags=False):
    """One setup+solve of the reference CPU build; returns (setup_s, solve_s, iterations)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/ref_dump missing: run __graft_entry__.build() where /root/reference exists")
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false")
    out = subprocess.run([exe, "-n", str(n1), str(n1), str(n1)] + (REF_ARGS_DIST if dist_flags else REF_ARGS), env=env, capture_output=True, text=True,
                         check=True).stdout
    m = re.search(r"iterations=(\d+) relres=(\S+) setup_s=(\S+) solve_s=(\S+)", out)
    return float(m.group(3)), float(m.group(4)), int(m.group(1))


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    # bounded sample: probe with 128^3, then use the full 256^3 workload only if K+W steps fit ~4 minutes
    df = a.gpus > 1
    s0, v0, _ = run_reference(128, threads, df)
    est_full = 8.5 * (s0 + v0)
    n1 = N1 if est_full * (a.steps + a.warmup) < 240 else 128
    for _ in range(a.warmup):
        run_reference(n1, threads, df)
    t_set = t_sol = 0.0
    its = 0
    for _ in range(a.steps):
        s, v, its = run_reference(n1, threads, df)
        t_set += s
        t_sol += v
    # AMG-PCG work is linear in the number of unknowns; the N-GPU arm is weak-scaled (256^3 unknowns per GPU),
    # so the same job on the host is N x 256^3 unknowns
    scale = (N1 / n1) ** 3 * max(1, a.gpus)
    per_step = (t_set + t_sol) / a.steps * scale
    sample = "%d^3 run per step on all host threads, seconds scaled by %g (= unknown ratio) to the %d x 256^3 workload" % (
        n1, scale, max(1, a.gpus))
    if n1 == N1 and a.gpus <= 1:
        sample = "full workload (256^3) per step"
    line = {
        "impl": "reference", "metric": "boomeramg_pcg_setup_plus_solve_seconds", "value": per_step, "unit": "s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD if a.gpus <= 1 else WORKLOAD.replace("256^3", "%d x 256^3 (weak-scaled)" % a.gpus),
                   "impl": "reference hypre 2.20 (SX-Aurora fork) CPU path, OpenMP, sequential MPI stubs"},
        "setup_s": t_set / a.steps * scale, "solve_s": t_sol / a.steps * scale, "iterations": its,
        "cpu_baseline": {"value": per_step, "unit": "s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": per_step, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def main_dist(a, rank, world, local_rank):
    """N > 1: row-partitioned BoomerAMG-PCG, 256^3 unknowns per GPU (weak scaling)."""
    import torch
    import torch.distributed as dist
    import hypre_ve_b200 as hb
    from hypre_ve_b200 import launch

    try:
        P, Q, R = launch.process_grid(world)
    except ValueError as e:
        raise SystemExit(str(e))
    n1 = a.n
    nx, ny, nz = n1 * P, n1 * Q, n1 * R
    h = hb.Handle(local_rank)
    # NCCL communicator of the library: rank 0 creates the id, torch.distributed broadcasts it
    uid = launch.broadcast_bytes(hb.Comm.nccl_unique_id() if rank == 0 else b"", 0, 128, "cuda")
    comm = hb.Comm.nccl(h, world, rank, uid)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    A = hb.DistMatrix.laplacian(h, comm, nx, ny, nz, P, Q, R, 7)
    inf = A.info
    n, nnz = inf["local_rows"], inf["local_nnz"]
    b = A.vector(1.0)
    x = A.vector(0.0)
    prm = hb.Amg(h, ModuleRAP2=0)          # driver default: fused BuildCoarseOperatorKT order, (R A) P

    def step():
        amg = hb.DistAmg(h, comm, prm, A)
        s_ms = amg.setup_ms
        h.fill(x, 0.0)
        h.timer_start()
        its, rel, _ = hb.dist_pcg(h, comm, A, amg, b, x, tol=1e-8, max_iter=100)
        v_ms = h.timer_stop_ms()
        amg.destroy()
        return s_ms, v_ms, its, rel

    for _ in range(a.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = h.launch_count()
    barrier()
    t_set = t_sol = 0.0
    for _ in range(a.steps):
        s_ms, v_ms, its, rel = step()
        t_set += s_ms
        t_sol += v_ms
    barrier()
    launches = h.launch_count() - launches0
    # end to end: the operator is generated on the device (no host matrix exists at 512^3); the host
    # buffers of this path are the rhs (H2D) and the solution (D2H) of every rank
    hb_host = torch.ones(n, dtype=torch.float64).pin_memory().numpy()
    hx_host = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    e2e_ms = 0.0
    for k in range(a.steps + 1):
        barrier()
        h.timer_start()
        hb._chk(hb._lib.b200_memcpy_h2d(h.p, b.ptr, hb._np_ptr(hb_host), hb_host.nbytes))
        amg = hb.DistAmg(h, comm, prm, A)
        h.fill(x, 0.0)
        hb.dist_pcg(h, comm, A, amg, b, x, tol=1e-8, max_iter=100)
        hb._chk(hb._lib.b200_memcpy_d2h(h.p, hb._np_ptr(hx_host), x.ptr, hx_host.nbytes))
        ms = h.timer_stop_ms()
        amg.destroy()
        if k > 0:
            e2e_ms += ms
    # distributed SpMV sweep (ij -solver -1 analogue): halo exchange + one kernel per repetition
    y = h.zeros(n)
    for _ in range(5):
        A.matvec(1.0, b, 0.0, None, y)
    barrier()
    reps = 100
    h.timer_start()
    for _ in range(reps):
        A.matvec(1.0, b, 0.0, None, y)
    spmv_ms = h.timer_stop_ms() / reps
    sampler.stop_flag = True
    sampler.join()
    per = launch.reduce_over_ranks([t_set / a.steps, t_sol / a.steps, e2e_ms / a.steps, spmv_ms], "max", "cuda")
    gn, gnnz = launch.reduce_over_ranks([float(n), float(nnz)], "sum", "cuda")
    set_s, sol_s, e2e_s, spmv_ms = per[0] / 1e3, per[1] / 1e3, per[2] / 1e3, per[3]
    spmv_bytes = 12.0 * gnnz + 4.0 * (gn + world) + 16.0 * gn
    peak, peak_src = peaks()
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    line = {
        "metric": "boomeramg_pcg_setup_plus_solve_seconds", "value": set_s + sol_s, "unit": "s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": (set_s + sol_s) * 1e3,
        "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ij 3D 7-pt Laplacian %dx%dx%d (-P %d %d %d, %d^3 per GPU) BoomerAMG-PCG, PMIS + ext+i(Pmx 4) + "
                               "l1-Jacobi, tol 1e-8" % (nx, ny, nz, P, Q, R, n1),
                   "rows_per_gpu": n, "nnz_per_gpu": nnz, "global_rows": int(gn), "global_nnz": int(gnnz),
                   "parallelism": "row-partitioned ParCSR, %d GPUs, halo + allreduce over NCCL" % world,
                   "l2": "per-GPU operator (1.7 GB) larger than the 126 MB L2; no flush needed"},
        "setup_s": set_s, "solve_s": sol_s, "iterations": its, "final_rel_res": rel,
        "spmv_gbs": achieved, "spmv_ms": spmv_ms,
        "roofline": {"bound": "hbm", "kernel": "halo exchange + spmv_pipe_kernel (y = A0*x, %d^3 per GPU)" % n1,
                     "achieved": achieved, "peak": peak * world, "peak_source": peak_src + " x n_gpus", "unit": "GB/s",
                     "frac": achieved / (peak * world), "algorithmic_bytes_per_launch": spmv_bytes / world, "traffic": None},
        "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": hb_host.nbytes * world, "d2h_bytes_per_step": hx_host.nbytes * world},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if rank == 0:
        emit(line)
    A.destroy()
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--edge", dest="n", type=int, default=N1, help="grid edge per GPU (default: config 2, 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    claim_stdout()
    if a.impl == "reference":
        return reference_arm(a)

    import torch
    import torch.distributed as dist
    import hypre_ve_b200 as hb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        return main_dist(a, rank, world, local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    h = hb.Handle(local_rank)
    n1 = a.n
    A = hb.ParCsr.laplacian(h, n1, n1, n1)          # synthetic config-2 operator, built on the device
    n, nnz, _, _ = A.local
    b = h.zeros(n)
    h.fill(b, 1.0)
    x = h.zeros(n)
    # host copies for the end-to-end leg (pinned)
    hi, hj, ha = A.diag.download()
    t_i, t_j, t_a = (torch.from_numpy(v).pin_memory() for v in (hi, hj, ha))
    t_b = torch.ones(n, dtype=torch.float64).pin_memory()
    t_x = torch.empty(n, dtype=torch.float64).pin_memory()
    hi, hj, ha, hb_, hx = (t.numpy() for t in (t_i, t_j, t_a, t_b, t_x))
    h2d = hi.nbytes + hj.nbytes + ha.nbytes + hb_.nbytes
    d2h = hx.nbytes

    def step_resident():
        amg = hb.Amg(h, ModuleRAP2=0)          # driver default: fused BuildCoarseOperatorKT order, (R A) P
        h.timer_start()
        amg.setup(A)
        s_ms = h.timer_stop_ms()
        h.fill(x, 0.0)
        h.timer_start()
        its, rel, _ = h.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
        v_ms = h.timer_stop_ms()
        ph = amg.setup_times()
        amg.destroy()
        return s_ms, v_ms, its, rel, ph

    def step_e2e():
        h.timer_start()
        A2 = hb.ParCsr.from_host(h, hi, hj, ha)
        b2 = h.array(hb_)
        x2 = h.zeros(n)
        amg = hb.Amg(h, ModuleRAP2=0)          # driver default: fused BuildCoarseOperatorKT order, (R A) P
        amg.setup(A2)
        its, rel, _ = h.pcg(A2, amg, b2, x2, tol=1e-8, max_iter=100)
        hb._chk(hb._lib.b200_memcpy_d2h(h.p, hb._np_ptr(hx), x2.ptr, hx.nbytes))
        ms = h.timer_stop_ms()
        amg.destroy(); A2.destroy(); b2.free(); x2.free()
        return ms, its

    for _ in range(a.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = h.launch_count()
    barrier()
    t_set = t_sol = 0.0
    phases = np.zeros(8)
    for _ in range(a.steps):
        s_ms, v_ms, its, rel, ph = step_resident()
        t_set += s_ms
        t_sol += v_ms
        phases += np.array(ph)
    barrier()
    launches = h.launch_count() - launches0
    # end to end (host buffers)
    step_e2e()
    barrier()
    e2e_ms = 0.0
    for _ in range(a.steps):
        ms, its_e = step_e2e()
        e2e_ms += ms
    barrier()
    # SpMV roofline leg: 100 x (y = A0 x), the `ij -solver -1` loop (test/ij.c:3206-3243); the 256^3
    # operands (1.7 GB) exceed the 126 MB L2, so every repetition streams from HBM
    y = h.zeros(n)
    for _ in range(5):
        A.matvec(1.0, b, 0.0, None, y)
    reps = 100
    h.timer_start()
    for _ in range(reps):
        A.matvec(1.0, b, 0.0, None, y)
    spmv_ms = h.timer_stop_ms() / reps
    sampler.stop_flag = True
    sampler.join()

    per = torch.tensor([t_set / a.steps, t_sol / a.steps, e2e_ms / a.steps, spmv_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(per, op=dist.ReduceOp.MAX)
    set_s, sol_s, e2e_s, spmv_ms = per[0].item() / 1e3, per[1].item() / 1e3, per[2].item() / 1e3, per[3].item()
    spmv_bytes = 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n
    peak, peak_src = peaks()
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    line = {
        "metric": "boomeramg_pcg_setup_plus_solve_seconds", "value": set_s + sol_s, "unit": "s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": (set_s + sol_s) * 1e3,
        "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD if n1 == N1 else WORKLOAD.replace("256^3", "%d^3" % n1),
                   "rows_per_gpu": n, "nnz_per_gpu": nnz,
                   "parallelism": "1 GPU" if world == 1 else "replicas only (row-partitioned multi-GPU path not built yet)",
                   "l2": "inputs (1.7 GB operator) larger than the 126 MB L2; no flush needed"},
        "setup_s": set_s, "solve_s": sol_s, "iterations": its, "final_rel_res": rel,
        "setup_phases_ms": dict(zip(["strength", "pmis", "interp", "trunc", "transpose", "rap", "l1_alloc", "total"],
                                    (phases / a.steps).round(3).tolist())),
        "spmv_gbs": achieved, "spmv_ms": spmv_ms,
        "roofline": {"bound": "hbm", "kernel": "spmv_pipe_kernel<1,2> (y = A0*x, 256^3 7-pt)", "achieved": achieved,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes_per_launch": spmv_bytes, "traffic": spmv_traffic()},
        "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            threads = host_threads()
            s, v, cits = run_reference(128, threads)
            line["cpu_baseline"] = {"value": (s + v) * 8.0, "unit": "s", "cores": threads, "kind": "reference",
                                    "sample": "oracle/_ref/ref_dump, 128^3 sample (1/8 of the unknowns), seconds scaled by 8",
                                    "setup_s_sample": s, "solve_s_sample": v, "iterations_sample": cits}
        except Exception as e:      # the baseline is reporting only; never fail the GPU line for it
            line["cpu_baseline"] = {"value": None, "unit": "s", "cores": host_threads(), "kind": "reference",
                                    "sample": "unavailable: %s" % e}
    if rank == 0:
        emit(line)
    A.destroy()
    h.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
