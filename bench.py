#!/usr/bin/env python3
"""bench.py -- BoomerAMG-PCG setup+solve seconds and SpMV HBM GB/s on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU build

A step = one full pass of the hot path on BASELINE config 2 (`ij -laplacian -n 256 256 256
-solver 1 -pmis -rlx 18`: PMIS, ext+i with Pmx 4, l1-Jacobi, the library's default fused Galerkin
order; rhs = 1, x0 = 0, tol 1e-8): BoomerAMG setup + PCG solve.  Prints ONE JSON line (rank 0).

  value        device-timed setup+solve seconds per step, matrix/rhs resident in HBM (CUDA events)
  e2e          the same through the C-ABI with HOST CSR/rhs buffers: H2D copy of A and b, setup,
               solve, D2H copy of x inside the timed region
  roofline     dominant kernel = the streaming CSR SpMV family (SpMV / residual / l1-Jacobi / P / R):
               timed live as a standalone y=A0*x loop (the `ij -solver -1` analogue),
               algorithmic bytes 12*nnz + 4*(N+1) + 16*N / CUDA-event time
  roofline_solve  algorithmic bytes of ONE PCG iteration computed from the actual hierarchy (formula in
               solve_bytes_per_iteration below = SURVEY.md 8d per operator application) / device time per iteration
  cpu_baseline the UNMODIFIED reference driver oracle/_ref/ij (the reference compiled in place), same flags,
               all host cores, "wall clock time" lines of test/ij.c:4301-4314

Multi-GPU (N>1): one process per GPU under torchrun; weak scaling with 256^3 unknowns per GPU on the global
grids 256x256x512 / 256x512x512 / 512^3 (config 5) for N = 2 / 4 / 8, rows partitioned like `ij -P 1 1 N`
(z-slabs: the global numbering stays lexicographic, so hierarchy and iteration count are comparable with -- and in
the tests bit-identical to -- the reference's np = 1 run of the same grid; `--grid box` selects 2x1x1 / 2x2x1 /
2x2x2 boxes instead).  Halo exchange and Krylov reductions run inside libhypre_b200.so over NVLink; torch only
launches the ranks and carries the communicator id.
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
REF_ARGS = ["-solver", "1", "-pmis", "-rlx", "18", "-keepT", "1"]
WORKLOAD = ("ij 3D 7-pt Laplacian 256^3 BoomerAMG-PCG, PMIS + ext+i(Pmx 4) interp + l1-Jacobi, tol 1e-8 "
            "(ij -n 256 256 256 -solver 1 -pmis -rlx 18)")


def global_dims(world, n1, grid):
    """global grid and process grid of the weak-scaled job: n1^3 unknowns per GPU"""
    if grid == "box":
        from hypre_ve_b200 import launch
        P, Q, R = launch.process_grid(world)
        return (n1 * P, n1 * Q, n1 * R), (P, Q, R)
    fx = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(world, (1, 1, world))
    return (n1 * fx[0], n1 * fx[1], n1 * fx[2]), (1, 1, world)


def reference_iterations(dims):
    """PCG iteration count of the reference CPU build (np = 1) on this grid, from the committed table
    tests/golden/reference_iterations.json (made by tests/golden/make_reference_iterations.py); None if not recorded"""
    try:
        tab = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_iterations.json")))["counts"]
        return tab.get("%d %d %d" % tuple(dims))
    except Exception:
        return None


def solve_bytes_per_iteration(levels):
    """Algorithmic HBM bytes of one AMG-PCG iteration (every input array read once, every output written
    once; FP64 values, int32 indices).  levels = [(N_l, nnz(A_l), nnz(P_l))...] (nnz(P) = 0 on the coarsest level).
      per level l < L-1 of the V(1,1) cycle:
        pre-smooth from a zero iterate   u = w f / l1                      24 N
        residual  r = f - A u                                             12 nnzA + 4 N + 24 N
        restrict  f_c = R r  (R = P^T stored)                             12 nnzP + 4 Nc + 8 N + 8 Nc
        prolong   u += P e                                                12 nnzP + 4 N + 8 Nc + 16 N
        post-smooth u = u + w (f - A u) / l1  (fused l1-Jacobi)           12 nnzA + 4 N + 32 N
      coarsest level: dense Gaussian elimination on <= 9 unknowns (counted as 0)
      PCG around it: s = A p (12 nnzA0 + 4 N + 16 N), <s,p> 16 N, x += a p 24 N, r -= a s 24 N, <r,z> and <r,r> 24 N,
        p = z + b p 24 N"""
    total = 0.0
    for l, (n, nnz_a, nnz_p) in enumerate(levels[:-1]):
        nc = levels[l + 1][0]
        total += 24.0 * n + (12.0 * nnz_a + 28.0 * n) + (12.0 * nnz_p + 12.0 * nc + 8.0 * n) + (12.0 * nnz_p + 20.0 * n + 8.0 * nc) \
            + (12.0 * nnz_a + 36.0 * n)
    n0, nnz0, _ = levels[0]
    total += 12.0 * nnz0 + 20.0 * n0 + (16.0 + 24.0 + 24.0 + 24.0 + 24.0) * n0
    return total


def setup_bytes(levels):
    """Compulsory HBM bytes of the setup phases, from the actual hierarchy (every input array read once, every
    output written once; FP64 values, int32 indices).  levels = [(N_l, nnz(A_l), nnz(P_l))...].  nnz(S_l) is bounded by
    nnz(A_l) - N_l (no diagonal); the formulas are LOWER bounds of what an ideal one-pass kernel would move:
      strength   per level: read A (12 nnzA + 4 N), write S (4 nnzS + 4 N)
      pmis       per level: one sweep over S (4 nnzS + 4 N) + measures / markers (20 N)
      interp     per level: read A (12 nnzA + 4 N), S (4 nnzS + 4 N), CF (4 N); write P (12 nnzP + 4 N)      [ext+i + truncation]
      transpose  per level: read P (12 nnzP + 4 N), write R = P^T (12 nnzP + 4 Nc)
      rap        per level: read R, A, P once (12 (nnzP + nnzA + nnzP) + 4 (Nc + 2 N)), write A_{l+1} (12 nnzA' + 4 Nc)"""
    out = {"strength": 0.0, "pmis": 0.0, "interp": 0.0, "transpose": 0.0, "rap": 0.0}
    for l, (n, za, zp) in enumerate(levels[:-1]):
        nc, zc = levels[l + 1][0], levels[l + 1][1]
        zs = max(za - n, 0)
        out["strength"] += 12.0 * za + 4.0 * n + 4.0 * zs + 4.0 * n
        out["pmis"] += 4.0 * zs + 4.0 * n + 20.0 * n
        out["interp"] += 12.0 * za + 4.0 * n + 4.0 * zs + 4.0 * n + 4.0 * n + 12.0 * zp + 4.0 * n
        out["transpose"] += 12.0 * zp + 4.0 * n + 12.0 * zp + 4.0 * nc
        out["rap"] += 12.0 * (2.0 * zp + za) + 4.0 * (nc + 2.0 * n) + 12.0 * zc + 4.0 * nc
    return out


def spmv_traffic(bytes_per_entry=12):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (None if absent)"""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "spmv_traffic.json")))
        return tab["traffic_bytes_per_launch"] if bytes_per_entry == 12 else tab.get("traffic_bytes_per_launch_%dB" % bytes_per_entry)
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
    """SM clocks / throttle reasons sampled during the timed region (what `nvidia-smi --query-gpu=clocks.sm,...` prints).
    Read through NVML inside this process, initialised BEFORE the timed region: spawning one nvidia-smi per rank every 200 ms
    meant N concurrent NVML start-ups enumerating all GPUs of the node, and on 4-GPU boxes the first timed solve intermittently
    stalled for ~0.6 s (ranks spin on each other's exchanges, so one held-up launch holds up all).  Falls back to the nvidia-smi
    command line when the NVML binding is missing."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False
        self.nvml = self.dev = None
        try:
            import pynvml
            pynvml.nvmlInit()
            dev = None
            try:        # the CUDA ordinal is not the NVML index when CUDA_VISIBLE_DEVICES hides or reorders devices: go by UUID
                import torch
                uuid = str(torch.cuda.get_device_properties(gpu).uuid)
                dev = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            except Exception:
                dev = None
            if dev is None:
                dev = pynvml.nvmlDeviceGetHandleByIndex(gpu)
            self.nvml, self.dev = pynvml, dev
        except Exception:
            self.nvml = self.dev = None

    def device_name(self):
        try:
            u = self.nvml.nvmlDeviceGetUUID(self.dev)
            return u.decode() if isinstance(u, bytes) else str(u)
        except Exception:
            return None

    def sample_nvml(self):
        nv = self.nvml
        sm = nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(self.dev, nv.NVML_CLOCK_SM)
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
        act = lambda bit: "Active" if (r & bit) else "Not Active"
        return [str(sm), str(mx), act(0x8), act(0x40), act(0x20), act(0x4)]     # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.rows.append(self.sample_nvml())
                else:
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
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi",
                "device": self.device_name()}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(dims, threads, timeout=None):
    """One setup+solve of the UNMODIFIED reference driver (oracle/_ref/ij = test/ij.c compiled in place against the
    reference's own library); returns (setup_s, solve_s, iterations) from its wall-clock lines (ij.c:4301-4314)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ij")
    if dims[0] * dims[1] * dims[2] > 100_000_000:
        # one process holds the whole problem (no MPI here): past ~10^8 unknowns the reference's 32-bit counters overflow
        # (`ij -n 512 512 512` segfaults), so the same driver from the reference's own --enable-bigint configuration runs it
        exe = os.path.join(ROOT, "oracle", "_ref", "big", "ij_big")
    if not os.path.exists(exe):
        raise RuntimeError("%s missing: run __graft_entry__.build() where /root/reference exists" % os.path.relpath(exe, ROOT))
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false")
    out = subprocess.run([exe, "-n"] + [str(d) for d in dims] + REF_ARGS, env=env, capture_output=True, text=True, check=True,
                         timeout=timeout).stdout
    m = re.search(r"PCG Setup:\s*\n\s*wall clock time = (\S+) seconds", out)
    v = re.search(r"PCG Solve:\s*\n\s*wall clock time = (\S+) seconds", out)
    k = re.search(r"^Iterations = (\d+)", out, re.M)
    return float(m.group(1)), float(v.group(1)), int(k.group(1))


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    world = max(1, a.gpus)
    dims, _ = global_dims(world, a.n, a.grid)
    # bounded: probe with 128^3 (1/8 of one GPU's share), estimate the job linearly in the unknowns
    s0, v0, _ = run_reference((128, 128, 128), threads)
    est = (s0 + v0) * (dims[0] * dims[1] * dims[2]) / 128.0 ** 3 * 1.1
    steps, warmup = a.steps, a.warmup
    if est * (steps + warmup) > 240:          # the whole run must end within a few minutes: ONE real job, stated in the line
        steps, warmup = 1, 0
    base = {"impl": "reference", "metric": "boomeramg_pcg_setup_plus_solve_seconds", "unit": "s", "n_gpus": a.gpus}
    if est > 400:
        emit(dict(base, unavailable="the reference job %dx%dx%d is estimated at %.0f s on %d host threads (128^3 probe: %.2f s); "
                                   "not run, and never extrapolated" % (dims + (est, threads, s0 + v0))))
        return 0
    try:
        for _ in range(warmup):
            run_reference(dims, threads, timeout=900)
        t_set = t_sol = 0.0
        its = 0
        for _ in range(steps):
            s_, v_, its = run_reference(dims, threads, timeout=900)
            t_set += s_
            t_sol += v_
    except Exception as e:      # killed for memory, timeout: say so, never substitute a scaled number
        emit(dict(base, unavailable="the reference job %dx%dx%d failed on this host: %s" % (dims + (str(e)[:200],))))
        return 0
    per_step = (t_set + t_sol) / steps
    sample = "full workload %dx%dx%d per step, stock driver `ij -n ... %s`%s, %d step(s) after %d warm-up" % (
        dims + (" ".join(REF_ARGS), " (the reference's --enable-bigint build: 64-bit counters)" if dims[0] * dims[1] * dims[2] > 100_000_000 else "",
                steps, warmup))
    line = dict(base, **{
        "value": per_step, "steps": steps, "warmup": warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(world, a.n, a.grid),
        "reference_impl": "hypre 2.20 (SX-Aurora fork) CPU path: unmodified test/ij.c, OpenMP, sequential MPI stubs (np = 1)",
        "setup_s": t_set / steps, "solve_s": t_sol / steps, "iterations": its, "reference_iterations": its,
        "cpu_baseline": {"value": per_step, "unit": "s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": per_step, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    emit(line)
    return 0


def workload_name(world, n1, grid):
    if world <= 1:
        return WORKLOAD if n1 == N1 else WORKLOAD.replace("256^3", "%d^3" % n1).replace("256 256 256", "%d %d %d" % (n1, n1, n1))
    dims, pg = global_dims(world, n1, grid)
    return ("ij 3D 7-pt Laplacian %dx%dx%d (%d^3 unknowns per GPU, -P %d %d %d) BoomerAMG-PCG, PMIS + ext+i(Pmx 4) interp + l1-Jacobi, "
            "tol 1e-8 (ij -n %d %d %d -solver 1 -pmis -rlx 18)" % (dims + (n1,) + pg + dims))


def bench_config(world, n1, grid):
    """the `config` object: identical in this arm's line and in the reference arm's"""
    dims, _ = global_dims(world, n1, grid)
    return {"workload": workload_name(world, n1, grid), "global_grid": "%dx%dx%d" % dims, "unknowns_per_gpu": n1 ** 3,
            "l2": "per-GPU operator (1.7 GB at 256^3) larger than the 126 MB L2; no flush needed"}


def cpu_baseline_inline(dims, note):
    """the reference driver on all host cores, one run of `dims` (reported baseline of this line)"""
    threads = host_threads()
    try:
        s_, v_, cits = run_reference(dims, threads, timeout=600)
        return {"value": s_ + v_, "unit": "s", "cores": threads, "kind": "reference",
                "sample": "oracle/_ref/ij (unmodified reference driver), %s" % note,
                "setup_s": s_, "solve_s": v_, "iterations": cits}
    except Exception as e:      # the baseline is reporting only; never fail the GPU line for it
        return {"value": None, "unit": "s", "cores": threads, "kind": "reference", "sample": "unavailable: %s" % str(e)[:200]}


def main_dist(a, rank, world, local_rank):
    """N > 1: row-partitioned BoomerAMG-PCG, n1^3 unknowns per GPU (weak scaling)."""
    import torch
    import torch.distributed as dist
    import hypre_ve_b200 as hb
    from hypre_ve_b200 import launch

    n1 = a.n
    (nx, ny, nz), (P, Q, R) = global_dims(world, n1, a.grid)
    h = hb.Handle(local_rank)
    # communicator of the library: rank 0 creates the id, torch.distributed broadcasts it
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
    # driver default Galerkin order (fused BuildCoarseOperatorKT, (R A) P); levels with at most --seq-threshold rows in total
    # are replicated on every rank (HYPRE_BoomerAMGSetSeqThreshold, `ij -seq_th`): same hierarchy bit for bit, no exchanges there
    prm = hb.Amg(h, ModuleRAP2=0, SeqThreshold=a.seq_threshold)

    levels = []

    def step(record=False):
        amg = hb.DistAmg(h, comm, prm, A)
        s_ms = amg.setup_ms
        h.fill(x, 0.0)
        h.timer_start()
        its, rel, _ = hb.dist_pcg(h, comm, A, amg, b, x, tol=1e-8, max_iter=100)
        v_ms = h.timer_stop_ms()
        if record and not levels:
            for l in range(amg.num_levels):
                ia = amg.level_A(l).info
                levels.append((ia["local_rows"], ia["local_nnz"], amg.level_P(l).info["local_nnz"] if l < amg.num_levels - 1 else 0))
        amg.destroy()
        return s_ms, v_ms, its, rel

    for k in range(a.warmup):
        step(record=(k == 0))
    if not levels:
        step(record=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = h.launch_count()
    barrier()
    t_set = t_sol = 0.0
    step_ms = []
    for _ in range(a.steps):
        s_ms, v_ms, its, rel = step()
        t_set += s_ms
        t_sol += v_ms
        step_ms.append((round(s_ms, 3), round(v_ms, 3)))
    barrier()
    launches = h.launch_count() - launches0
    # end to end through the rows-from-the-caller entry point (b200_dist_matrix_create_from_host = IJ assembly per rank,
    # hypre_IJMatrixAssembleParCSR): every rank's CSR rows (global column ids) and rhs start in pinned HOST memory,
    # are uploaded inside the timed region, and x comes back to the host
    gi, gj, ga = A.download()
    t_i, t_j, t_a = (torch.from_numpy(v).pin_memory() for v in (gi, gj, ga))
    hi, hj, ha = (t.numpy() for t in (t_i, t_j, t_a))
    hb_host = torch.ones(n, dtype=torch.float64).pin_memory().numpy()
    hx_host = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    h2d = hi.nbytes + hj.nbytes + ha.nbytes + hb_host.nbytes
    e2e_ms = 0.0
    e2e_its = 0
    for k in range(a.steps + 1):
        barrier()
        h.timer_start()
        A2 = hb.DistMatrix.from_rows(h, comm, hi, hj, ha)
        b2 = A2.vector(0.0)
        x2 = A2.vector(0.0)
        hb._chk(hb._lib.b200_memcpy_h2d(h.p, b2.ptr, hb._np_ptr(hb_host), hb_host.nbytes))
        amg = hb.DistAmg(h, comm, prm, A2)
        e2e_its, _, _ = hb.dist_pcg(h, comm, A2, amg, b2, x2, tol=1e-8, max_iter=100)
        hb._chk(hb._lib.b200_memcpy_d2h(h.p, hb._np_ptr(hx_host), x2.ptr, hx_host.nbytes))
        ms = h.timer_stop_ms()
        amg.destroy(); A2.destroy(); b2.free(); x2.free()
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
    flat = [float(v) for lv in levels for v in lv] + [0.0] * (3 * 32 - 3 * len(levels))
    glev = launch.reduce_over_ranks(flat, "sum", "cuda")
    glevels = [tuple(glev[3 * l:3 * l + 3]) for l in range(len(levels))]
    set_s, sol_s, e2e_s, spmv_ms = per[0] / 1e3, per[1] / 1e3, per[2] / 1e3, per[3]
    spmv_bytes = 12.0 * gnnz + 4.0 * (gn + world) + 16.0 * gn
    bpe = A.stream_bytes_per_entry               # 12, or 9 / 5 / 2 with the dictionary-compressed solve copy of A0 (this rank's block)
    peak, peak_src = peaks()
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    it_bytes = solve_bytes_per_iteration(glevels)
    it_gbs = it_bytes * its / sol_s / 1e9
    ref_its = reference_iterations((nx, ny, nz)) if a.grid == "slab" else None
    line = {
        "metric": "boomeramg_pcg_setup_plus_solve_seconds", "value": set_s + sol_s, "unit": "s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": (set_s + sol_s) * 1e3,
        "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(world, n1, a.grid),
        "problem": {"rows_per_gpu": n, "nnz_per_gpu": nnz, "global_rows": int(gn), "global_nnz": int(gnnz),
                    "parallelism": "row-partitioned ParCSR (-P %d %d %d), %d GPUs, halo + reductions over NVLink; levels <= %d rows "
                                   "replicated (seq_threshold)" % (P, Q, R, world, a.seq_threshold)},
        "setup_s": set_s, "solve_s": sol_s, "iterations": its, "final_rel_res": rel,
        "rank0_step_ms": [{"setup": sm, "solve": vm} for sm, vm in step_ms],
        "reference_iterations": ref_its, "iterations_match_reference": (its == ref_its) if ref_its is not None else None,
        "e2e_iterations": e2e_its,
        "spmv_gbs": achieved, "spmv_ms": spmv_ms,
        "roofline": {"bound": "hbm", "kernel": "halo exchange + spmv_pipe_kernel (y = A0*x, %d^3 per GPU)" % n1,
                     "achieved": achieved, "peak": peak * world, "peak_source": peak_src + " x n_gpus", "unit": "GB/s",
                     "frac": achieved / (peak * world), "algorithmic_bytes_per_launch": spmv_bytes / world, "traffic": None,
                     "stream_bytes_per_entry": bpe, "format_bytes_per_launch": (spmv_bytes - (12.0 - bpe) * gnnz) / world},
        "roofline_solve": {"bound": "hbm", "what": "whole PCG iteration (V(1,1) cycle over %d levels + SpMV + BLAS-1), all GPUs" % len(glevels),
                           "algorithmic_bytes_per_iteration": it_bytes, "ms_per_iteration": sol_s * 1e3 / max(1, its),
                           "achieved": it_gbs, "peak": peak * world, "unit": "GB/s", "frac": it_gbs / (peak * world)},
        "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": hx_host.nbytes * world},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if rank == 0:
        if not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_inline((n1, n1, n1), "%d^3 = ONE GPU's share of this job (1/%d of the unknowns), "
                                                       "seconds NOT scaled" % (n1, world))
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
    ap.add_argument("--grid", choices=["slab", "box"], default="slab",
                    help="N > 1: z-slabs (-P 1 1 N, lexicographic numbering = the reference's np = 1 job) or boxes (-P 2 2 2 ...)")
    ap.add_argument("--seq-threshold", type=int, default=200000,
                    help="N > 1: levels with at most this many rows in total are replicated on every rank (ij -seq_th); 0 = off")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    claim_stdout()
    if a.impl == "reference":
        return reference_arm(a)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # kernels of one rank wait on flags that kernels of its neighbours raise (direct NVLink halos): load every kernel
        # before the first launch, so that no first-time module load has to wait for a kernel that is itself waiting
        os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
    import torch
    import torch.distributed as dist
    import hypre_ve_b200 as hb

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

    levels = []

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
        if not levels:
            for l in range(amg.num_levels):
                na, _, za = amg.level_A(l).dims
                levels.append((na, za, amg.level_P(l).dims[2] if l < amg.num_levels - 1 else 0))
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
    bpe = A.diag.stream_bytes_per_entry          # 12, or 9 / 5 / 2 with the dictionary-compressed solve copy of A0
    fmt_bytes = spmv_bytes - (12.0 - bpe) * nnz
    peak, peak_src = peaks()
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    it_bytes = solve_bytes_per_iteration(levels)
    it_gbs = it_bytes * its / sol_s / 1e9
    it_fmt_bytes = it_bytes - 3.0 * (12.0 - bpe) * nnz       # A0 is applied three times per iteration
    ph_ms = dict(zip(["strength", "pmis", "interp", "trunc", "transpose", "rap", "l1_alloc", "total"], (phases / a.steps).tolist()))
    ref_its = reference_iterations((n1, n1, n1))
    line = {
        "metric": "boomeramg_pcg_setup_plus_solve_seconds", "value": set_s + sol_s, "unit": "s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": (set_s + sol_s) * 1e3,
        "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(1, n1, a.grid),
        "problem": {"rows_per_gpu": n, "nnz_per_gpu": nnz, "parallelism": "1 GPU"},
        "setup_s": set_s, "solve_s": sol_s, "iterations": its, "final_rel_res": rel,
        "reference_iterations": ref_its, "iterations_match_reference": (its == ref_its) if ref_its is not None else None,
        "e2e_iterations": its_e,
        "roofline_solve": {"bound": "hbm", "what": "whole PCG iteration (V(1,1) cycle over %d levels + SpMV + BLAS-1)" % len(levels),
                           "algorithmic_bytes_per_iteration": it_bytes, "ms_per_iteration": sol_s * 1e3 / max(1, its),
                           "achieved": it_gbs, "peak": peak, "unit": "GB/s", "frac": it_gbs / peak,
                           "format_bytes_per_iteration": it_fmt_bytes, "frac_of_format_bytes": it_fmt_bytes * its / sol_s / 1e9 / peak},
        "setup_phases_ms": dict(zip(["strength", "pmis", "interp", "trunc", "transpose", "rap", "l1_alloc", "total"],
                                    (phases / a.steps).round(3).tolist())),
        "setup_roofline": {k: {"ms": float(ph_ms[k]), "algorithmic_bytes": v, "achieved_gbs": v / (float(ph_ms[k]) * 1e-3) / 1e9 if ph_ms[k] > 0 else None,
                               "frac": v / (float(ph_ms[k]) * 1e-3) / 1e9 / peak if ph_ms[k] > 0 else None}
                           for k, v in setup_bytes(levels).items()},
        "spmv_gbs": achieved, "spmv_ms": spmv_ms,
        "roofline": {"bound": "hbm", "kernel": ("spmv_dict_kernel<2,1,1>" if bpe == 2 else "spmv_pipe_kernel<1,2>") + " (y = A0*x, 256^3 7-pt)",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes_per_launch": spmv_bytes, "traffic": spmv_traffic(bpe),
                     "stream_bytes_per_entry": bpe, "format_bytes_per_launch": fmt_bytes,
                     "frac_of_format_bytes": fmt_bytes / (spmv_ms * 1e-3) / 1e9 / peak,
                     "note": "algorithmic bytes = SURVEY 8d CSR figure (12 B per entry); the solve copy of a stencil-structured operator "
                             "stores one byte per column offset and per value (lossless, bit-identical products), so DRAM traffic is below "
                             "the algorithmic bytes and `frac` can exceed 1" if bpe != 12 else None},
        "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_inline((n1, n1, n1), "full workload %d^3, one run" % n1)
    if rank == 0:
        emit(line)
    A.destroy()
    h.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
