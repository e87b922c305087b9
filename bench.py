#!/usr/bin/env python3
"""bench.py -- headline benchmark of the PINN-FEM hot path on B200.

Metric (BASELINE.json): element assembly+residual evaluations per second.
Workload (BASELINE.json configs[4], the synthetic C5 case of SURVEY.md 8d):
2-D lattice truss, 578 x 578 nodes = 999,941 elements, batched independent
problems sharing the mesh, 512 problems per GPU (4096 over 8 GPUs: weak
scaling, problems are sharded by batch with no data-path collective).
One "step" = one residual evaluation r = f_int(u; E, A) - lambda * f_ext of
every problem of the shard = 512 x 999,941 element evaluations per GPU.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N ...            # CPU restatement of the reference (rank 0)

Rank 0 prints ONE JSON line (see README / DESIGN.md for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

NX = 578
PROBLEMS_PER_GPU = 512
CPU_SAMPLE_PROBLEMS_PER_THREAD = 8


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks and throttle reasons of one GPU with nvidia-smi while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.t_mark = 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Start of the timed region: samples before it (warm-up) are reported separately."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, timed = [], [], set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            timed += ts >= self.t_mark
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_in_timed_region": timed, "reasons": sorted(reasons),
                "window": "warm-up + timed region (nvidia-smi -lms 20)"}


def algorithmic_bytes(nelem, nnode, B):
    """per problem 16*nelem (E, A) + 32*nnode (u read, r write); the mesh
    indices/coordinates (8*nelem + 16*nnode) are counted once per sweep."""
    return B * (16 * nelem + 32 * nnode) + 8 * nelem + 16 * nnode


def synthetic_problem(p, ndof, nelem):
    """Problem p of the synthetic batch: np.random.default_rng(p), u ~ U(-1e-3, 1e-3),
    then E, then A ~ U(0.5, 1.5)."""
    import numpy as np

    rng = np.random.default_rng(p)
    return rng.uniform(-1e-3, 1e-3, ndof), rng.uniform(0.5, 1.5, nelem), rng.uniform(0.5, 1.5, nelem)


def synthetic_inputs(ndof, nelem, B, first_problem, device, group=16):
    """This rank's shard [first_problem, first_problem + B) of the synthetic batch in the kernels' layout
    ([row][problem]): generated on the host problem by problem (a few threads), staged through pinned memory in
    groups and transposed on the device.  f_ext is shared by all problems (default_rng(2**31))."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np
    import torch

    u = torch.empty((ndof, B), dtype=torch.float64, device=device)
    E = torch.empty((nelem, B), dtype=torch.float64, device=device)
    A = torch.empty((nelem, B), dtype=torch.float64, device=device)
    on_gpu = torch.device(device).type == "cuda"
    stage = [torch.empty((group, n), dtype=torch.float64, pin_memory=on_gpu) for n in (ndof, nelem, nelem)]
    try:
        workers = max(1, min(8, len(os.sched_getaffinity(0))))
    except AttributeError:
        workers = 4
    with ThreadPoolExecutor(workers) as pool:
        for g0 in range(0, B, group):
            gb = min(group, B - g0)
            probs = list(pool.map(lambda q: synthetic_problem(first_problem + q, ndof, nelem), range(g0, g0 + gb)))
            for j, (pu, pe, pa) in enumerate(probs):
                stage[0][j].copy_(torch.from_numpy(pu))
                stage[1][j].copy_(torch.from_numpy(pe))
                stage[2][j].copy_(torch.from_numpy(pa))
            for dst, st in zip((u, E, A), stage):
                dst[:, g0:g0 + gb] = st[:gb].to(device, non_blocking=True).t()
            if on_gpu:
                torch.cuda.synchronize(device)  # the staging buffers are reused by the next group
    fx = torch.from_numpy(np.random.default_rng(2 ** 31).normal(scale=1e-3, size=ndof)).to(device)
    return u, E, A, fx


def oracle_parity(nodes, el, fixed, u, E, A, fx, r, cols):
    """Columns `cols` of the device residual against the C restatement of the reference's element loop
    (oracle/pf_oracle.c; fem/assembly.py:52-73 + fem/solver.py:267-269) on the same inputs."""
    import numpy as np

    from oracle import c_oracle

    idx = list(cols)
    uh, Eh, Ah = (np.ascontiguousarray(t[:, idx].cpu().numpy()) for t in (u, E, A))
    _, r_ref = c_oracle.residual(nodes, el, Eh, Ah, uh, fx.cpu().numpy(), 1.0, fixed, want_f=False, want_r=True)
    r_dev = r[:, idx].cpu().numpy()
    return float(np.max(np.abs(r_dev - r_ref)) / np.max(np.abs(r_ref)))


def tangent_algorithmic_bytes(nelem, nnode, nnzb, B):
    """SURVEY.md 8(d), tangent only: per problem 16*nelem (E, A) + 32*nnzb (2x2 fp64 blocks written);
    mesh indices/geometry once per sweep: 8*nelem (conn) + 16*nnode (coords) + 16*nelem (block slots)."""
    return B * (16 * nelem + 32 * nnzb) + 24 * nelem + 16 * nnode


def tangent_leg(plan, E, A, dev, Bt, steps, peak):
    """Secondary line: tangent stiffness K_t into node-block CSR for Bt problems of the shard."""
    import torch

    Et, At = E[:, :Bt].contiguous(), A[:, :Bt].contiguous()
    vals = plan.tangent_bsr(Et, At, B=Bt)
    for _ in range(2):
        plan.tangent_bsr(Et, At, B=Bt, out=vals)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        plan.tangent_bsr(Et, At, B=Bt, out=vals)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    ab = tangent_algorithmic_bytes(plan.nelem, plan.nnode, plan.nnzb, Bt)
    return {"problems": Bt, "ms_per_launch": ms, "element_evals_per_s": Bt * plan.nelem / (ms * 1e-3),
            "roofline": {"bound": "hbm", "kernel": "tangent_bsr_kernel<2,0>", "achieved": ab / (ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": ab / (ms * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_launch": ab}, "gpu_launches": steps + 3}


def jtj_leg(dev, n, m, peak_note="fp64 peaks are not in MEASURED_PEAKS.json: measured in this run with register-resident DFMA / DMMA loops"):
    """Secondary line: Gauss-Newton normal equations J^T J + damping (fem/nn_solver.py:268-274) on a random
    full-rank J ~ N(0,1) (SURVEY 8d, C4 scaled variant) and the LU solve of the damped system."""
    import torch

    from pinn_fem_b200 import ops

    g = torch.Generator(device=dev).manual_seed(0)
    J = torch.randn((m, n), generator=g, device=dev, dtype=torch.float64)
    R = torch.randn(m, generator=g, device=dev, dtype=torch.float64)
    out = {"m": m, "n": n}
    for name, fn in (("jtj", lambda: ops.gn_normal_equations(J, R)),):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            jtj, jtr, _ = fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / 3
        out["jtj_ms"] = ms
        out["jtj_tflops_full_2mn2"] = 2.0 * m * n * n / (ms * 1e-3) / 1e12
        out["jtj_tflops_executed"] = out["jtj_tflops_full_2mn2"] * (n / 64 + 1) / (2 * (n / 64))  # upper tiles only
    # damped LM system (SPD by construction): blocked Cholesky (product path) and the pivoted LU of
    # np.linalg.solve / torch.linalg.solve (kept for K_ff), both warm
    for name, fn, flops in (("spd_solve", ops.solve_spd, n ** 3 / 3.0), ("lu_solve", ops.solve_dense, 2.0 * n ** 3 / 3.0)):
        fn(jtj, -jtr)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dx = fn(jtj, -jtr)
        e1.record()
        torch.cuda.synchronize(dev)
        out[f"{name}_ms"] = e0.elapsed_time(e1)
        out[f"{name}_tflops"] = flops / (out[f"{name}_ms"] * 1e-3) / 1e12
    import ctypes

    from pinn_fem_b200 import _lib

    peaks = {}
    for kind, name in ((0, "dfma_tflops"), (1, "dmma_tflops")):
        v = ctypes.c_double(0.0)
        _lib.check(_lib.load().pf_measure_fp64_peak(kind, ctypes.byref(v)))
        peaks[name] = v.value
    out["fp64_peaks_measured"] = peaks
    out["roofline"] = {"bound": "tensor", "achieved": out["jtj_tflops_executed"], "peak": peaks["dmma_tflops"],
                       "unit": "TFLOP/s", "frac": out["jtj_tflops_executed"] / peaks["dmma_tflops"],
                       "peak_source": "pf_measure_fp64_peak(DMMA) in this run"}
    out["peak_note"] = peak_note
    out["gpu_launches"] = 4 * 3 + 4 * (n // 32 + 1)
    return out


def cg_leg(plan, E, A, fx, dev, Bc, iters, peak):
    """Secondary line: batched matrix-free Jacobi-CG (the linear solve of the Newton step on large meshes),
    a fixed number of iterations.  Bytes per iteration and problem: the mat-vec (16*nelem + 32*nnode) plus
    13 vector passes of the CG recurrences (16*nnode each)."""
    import torch

    from pinn_fem_b200 import ops

    Ec, Ac = E[:, :Bc].contiguous(), A[:, :Bc].contiguous()
    rhs = fx[:, None].expand(-1, Bc).contiguous()
    ops.cg_solve(plan, Ec, Ac, rhs, rel_tol=0.0, max_iters=8)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, n_it, resid = ops.cg_solve(plan, Ec, Ac, rhs, rel_tol=0.0, max_iters=iters)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / n_it
    ab = Bc * (16 * plan.nelem + 32 * plan.nnode + 13 * 16 * plan.nnode)
    return {"problems": Bc, "iterations": n_it, "ms_per_iteration": ms,
            "roofline": {"bound": "hbm", "achieved": ab / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": ab / (ms * 1e-3) / 1e9 / peak, "bytes_per_iteration": ab},
            "gpu_launches": 7 * (n_it + 8) + 8}


def cpu_reference_leg(steps, warmup, sample_problems=None):
    """Times the CPU restatement of the reference's element loop (oracle/pf_oracle.c, OpenMP over
    problems) on a bounded sample of the same workload: the full 999,941-element mesh, a slice of
    the batch."""
    import numpy as np

    from oracle import c_oracle
    from oracle import pinnfem_oracle as O

    # every host core this process may run on: torchrun exports OMP_NUM_THREADS=1, which is not a
    # property of the machine, so the thread count is passed to the OpenMP region explicitly
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    threads = max(threads, c_oracle.max_threads())
    nodes, el, fixed = O.lattice_truss(NX)
    Bs = sample_problems or threads * CPU_SAMPLE_PROBLEMS_PER_THREAD
    u, E, A = (np.empty((n, Bs)) for n in (2 * len(nodes), len(el), len(el)))
    for q in range(Bs):  # the first Bs problems of the synthetic batch (same seeds as the GPU arm)
        u[:, q], E[:, q], A[:, q] = synthetic_problem(q, 2 * len(nodes), len(el))
    fx = np.random.default_rng(2 ** 31).normal(scale=1e-3, size=2 * len(nodes))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        c_oracle.residual(nodes, el, E, A, u, fx, 1.0, fixed, want_f=False, want_r=True, nthreads=threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    value = Bs * len(el) * len(times) / total
    # the reference's own interpreter-bound loop, restated line by line (oracle.assemble_system_loop), on a
    # tiny sample for context: this is what fem/assembly.py actually costs per element in Python
    nodes_s, el_s, _ = O.lattice_truss(24)
    t0 = time.perf_counter()
    O.assemble_system_loop(nodes_s, el_s, 1.0, 1.0, np.zeros(2 * len(nodes_s)))
    py_rate = len(el_s) / (time.perf_counter() - t0)
    return {"value": value, "unit": "element_evals/s", "cores": threads, "kind": "port",
            "sample": f"{Bs} of {PROBLEMS_PER_GPU} problems x full {len(el)}-element mesh per step, "
                      f"{len(times)} steps, C restatement with OpenMP over problems",
            "python_loop_element_evals_per_s": py_rate, "ms_per_step": 1e3 * total / len(times)}


def _parse_cpu_list(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs next to its GPU before any pinned host buffer is allocated (first touch then
    places the buffers on that socket): the end-to-end leg is a host-memory -> PCIe stream and crossing sockets halves
    it.  Sources, in order: sysfs numa_node of the GPU's PCI function; the CPU-affinity column of `nvidia-smi topo -m`
    (containers often report numa_node = -1); else an even split of the allowed CPUs over the local ranks."""
    allowed = os.sched_getaffinity(0)
    try:
        import torch

        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text())
        if node >= 0:
            cpus = _parse_cpu_list(Path(f"/sys/devices/system/node/node{node}/cpulist").read_text()) & allowed
            if cpus:
                os.sched_setaffinity(0, cpus)
                return {"source": "sysfs", "numa_node": node, "cpus": len(cpus)}
    except Exception:
        pass
    try:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        header = next(l for l in topo.splitlines() if "CPU Affinity" in l)
        col = header.split("\t").index(next(h for h in header.split("\t") if "CPU Affinity" in h))
        row = next(l for l in topo.splitlines() if l.startswith(f"GPU{local_rank}\t") or l.startswith(f"GPU{local_rank} "))
        cells = [c.strip() for c in row.split("\t")]
        cpus = _parse_cpu_list(cells[col]) & allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            numa = cells[col + 1] if col + 1 < len(cells) else None
            return {"source": "nvidia-smi topo -m", "numa_node": numa, "cpus": len(cpus)}
    except Exception:
        pass
    try:
        n_local = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        if n_local > 1:
            ordered = sorted(allowed)
            per = max(1, len(ordered) // n_local)
            cpus = set(ordered[local_rank * per:(local_rank + 1) * per])
            if cpus:
                os.sched_setaffinity(0, cpus)
                return {"source": "even split of the allowed CPUs over the local ranks", "numa_node": None, "cpus": len(cpus)}
    except Exception:
        pass
    return None


def _claim_stdout():
    """Rank 0 must print exactly ONE line on stdout.  Libraries write banners there too (NCCL prints its
    version on the first collective), so file descriptor 1 is pointed at stderr for the duration of the run and
    the JSON line goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    out_stream = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--problems-per-gpu", type=int, default=PROBLEMS_PER_GPU)
    ap.add_argument("--nx", type=int, default=NX)
    ap.add_argument("--e2e-problems", type=int, default=0, help="problems in the host-buffer leg (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gd", action="store_true")
    ap.add_argument("--gd-problems-per-gpu", type=int, default=64, help="batched inverse problems per GPU (config 5)")
    ap.add_argument("--gd-iters", type=int, default=20)
    ap.add_argument("--no-tangent", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nelem_full = 2 * (args.nx - 1) * args.nx + (args.nx - 1) ** 2
    config = {"workload": f"synthetic 2D lattice truss {args.nx}x{args.nx} nodes, {nelem_full} elements x "
                          f"{args.problems_per_gpu} batched problems per GPU "
                          f"({args.problems_per_gpu * max(world, args.gpus)} total), linear bar element, "
                          f"E and A per (element, problem), residual r = f_int - lambda*f_ext",
              "sharding": "batch (independent problems), no data-path collective",
              "cache": "inputs per step (13.7 GB) are far larger than the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        # every step is a bounded sample of the workload (8 problems per host thread x the full mesh, ~0.5 s), so
        # --steps / --warmup are honoured as given
        steps, warm = max(1, args.steps), max(0, args.warmup)
        base = cpu_reference_leg(steps, warm)
        line = {"impl": "reference", "metric": "element assembly+residual evals/sec", "value": base["value"],
                "unit": "element_evals/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": "element_evals/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "reference is pure Python (no compiled code to build); this arm times the C/OpenMP "
                        "restatement of its element loop on all host threads. The literal Python loop runs at "
                        f"{base['python_loop_element_evals_per_s']:.0f} element_evals/s on one core."}
        print(json.dumps(line), file=out_stream, flush=True)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist

    from pinn_fem_b200 import AssemblyPlan
    from pinn_fem_b200.meshes import lattice_truss
    from pinn_fem_b200.sharding import max_over_ranks

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pinn_fem_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    nodes, el, fixed = lattice_truss(args.nx)
    plan = AssemblyPlan(nodes, el, fixed, device=dev)
    B = args.problems_per_gpu
    u, E, A, fx = synthetic_inputs(plan.ndof, plan.nelem, B, rank * B, dev)
    r = torch.empty_like(u)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        plan.residual_into(u, E, A, fx, 1.0, r)
    barrier()
    sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        plan.residual_into(u, E, A, fx, 1.0, r)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    ms_max = max_over_ranks(ms, dev)
    ms_per_step = ms_max / args.steps
    value = world * B * plan.nelem * args.steps / (ms_max * 1e-3)

    # ---- parity of the benchmarked configuration itself: columns of r against the CPU oracle ----------------
    cols = sorted({0, 1, B // 2, B - 1})
    perr = oracle_parity(nodes, el, fixed, u, E, A, fx, r, cols)
    perr = max_over_ranks(perr, dev)
    parity = {"rel_err": perr, "tolerance": 1e-10, "ok": bool(perr < 1e-10), "columns_per_gpu": cols,
              "against": "oracle/pf_oracle.c (C restatement of fem/assembly.py:52-73 + fem/solver.py:267-269), same inputs, "
                         "full mesh; max over ranks"}

    # ---- roofline of the dominant (only) kernel of the step: one patch_gather launch per step -------------
    peak, peak_src = measured_peaks()
    ab = algorithmic_bytes(plan.nelem, plan.nnode, B)
    kernel_ms = ms / args.steps  # this rank's average launch duration (launches are back to back on one stream)
    achieved = ab / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "r1_traffic.json"
    if tp.exists():
        with open(tp) as f:
            t = json.load(f)
        if t.get("nx") == args.nx and t.get("problems") == B:
            traffic = t.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "patch_gather_kernel<2,0,true,false,false>", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ab, "kernel_ms": kernel_ms}

    # ---- end to end through the public host-buffer API (pinned host -> device -> pinned host) --------------
    Be = args.e2e_problems
    if Be <= 0:
        try:
            with open("/proc/meminfo") as f:
                avail_kb = next(int(l.split()[1]) for l in f if l.startswith("MemAvailable"))
        except Exception:
            avail_kb = 0
        full_bytes = (2 * plan.ndof + 2 * plan.nelem) * B * 8
        Be = B if avail_kb * 1024 > 4 * full_bytes * max(world, 1) else min(B, 128)
    uh = torch.empty((plan.ndof, Be), dtype=torch.float64, pin_memory=True)
    Eh = torch.empty((plan.nelem, Be), dtype=torch.float64, pin_memory=True)
    Ah = torch.empty((plan.nelem, Be), dtype=torch.float64, pin_memory=True)
    rh = torch.empty((plan.ndof, Be), dtype=torch.float64, pin_memory=True)
    uh.copy_(u[:, :Be])
    Eh.copy_(E[:, :Be])
    Ah.copy_(A[:, :Be])
    fxh = fx.cpu()
    torch.cuda.synchronize(dev)
    e2e_steps = max(1, min(args.steps, 3))
    plan.residual_host(uh, Eh, Ah, fxh, 1.0, rh)  # warm-up: allocates the staging buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.residual_host(uh, Eh, Ah, fxh, 1.0, rh)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    e2e_value = world * Be * plan.nelem * e2e_steps / max_over_ranks(e2e_s, dev)
    e2e_ok = bool(torch.equal(rh[:, :8], r[:, :8].cpu()))
    e2e = {"value": e2e_value, "unit": "element_evals/s",
           "h2d_bytes_per_step": (plan.ndof + 2 * plan.nelem) * Be * 8 + plan.ndof * 8,
           "d2h_bytes_per_step": plan.ndof * Be * 8, "problems_per_gpu": Be, "steps": e2e_steps,
           "api": "AssemblyPlan.residual_host -> pf_residual_host (pinned host buffers, chunked H2D/compute/D2H "
                  "pipeline on 3 streams)", "matches_device_result": e2e_ok, "numa_binding": numa}
    launches = args.steps + e2e_steps * ((Be + 63) // 64)

    extra = {}
    if not args.no_tangent:
        try:
            extra["tangent_bsr"] = tangent_leg(plan, E, A, dev, min(B, 64), max(3, min(args.steps, 10)), peak)
            t_ms = max_over_ranks(extra["tangent_bsr"]["ms_per_launch"], dev)
            extra["tangent_bsr"]["element_evals_per_s"] = world * min(B, 64) * plan.nelem / (t_ms * 1e-3)
            launches += extra["tangent_bsr"].pop("gpu_launches")
        except Exception as exc:
            extra["tangent_bsr"] = {"error": f"{type(exc).__name__}: {exc}"}
    if not args.no_tangent:
        try:
            extra["cg_solve"] = cg_leg(plan, E, A, fx, dev, min(B, 128), 40, peak)
            launches += extra["cg_solve"].pop("gpu_launches")
        except Exception as exc:
            extra["cg_solve"] = {"error": f"{type(exc).__name__}: {exc}"}
    if not args.no_tangent:
        try:
            extra["gauss_newton"] = jtj_leg(dev, 4096, 4096)
            launches += extra["gauss_newton"].pop("gpu_launches")
        except Exception as exc:
            extra["gauss_newton"] = {"error": f"{type(exc).__name__}: {exc}"}
    if not args.no_gd:
        try:
            from bench_gd import gd_iterations_per_second

            extra["pinn_gd"] = gd_iterations_per_second(dev, world)
            launches += extra["pinn_gd"].get("gpu_launches", 0)
            from bench_gd import gd_large_mesh_iterations_per_second

            extra["pinn_gd_large_mesh"] = gd_large_mesh_iterations_per_second(dev, plan, world)
            launches += extra["pinn_gd_large_mesh"].pop("gpu_launches", 0)
            from bench_gd import gd_batched_large_mesh

            dmma_peak = (extra.get("gauss_newton", {}).get("fp64_peaks_measured") or {}).get("dmma_tflops")
            del E, A, r  # the batched inverse problems need the memory of the residual batch
            torch.cuda.empty_cache()
            extra["pinn_gd_batched_large"] = gd_batched_large_mesh(dev, plan, world, rank, args.gd_problems_per_gpu,
                                                                   args.gd_iters, dmma_peak)
            launches += extra["pinn_gd_batched_large"].pop("gpu_launches", 0)
            if rank == 0:
                from bench_gd import example_runs, gn_iteration_leg

                extra["examples"] = example_runs(ROOT / "tests" / "golden" / "inputs")
                extra["gauss_newton_example10"] = gn_iteration_leg(ROOT / "tests" / "golden" / "inputs")
            if world > 1:
                from bench_gd import gd_element_sharded_iterations_per_second

                extra["pinn_gd_element_sharded"] = gd_element_sharded_iterations_per_second(dev, nodes, el, fixed, world)
        except Exception as exc:  # the headline metric must not depend on the secondary one
            extra["pinn_gd"] = {"error": f"{type(exc).__name__}: {exc}"}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b = cpu_reference_leg(2, 1)
        cpu_base = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu_base["python_loop_element_evals_per_s"] = b["python_loop_element_evals_per_s"]

    if rank == 0:
        line = {"metric": "element assembly+residual evals/sec", "value": value, "unit": "element_evals/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (SURVEY 8d: problem p from np.random.default_rng(p): u~U(-1e-3,1e-3), E,A~U(0.5,1.5), "
                        "lambda=1; rank k holds problems [k*B, (k+1)*B))",
                "config": config, "roofline": roofline, "parity": parity, "cpu_baseline": cpu_base, "e2e": e2e,
                "gpu_launches": args.steps,  # timed region: exactly one patch_gather_kernel launch per step
                "gpu_launches_all_legs": launches, "clocks": clocks}
        line.update(extra)
        print(json.dumps(line), file=out_stream, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
