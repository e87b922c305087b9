"""Secondary benchmark metric: PINN-GD iterations per second (BASELINE.json configs[1] and configs[4]).
Bench helpers (not part of the product package); imported by bench.py, scripts/ and the multi-GPU tests.

Workload: the example 4-P model (4-node bar, 3 elements, E/A/rho = three tanh MLPs with
521/316/161 parameters, 6 measured DOFs, Adam on u and theta) -- every iteration is the
full fem/solver.py:252-355 loop body including the history row.  ``single`` runs one
problem (the reference's use case: latency bound), ``batched`` runs 512 independent
inverse problems per GPU (one CTA each), the batch-sharded configuration of BASELINE.json."""
from __future__ import annotations

import numpy as np
import torch

from pinn_fem_b200 import ops
from pinn_fem_b200.plan import AssemblyPlan

NODES = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [3.0, 0.0]])
ELEMENTS = np.array([[0, 1], [1, 2], [2, 3]])
FIXED = np.array([0, 1, 3, 5, 7])
LOADS = np.array([0.0, 0, 0, 0, 0, 0, 1.0, 0])
MEAS_DOFS = np.array([2, 3, 4, 5, 6, 7])
MEAS_VALS = np.array([1.0, 0, 2, 0, 3, 0])


def _theta0(nprob, device, seed=0):
    from pinn_fem_b200.examples.json.generic import SimpleNN

    torch.manual_seed(seed)
    nets = [SimpleNN(2, w, 3) for w in (20, 15, 10)]
    base = torch.cat([p.detach().reshape(-1) for n in nets for p in n.parameters()]).to(device)
    g = torch.Generator(device=device).manual_seed(seed)
    noise = 0.01 * torch.randn((nprob, base.numel()), generator=g, device=device, dtype=torch.float64)
    noise[0] = 0.0
    return (base[None, :] + noise).contiguous()


def gd_iterations_per_second(device, world=1, iters=2000, batched_problems=512, batched_iters=300):
    plan = AssemblyPlan(NODES, ELEMENTS, FIXED, device=device)
    nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), ops.NetSpec(3, 2, 10)]
    f_ext = torch.as_tensor(LOADS).to(device)
    out = {"workload": "example4-P model: 3 elements, 3 MLPs (998 parameters), 6 measured DOFs, Adam on u and theta, "
                       "history recorded; fixed iteration count (tolerance 0)", "dtype": "f64"}
    launches = 0
    for label, nprob, n_it in (("single", 1, iters), ("batched", batched_problems, batched_iters)):
        kw = dict(max_iterations=n_it, tolerance=0.0, learning_rate_u=0.01, learning_rate_theta=5e-4,
                  alpha_physics=1.0, alpha_data=100.0, load_factor=1.0)
        best = None
        for rep in range(3):
            theta = _theta0(nprob, device)
            u = torch.zeros((nprob, 8), dtype=torch.float64, device=device)
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = ops.gd_solve(plan, nets, [1.0, 1.0, 1.0], theta, u, f_ext, MEAS_DOFS, MEAS_VALS, **kw)
            e1.record()
            torch.cuda.synchronize(device)
            launches += 1
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        assert int(res.n_iters.min()) == n_it
        t = torch.tensor([best], device=device, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[label] = {"problems_per_gpu": nprob, "iterations_per_problem": n_it, "ms": float(t.item()),
                      "iters_per_s": world * nprob * n_it / (float(t.item()) * 1e-3),
                      "us_per_iteration": 1e3 * float(t.item()) / n_it}
    out["reference_iters_per_s"] = {"value": 55.0, "source": "BASELINE.md section 2: example4-P, 1874 iterations in 33.8 s "
                                                           "on the survey container's CPU (indicative)"}
    out["gpu_launches"] = launches
    return out


def gd_large_mesh_iterations_per_second(device, plan, world=1, iters=100):
    """PINN-GD on the C5 lattice itself (one inverse problem per GPU, 999,941 elements): E and A are the
    521- and 316-parameter MLPs evaluated at every element centroid each iteration, the loop is the
    device-resident multi-kernel sequence of pf_gd_large.cu (MLP backward on the fp64 tensor pipe)."""
    nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), None]
    g = torch.Generator(device=device).manual_seed(7)
    f_ext = torch.randn(plan.ndof, generator=g, device=device, dtype=torch.float64) * 1e-3
    theta0 = _theta0(1, device)[:, : nets[0].n_params + nets[1].n_params].contiguous()
    md = np.arange(2 * plan.nnode - 64, 2 * plan.nnode, dtype=np.int64)
    mv = np.linspace(-1e-3, 1e-3, md.size)
    kw = dict(max_iterations=iters, tolerance=0.0, learning_rate_u=1e-5, learning_rate_theta=5e-4, alpha_physics=1.0,
              alpha_data=100.0, load_factor=1.0)
    best = None
    for rep in range(2):
        theta = theta0.clone()
        u = torch.zeros((1, plan.ndof), dtype=torch.float64, device=device)
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = ops.gd_solve(plan, nets, [1.0, 1.0, 1.0], theta, u, f_ext, md, mv, **kw)
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    assert int(res.n_iters[0]) == iters
    t = torch.tensor([best], device=device, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"workload": f"{plan.nelem}-element lattice, one inverse problem per GPU, E and A = MLPs (837 parameters) at "
                        f"every centroid, 64 measured DOFs, {iters} iterations (tolerance 0), history recorded",
            "dtype": "f64", "ms_per_iteration": ms / iters, "iters_per_s": world * iters / (ms * 1e-3),
            "element_evals_per_s": world * iters * plan.nelem / (ms * 1e-3),
            "gpu_launches": 2 * iters * 12}


def _batched_large_inputs(plan, device, B, first_problem):
    """Problem p (global index) is seeded by p: theta_0 from torch.manual_seed(p) through the
    reference's SimpleNN initialisation, u_0 ~ U(-1e-3, 1e-3) on the free DOFs from default_rng(p)."""
    from pinn_fem_b200.examples.json.generic import SimpleNN

    thetas, us = [], []
    free = np.ones(plan.ndof, dtype=bool)
    free[np.asarray(plan.fixed_dofs)] = False
    for p in range(first_problem, first_problem + B):
        torch.manual_seed(p)
        nets = [SimpleNN(2, w, 3) for w in (20, 15)]
        thetas.append(torch.cat([q.detach().reshape(-1).double() for n in nets for q in n.parameters()]))
        us.append(np.random.default_rng(p).uniform(-1e-3, 1e-3, plan.ndof) * free)
    return torch.stack(thetas).pin_memory(), torch.as_tensor(np.stack(us)).pin_memory()


def gd_batched_large_mesh(device, plan, world=1, rank=0, problems_per_gpu=64, iters=20, dmma_peak_tflops=None):
    """BASELINE.json configs[4] as stated: batched inverse problems on the 999,941-element lattice.  Every GPU
    advances `problems_per_gpu` independent PINN-GD problems together (E and A = the 521- and 316-parameter MLPs at
    every centroid, per-problem theta, u, Adam state): fragment MLP kernels + patch-staged residual / K r / material
    VJP.  Also measured end to end from pinned host buffers (theta_0, u_0 in; theta, u, history out)."""
    nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), None]
    B = problems_per_gpu
    g = torch.Generator(device=device).manual_seed(7)
    f_ext = torch.randn(plan.ndof, generator=g, device=device, dtype=torch.float64) * 1e-3
    theta_h, u_h = _batched_large_inputs(plan, device, B, rank * B)
    md = np.arange(2 * plan.nnode - 64, 2 * plan.nnode, dtype=np.int64)
    mv_h = torch.as_tensor(np.stack([np.linspace(-1e-3, 1e-3, md.size) * (1.0 + 0.01 * p) for p in range(rank * B, rank * B + B)])).pin_memory()
    kw = dict(tolerance=0.0, learning_rate_u=1e-5, learning_rate_theta=5e-4, alpha_physics=1.0, alpha_data=100.0,
              load_factor=1.0)

    def sync_max(x):
        t = torch.tensor([x], device=device, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # device-resident: the per-iteration cost is the difference of two solves of different length (removes set-up,
    # layout changes and the reactions pass)
    theta_d, u_d, mv_d = theta_h.to(device), u_h.to(device), mv_h.to(device)
    ops.gd_solve(plan, nets, [1.0, 1.0, 1.0], theta_d.clone(), u_d.clone(), f_ext, md, mv_d, max_iterations=3, **kw)
    times = {}
    for n_it in (iters, 2 * iters):
        th, uu = theta_d.clone(), u_d.clone()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = ops.gd_solve(plan, nets, [1.0, 1.0, 1.0], th, uu, f_ext, md, mv_d, max_iterations=n_it, **kw)
        e1.record()
        torch.cuda.synchronize(device)
        times[n_it] = sync_max(e0.elapsed_time(e1))
    assert int(res.n_iters.min()) == 2 * iters
    ms_group_it = (times[2 * iters] - times[iters]) / iters
    del res, th, uu

    # end to end through the host-buffer entry point: pinned theta_0, u_0, targets in; theta, u, history out
    out_theta = torch.empty_like(theta_h).pin_memory()
    out_u = torch.empty_like(u_h).pin_memory()
    out_hist = torch.empty((B, iters, 7), dtype=torch.float64).pin_memory()
    torch.cuda.synchronize(device)
    import time

    t0 = time.perf_counter()
    res = ops.gd_solve_host(plan, nets, [1.0, 1.0, 1.0], theta_h, u_h, f_ext, md, mv_h, out_theta=out_theta, out_u=out_u,
                            out_history=out_hist, max_iterations=iters, **kw)
    e2e_s = sync_max(time.perf_counter() - t0)
    flops = sum(_mlp_flops(n) for n in nets[:2]) * plan.nelem  # GEMM flops of one problem-iteration (tanh not counted)
    tfl = flops / (ms_group_it / B * 1e-3) / 1e12
    out = {"workload": f"{plan.nelem}-element lattice x {B} inverse problems per GPU ({world * B} total), E and A = MLPs "
                       f"(837 parameters per problem) at every centroid, 64 measured DOFs, Adam on u and theta, history "
                       f"recorded; problem p seeded by p (theta_0: torch.manual_seed(p), u_0: default_rng(p))",
           "dtype": "f64", "scaling": "weak", "problems_per_gpu": B,
           "ms_per_iteration_of_the_group": ms_group_it, "ms_per_problem_iteration": ms_group_it / B,
           "problem_iters_per_s": world * B / (ms_group_it * 1e-3),
           "element_evals_per_s": world * B * plan.nelem / (ms_group_it * 1e-3),
           "roofline": {"bound": "fp64 pipe (DMMA and DFMA share it)", "achieved": tfl, "unit": "TFLOP/s",
                        "peak": dmma_peak_tflops, "frac": (tfl / dmma_peak_tflops) if dmma_peak_tflops else None,
                        "flops_per_problem_iteration": flops,
                        "note": "algorithmic multiply-adds of the two networks' forward and backward passes only; the 70 "
                                "tanh per point and iteration (about half of the executed fp64 work) are not counted"},
           "e2e_inverse": {"value": world * B * iters / e2e_s, "unit": "problem_iterations/s", "iterations": iters,
                           "h2d_bytes_per_solve": int(theta_h.numel() + u_h.numel() + mv_h.numel()) * 8,
                           "d2h_bytes_per_solve": int(out_theta.numel() + out_u.numel() + out_hist.numel()) * 8,
                           "api": "ops.gd_solve_host -> pf_gd_solve (pinned host theta_0, u_0, targets in; theta, u, "
                                  "history out; includes set-up, layout changes and the copies)",
                           "device_resident_equivalent": world * B * iters / (times[iters] * 1e-3)},
           "gpu_launches": 3 * 14 * iters + 14 * iters}
    return out


def _mlp_flops(spec):
    """2 * MACs of forward + backward of one SimpleNN evaluation (examples/json/generic.py:118-142)."""
    w, i = spec.width, spec.input_dim
    fwd = i * w + (spec.hidden_layers - 1) * w * w + w
    bwd = fwd + (spec.hidden_layers - 1) * w * w  # weight gradients of every layer + back-propagation through the hidden ones
    return 2 * (fwd + bwd)


def gd_element_sharded_iterations_per_second(device, nodes, elements, fixed, world, iters=100):
    """The same large-mesh PINN-GD problem as ``gd_large_mesh_iterations_per_second`` but ONE problem sharded by
    element over all ranks (strong scaling): row bands of the lattice, halo exchange of u and r with the two
    neighbours and one all-reduce of [dL/dtheta | losses] per iteration over NCCL/NVLink."""
    import torch.distributed as dist

    from pinn_fem_b200.element_sharding import Communicator, ShardedMesh, gd_solve_element_sharded

    comm = Communicator(device)
    mesh = ShardedMesh(nodes, elements, fixed, comm)
    nets = [ops.NetSpec(3, 2, 20), ops.NetSpec(3, 2, 15), None]
    nnode = len(nodes)
    g = torch.Generator(device=device).manual_seed(7)
    f_ext = (torch.randn(2 * nnode, generator=g, device=device, dtype=torch.float64) * 1e-3).cpu().numpy()
    theta0 = _theta0(1, device)[0, : nets[0].n_params + nets[1].n_params].contiguous()
    md = np.arange(2 * nnode - 64, 2 * nnode, dtype=np.int64)
    mv = np.linspace(-1e-3, 1e-3, md.size)
    kw = dict(max_iterations=iters, tolerance=0.0, learning_rate_u=1e-5, learning_rate_theta=5e-4, alpha_physics=1.0,
              alpha_data=100.0, load_factor=1.0)
    u0 = np.zeros(2 * nnode)
    best = None
    for rep in range(2):
        dist.barrier()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = gd_solve_element_sharded(mesh, nets, [1.0, 1.0, 1.0], theta0, u0, f_ext, md, mv, **kw)
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1)
        if best is None or ms < best:
            best, loop_ms = ms, out["solve_ms"]
    assert out["n_iters"] == iters
    # correctness at this rank count: the same problem through the single-GPU loop on this rank's own GPU
    # (every rank holds the whole mesh for this check); history is replicated, u is compared on the owned rows
    full = AssemblyPlan(nodes, elements, fixed, device=device)
    one = ops.gd_solve(full, nets, [1.0, 1.0, 1.0], theta0[None].clone(),
                       torch.zeros((1, 2 * nnode), dtype=torch.float64, device=device),
                       torch.as_tensor(f_ext).to(device), md, mv, **kw)
    relerr = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
    owned = torch.as_tensor(out["owned_dofs"]).to(device)
    errs = torch.tensor([relerr(out["history"][:, 1:], one.history[0, :iters, 1:]), relerr(out["u_owned"], one.u[0][owned]),
                         relerr(out["theta"], one.theta[0])], device=device, dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    del full, one
    t = torch.tensor([best], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    t2 = torch.tensor([loop_ms], device=device, dtype=torch.float64)
    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    res = {"workload": f"{len(elements)}-element lattice, ONE inverse problem sharded by element over {world} GPUs "
                       f"({mesh.local.n_owned} owned + {mesh.local.halo.size} halo nodes on rank {comm.rank}), "
                       f"{iters} iterations", "scaling": "strong", "dtype": "f64", "ms_per_iteration": ms / iters,
           "iters_per_s": iters / (ms * 1e-3), "loop_only_ms_per_iteration": float(t2.item()) / iters,
           "transport": comm.transport,
           "parity_vs_single_gpu_loop": {"history_rel_err": float(errs[0]), "u_owned_rel_err": float(errs[1]),
                                         "theta_rel_err": float(errs[2]), "tolerance": 1e-10,
                                         "ok": bool(float(errs.max()) < 1e-10),
                                         "what": f"{iters} iterations of the same problem on one GPU (all history columns, "
                                                 "owned rows of u, theta), max over ranks"},
           "collectives_per_iteration": ("2 halo exchanges (one kernel each: stores into the neighbours' mailboxes over "
                                         "NVLink + epoch flag) + 1 all-reduce of 840 doubles fused into the Adam kernel"
                                         if comm.transport == "peer" else
                                         "2 halo exchanges (ncclSend/Recv group) + 1 all-reduce of 840 doubles"),
           "timed_region_includes": "host->device staging of the local vectors and the final halo exchange"}
    mesh.close()
    comm.close()
    return res


def gn_iteration_leg(golden_inputs_dir, iters=10):
    """seconds per Gauss-Newton / LM iteration on example 10's model (4-node bar, three
    networks, J 6 x 1001, targets u = [1, 2, 3] at DOFs [2, 4, 6], line search on) through
    fem.nn_solver.solve_pinn_newton_raphson, next to the CPU restatement of the reference's
    compute_jacobian_blocks + nn_solver.py:266-277 (the reference itself is not on the GPU box; BASELINE.md
    records 0.12 s per compute_jacobian_blocks call for it on the survey container)."""
    import io
    import json
    import tempfile
    import time
    from contextlib import redirect_stdout
    from pathlib import Path

    from oracle import pinnfem_oracle as O
    from pinn_fem_b200.examples.json import generic
    from pinn_fem_b200.fem import PINNSolverConfig, solve_pinn_newton_raphson

    src = Path(golden_inputs_dir) / "example10.json"
    mv, md = np.array([1.0, 2.0, 3.0]), [2, 4, 6]
    out = {"workload": "example 10's model: 3 elements, 3 networks (998 parameters, 837 reach the loss), J 6 x 1001, "
                       "line search with 15 trial steps evaluated as one batch, LM step through the 6 x 6 dual system"}
    with tempfile.TemporaryDirectory() as tmp:
        dst = Path(tmp) / src.name
        dst.write_text(src.read_text())
        best = None
        for rep in range(3):
            torch.manual_seed(0)
            with redirect_stdout(io.StringIO()):
                model = generic.parse_problem(str(dst))["model"]
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res = solve_pinn_newton_raphson(model, model.loads, mv, md, PINNSolverConfig(max_iterations=iters, tolerance=0.0))
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out["gpu_s_per_iteration"] = best / len(res.history)
        out["iterations"] = len(res.history)
        out["host_syncs_per_iteration"] = 1
    # CPU: the oracle's restatement of compute_jacobian_blocks (closed form, no autograd) + the n x n LM solve
    SPECS = [O.NetSpec(3, 2, 20), O.NetSpec(3, 2, 15), O.NetSpec(3, 2, 10)]
    torch.manual_seed(0)
    nets = [generic.SimpleNN(2, w, 3) for w in (20, 15, 10)]
    th = [torch.cat([q.detach().reshape(-1).double() for q in n.parameters()]).numpy() for n in nets]
    mesh = O.Mesh(NODES, ELEMENTS, LOADS, FIXED)
    mat = O.MaterialNets(*[(sp, t.copy(), 1.0) for sp, t in zip(SPECS, th)])
    u = np.array([0, 0, 0.5, 0, 1.0, 0, 1.5, 0])
    cpu = None
    for rep in range(5):
        t0 = time.perf_counter()
        j_uu, j_ut, r_p, j_du = O.jacobian_blocks(mesh, mat, u, mesh.loads, np.array(md), 1.0)
        J, R = O.gauss_newton_system(j_uu, j_ut, r_p, j_du, mv - u[md], 1.0, 1.0)
        O.lm_step(J, R)
        dt = time.perf_counter() - t0
        cpu = dt if cpu is None else min(cpu, dt)
    out["cpu_port_s_per_jacobian_and_lm_solve"] = cpu
    out["cpu_note"] = ("numpy restatement (closed-form Jacobian, 1001 x 1001 LU), one core, no line search; the reference's own "
                       "compute_jacobian_blocks (998 + 3 autograd passes) took 0.12 s per call on the survey container (BASELINE.md)")
    return out


def example_runs(golden_inputs_dir, names=("example1", "example4-P", "example7-P")):
    """BASELINE.json configs[0..2]: whole ``generic.py`` solves (parse + solve + post-processing, seed 0) of the
    reference's own example files, timed warm (second run) on the wall clock, next to the reference's timings
    recorded in BASELINE.md (CPU)."""
    import io
    import json
    import time
    from contextlib import redirect_stdout
    from pathlib import Path
    import logging
    import tempfile

    from pinn_fem_b200.examples.json import generic
    from pinn_fem_b200.fem import solver

    ref = {"example1": {"wall_s": 7.5, "note": "BASELINE.md: 7.5 s here, ~all interpreter start-up; README ~1 s"},
           "example4-P": {"wall_s": 33.8, "iterations": 1874, "iters_per_s": 55.0},
           "example7-P": {"wall_s": 39.7, "iterations": 2140, "iters_per_s": 54.0}}
    out = {}
    logging.disable(logging.CRITICAL)
    try:
        with tempfile.TemporaryDirectory() as tmp:
            for name in names:
                src = Path(golden_inputs_dir) / f"{name}.json"
                if not src.exists():
                    continue
                dst = Path(tmp) / src.name
                dst.write_text(src.read_text())
                best = None
                for rep in range(2):
                    torch.manual_seed(0)
                    solver.COUNTERS["gd_iterations"] = 0
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    with redirect_stdout(io.StringIO()):
                        res = generic.solve_problem(generic.parse_problem(str(dst)))
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                total = solver.COUNTERS["gd_iterations"]  # GD iterations over all increments and phases
                out[name] = {"wall_s": best, "converged": bool(res["converged"]),
                             "result_iterations": res["iterations"], "gd_iterations_total": total,
                             "gd_iters_per_s": (total / best) if total else None, "reference_cpu": ref.get(name)}
    finally:
        logging.disable(logging.NOTSET)
    return out
