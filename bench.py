#!/usr/bin/env python
"""bench.py -- FCT steps/s of the drift-control advection PDECO on the 4096^2 P1 mesh (BASELINE.json config 5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells 4096] [--nt 4]

One bench "step" = one gradient-iteration pass of the hot path over N_t time levels of the synthetic
4097^2-DoF problem: state sweep (N_t x [drift-operator assembly + FCT step]) + adjoint sweep (N_t x [assembly
+ M(uhat-u) + FCT step]) + gradient (N_t+1 x [load vector + SpMV + 20 Chebyshev iterations]) + cost functional,
all device-resident.  `value` = FCT steps (state + adjoint) per second, whole job.  `e2e` = the same metric
through the host-buffer C-ABI call (fct_advdrift_state_host: control slices H2D, state slices D2H inside the
timed region, pinned host memory).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def step_bytes(n, nnz, cells, k_j):
    """algorithmic bytes of one FCT step, SURVEY.md App. E / DESIGN.md"""
    return (344 + 12 * k_j) * nnz + (1108 + 28 * k_j) * n + 12 * cells


def cheb_iter_bytes(n, nnz):
    """one Chebyshev iteration: M values + column indices + rowptr + 5 vectors (App. E, P4)"""
    return 12 * nnz + 4 * n + 5 * 8 * n


class ClockSampler:
    """samples nvidia-smi SM clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def rd():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=rd, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """one host core: `nsteps` FCT state steps of the oracle port on a sample mesh; returns the step-loop seconds"""
    sample_cells, nsteps = args
    from oracle import pdeco_numpy as drv
    orc = drv.AdvectionDriftPDECO(sample_cells, 0.0, 1.0, solver="jacobi")
    h = 1.0 / sample_cells
    dt = 0.25 * h / (2 * np.sqrt(2))
    u0 = orc.gaussian_ic()
    c = np.full((nsteps + 1, orc.nodes), 1.0)
    xy = orc.mesh.dof_xy
    c += 0.5 * np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1])
    t0 = time.perf_counter()
    orc.state(c, u0, nsteps, dt)
    return time.perf_counter() - t0


def cpu_port_steps_per_s(cells_full, sample_cells, nsteps, cores=None):
    """CPU arm.  The reference's FCT step is single-threaded Python/scipy (and its `FCT_alg_ref` cannot run beyond
    ~1e5 DoF because of an O(n^2) todense(), a direct solve is impractical beyond ~1M DoF: BASELINE.md), so the
    baseline is the oracle's vectorised port with the Jacobi twin of the GPU solver, run as `cores` independent
    replicas (one per host core, i.e. perfect-scaling credit for the host) on a `sample_cells`^2 mesh and scaled
    linearly in DoF to the full mesh (every pass of the Jacobi-based step is O(nnz)).
    Returns (steps/s at full size using all cores, steps/s of the sample on all cores, wall seconds, cores)."""
    import multiprocessing as mp
    if cores is None:
        try:
            cores = len(os.sched_getaffinity(0))
        except Exception:
            cores = os.cpu_count() or 1
        cores = max(1, min(cores, 32))
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    t0 = time.perf_counter()
    if cores == 1:
        times = [_cpu_worker((sample_cells, nsteps))]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            times = pool.map(_cpu_worker, [(sample_cells, nsteps)] * cores)
    wall = time.perf_counter() - t0
    sps_sample = cores * nsteps / max(times)          # replicas run concurrently: aggregate over the slowest one
    scale = (sample_cells + 1) ** 2 / float((cells_full + 1) ** 2)
    return sps_sample * scale, sps_sample, wall, cores


CPU_SAMPLE_CELLS, CPU_SAMPLE_STEPS = 1024, 3
CPU_C_STEPS = 3
_C_PROBLEM = {}
BETA, C_LOWER, C_UPPER = 0.01, 0.0, 5.0      # advection_solidbody_FCT_PDECO_alltime.py:47-50


def workload_string(cells):
    """ONE workload name for both arms (the driver compares the strings)"""
    N = cells + 1
    n = N * N
    nnz = n + 2 * (2 * N * (N - 1) + (N - 1) ** 2)
    return (f"synthetic {cells}^2-cell unit-square drift-control advection FCT PDECO (BASELINE config 5): "
            f"{n} DoF, {nnz} nnz, fp64")


def _cpu_sample_text(cells_full, sample, nsteps, cores, sps_sample, wall):
    return (f"oracle port (numpy/scipy CSR, Jacobi low-order solve): {cores} independent replicas (one per host core) of "
            f"{nsteps} FCT state steps on a {sample}^2-cell mesh = {sps_sample:.3f} steps/s aggregate, {wall:.1f} s wall "
            f"incl. set-up; scaled by the DoF ratio to {cells_full}^2")


def synth_fields(xy):
    """initial condition and (time-constant) initial control of config 5 in DoF order"""
    x, y = 2 * xy[:, 0] - 1, 2 * xy[:, 1] - 1
    u0 = np.exp(-20 * ((x + 2 / 3) ** 2 + 5 * (y + 5 / 6) ** 2))
    c = 1.0 + 0.25 * np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1])
    return u0, c


def synth_target_slice(xy, k, dt):
    """target state at time level k: the IC translated by the exact drift of the constant control c = 2"""
    x, y = 2 * xy[:, 0] - 1, 2 * xy[:, 1] - 1
    s = 2.0 * k * dt * 2          # shift in [-1,1] coordinates
    return np.exp(-20 * ((x - s + 2 / 3) ** 2 + 5 * (y - s + 5 / 6) ** 2))


def cpu_c_port(cells_full, nsteps, keep=False):
    """CPU arm, preferred: the C/OpenMP twin of the oracle (oracle/fct_c.c, pinned on the numpy oracle by
    tests/test_oracle_c.py) runs `nsteps` FCT state steps of the FULL-size problem on all host cores -- the same
    workload as the GPU arm (drift-operator assembly + FCT step with the Jacobi low-order solve), no extrapolation.
    Returns (steps/s, seconds of the step loop, threads, jacobi sweeps per step, problem, trajectory | None) or None
    when no C compiler works."""
    try:
        from oracle import fct_c
        fct_c.lib(native=True)
        threads = fct_c.use_all_host_threads()          # not OMP_NUM_THREADS: torchrun exports 1 to its ranks
        prob = _C_PROBLEM.get(cells_full)
        if prob is None:
            prob = _C_PROBLEM[cells_full] = fct_c.CDriftProblem(cells_full, 0.0, 1.0)
    except Exception as e:  # noqa: BLE001
        sys.stderr.write(f"bench.py: C/OpenMP oracle unavailable ({e}); using the numpy port\n")
        return None
    u0, c = synth_fields(prob.dof_xy)
    dt = 0.25 * (1.0 / cells_full) / (2 * np.sqrt(2))
    t0 = time.perf_counter()
    traj, sweeps = prob.state(np.tile(c, nsteps + 1), u0, nsteps, dt)
    el = time.perf_counter() - t0
    return nsteps / el, el, threads, sweeps / nsteps, prob, (traj if keep else None)


def cpu_baseline_record(cells_full, nsteps=CPU_C_STEPS, keep=False):
    """(value steps/s, dict for the `cpu_baseline` key, wall seconds of the sample, (problem, trajectory) | None)"""
    r = cpu_c_port(cells_full, nsteps, keep)
    if r is not None:
        sps, el, threads, kj, prob, traj = r
        txt = (f"C/OpenMP port of the oracle (oracle/fct_c.c: drift-operator assembly + FCT step, Jacobi low-order solve, "
               f"{kj:.1f} sweeps/step) on {threads} host threads: {nsteps} FCT state steps of the full {cells_full}^2-cell "
               f"problem in {el:.1f} s, no extrapolation")
        return sps, {"value": sps, "unit": "steps/s", "cores": threads, "kind": "port", "sample": txt}, el, \
            ((prob, traj) if keep else None)
    sample = min(cells_full, CPU_SAMPLE_CELLS)
    sps, sps_sample, wall, cores = cpu_port_steps_per_s(cells_full, sample, CPU_SAMPLE_STEPS)
    return sps, {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port",
                 "sample": _cpu_sample_text(cells_full, sample, CPU_SAMPLE_STEPS, cores, sps_sample, wall)}, wall, None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = []
    for i in range(args.warmup + args.steps):
        # every bench step is one bounded sample of the workload; warm-up samples are one FCT step (page the code in)
        r = cpu_baseline_record(args.cells, CPU_C_STEPS if i >= args.warmup else 1)
        if i >= args.warmup:
            res.append(r)
    sps = float(np.mean([r[0] for r in res]))
    cb = dict(res[-1][1])
    cb["value"] = sps
    line = {"impl": "reference", "metric": "FCT steps/sec", "value": sps, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([r[2] for r in res])),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(args.cells)},
            "cpu_baseline": cb,
            "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
PARITY_TOL_FIELD, PARITY_TOL_COST = 1e-12, 1e-9      # BASELINE.json north_star: rel-L2 per time step, final cost


def parity_vs_cpu_port(ctx, prob, cpu_traj, dt):
    """the GPU state trajectory against the C/OpenMP oracle port on the SAME full-size inputs
    (u0, c, dt of the CPU leg): rel-L2 per time step <= 1e-12, cost functional <= 1e-9 (relative)."""
    ns = cpu_traj.shape[0] - 1
    n = prob.nodes
    u0, c = synth_fields(prob.dof_xy)
    d_c = ctx.array(np.tile(c, ns + 1))
    utr = np.zeros((ns + 1) * n); utr[:n] = u0
    d_u = ctx.array(utr)
    ctx.advdrift_state(d_c, d_u, ns, dt)
    gpu = d_u.download().reshape(ns + 1, n)
    errs = [float(np.linalg.norm(gpu[k] - cpu_traj[k]) / np.linalg.norm(cpu_traj[k])) for k in range(1, ns + 1)]
    uhat = np.array([synth_target_slice(prob.dof_xy, k, dt) for k in range(ns + 1)])
    d_uhat = ctx.array(uhat.ravel())
    M = ctx.static()[0]
    J_gpu = 0.5 * ctx.norm_sq_Q(M, d_u, ns, dt, target=d_uhat) + BETA / 2 * ctx.norm_sq_Q(M, d_c, ns, dt)
    J_cpu = 0.5 * prob.norm_sq_Q(cpu_traj, ns, dt, target=uhat) + BETA / 2 * prob.norm_sq_Q(np.tile(c, ns + 1), ns, dt)
    for a in (d_c, d_u, d_uhat):
        a.free()
    cost_rel = abs(J_gpu - J_cpu) / abs(J_cpu)
    ok = bool(max(errs) <= PARITY_TOL_FIELD and cost_rel <= PARITY_TOL_COST)
    return {"against": "oracle/fct_c.c (C/OpenMP port of the oracle) on the same u0, c, dt at full size",
            "time_steps": ns, "rel_l2_per_step": errs, "tol_rel_l2": PARITY_TOL_FIELD, "cost_gpu": J_gpu, "cost_cpu": J_cpu,
            "cost_rel_diff": cost_rel, "tol_cost": PARITY_TOL_COST, "ok": ok}


def load_kernel_profile():
    """per-kernel ncu figures of one FCT step at HEAD (profiles/r2_kernels.json, written by tools/kernels_table.py
    from an `ncu` capture of tools/kprof_step.py): {kernel: {launches, ms, dram_bytes, GBs, pct_of_peak}}"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_kernels.json")))
    except Exception:  # noqa: BLE001
        return None


# kernels of ONE state FCT step on the fused-tile path and how often each runs (16 Jacobi sweeps = 4 launches of 4, 20 Chebyshev
# iterations = 1 + 4 launches of 5/5/5/4): with the per-launch DRAM bytes of profiles/r2_kernels.json this gives the bytes a
# step really moves, the whole-step counterpart of the per-kernel roofline
STATE_STEP_KERNELS = {"k_drift_low_build": 1, "k_tile": 4, "k_spmv": 1, "k_cheb_first": 1, "k_cheb_tile": 4, "k_flux_limits": 1,
                      "k_flux_apply": 1}


def step_roofline(kprof, state_ms, peak, appE_gb):
    out = {"appE_GB_per_fct_step": appE_gb, "note": "appE = SURVEY App. E accounting unit, for reference only"}
    ks = (kprof or {}).get("kernels") or {}
    if all(k in ks for k in STATE_STEP_KERNELS):
        gb = sum(ks[k]["dram_bytes_per_launch"] * c for k, c in STATE_STEP_KERNELS.items()) / 1e9
        kernel_ms = sum(ks[k]["ms_per_launch"] * c for k, c in STATE_STEP_KERNELS.items())
        out.update({"dram_GB_per_state_step": gb, "GBs": gb / (state_ms * 1e-3), "frac_of_peak": gb / (state_ms * 1e-3) / peak,
                    "ncu_kernel_ms_per_state_step": kernel_ms,
                    "how": "sum over the kernels of one state step of ncu dram bytes per launch (profiles/r2_kernels.json) / "
                           "the CUDA-event time of a state step measured in this run; the fused tile kernels moved the step "
                           "off the HBM roofline (they are bound by shared-memory latency and instruction issue, DESIGN.md 4)"})
    return out


class GradientIteration:
    """one projected-gradient iteration of advection_solidbody_FCT_PDECO_alltime.py:196-303 on device trajectories:
    state sweep, adjoint sweep, gradient (N_t+1 load vectors + ChebSI), ONE Armijo trial (clip, forward solve, cost,
    ||c_inc - c||^2), cost functional.  The control is not updated, so every bench step does identical work."""

    def __init__(self, ctx, n, nt, dt, d_c, d_u, d_uhat):
        self.ctx, self.n, self.nt, self.dt = ctx, n, nt, dt
        L = (nt + 1) * n
        self.d_c, self.d_u, self.d_uhat = d_c, d_u, d_uhat
        self.d_p, self.d_d = ctx.empty(L), ctx.empty(L)
        self.d_cinc, self.d_utrial = ctx.empty(L), ctx.empty(L)
        ctx.axpby(1.0, d_u, 0.0, None, self.d_utrial, length=n)      # IC of the trial trajectory
        self.M = ctx.static()[0]
        self.sweeps = []
        self.J = self.J_trial = None

    def cost(self, d_u, d_c):
        ctx = self.ctx
        return (0.5 * ctx.norm_sq_Q(self.M, d_u, self.nt, self.dt, target=self.d_uhat)
                + BETA / 2 * ctx.norm_sq_Q(self.M, d_c, self.nt, self.dt))

    def __call__(self):
        ctx, nt, dt = self.ctx, self.nt, self.dt
        s1 = ctx.advdrift_state(self.d_c, self.d_u, nt, dt)
        s2 = ctx.advdrift_adjoint(self.d_c, self.d_u, self.d_uhat, self.d_p, nt, dt)
        ctx.advdrift_gradient(self.d_c, self.d_u, self.d_p, self.d_d, nt, BETA)
        self.J = self.cost(self.d_u, self.d_c)
        # Armijo trial k = 0 (old_helpers.py:40-80): s = 1, c_inc = clip(c + s d), forward solve, cost, ||c_inc - c||^2_Q
        ctx.clip_axpy(self.d_c, 1.0, self.d_d, C_LOWER, C_UPPER, self.d_cinc)
        s3 = ctx.advdrift_state(self.d_cinc, self.d_utrial, nt, dt)
        self.J_trial = self.cost(self.d_utrial, self.d_cinc)
        self.cdiff = ctx.norm_sq_Q(self.M, self.d_cinc, nt, dt, target=self.d_c)
        self.sweeps.append((s1, s2, s3))
        return self.J

    fct_steps_per_pass = property(lambda self: 3 * self.nt)


def run_gpu_arm(args):
    import fem_fct_pdeco_b200 as fp
    from fem_fct_pdeco_b200.mesh import RectMeshP1

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from fem_fct_pdeco_b200 import distributed as dist_mod
        return dist_mod.bench_multi(args, rank, world, local_rank)

    if fp.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    n_cells, nt = args.cells, args.nt
    h = 1.0 / n_cells
    dt = 0.25 * h / (2 * np.sqrt(2))
    t_setup = time.perf_counter()
    mesh = RectMeshP1(n_cells, 0.0, 1.0)
    ctx = mesh.context(device=local_rank)
    n, nnz, ncell = mesh.nodes, mesh.nnz, mesh.ncells
    u0, c0 = synth_fields(mesh.dof_xy)
    L = (nt + 1) * n
    d_c, d_u, d_uhat = ctx.empty(L), ctx.empty(L), ctx.empty(L)
    for k in range(nt + 1):          # slice by slice: no (nt+1) x n host arrays
        d_c.slice(k * n, n).upload(c0)
        d_uhat.slice(k * n, n).upload(synth_target_slice(mesh.dof_xy, k, dt))
    d_u.slice(0, n).upload(u0)
    it = GradientIteration(ctx, n, nt, dt, d_c, d_u, d_uhat)
    t_setup = time.perf_counter() - t_setup

    for _ in range(args.warmup):
        it()
    ctx.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = ctx.event(), ctx.event()
    del it.sweeps[:]
    ctx.record(e0)
    for _ in range(args.steps):
        J = it()
    ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1)
    launches = ctx.launch_count() - l0
    fct_steps = it.fct_steps_per_pass * args.steps
    value = fct_steps / (ms * 1e-3)
    k_state = np.mean([s[0] for s in it.sweeps]) / nt
    k_adj = np.mean([s[1] for s in it.sweeps]) / nt
    k_trial = np.mean([s[2] for s in it.sweeps]) / nt

    # ---- a state sweep on its own (what `e2e` is the host-buffer twin of) ---------------------------------------
    ctx.record(e0)
    ctx.advdrift_state(d_c, d_u, nt, dt)
    ctx.record(e1)
    state_ms = ctx.elapsed_ms(e0, e1) / nt

    # ---- dominant kernel timed live with CUDA events on the library's stream -----------------------------------
    peak, peak_src = _peaks()
    kern = ctx.bench_dominant_kernel(d_c.slice(0, n), d_u.slice(0, n), dt)
    kprof = load_kernel_profile()
    dom = kern["name"]
    traffic = None
    if kprof and dom in kprof.get("kernels", {}):
        traffic = kprof["kernels"][dom].get("dram_bytes_per_launch")
    ach = kern["bytes_per_launch"] / (kern["ms_per_launch"] * 1e-3) / 1e9

    # ---- e2e: host buffers through the C-ABI (H2D of control slices, D2H of state slices inside) -----
    nt_e2e = min(nt, args.e2e_nt)
    Le = (nt_e2e + 1) * n
    hc = ctx.pinned(Le); hu = ctx.pinned(Le)
    hc.reshape(nt_e2e + 1, n)[:] = c0
    hu[:] = 0.0; hu[:n] = u0
    ctx.advdrift_state_host(hc, hu, nt_e2e, dt)          # warm-up
    t0 = time.perf_counter()
    reps_e2e = max(1, args.steps // 2)
    for _ in range(reps_e2e):
        ctx.advdrift_state_host(hc, hu, nt_e2e, dt)      # synchronises before returning
    e2e_s = time.perf_counter() - t0
    e2e_value = nt_e2e * reps_e2e / e2e_s
    clocks = sampler.stop()

    k_mean = (k_state + k_adj + k_trial) / 3
    step_gb = step_bytes(n, nnz, ncell, k_mean) / 1e9
    line = {
        "metric": "FCT steps/sec", "value": value, "unit": "steps/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(n_cells),
                   "bench_step": f"one projected-gradient iteration over {nt} time levels: state sweep + adjoint sweep + "
                                 f"gradient ({nt + 1} load vectors + ChebSI) + one Armijo trial (clip, forward sweep, cost) "
                                 f"+ cost functional = {it.fct_steps_per_pass} FCT steps, device-resident",
                   "time_levels": nt, "dt": dt, "l2_flush": "working set per FCT step >> L2 (matrix values alone are "
                   f"{8 * nnz / 1e6:.0f} MB per array)",
                   "jacobi_sweeps_per_step": {"state": k_state, "adjoint": k_adj, "armijo_trial": k_trial},
                   "cost_functional": J, "cost_trial": it.J_trial, "setup_s": t_setup},
        "roofline": {"bound": "hbm", "kernel": dom, "what": kern["what"],
                     "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                     "peak_source": peak_src, "bytes_per_launch": kern["bytes_per_launch"],
                     "bytes_model": kern["bytes_model"], "ms_per_launch": kern["ms_per_launch"],
                     "traffic_frac_of_peak": (traffic / (kern["ms_per_launch"] * 1e-3) / 1e9 / peak) if traffic else None,
                     "appE_bytes_per_launch": kern["appE_bytes"],
                     "appE_GBs": kern["appE_bytes"] / (kern["ms_per_launch"] * 1e-3) / 1e9,
                     "note": "achieved = the bytes this kernel has to move per launch (bytes_model; DESIGN.md 4) / CUDA-event "
                             "time; traffic = ncu dram__bytes_read+write of one launch (profiles/r2_kernels.json); appE_* = the "
                             "SURVEY App. E accounting unit for the work this launch replaces (not a bound for this "
                             "implementation: templates and fusion delete bytes from it)"},
        "state_step": {"ms": state_ms, "steps_per_s": 1e3 / state_ms,
                       "note": "one FCT state step (assembly + low-order solve + ChebSI + limiter), device-resident"},
        "step_roofline": step_roofline(kprof, state_ms, peak, step_gb),
        "kernels": (kprof or {}).get("kernels"),
        "kernels_source": (kprof or {}).get("source"),
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n,
                "time_levels": nt_e2e,
                "call": "fct_advdrift_state_host (state sweep, pinned host trajectories)"},
        "templates": {"mass_rows": ctx.template_count(), "geometry": ctx.geom_template_count()},
        "gpu_launches": int(launches),
        "gpu_launches_note": "host-enqueued kernels of libfctpdeco in the timed region; CUDA-graph WHILE bodies are counted "
                             "once per solve (executed sweeps: jacobi_sweeps_per_step)",
        "clocks": clocks,
    }
    rc = 0
    if not args.no_cpu:
        _, cb, _, kept = cpu_baseline_record(n_cells, keep=True)
        line["cpu_baseline"] = cb
        if kept is not None:
            prob, traj = kept
            line["parity"] = parity_vs_cpu_port(ctx, prob, traj, dt)
            if not line["parity"]["ok"]:
                sys.stderr.write("bench.py: PARITY VIOLATED against the oracle port at full size\n")
                rc = 3
        else:
            line["parity"] = None
    print(json.dumps(line))
    if rc:
        raise SystemExit(rc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=4096, help="cells per side (4096 = BASELINE config 5)")
    ap.add_argument("--nt", type=int, default=50, help="time levels per bench step (SURVEY.md 8d: 50)")
    ap.add_argument("--e2e-nt", type=int, default=50, help="time levels of the host-trajectory (e2e) sweep")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
