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


def pass_bytes(n, nnz, cells, nt, k_state, k_adj):
    """algorithmic bytes of one bench step (gradient-iteration pass over nt time levels), same accounting unit:
    state + adjoint FCT steps, the adjoint right-hand side M(uhat-u), the gradient slices (load vector + SpMV + 20
    Chebyshev iterations each) and the two cost-functional norms"""
    V = 8 * n
    spmv = 12 * nnz + 4 * n + 3 * V
    cheb20 = 20 * (12 * nnz + 4 * n + 5 * V)
    adj_rhs = 3 * V + spmv
    grad = (12 * cells + 3 * V) + spmv + cheb20
    dots = 2 * (nt + 1) * (12 * nnz + 4 * n + 4 * V)
    return (nt * step_bytes(n, nnz, cells, k_state) + nt * (step_bytes(n, nnz, cells, k_adj) + adj_rhs)
            + (nt + 1) * grad + dots)


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


def _cpu_sample_text(cells_full, sample, nsteps, cores, sps_sample, wall):
    return (f"oracle port (numpy/scipy CSR, Jacobi low-order solve): {cores} independent replicas (one per host core) of "
            f"{nsteps} FCT state steps on a {sample}^2-cell mesh = {sps_sample:.3f} steps/s aggregate, {wall:.1f} s wall "
            f"incl. set-up; scaled by the DoF ratio to {cells_full}^2")


def cpu_c_port(cells_full, nsteps):
    """CPU arm, preferred: the C/OpenMP twin of the oracle (oracle/fct_c.c, pinned on the numpy oracle by
    tests/test_oracle_c.py) runs `nsteps` FCT state steps of the FULL-size problem on all host cores -- the same
    workload as the GPU arm (drift-operator assembly + FCT step with the Jacobi low-order solve), no extrapolation.
    Returns (steps/s, seconds of the step loop, threads, jacobi sweeps per step) or None when no C compiler works."""
    try:
        from oracle.fct_c import CDriftProblem
        prob = _C_PROBLEM.get(cells_full)
        if prob is None:
            prob = _C_PROBLEM[cells_full] = CDriftProblem(cells_full, 0.0, 1.0)
    except Exception as e:  # noqa: BLE001
        sys.stderr.write(f"bench.py: C/OpenMP oracle unavailable ({e}); using the numpy port\n")
        return None
    xy = prob.dof_xy
    u0 = np.exp(-20 * ((2 * xy[:, 0] - 1 + 2 / 3) ** 2 + 5 * (2 * xy[:, 1] - 1 + 5 / 6) ** 2))
    c = 1.0 + 0.25 * np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1])
    dt = 0.25 * (1.0 / cells_full) / (2 * np.sqrt(2))
    t0 = time.perf_counter()
    _, sweeps = prob.state(np.tile(c, nsteps + 1), u0, nsteps, dt)
    el = time.perf_counter() - t0
    return nsteps / el, el, prob.threads(), sweeps / nsteps


def cpu_baseline_record(cells_full, nsteps=CPU_C_STEPS):
    """(value steps/s, dict for the `cpu_baseline` key, wall seconds of the sample)"""
    r = cpu_c_port(cells_full, nsteps)
    if r is not None:
        sps, el, threads, kj = r
        txt = (f"C/OpenMP port of the oracle (oracle/fct_c.c: drift-operator assembly + FCT step, Jacobi low-order solve, "
               f"{kj:.1f} sweeps/step) on {threads} host threads: {nsteps} FCT state steps of the full {cells_full}^2-cell "
               f"problem in {el:.1f} s, no extrapolation")
        return sps, {"value": sps, "unit": "steps/s", "cores": threads, "kind": "port", "sample": txt}, el
    sample = min(cells_full, CPU_SAMPLE_CELLS)
    sps, sps_sample, wall, cores = cpu_port_steps_per_s(cells_full, sample, CPU_SAMPLE_STEPS)
    return sps, {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port",
                 "sample": _cpu_sample_text(cells_full, sample, CPU_SAMPLE_STEPS, cores, sps_sample, wall)}, wall


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
            "config": {"workload": f"synthetic {args.cells}^2-cell unit-square advection FCT PDECO (BASELINE config 5)"},
            "cpu_baseline": cb,
            "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def synth_problem(mesh, nt, dt):
    """config 5: Gaussian IC, initial control c = 1 (+ a smooth perturbation so that the
    drift-mass term is exercised), target = the IC translated by the exact drift of c = 2."""
    xy = mesh.dof_xy
    x, y = 2 * xy[:, 0] - 1, 2 * xy[:, 1] - 1
    u0 = np.exp(-20 * ((x + 2 / 3) ** 2 + 5 * (y + 5 / 6) ** 2))
    c = 1.0 + 0.25 * np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1])
    uhat = []
    for k in range(nt + 1):
        s = 2.0 * k * dt * 2          # shift in [-1,1] coordinates
        uhat.append(np.exp(-20 * ((x - s + 2 / 3) ** 2 + 5 * (y - s + 5 / 6) ** 2)))
    return u0, c, np.array(uhat)


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
    beta = 0.01
    t_setup = time.perf_counter()
    mesh = RectMeshP1(n_cells, 0.0, 1.0)
    ctx = mesh.context(device=local_rank)
    n, nnz, ncell = mesh.nodes, mesh.nnz, mesh.ncells
    u0, c0, uhat = synth_problem(mesh, nt, dt)
    M = ctx.static()[0]
    L = (nt + 1) * n
    d_c = ctx.array(np.tile(c0, nt + 1))
    utr = np.zeros(L); utr[:n] = u0
    d_u = ctx.array(utr)
    d_uhat = ctx.array(uhat.ravel())
    d_p, d_d = ctx.empty(L), ctx.empty(L)
    del utr
    t_setup = time.perf_counter() - t_setup

    sweeps_hist = []

    def gradient_pass():
        s1 = ctx.advdrift_state(d_c, d_u, nt, dt)
        s2 = ctx.advdrift_adjoint(d_c, d_u, d_uhat, d_p, nt, dt)
        ctx.advdrift_gradient(d_c, d_u, d_p, d_d, nt, beta)
        J = 0.5 * ctx.norm_sq_Q(M, d_u, nt, dt, target=d_uhat) + beta / 2 * ctx.norm_sq_Q(M, d_c, nt, dt)
        sweeps_hist.append((s1, s2))
        return J

    for _ in range(args.warmup):
        gradient_pass()
    ctx.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = ctx.event(), ctx.event()
    del sweeps_hist[:]
    ctx.record(e0)
    for _ in range(args.steps):
        J = gradient_pass()
    ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1)
    launches = ctx.launch_count() - l0
    fct_steps = 2 * nt * args.steps
    value = fct_steps / (ms * 1e-3)
    k_state = np.mean([s[0] for s in sweeps_hist]) / nt
    k_adj = np.mean([s[1] for s in sweeps_hist]) / nt

    # ---- kernels timed live with CUDA events on the library's stream ---------------------------------
    # dominant kernel: the Jacobi sweep of the low-order solve (k_J ~ 14 launches per FCT step).  Algorithmic bytes per
    # sweep = 12 nnz + 4 n + 3 V (App. E, P3).
    peak, peak_src = _peaks()
    A_tmp = ctx.empty(nnz)
    ctx.assemble_matrix(2, A_tmp, c0=d_c.slice(0, n), s0=1.0, s1=1.0, scale=-1.0)       # FCT_FORM_DRIFT, state sign
    jac_ms = ctx.bench_jacobi_sweeps(A_tmp, d_u.slice(0, n), dt, reps=20)
    A_tmp.free()
    jb = 12 * nnz + 4 * n + 3 * 8 * n
    jac_gbs = jb / (jac_ms * 1e-3) / 1e9
    # what the template-column, row-scaled sweep actually moves: 8 B/nnz values + 2 B/row code + 4 B/row rowptr + b, x, x_new
    jb_actual = (8 * nnz + 6 * n + 3 * 8 * n) if ctx.template_count() else jb
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["k_jacobi_sweep_tpl"]
        traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"]) if ctx.template_count() else None
    except Exception:
        pass
    # second kernel: one Chebyshev iteration.  With the mass matrix on row templates it moves 2 B/row + 5 V instead of
    # the 12 B/nnz of the accounting unit, so its "algorithmic" rate exceeds the HBM peak; both figures are given.
    Md = ctx.static()[2]
    b, y = d_d.slice(0, n), d_p.slice(0, n)        # scratch slices (overwritten by the next pass anyway)
    reps = 5
    ctx.chebsi(M, Md, b, y, 20)
    ctx.record(e0)
    for _ in range(reps):
        ctx.chebsi(M, Md, b, y, 20)
    ctx.record(e1)
    t20 = ctx.elapsed_ms(e0, e1)
    ctx.record(e0)
    for _ in range(reps):
        ctx.chebsi(M, Md, b, y, 1)                   # the vector-only first iteration (k_cheb_first)
    ctx.record(e1)
    t1 = ctx.elapsed_ms(e0, e1)
    cheb_ms = (t20 - t1) / (reps * 19)               # 19 matrix iterations per ChebSI call
    cb = cheb_iter_bytes(n, nnz)
    ntpl = ctx.template_count()
    cb_actual = (2 * n + 4 * 8 * n) if ntpl else cb     # code + g, y_k (gathered), y_{k-1}, y_{k+1}; diag(M) from the table

    # ---- e2e: host buffers through the C-ABI (H2D of control slices, D2H of state slices inside) -----
    hc = ctx.pinned(L); hu = ctx.pinned(L)
    hc[:] = np.tile(c0, nt + 1); hu[:] = 0.0; hu[:n] = u0
    ctx.advdrift_state_host(hc, hu, nt, dt)          # warm-up
    t0 = time.perf_counter()
    reps_e2e = max(1, args.steps // 2)
    for _ in range(reps_e2e):
        ctx.advdrift_state_host(hc, hu, nt, dt)      # synchronises before returning
    e2e_s = time.perf_counter() - t0
    e2e_value = nt * reps_e2e / e2e_s
    clocks = sampler.stop()

    # whole-step roofline (algorithmic bytes of an FCT step with the sweeps actually executed)
    k_mean = 0.5 * (k_state + k_adj)
    step_gb = step_bytes(n, nnz, ncell, k_mean) / 1e9
    pass_gb = pass_bytes(n, nnz, ncell, nt, k_state, k_adj) / 1e9
    line = {
        "metric": "FCT steps/sec", "value": value, "unit": "steps/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic {n_cells}^2-cell unit-square drift-control advection FCT PDECO "
                               f"(BASELINE config 5): {n} DoF, {nnz} nnz; bench step = state+adjoint sweeps over "
                               f"{nt} time levels + gradient + cost",
                   "time_levels": nt, "dt": dt, "l2_flush": "working set per pass >> L2 (matrix values alone are "
                   f"{8 * nnz / 1e6:.0f} MB)", "jacobi_sweeps_per_step": {"state": k_state, "adjoint": k_adj},
                   "cost_functional": J, "setup_s": t_setup},
        "roofline": {"bound": "hbm", "kernel": "k_jacobi_sweep_tpl (low-order solve, ~14 launches per FCT step)",
                     "achieved": jac_gbs, "peak": peak, "unit": "GB/s", "frac": jac_gbs / peak, "traffic": traffic,
                     "peak_source": peak_src, "bytes_per_launch": jb, "ms_per_launch": jac_ms,
                     "actual_bytes_per_launch": jb_actual, "actual_GBs": jb_actual / (jac_ms * 1e-3) / 1e9,
                     "actual_frac_of_peak": jb_actual / (jac_ms * 1e-3) / 1e9 / peak,
                     "note": "achieved = algorithmic bytes of SURVEY App. E (12 B/nnz CSR sweep) / measured time; the kernel "
                             "takes the column pattern from the row templates and moves fewer bytes (actual_*; traffic = "
                             "ncu dram bytes of one launch, profiles/r1_traffic.json)"},
        "roofline_chebsi": {"kernel": "k_cheb_iter_tpl" if ntpl else "k_cheb_iter", "ms_per_launch": cheb_ms,
                            "row_templates": ntpl, "algorithmic_bytes_per_launch": cb,
                            "algorithmic_GBs": cb / (cheb_ms * 1e-3) / 1e9,
                            "actual_bytes_per_launch": cb_actual, "actual_GBs": cb_actual / (cheb_ms * 1e-3) / 1e9,
                            "actual_frac_of_peak": cb_actual / (cheb_ms * 1e-3) / 1e9 / peak,
                            "note": "row templates replace the 12 B/nnz CSR read of the static mass matrix by a 16-bit "
                                    "code per row; results are bit-identical to the CSR kernel"},
        "step_roofline": {"algorithmic_GB_per_fct_step": step_gb, "achieved_GBs": step_gb * value,
                          "frac": step_gb * value / peak,
                          "note": "FCT-step bytes only, against the whole pass time (which also contains the gradient "
                                  "and cost work)"},
        "pass_roofline": {"algorithmic_GB_per_pass": pass_gb, "achieved_GBs": pass_gb / (ms / args.steps * 1e-3),
                          "frac": pass_gb / (ms / args.steps * 1e-3) / peak,
                          "note": "all work of the pass (state + adjoint sweeps, gradient, cost) in the accounting "
                                  "unit of SURVEY.md App. E; fusions move fewer actual bytes than that"},
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n,
                "call": "fct_advdrift_state_host (state sweep, pinned host trajectories)"},
        "templates": {"mass_rows": ntpl, "geometry": ctx.geom_template_count()},
        "gpu_launches": int(launches),
        "gpu_launches_note": "host-enqueued kernels of libfctpdeco in the timed region; the Jacobi sweeps run as a CUDA-graph "
                             "WHILE body and are counted once per solve (executed sweeps: jacobi_sweeps_per_step)",
        "clocks": clocks,
    }
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_record(n_cells)[1]
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=4096, help="cells per side (4096 = BASELINE config 5)")
    ap.add_argument("--nt", type=int, default=4, help="time levels per bench step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
