#!/usr/bin/env python
"""Secondary measurements (not the driver's bench.py): BASELINE configs 1 and 4 on the device next to the CPU oracle port.

  steady : config-1 topology (24-element swept-tip rotating blade, n = 444), batch of rotor-speed variants, steady_state_analysis
  gust   : config 4 — 512-member gust ensemble, time_domain_analysis of 2000 Newmark steps (dt = 0.25 ms) on the same topology,
           initial state from initial_condition_analysis like benchmark/benchmarks.jl:143-216 (or --gust-init steady), structural
           damping on (the reference's default)
Prints one JSON line per workload; bench.py imports the three functions and attaches their dicts to its JSON line as `secondary`
(time-boxed CPU samples), so that the driver's own run carries these numbers with clocks.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

PERSISTENT = 0
SAVE_EVERY = 0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gxbeam_b200 import api, workloads  # noqa: E402
from gxbeam_b200.parity import state_error  # noqa: E402


def make_ctx(w, nb):
    ctx = api.Context(0)
    ctx.topology(w["start"], w["stop"], w["pd"], w["pl"], kind=1)
    ctx.set_batch(nb)
    ctx.set_elements(w["elems"]); ctx.set_points(w["points_b"]); ctx.set_bc(w["bc"]); ctx.set_body(None, w["force_scaling"])
    ctx.set_motion(w["motion"])
    return ctx


def steady(nb, cpu_sample):
    from oracle import pyoracle as O
    w = workloads.rotating_blade_batch(nb, rpm_jitter=0.5)
    ctx = make_ctx(w, nb)
    for _ in range(3):
        ctx.steady_solve_resident(api.make_opts(persistent=PERSISTENT))
    ms, its = 0.0, 0
    for _ in range(10):
        ctx.steady_solve_resident(api.make_opts(persistent=PERSISTENT)); st = ctx.stats(); ms += st["ms_total"]; its += st["newton_iters"]
    res = ctx.fetch(); ctx.close()
    ns = min(nb, cpu_sample)
    pk = O.Packed(w["start"], w["stop"], w["elems"][:ns], w["points_b"][:ns], w["pd"], w["pl"], w["bc"][:ns])
    dyns = [O.make_dyn(vb=w["motion"][b, 0:3], wb=w["motion"][b, 3:6]) for b in range(ns)]
    t0 = time.perf_counter(); ref = O.steady_solve_batch(pk, ns, dyns, force_scaling=w["force_scaling"][:ns]); dt = time.perf_counter() - t0
    err = state_error(res["x"][:ns], ref["x"], 1)
    # sanity anchor: the restatement's one-core time per instance next to the published single-core Julia figure
    n1 = min(ns, 16)
    pk1 = O.Packed(w["start"], w["stop"], w["elems"][:n1], w["points_b"][:n1], w["pd"], w["pl"], w["bc"][:n1])
    t0 = time.perf_counter(); O.steady_solve_batch(pk1, n1, dyns[:n1], force_scaling=w["force_scaling"][:n1], nthreads=1); dt1 = time.perf_counter() - t0
    return {"workload": f"steady_state_analysis, {nb} rotating blades (config-1 topology, n=444)", "gpu_newton_iters_per_s": its / (ms * 1e-3),
            "gpu_ms_per_batch": ms / 10, "cpu_newton_iters_per_s": float(ref["iters"].sum() / dt), "cpu_cores": O.num_threads(),
            "cpu_sample": ns, "iters_equal": bool(np.array_equal(ref["iters"], res["iters"][:ns])), "max_componentwise_rel_diff": err,
            "cpu_ms_per_instance_1core": 1e3 * dt1 / n1,
            "published_julia_ms_per_instance_1core": 4.716, "published_source": "benchmark/benchmarks.jl:221-223 (i7-5500U, whole analysis)"}


def gust(nb, nt, cpu_sample, damping=True, init="ic"):
    """init="ic": exactly benchmark/benchmarks.jl:143-216 -- time_domain_analysis without initial_state, i.e. initial_condition_analysis
    of the undeformed spinning blade followed by the Newmark march; init="steady": start from the steady rotating equilibrium."""
    from oracle import pyoracle as O
    rng = np.random.default_rng(1234 + 4)
    w = workloads.rotating_blade_batch(nb)
    tvec = np.linspace(0.0, 0.5 * (nt - 1) / 2000, nt)
    # amplitude A_i = 100 U(0.5, 1.5), duration and onset of the 1 - cos gust (benchmarks.jl:198) jittered by 10 %
    amp = 100 * rng.uniform(0.5, 1.5, nb)
    period = 0.4 * rng.uniform(0.9, 1.1, nb)[:, None]
    onset = 0.02 * rng.uniform(0.0, 1.0, nb)[:, None]
    tt = tvec[None, :] - onset
    tf1 = np.where((tt >= 0) & (tt < 0.5 * period), 0.1 * (1 - np.sin(2 * np.pi * (tt / period + 0.25))), 0.0)
    series = (amp[:, None] * tf1)[:, :, None]
    ctx = make_ctx(w, nb)
    opts = api.make_opts(structural_damping=damping, persistent=PERSISTENT)
    rates0 = np.zeros((nb, 25, 12))
    ms_init, it_init, launches_init = 0.0, 0, 0

    def start():
        nonlocal ms_init, it_init, launches_init
        ctx.set_bc(w["bc"])                                  # the march overwrites the prescribed values with the series
        if init == "ic":
            ini = ctx.initial_condition(opts=opts)
            s0 = ctx.stats(); ms_init, it_init, launches_init = s0["ms_total"], s0["newton_iters"], s0["launches"]
            assert ini["converged"].all()
            return ini["x"], ini["rates"]
        st = ctx.steady_solve()
        return st["x"], rates0

    ctx.set_bc_series([21], [7], series)
    x0, r0 = start()
    ctx.newmark_run(tvec[:21], x0, r0, opts, want_history=False)      # warm-up
    t0 = time.perf_counter()
    x0, r0 = start()
    # the reference saves every time level (save = eachindex(tvec)): so does this run -- 3.6 GB of states stay on the device during the
    # march and are copied back once (included in gpu_wall_s, not in the event-timed gpu_s_per_run)
    se = SAVE_EVERY or 1
    res = ctx.newmark_run(tvec, None, None, opts, save_every=se) if init == "ic" else ctx.newmark_run(tvec, x0, r0, opts, save_every=se)
    wall = time.perf_counter() - t0
    s = ctx.stats(); ctx.close()
    ns = min(nb, cpu_sample)
    t0 = time.perf_counter(); cpu_iters = 0; err = 0.0; iters_equal = True
    for b in range(ns):
        pk = O.Packed(w["start"], w["stop"], w["elems"][b], w["points"], w["pd"], w["pl"], w["bc"][b], force_scaling=w["force_scaling"][b])
        dyn = O.make_dyn(vb=w["motion"][b, 0:3], wb=w["motion"][b, 3:6], structural_damping=damping)
        bct = np.tile(w["bc"][b][None], (nt, 1, 1)); bct[:, 20, 7] = series[b, :, 0]
        if init == "ic":
            ri = O.initial_condition(pk, dyn, O.InitialConditions(25, O.nstates(pk, static=False)))
            cx0, cr0 = ri["x"], ri["rates"]; cpu_iters += ri["iters"]
        else:
            cx0, cr0 = x0[b], rates0[b]
        ref = O.newmark_run(pk, dyn, tvec, cx0, cr0, bct)
        cpu_iters += int(ref["iters"].sum())
        iters_equal = iters_equal and int(res["iters_total"][b]) == int(ref["iters"].sum())
        err = max(err, state_error(res["x"][b, -1], ref["x"][-1], 1))
    dt = time.perf_counter() - t0
    gpu_ms = s["ms_total"] + ms_init
    return {"workload": f"time_domain_analysis, {nb}-member gust ensemble, {nt - 1} Newmark steps (config 4), structural_damping={damping}, "
                        f"initial state: {'initial_condition_analysis (as benchmarks.jl:143-216)' if init == 'ic' else 'steady state'}",
            "gpu_newton_iters_per_s": (s["newton_iters"] + it_init) / (gpu_ms * 1e-3), "gpu_s_per_run": gpu_ms * 1e-3,
            "gpu_s_initial_condition": ms_init * 1e-3, "gpu_wall_s": wall, "saved_time_levels": int(res["x"].shape[1]),
            "gpu_launches": s["launches"] + launches_init, "newton_iters": s["newton_iters"] + it_init,
            "all_converged": bool(res["converged"].all()), "iters_equal": bool(iters_equal),
            "cpu_newton_iters_per_s_1core": cpu_iters / dt, "cpu_s_per_instance_1core": dt / ns, "cpu_sample": ns,
            "published_julia_s_per_instance_1core": 9.019, "published_source": "benchmark/benchmarks.jl:235-237 (i7-5500U, whole analysis)",
            "max_componentwise_rel_diff_final_state": err}


def blades(nb, nelem, nev, m, cpu_sample):
    """BASELINE config 3: steady_state_analysis + linearised eigenvalue analysis of rotating wind-turbine blades."""
    from oracle import pyoracle as O
    w = workloads.wind_turbine_blade_batch(nb, nelem)
    ctx = make_ctx(w, nb)
    # the reference's defaults (ftol = 1e-9, iterations = 1000); every design x speed of the generator converges (workloads.py)
    opts = api.make_opts(persistent=PERSISTENT)
    ctx.steady_solve_resident(opts)
    ms, its = 0.0, 0
    for _ in range(3):
        ctx.steady_solve_resident(opts); st = ctx.stats(); ms += st["ms_total"]; its += st["newton_iters"]
    res = ctx.fetch()
    t0 = time.perf_counter()
    lam, vec, rr = ctx.eigenvalue_analysis(res["x"], nev=nev, m=m)
    t_eig = time.perf_counter() - t0
    es = ctx.stats(); ctx.close()
    ns = min(nb, cpu_sample)
    pk = O.Packed(w["start"], w["stop"], w["elems"][:ns], w["points_b"][:ns], w["pd"], w["pl"], w["bc"][:ns])
    dyns = [O.make_dyn(wb=w["motion"][b, 3:6]) for b in range(ns)]
    t0 = time.perf_counter(); ref = O.steady_solve_batch(pk, ns, dyns, force_scaling=w["force_scaling"][:ns]); dt = time.perf_counter() - t0
    err = state_error(res["x"][:ns], ref["x"], 1)
    return ({"workload": f"config 3: {nb} rotating blades x {nelem} elements (n={18 * nelem + 12}), steady_state_analysis + eigen (nev={nev}, m={m})",
                      "steady_gpu_newton_iters_per_s": its / (ms * 1e-3), "steady_gpu_ms_per_batch": ms / 3, "steady_converged": int(res["converged"].sum()), "steady_converged_cpu_sample": int(ref["converged"].sum()),
                      "steady_cpu_newton_iters_per_s": float(ref["iters"].sum() / dt), "cpu_cores": O.num_threads(), "cpu_sample": ns,
                      "steady_iters_equal": bool(np.array_equal(ref["iters"], res["iters"][:ns])), "steady_max_componentwise_rel_diff": err,
                      "eigen_device_s": es["ms_total"] * 1e-3, "eigen_wall_s_incl_host_ritz": t_eig, "eigen_max_ritz_residual": float(rr.max()),
                      "lowest_freq_hz_first_instance": float(np.sort(np.abs(lam[0].imag))[0] / (2 * np.pi))})


def sections(nb, cpu_sample):
    """BASELINE config 5: compliance_matrix + mass_matrix (src/section.jl) of `nb` parameterised airfoil layups sharing one quad mesh."""
    from oracle import section_oracle as SO
    w = workloads.airfoil_layup_batch(nb)
    sec = api.Section(w["nn"], w["conn"])
    sec.properties(w["nodes"][:64], w["materials"][:64], w["theta"][:64])           # warm-up
    t0 = time.perf_counter()
    r = sec.properties(w["nodes"], w["materials"], w["theta"])
    wall = time.perf_counter() - t0
    hb = sec.half_bandwidth
    sec.close()
    ns = min(nb, cpu_sample)
    t0 = time.perf_counter(); err = 0.0
    for b in range(ns):
        Sref, sc, tc = SO.compliance_matrix(w["nodes"][b], w["conn"] - 1, w["materials"][b], w["theta"][b])
        d = np.sqrt(np.abs(np.diag(Sref)))
        err = max(err, float((np.abs(r["S"][b] - Sref) / np.outer(d, d)).max()))
    dt = time.perf_counter() - t0
    return {"workload": f"config 5: compliance_matrix + mass_matrix of {nb} airfoil layups (quad mesh: {w['nn']} nodes, {w['ne']} elements, "
                        f"KKT system of {3 * w['nn'] + 12}, half bandwidth {hb} after reverse Cuthill-McKee)",
            "gpu_layups_per_s": nb / (r["device_ms"] * 1e-3), "gpu_device_s": r["device_ms"] * 1e-3, "gpu_wall_s_incl_transfers": wall,
            "all_ok": bool(r["ok"].all()), "cpu_s_per_layup_numpy_oracle_1core": dt / ns, "cpu_sample": ns,
            "max_scaled_diff_S_vs_oracle": err}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steady-batch", type=int, default=4096)
    ap.add_argument("--gust-batch", type=int, default=512)
    ap.add_argument("--gust-steps", type=int, default=2000)
    ap.add_argument("--gust-init", default="ic", choices=["ic", "steady"])
    ap.add_argument("--gust-save-every", type=int, default=1, help="save every k-th time level of the gust run (reference default: 1)")
    ap.add_argument("--persistent", type=int, default=0, help="0 auto, 1 force, 2 forbid the persistent per-instance kernel")
    ap.add_argument("--cpu-sample", type=int, default=4)
    ap.add_argument("--blades-batch", type=int, default=1024)
    ap.add_argument("--blades-nelem", type=int, default=200)
    ap.add_argument("--sections-batch", type=int, default=2048)
    ap.add_argument("--only", default="all", choices=["all", "steady", "gust", "blades", "sections"])
    a = ap.parse_args()
    globals()["PERSISTENT"] = a.persistent
    globals()["SAVE_EVERY"] = a.gust_save_every
    ap2 = a
    if a.only in ("all", "steady"):
        print(json.dumps(steady(a.steady_batch, 1024)), flush=True)
    if a.only in ("all", "gust"):
        print(json.dumps(gust(a.gust_batch, a.gust_steps + 1, a.cpu_sample, init=a.gust_init)), flush=True)
    if a.only in ("all", "blades"):
        print(json.dumps(blades(a.blades_batch, a.blades_nelem, 20, 70, 256)), flush=True)
    if a.only in ("all", "sections"):
        print(json.dumps(sections(a.sections_batch, 2)), flush=True)
