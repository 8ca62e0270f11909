"""GPU tests of the persistent per-instance kernel (one CTA runs an instance's whole Newton loop / time loop) against the
multi-kernel rounds and the CPU oracle.  Both paths call the same device functions in the same order per instance; only the
partition of the line search's sums over threads differs, so iteration counts must match and states agree to round-off."""
import numpy as np
import pytest

from gxbeam_b200 import api, workloads
from test_gpu_newmark import _ctx, _groups, _gust

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import pyoracle as O
    return O


@pytest.mark.parametrize("damping", [False, True])
def test_steady_persistent_equals_multikernel_and_oracle(damping):
    O = _oracle()
    nb = 48
    w = workloads.rotating_blade_batch(nb, rpm_jitter=0.5)
    out = {}
    for name, flag in (("persistent", 1), ("rounds", 2)):
        ctx = _ctx(w, w["motion"], nb, w["points_b"])
        out[name] = ctx.steady_solve(api.make_opts(structural_damping=damping, persistent=flag))
        out[name + "_launches"] = ctx.stats()["launches"]
        ctx.close()
    a, b = out["persistent"], out["rounds"]
    assert a["converged"].all() and b["converged"].all()
    assert np.array_equal(a["iters"], b["iters"])
    assert np.abs(a["x"] - b["x"]).max() <= 1e-11 * np.abs(b["x"]).max()
    # the persistent kernel + the kernel that sums the iteration counts for the statistics
    assert out["persistent_launches"] == 2 and out["rounds_launches"] > 5
    for i in (0, nb - 1):
        pk = O.Packed(w["start"], w["stop"], w["elems"][i], w["points"], w["pd"], w["pl"], w["bc"][i], force_scaling=w["force_scaling"][i])
        ref = O.steady_solve(pk, O.make_dyn(vb=w["motion"][i, 0:3], wb=w["motion"][i, 3:6], structural_damping=damping))
        assert ref["iters"] == int(a["iters"][i])
        for name, ids in _groups(O, w).items():
            assert np.abs(a["x"][i][ids] - ref["x"][ids]).max() <= 1e-9 * np.abs(ref["x"][ids]).max(), name


def test_time_domain_persistent_equals_multikernel_and_oracle():
    O = _oracle()
    nb, nt = 8, 61
    w, tvec, series = _gust(nb, nt, 0.015)
    res = {}
    for name, flag in (("persistent", 1), ("rounds", 2)):
        opts = api.make_opts(structural_damping=True, persistent=flag)
        ctx = _ctx(w, w["motion"], nb, w["points_b"])
        ini = ctx.initial_condition(opts=opts)
        assert ini["converged"].all()
        ctx.set_bc_series([21], [7], series)
        res[name] = ctx.newmark_run(tvec, opts=opts, save_every=5)
        res[name + "_ini"] = ini
        res[name + "_launches"] = ctx.stats()["launches"]
        ctx.close()
    a, b = res["persistent"], res["rounds"]
    assert a["converged"].all() and b["converged"].all()
    assert np.array_equal(res["persistent_ini"]["iters"], res["rounds_ini"]["iters"])
    assert np.array_equal(a["iters_total"], b["iters_total"]) and np.array_equal(a["steps_done"], b["steps_done"])
    assert np.abs(a["x"] - b["x"]).max() <= 1e-10 * np.abs(b["x"]).max()
    assert res["persistent_launches"] == 1 and res["rounds_launches"] > 100
    i = 3
    pk = O.Packed(w["start"], w["stop"], w["elems"][i], w["points"], w["pd"], w["pl"], w["bc"][i], force_scaling=w["force_scaling"][i])
    dyn = O.make_dyn(vb=w["motion"][i, 0:3], wb=w["motion"][i, 3:6], structural_damping=True)
    ri = O.initial_condition(pk, dyn, O.InitialConditions(25, O.nstates(pk, static=False)))
    bct = np.tile(w["bc"][i][None], (nt, 1, 1)); bct[:, 20, 7] = series[i, :, 0]
    ref = O.newmark_run(pk, dyn, tvec, ri["x"], ri["rates"], bct)
    assert int(a["iters_total"][i]) == int(ref["iters"].sum())
    for name, ids in _groups(O, w).items():
        err = np.abs(a["x"][i][:, ids] - ref["x"][::5][:, ids]).max()
        assert err <= 1e-9 * np.abs(ref["x"][:, ids]).max(), (name, err)


def test_persistent_instances_fail_independently():
    """An instance that cannot converge (iteration cap) must not affect the others, and must stop marching on its own."""
    nb, nt = 6, 21
    w, tvec, series = _gust(nb, nt, 0.005)
    series = series.copy(); series[2] *= 1e6                       # absurd load on one member: its Newton iteration hits the cap
    opts = api.make_opts(structural_damping=True, persistent=1, iterations=4)
    ctx = _ctx(w, w["motion"], nb, w["points_b"])
    st = ctx.steady_solve(api.make_opts(persistent=1))
    ctx.set_bc_series([21], [7], series)
    res = ctx.newmark_run(tvec, st["x"], np.zeros((nb, 25, 12)), opts)
    ctx.close()
    ok = np.ones(nb, bool); ok[2] = False
    assert np.array_equal(res["converged"], ok)
    assert (res["steps_done"][ok] == nt).all() and res["steps_done"][2] < nt


def test_straggler_handoff_matches_plain_rounds():
    """Default policy: multi-kernel rounds, then the last few unfinished instances are handed to the persistent kernel.  Iteration
    counts, convergence flags and states must equal the pure multi-kernel run (persistent = 2)."""
    nb = 96
    w = workloads.rotating_blade_batch(nb, rpm_jitter=0.9)
    out = {}
    for name, flag in (("auto", 0), ("rounds", 2)):
        ctx = _ctx(w, w["motion"], nb, w["points_b"])
        out[name] = ctx.steady_solve(api.make_opts(persistent=flag, iterations=12))
        out[name + "_stats"] = ctx.stats()
        ctx.close()
    a, b = out["auto"], out["rounds"]
    assert len(set(b["iters"].tolist())) > 1                       # the batch really has instances that need more rounds than others
    assert np.array_equal(a["iters"], b["iters"]) and np.array_equal(a["converged"], b["converged"])
    assert np.abs(a["x"] - b["x"]).max() <= 1e-11 * np.abs(b["x"]).max()
    assert out["auto_stats"]["launches"] < out["rounds_stats"]["launches"]


def test_handoff_in_the_initial_condition_solve():
    """One member of nine crawls under the line search's maxstep clamp (cf. test_gpu_initial.py) while the others converge in one
    iteration: the default policy hands the crawler to the persistent kernel (MODE 2, resume).  Same iterates as the plain rounds."""
    nb, cap = 9, 12
    w = workloads.rotating_blade_batch(nb)
    V0 = np.zeros((nb, 25, 3)); V0[4, :, 2] = np.linspace(0, 1, 25) ** 2
    out = {}
    for name, flag in (("auto", 0), ("rounds", 2)):
        ctx = _ctx(w, w["motion"], nb, w["points_b"])
        out[name] = ctx.initial_condition(opts=api.make_opts(structural_damping=True, iterations=cap, persistent=flag), V0=V0)
        out[name + "_launches"] = ctx.stats()["launches"]
        ctx.close()
    a, b = out["auto"], out["rounds"]
    ok = np.ones(nb, bool); ok[4] = False
    assert np.array_equal(a["converged"], ok) and np.array_equal(b["converged"], ok)
    assert np.array_equal(a["iters"], b["iters"]) and a["iters"][4] == cap
    assert np.abs(a["rates"] - b["rates"]).max() <= 1e-9 * np.abs(b["rates"]).max()
    assert out["auto_launches"] < out["rounds_launches"]


def test_contexts_release_all_device_memory():
    """Create / use / destroy every kind of handle several times (static with every factorisation kernel, dynamic with steady, eigen,
    initial-condition and Newmark runs, a two-shard gxb_multi, a section topology): free device memory returns to its starting value --
    the RAII holders of the C-ABI leave nothing behind on any path."""
    import torch
    from gxbeam_b200 import workloads

    def free():
        torch.cuda.synchronize()
        return torch.cuda.mem_get_info()[0]
    torch.zeros(1, device="cuda")
    w = workloads.cantilever_tipmoment_batch(64, 40, tip_force=True)
    wb = workloads.rotating_blade_batch(16, rpm_jitter=0.2)
    ws = workloads.airfoil_layup_batch(4)

    def cycle():
        ctx = api.Context(0); ctx.topology(w["start"], w["stop"], w["pd"], w["pl"]); ctx.set_batch(64)
        ctx.set_elements(w["elems"]); ctx.set_bc(w["bc"]); ctx.set_body(None, w["force_scaling"])
        for lp in (0, 1, 2, 4):
            assert ctx.static_solve(api.make_opts(lu_parts=lp))["converged"].all()
        ctx.assembly_state(); ctx.close()
        c = api.Context(0); c.topology(wb["start"], wb["stop"], wb["pd"], wb["pl"], kind=1); c.set_batch(16)
        c.set_elements(wb["elems"]); c.set_points(wb["points_b"]); c.set_bc(wb["bc"]); c.set_body(None, wb["force_scaling"]); c.set_motion(wb["motion"])
        st = c.steady_solve(); c.eigenvalue_analysis(st["x"], nev=4); c.initial_condition(); c.newmark_run(np.linspace(0, 0.005, 6)); c.close()
        mc = api.MultiContext((0, 0)); mc.topology(w["start"], w["stop"], w["pd"], w["pl"]); mc.set_batch(64)
        mc.solve(w["elems"], w["bc"], force_scaling=w["force_scaling"]); mc.close()
        s = api.Section(ws["nn"], ws["conn"]); s.properties(ws["nodes"], ws["materials"], ws["theta"]); s.close()
    cycle()
    f0 = free()
    for _ in range(3):
        cycle()
    assert free() == f0


def test_batch_that_does_not_fit_fails_loudly_and_leaves_the_context_usable():
    """gxb_set_batch with a batch far beyond device memory: an error code with CUDA's message, every partial allocation given back, no stale
    error in CUDA's last-error slot -- the same context then takes a normal batch and solves it."""
    import torch
    from gxbeam_b200 import workloads
    torch.zeros(1, device="cuda")

    def free():
        torch.cuda.synchronize()
        return torch.cuda.mem_get_info()[0]
    w = workloads.cantilever_tipmoment_batch(8, 100)
    ctx = api.Context(0)
    ctx.topology(w["start"], w["stop"], w["pd"], w["pl"])
    f0 = free()
    with pytest.raises(api.GXBError, match="out of memory"):
        ctx.set_batch(2_000_000)                       # ~1.4 TB of Jacobian and U storage
    assert f0 - free() < (64 << 20)                    # nothing of the failed batch is left
    assert ctx.lib.gxb_set_elements(ctx.h, api._ptr(w["elems"])) != 0 and b"gxb_set_batch first" in ctx.lib.gxb_last_error()
    ctx.set_batch(8)
    ctx.set_elements(w["elems"]); ctx.set_bc(w["bc"]); ctx.set_body(None, w["force_scaling"])
    assert ctx.static_solve()["converged"].all()
    ctx.close()
