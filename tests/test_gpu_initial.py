"""GPU parity tests for the initial-condition analysis (initial_system_residual!/jacobian!, src/system.jl:466-521, 731-817, and
initial_condition_analysis!, src/analyses.jl:1695-1945) vs the CPU oracle.

The residual/Jacobian test follows test/jacobians.jl:207-236: random state vector, random u0 .. Omegadot0, RANDOM rate_vars1/2
(so every combination of displacement / velocity-rate slots and replaced rows is exercised), structural damping, body-frame
velocities and accelerations.  Tolerances: 1e-12 relative for residuals and Jacobian entries, 1e-9 for solved states."""
import numpy as np
import pytest

import cases
from gxbeam_b200 import api, workloads
from test_gpu_newmark import _ctx, _groups

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import pyoracle as O
    return O


@pytest.mark.parametrize("seed,nelem", [(0, 2), (1, 6), (2, 9)])
@pytest.mark.parametrize("damping", [False, True])
@pytest.mark.parametrize("two_d", [False, True])
def test_initial_residual_and_jacobian_parity(seed, nelem, damping, two_d):
    O = _oracle()
    case = cases.random_anisotropic(seed, nelem=nelem, pmass=(seed % 2 == 0))
    pk = cases.oracle_packed(case, two_dimensional=two_d)
    rng = np.random.default_rng(seed + 90)
    mo = rng.random(12)
    dyn = O.make_dyn(vb=mo[0:3], wb=mo[3:6], ab=mo[6:9], alb=mo[9:12], structural_damping=damping)
    ctx = _ctx(case, mo[None])
    n, npnt = ctx.n, len(case["points"])
    brow, bcol = ctx.pattern()
    vals = [rng.random((npnt, 3)) for _ in range(6)]
    rv1, rv2 = rng.integers(0, 2, n).astype(np.uint8), rng.integers(0, 2, n).astype(np.uint8)
    ic = O.InitialConditions(npnt, n, *vals, rv1=rv1, rv2=rv2)
    kw = dict(zip(("u0", "theta0", "V0", "Omega0", "Vdot0", "Omegadot0"), vals))
    opts = api.make_opts(structural_damping=damping, two_dimensional=two_d)
    for scale in (1.0, 10.0):
        x = scale * rng.random((1, n))
        r, Jb = ctx.initial_residual_jacobian(x, opts, rate_vars1=rv1, rate_vars2=rv2, **kw)
        r_ref = O.initial_residual(pk, dyn, ic, x[0])
        J_ref = O.initial_jacobian(pk, dyn, ic, x[0])
        assert np.abs(r[0] - r_ref).max() <= 1e-12 * np.abs(r_ref).max()
        J = api.blocks_to_dense(n, brow, bcol, Jb[0])
        assert np.abs(J - J_ref).max() <= 1e-12 * np.abs(J_ref).max()
    ctx.close()


def test_rate_vars_match_oracle_massless_elements():
    """test/issues/zeros.jl:170-222: massless elements, point masses on some points -> mixed rate_vars1, full rate_vars2."""
    O = _oracle()
    case = cases.massless_initial_jl()
    pk = cases.oracle_packed(case)
    ctx = _ctx(case, np.zeros((1, 12)))
    res = ctx.initial_condition(opts=api.make_opts(structural_damping=True), want_rate_vars=True)
    ctx.close()
    n = O.nstates(pk, static=False)
    ref = O.initial_condition(pk, O.make_dyn(structural_damping=True), O.InitialConditions(pk.np, n))
    idx = O.indices(pk.start, pk.stop, False)
    for ip in range(pk.np):
        c0 = idx["icol_point"][ip] - 1
        assert np.array_equal(res["rate_vars1"][0, ip], ref["rv1"][c0 + 6:c0 + 12])
        assert np.array_equal(res["rate_vars2"][0, ip], ref["rv2"][c0 + 6:c0 + 12])
    assert res["converged"][0] and ref["converged"]
    assert int(res["iters"][0]) == ref["iters"]
    assert np.abs(res["x"][0] - ref["x"]).max() <= 1e-9 * np.abs(ref["x"]).max()
    assert np.abs(res["rates"][0] - ref["rates"]).max() <= 1e-9 * max(1.0, np.abs(ref["rates"]).max())


def test_initial_jl_known_answer_and_parity():
    """test/initial.jl:72-104: after the initial-condition solve the general dynamic residual (evaluated here by the oracle from
    the DEVICE's states and rates) has norm <= 1e-9."""
    O = _oracle()
    from test_oracle_initial import dynamic_residual
    case = cases.initial_jl()
    pk = cases.oracle_packed(case)
    mo = np.zeros((1, 12)); mo[0, 0:3] = case["linear_velocity"]; mo[0, 3:6] = case["angular_velocity"]
    ctx = _ctx(case, mo)
    res = ctx.initial_condition(opts=api.make_opts(structural_damping=False))
    ctx.close()
    assert res["converged"][0]
    idx = O.indices(pk.start, pk.stop, False)
    dyn = O.make_dyn(vb=mo[0, 0:3], wb=mo[0, 3:6], structural_damping=False)
    r = dynamic_residual(pk, dyn, idx, res["x"][0], res["rates"][0])
    assert np.linalg.norm(r) <= 1e-9
    ref = O.initial_condition(pk, dyn, O.InitialConditions(pk.np, idx["nstates"]))
    assert int(res["iters"][0]) == ref["iters"]
    for name, ids in _groups(O, case).items():
        err = np.abs(res["x"][0][ids] - ref["x"][ids]).max()
        mag = np.abs(ref["x"][ids]).max()
        # an undeformed spinning beam: displacements are exactly zero and the internal loads are round-off zeros (~1e-17) -- the
        # unknowns of this solve are the accelerations; groups that are numerically zero are compared absolutely
        assert err <= (1e-9 * mag if mag > 1e-9 else 1e-12), (name, err)
    assert np.abs(res["rates"][0] - ref["rates"]).max() <= 1e-9 * np.abs(ref["rates"]).max()


def test_time_domain_analysis_end_to_end_from_initial_conditions():
    """time_domain_analysis (analyses.jl:2510-2790) = initial_condition_analysis + Newmark march, batched: blades start undeformed
    with per-instance initial tip velocities, the Newmark march continues from the state left in HBM."""
    O = _oracle()
    nb, nt = 6, 21
    rng = np.random.default_rng(77)
    w = workloads.rotating_blade_batch(nb)
    tvec = np.linspace(0, 0.005, nt)
    # (larger initial velocities make |Newton step| exceed the line search's maxstep = 1e6 through the tiny rotary inertia of this
    # blade -- see test_maxstep_clamped_newton_matches_oracle)
    V0 = np.zeros((nb, 25, 3)); V0[:, :, 2] = 2e-4 * rng.uniform(0.5, 2.0, nb)[:, None] * np.linspace(0, 1, 25)[None, :] ** 2
    opts = api.make_opts(structural_damping=True)
    ctx = _ctx(w, w["motion"], nb, w["points_b"])
    ini = ctx.initial_condition(opts=opts, V0=V0)
    assert ini["converged"].all()
    res = ctx.newmark_run(tvec, opts=opts)                       # x0 = rates0 = None: resident initial state
    ctx.close()
    assert res["converged"].all() and (res["steps_done"] == nt).all()
    for b in range(nb):
        pk = O.Packed(w["start"], w["stop"], w["elems"][b], w["points"], w["pd"], w["pl"], w["bc"][b], force_scaling=w["force_scaling"][b])
        dyn = O.make_dyn(vb=w["motion"][b, 0:3], wb=w["motion"][b, 3:6], structural_damping=True)
        n = O.nstates(pk, static=False)
        ri = O.initial_condition(pk, dyn, O.InitialConditions(25, n, V0=V0[b]))
        assert ri["converged"] and int(ini["iters"][b]) == ri["iters"]
        ref = O.newmark_run(pk, dyn, tvec, ri["x"], ri["rates"])
        assert ref["converged"]
        assert int(res["iters_total"][b]) == int(ref["iters"].sum())
        for name, ids in _groups(O, w).items():
            err = np.abs(res["x"][b][:, ids] - ref["x"][:, ids]).max()
            assert err <= 1e-9 * np.abs(ref["x"][:, ids]).max(), (name, err)


def test_maxstep_clamped_newton_matches_oracle():
    """With structural damping and O(1) initial velocities the first Newton step is ~2e9 (angular accelerations of a section whose
    rotary inertia is 5e-9), so BackTracking's alpha_0 = min(1, maxstep/|p|_inf) (maxstep = 1e6, analyses.jl:94) clamps every step: the
    iteration crawls and does not converge within the cap -- identically on the device and in the oracle."""
    O = _oracle()
    nb, cap = 3, 6
    w = workloads.rotating_blade_batch(nb)
    V0 = np.zeros((nb, 25, 3)); V0[:, :, 2] = np.array([0.5, 1.0, 2.0])[:, None] * np.linspace(0, 1, 25)[None, :] ** 2
    opts = api.make_opts(structural_damping=True, iterations=cap)
    ctx = _ctx(w, w["motion"], nb, w["points_b"])
    ini = ctx.initial_condition(opts=opts, V0=V0)
    ctx.close()
    assert not ini["converged"].any() and (ini["iters"] == cap).all()
    for b in range(nb):
        pk = O.Packed(w["start"], w["stop"], w["elems"][b], w["points"], w["pd"], w["pl"], w["bc"][b], force_scaling=w["force_scaling"][b])
        dyn = O.make_dyn(vb=w["motion"][b, 0:3], wb=w["motion"][b, 3:6], structural_damping=True)
        ri = O.initial_condition(pk, dyn, O.InitialConditions(25, O.nstates(pk, static=False), V0=V0[b]), iterations=cap)
        assert not ri["converged"] and ri["iters"] == cap
        assert abs(ini["resid_inf"][b] - ri["resid_inf"]) <= 1e-9 * ri["resid_inf"]
        assert np.abs(ini["rates"][b] - ri["rates"]).max() <= 1e-8 * np.abs(ri["rates"]).max()
