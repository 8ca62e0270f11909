"""GPU parity tests for the steady-state (DynamicSystem) path: CUDA through the C-ABI vs the CPU oracle."""
import numpy as np
import pytest

import cases
from gxbeam_b200 import api
from gxbeam_b200 import workloads
from gxbeam_b200.parity import componentwise_error, state_error

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _oracle():
    from oracle import pyoracle as O
    return O


def _ctx(case, motion, nb=1, pts=None):
    ctx = api.Context(0)
    ctx.topology(case["start"], case["stop"], case["pd"], case["pl"], kind=1)
    ctx.set_batch(nb)
    ctx.set_elements(case["elems"][:nb])
    ctx.set_points(np.broadcast_to(case["points"], (nb,) + case["points"].shape).copy() if pts is None else pts)
    ctx.set_bc(case["bc"][:nb])
    if case.get("dload") is not None:
        ctx.set_dloads(case["dload"][:nb])
    if case.get("pmass") is not None:
        ctx.set_pmass(case["pmass"][:nb])
    g = case.get("gravity")
    if g is not None:
        g = np.broadcast_to(np.asarray(g, float).reshape(-1, 3), (nb, 3)).copy()
    ctx.set_body(g, np.broadcast_to(np.asarray(case["force_scaling"], float).reshape(-1), (nb,)).copy())
    ctx.set_motion(motion)
    return ctx


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_dynamic_indices_bit_exact():
    O = _oracle()
    for n in (1, 3, 24):
        start, stop = np.arange(1, n + 1), np.arange(2, n + 2)
        ctx = api.Context(0)
        idx = ctx.topology(start, stop, kind=1)
        ref = O.indices(start, stop, False)
        assert idx["nstates"] == ref["nstates"] == 18 * n + 12
        for k in ("irow_point", "irow_elem", "icol_point", "icol_elem"):
            assert np.array_equal(idx[k], ref[k])
        ctx.close()


def test_body_acceleration_state_is_refused_by_the_eigen_path_only():
    """Body-frame accelerations as states are supported by the steady / initial-condition solves (tests/test_gpu_body.py); the
    eigen path still refuses them loudly."""
    case = cases.jacobians_jl(0)
    ctx = api.Context(0)
    ctx.topology(case["start"], case["stop"], case["pd"], case["pl"], kind=1)
    ctx.set_batch(1)
    ctx.set_elements(case["elems"]); ctx.set_points(case["points"][None].copy()); ctx.set_bc(case["bc"]); ctx.set_body(None, np.array([case["force_scaling"]]))
    with pytest.raises(api.GXBError, match="body-frame acceleration"):
        ctx.eigen_arnoldi(np.zeros((1, ctx.n)), 4)
    ctx.close()


@pytest.mark.parametrize("seed,nelem", [(0, 2), (1, 5), (2, 9)])
@pytest.mark.parametrize("two_d", [False, True])
@pytest.mark.parametrize("damping", [False, True])
def test_steady_residual_and_jacobian_parity(seed, nelem, two_d, damping):
    O = _oracle()
    case = cases.random_anisotropic(seed, nelem=nelem, pmass=(seed % 2 == 0))
    pk = cases.oracle_packed(case, two_dimensional=two_d)
    rng = np.random.default_rng(seed + 40)
    mo = 1e2 * rng.random(12)
    dyn = O.make_dyn(vb=mo[0:3], wb=mo[3:6], ab=mo[6:9], alb=mo[9:12], structural_damping=damping)
    ctx = _ctx(case, mo[None])
    n = ctx.n
    brow, bcol = ctx.pattern()
    mask = api.blocks_to_dense(n, brow, bcol, np.ones((len(brow), 9))) != 0
    for scale in (1.0, 1e2):
        x = scale * rng.random((1, n))
        r, Jb = ctx.residual_jacobian(x, api.make_opts(two_dimensional=two_d, structural_damping=damping))
        r_ref = O.steady_residual(pk, dyn, x[0])
        J_ref = O.steady_jacobian(pk, dyn, x[0])
        assert _relerr(r[0], r_ref) < 1e-12
        J = api.blocks_to_dense(n, brow, bcol, Jb[0])
        assert np.abs(J - J_ref).max() <= 1e-12 * np.abs(J_ref).max()
        assert np.all(J_ref[~mask] == 0.0)
    ctx.close()


def test_steady_newton_direction_parity():
    O = _oracle()
    case = cases.random_anisotropic(4, nelem=8, pmass=True)
    pk = cases.oracle_packed(case)
    rng = np.random.default_rng(3)
    mo = 10 * rng.random(12)
    dyn = O.make_dyn(vb=mo[0:3], wb=mo[3:6], ab=mo[6:9], alb=mo[9:12])
    ctx = _ctx(case, mo[None])
    x = rng.random((1, ctx.n))
    p = ctx.newton_direction(x)
    J = O.steady_jacobian(pk, dyn, x[0])
    r = O.steady_residual(pk, dyn, x[0])
    assert _relerr(p[0], -np.linalg.solve(J, r)) < 1e-9
    ctx.close()


def test_rotating_blade_config1_parity():
    """BASELINE config 1 (benchmark/benchmarks.jl steady_benchmark): single Assembly, n = 444."""
    O = _oracle()
    w = workloads.rotating_blade_batch(1)
    ctx = _ctx(w, w["motion"], 1, w["points_b"])
    res = ctx.steady_solve()
    assert ctx.n == 444
    ctx.close()
    pk = O.Packed(w["start"], w["stop"], w["elems"][0], w["points"], w["pd"], w["pl"], w["bc"][0], force_scaling=w["force_scaling"][0])
    dyn = O.make_dyn(vb=w["motion"][0, 0:3], wb=w["motion"][0, 3:6])
    ref = O.steady_solve(pk, dyn)
    assert ref["converged"] and res["converged"][0]
    assert int(res["iters"][0]) == ref["iters"]
    # compare per physical group (displacement-like, velocity-like, load-like)
    idx = O.indices(w["start"], w["stop"], False)
    xg, xr = res["x"][0], ref["x"]
    groups = {"disp": [], "vel": [], "load": []}
    for ic in idx["icol_point"]:
        groups["disp"] += list(range(ic - 1, ic + 5)); groups["vel"] += list(range(ic + 5, ic + 11))
    for ic in idx["icol_elem"]:
        groups["load"] += list(range(ic - 1, ic + 5))
    # root point holds reactions, not displacements
    root = list(range(0, 6))
    groups["disp"] = [k for k in groups["disp"] if k not in root]; groups["load"] += root
    for name, ids in groups.items():
        err = componentwise_error(xg[ids], xr[ids])          # every component against its own magnitude (gxbeam_b200/parity.py)
        assert err <= RTOL, (name, err)
    # centrifugal stiffening sanity: the blade stretches outward (tip u_x > 0)
    tip = idx["icol_point"][-1] - 1
    assert xg[tip] > 0


def test_rotor_speed_sweep_batch_parity():
    """A design-sweep style batch: 32 blades at different rotor speeds, every instance against the oracle."""
    O = _oracle()
    nb = 32
    w = workloads.rotating_blade_batch(nb, rpm_jitter=0.5)
    ctx = _ctx(w, w["motion"], nb, w["points_b"])
    res = ctx.steady_solve()
    st = ctx.stats()
    ctx.close()
    pk = O.Packed(w["start"], w["stop"], w["elems"], w["points_b"], w["pd"], w["pl"], w["bc"])
    dyns = [O.make_dyn(vb=w["motion"][b, 0:3], wb=w["motion"][b, 3:6]) for b in range(nb)]
    ref = O.steady_solve_batch(pk, nb, dyns, force_scaling=w["force_scaling"])
    assert res["converged"].all() and ref["converged"].all()
    assert np.array_equal(res["iters"], ref["iters"])
    assert st["newton_iters"] == int(res["iters"].sum())
    for b in range(nb):
        assert np.abs(res["x"][b] - ref["x"][b]).max() <= 1e-9 * np.abs(ref["x"][b]).max()


def test_steady_solve_with_structural_damping_parity():
    """steady_state_analysis(...; structural_damping=true): with zero strain rates at equilibrium the solution equals the
    undamped one, but the Newton path goes through the damped residual/Jacobian on both sides."""
    O = _oracle()
    nb = 6
    w = workloads.rotating_blade_batch(nb, rpm_jitter=0.3)
    ctx = _ctx(w, w["motion"], nb, w["points_b"])
    res = ctx.steady_solve(api.make_opts(structural_damping=True))
    ctx.close()
    pk = O.Packed(w["start"], w["stop"], w["elems"], w["points_b"], w["pd"], w["pl"], w["bc"])
    dyns = [O.make_dyn(vb=w["motion"][b, 0:3], wb=w["motion"][b, 3:6], structural_damping=True) for b in range(nb)]
    ref = O.steady_solve_batch(pk, nb, dyns, force_scaling=w["force_scaling"])
    assert res["converged"].all() and ref["converged"].all()
    assert np.array_equal(res["iters"], ref["iters"])
    for b in range(nb):
        assert np.abs(res["x"][b] - ref["x"][b]).max() <= 1e-9 * np.abs(ref["x"][b]).max()


def test_pointmass_model_steady_jacobian_parity():
    """test/issues/pointmass.jl model (point masses at every point, massless elements): K = steady Jacobian at the converged state
    and at a random state, against the oracle that reproduces the reference's recorded eigen-frequencies."""
    O = _oracle()
    case = cases.pointmass_jl()
    pk = cases.oracle_packed(case)
    rng = np.random.default_rng(11)
    mo = np.concatenate([rng.random(6), np.zeros(6)])
    dyn = O.make_dyn(vb=mo[0:3], wb=mo[3:6])
    ctx = _ctx(case, mo[None])
    n = ctx.n
    brow, bcol = ctx.pattern()
    for x in (np.zeros((1, n)), rng.random((1, n))):
        r, Jb = ctx.residual_jacobian(x)
        J_ref = O.steady_jacobian(pk, dyn, x[0])
        r_ref = O.steady_residual(pk, dyn, x[0])
        assert np.abs(r[0] - r_ref).max() <= 1e-12 * max(np.abs(r_ref).max(), 1e-30)
        J = api.blocks_to_dense(n, brow, bcol, Jb[0])
        assert np.abs(J - J_ref).max() <= 1e-12 * np.abs(J_ref).max()
    ctx.close()


def test_long_chain_chunked_tiles_parity():
    """N = 150 elements: the shared-memory tile holds fewer stations than the chain (chunked flush); BASELINE config 3 style blade."""
    O = _oracle()
    nb = 2
    w = workloads.wind_turbine_blade_batch(nb, nelem=150, nspeeds=2)
    w["motion"][:, 5] = [0.7, 1.9]
    ctx = _ctx(w, w["motion"], nb, w["points_b"])
    n = ctx.n
    assert n == 18 * 150 + 12
    rng = np.random.default_rng(2)
    x = 0.05 * rng.random((nb, n))
    brow, bcol = ctx.pattern()
    r, Jb = ctx.residual_jacobian(x)
    for b in range(nb):
        pk = O.Packed(w["start"], w["stop"], w["elems"][b], w["points"], w["pd"], w["pl"], w["bc"][b], force_scaling=w["force_scaling"][b])
        dyn = O.make_dyn(wb=w["motion"][b, 3:6])
        r_ref = O.steady_residual(pk, dyn, x[b])
        J_ref = O.steady_jacobian(pk, dyn, x[b])
        assert np.abs(r[b] - r_ref).max() <= 1e-12 * np.abs(r_ref).max()
        J = api.blocks_to_dense(n, brow, bcol, Jb[b])
        assert np.abs(J - J_ref).max() <= 1e-12 * np.abs(J_ref).max()
    res = ctx.steady_solve()
    ctx.close()
    pk = O.Packed(w["start"], w["stop"], w["elems"], w["points_b"], w["pd"], w["pl"], w["bc"])
    ref = O.steady_solve_batch(pk, nb, [O.make_dyn(wb=w["motion"][b, 3:6]) for b in range(nb)], force_scaling=w["force_scaling"])
    assert res["converged"].all() and ref["converged"].all() and np.array_equal(res["iters"], ref["iters"])
    for b in range(nb):
        assert np.abs(res["x"][b] - ref["x"][b]).max() <= 1e-9 * np.abs(ref["x"][b]).max()


@pytest.mark.parametrize("builder", ["zero_mass_jl", "zero_length_element_jl"])
def test_zeros_jl_steady_cases_parity(builder):
    """test/issues/zeros.jl:3-168 on the device: zero mass matrix and a zero-length element (L = 0: compliance and mass blocks
    vanish, the duplicated joint point is tied by the compatibility rows alone) converge like the oracle."""
    O = _oracle()
    case = getattr(cases, builder)()
    pk = cases.oracle_packed(case)
    mo = np.zeros((1, 12)); mo[0, 3:6] = case["angular_velocity"]
    ctx = _ctx(case, mo)
    res = ctx.steady_solve()
    ctx.close()
    ref = O.steady_solve(pk, O.make_dyn(wb=case["angular_velocity"]))
    assert ref["converged"] and res["converged"][0] and int(res["iters"][0]) == ref["iters"]
    scale = np.abs(ref["x"]).max()
    assert np.abs(res["x"][0] - ref["x"]).max() <= 1e-9 * scale


def test_edge_sizes_and_failure_isolation():
    """Smallest chain (1 element), a batch of one, an iteration cap of one, and a NaN in one member's inputs: the bad member is
    reported unconverged, everyone else is untouched."""
    O = _oracle()
    w = workloads.rotating_blade_batch(5, rpm_jitter=0.3)
    elems = w["elems"].copy()
    elems[3, 7, 10] = np.nan                                       # a NaN compliance entry in member 3
    ctx = _ctx({**w, "elems": elems}, w["motion"], 5, w["points_b"])
    res = ctx.steady_solve()
    capped = ctx.steady_solve(api.make_opts(iterations=1))
    ctx.close()
    ok = np.array([True, True, True, False, True])
    assert np.array_equal(res["converged"], ok)
    assert not capped["converged"][ok].any() and (capped["iters"][ok] == 1).all()
    pk = O.Packed(w["start"], w["stop"], w["elems"], w["points_b"], w["pd"], w["pl"], w["bc"])
    dyns = [O.make_dyn(vb=w["motion"][b, 0:3], wb=w["motion"][b, 3:6]) for b in range(5)]
    ref = O.steady_solve_batch(pk, 5, dyns, force_scaling=w["force_scaling"])
    for b in np.nonzero(ok)[0]:
        assert res["iters"][b] == ref["iters"][b]
        assert np.abs(res["x"][b] - ref["x"][b]).max() <= 1e-9 * np.abs(ref["x"][b]).max()
    # one element, one instance
    one = cases.tipmoment(0.2, nelem=1)
    pk1 = cases.oracle_packed(one)
    mo = np.zeros((1, 12)); mo[0, 3:6] = [0.1, 0.2, 0.3]
    ctx = _ctx(one, mo)
    r1 = ctx.steady_solve()
    ctx.close()
    ref1 = O.steady_solve(pk1, O.make_dyn(wb=mo[0, 3:6]))
    assert bool(r1["converged"][0]) == ref1["converged"] and int(r1["iters"][0]) == ref1["iters"]
    assert np.abs(r1["x"][0] - ref1["x"]).max() <= 1e-9 * np.abs(ref1["x"]).max()


def test_device_assembly_state_dynamic_system():
    """gxb_assembly_state for a DynamicSystem: velocities from the states, rates NaN without dx (the reference's Fill(NaN), output.jl:140)
    and taken from the given point rates otherwise (prescribed displacements: zero rate, point.jl:265-281)."""
    from gxbeam_b200.state import assembly_state_records
    w = workloads.rotating_blade_batch(3, rpm_jitter=0.3)
    ctx = _ctx(w, w["motion"], 3, w["points_b"])
    res = ctx.steady_solve()
    pts, els = ctx.assembly_state()
    rates = np.random.default_rng(0).random((3, 25, 12))
    pts2, els2 = ctx.assembly_state(rates)
    ctx.close()
    for b in range(3):
        rp, re_ = assembly_state_records(res["x"][b], 1, w["pd"], w["pl"], w["bc"][b], w["force_scaling"][b])
        assert np.array_equal(np.isnan(pts[b]), np.isnan(rp)) and np.array_equal(np.isnan(els[b]), np.isnan(re_))
        assert np.nanmax(np.abs(pts[b] - rp)) <= 1e-13 * np.nanmax(np.abs(rp)) and np.nanmax(np.abs(els[b] - re_)) <= 1e-13 * np.nanmax(np.abs(re_))
        rp, re_ = assembly_state_records(res["x"][b], 1, w["pd"], w["pl"], w["bc"][b], w["force_scaling"][b], rates[b])
        assert np.abs(pts2[b] - rp).max() <= 1e-13 * np.abs(rp).max() and np.abs(els2[b] - re_).max() <= 1e-13 * np.abs(re_).max()
