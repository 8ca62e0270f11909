"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C-ABI, against the CPU oracle.

Bars (BASELINE.json north_star): indices and sparsity pattern bit-exact; states / reactions within 1e-9 relative;
`converged` flags and Newton iteration counts identical on well-conditioned cases.
"""
import numpy as np
import pytest

import cases
from gxbeam_b200 import api
from gxbeam_b200 import workloads
from gxbeam_b200.parity import componentwise_error, state_error
from gxbeam_b200.state import assembly_state

pytestmark = pytest.mark.gpu

RTOL = 1e-9  # north_star tolerance on states and reactions


def _oracle():
    from oracle import pyoracle as O
    return O


def _ctx_for(case, nbatch=None):
    ctx = api.Context(0)
    ctx.topology(case["start"], case["stop"], case["pd"], case["pl"])
    nb = case["elems"].shape[0] if nbatch is None else nbatch
    ctx.set_batch(nb)
    ctx.set_elements(case["elems"][:nb])
    ctx.set_bc(case["bc"][:nb])
    if case.get("dload") is not None:
        ctx.set_dloads(case["dload"][:nb])
    if case.get("pmass") is not None:
        ctx.set_pmass(case["pmass"][:nb])
    g = case.get("gravity")
    if g is not None:
        g = np.broadcast_to(np.asarray(g, float).reshape(-1, 3), (nb, 3)).copy()
    fs = np.broadcast_to(np.asarray(case["force_scaling"], float).reshape(-1), (nb,)).copy()
    ctx.set_body(g, fs)
    return ctx


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_indices_bit_exact():
    O = _oracle()
    for n in (1, 2, 16, 100):
        start, stop = np.arange(1, n + 1), np.arange(2, n + 2)
        ctx = api.Context(0)
        idx = ctx.topology(start, stop)
        ref = O.indices(start, stop, True)
        assert idx["nstates"] == ref["nstates"]
        for k in ("irow_point", "irow_elem", "icol_point", "icol_elem"):
            assert np.array_equal(idx[k], ref[k])
        ctx.close()


def test_unsupported_topology_fails_loudly():
    ctx = api.Context(0)
    ctx.topology([1, 2, 2], [2, 3, 4])          # indices are still produced (bit-exact, general connectivity)
    assert list(ctx.indices["icol_point"]) == [1, 13, 25, 37]
    with pytest.raises(api.GXBError, match="chain"):
        ctx.set_batch(4)
    ctx.close()


@pytest.mark.parametrize("seed,nelem", [(0, 2), (1, 3), (2, 7)])
@pytest.mark.parametrize("two_d", [False, True])
def test_residual_and_jacobian_parity_random_states(seed, nelem, two_d):
    O = _oracle()
    case = cases.random_anisotropic(seed, nelem=nelem)
    pk = cases.oracle_packed(case, two_dimensional=two_d)
    ctx = _ctx_for(case)
    n = ctx.n
    rng = np.random.default_rng(seed)
    brow, bcol = ctx.pattern()
    for scale in (1.0, 1e2):                       # 1e2 exercises rotation_parameter_scaling != 1
        x = scale * rng.random((1, n))
        r, Jb = ctx.static_residual_jacobian(x, api.make_opts(two_dimensional=two_d))
        r_ref = O.static_residual(pk, x[0])
        J_ref = O.static_jacobian(pk, x[0])
        assert _relerr(r[0], r_ref) < 1e-12
        J = api.blocks_to_dense(n, brow, bcol, Jb[0])
        assert np.abs(J - J_ref).max() <= 1e-12 * np.abs(J_ref).max()
        # sparsity pattern: everything the reference can touch lies inside the canonical pattern ...
        mask = api.blocks_to_dense(n, brow, bcol, np.ones((len(brow), 9))) != 0
        assert np.all(J_ref[~mask] == 0.0)
    ctx.close()


def test_pattern_is_canonical_and_bit_exact():
    # 24 blocks / element + 3 / point, minus the shared point θ-blocks counted twice
    O = _oracle()
    case = cases.random_anisotropic(0, nelem=5)
    ctx = _ctx_for(case)
    brow, bcol = ctx.pattern()
    n = ctx.n
    # oracle-side structural pattern: union over random states of the nonzeros of the oracle Jacobian, grown to 3x3 blocks
    pk = cases.oracle_packed(case)
    rng = np.random.default_rng(1)
    nz = np.zeros((n, n), bool)
    for _ in range(3):
        nz |= O.static_jacobian(pk, 10 * rng.random(n)) != 0
    blocks = {(3 * (i // 3), 3 * (j // 3)) for i, j in zip(*np.nonzero(nz))}
    ours = set(zip(brow.tolist(), bcol.tolist()))
    assert blocks <= ours
    # blocks of ours that are numerically empty here must be structurally reachable ones only (zero 3x3 writes)
    assert len(ours) == 24 * 5 + 3 * 6 - 4 * 5
    assert sorted(ours) == list(zip(brow.tolist(), bcol.tolist()))
    ctx.close()


@pytest.mark.parametrize("seed", [0, 3])
def test_newton_direction_parity(seed):
    """Batched pivoted block-banded LU vs a dense LAPACK solve of the oracle's Jacobian."""
    O = _oracle()
    case = cases.random_anisotropic(seed, nelem=9)
    pk = cases.oracle_packed(case)
    ctx = _ctx_for(case)
    rng = np.random.default_rng(seed)
    x = rng.random((1, ctx.n))
    p = ctx.static_newton_direction(x)
    J = O.static_jacobian(pk, x[0])
    r = O.static_residual(pk, x[0])
    p_ref = -np.linalg.solve(J, r)
    assert _relerr(p[0], p_ref) < 1e-9
    ctx.close()


def _solve_case(case, **kw):
    ctx = _ctx_for(case)
    res = ctx.static_solve(api.make_opts(**kw))
    st = ctx.stats()
    idx = ctx.indices
    ctx.close()
    return res, st, idx


def _check_states(case, x_gpu, x_ref, idx, b=0):
    a = assembly_state(x_gpu, case["start"], case["stop"], idx["icol_point"], idx["icol_elem"], case["pd"], case["pl"], case["bc"][b],
                       float(np.ravel(case["force_scaling"])[b] if np.ndim(case["force_scaling"]) else case["force_scaling"]))
    r = assembly_state(x_ref, case["start"], case["stop"], idx["icol_point"], idx["icol_elem"], case["pd"], case["pl"], case["bc"][b],
                       float(np.ravel(case["force_scaling"])[b] if np.ndim(case["force_scaling"]) else case["force_scaling"]))
    # relative to the magnitude of each physical group (a reaction that is exactly zero in exact arithmetic is compared
    # against the size of the loads in the problem, not against its own round-off)
    # COMPONENT-wise: every entry against its own magnitude (floored at 1e-5 of its group's scale, gxbeam_b200/parity.py)
    groups = {"kinematic": ("point_u", "elem_u", "point_theta", "elem_theta"), "load": ("point_F", "elem_Fi", "point_M", "elem_Mi")}
    for g, names in groups.items():
        av = np.concatenate([np.ravel(getattr(a, f)) for f in names]); rv = np.concatenate([np.ravel(getattr(r, f)) for f in names])
        err = componentwise_error(av, rv)
        assert err <= RTOL, (g, err)


@pytest.mark.parametrize("lam", [0.0, 0.4, 1.2, 2.0])
def test_tipmoment_solve_parity(lam):
    O = _oracle()
    case = cases.tipmoment(lam)
    ref = O.static_solve(cases.oracle_packed(case))
    res, st, idx = _solve_case(case)
    assert bool(res["converged"][0]) == ref["converged"]
    assert int(res["iters"][0]) == ref["iters"]
    _check_states(case, res["x"][0], ref["x"], idx)


def test_curved_beam_parity_and_recorded_reference_output():
    O = _oracle()
    case = cases.curved()
    ref = O.static_solve(cases.oracle_packed(case))
    res, st, idx = _solve_case(case)
    assert res["converged"][0] and ref["converged"]
    _check_states(case, res["x"][0], ref["x"], idx)
    ic = idx["icol_point"][-1] - 1
    assert np.allclose(res["x"][0, ic:ic + 3], case["ref_tip_u"], rtol=1e-9, atol=0)   # curved.jl:56-58


@pytest.mark.parametrize("builder", [cases.cantilever_partial_load, cases.overdetermined])
def test_linear_analysis_parity(builder):
    O = _oracle()
    case = builder()
    ref = O.static_solve(cases.oracle_packed(case), linear=True)
    res, st, idx = _solve_case(case, linear=True)
    assert res["converged"][0] and res["iters"][0] == 1
    _check_states(case, res["x"][0], ref["x"], idx)
    a = assembly_state(res["x"][0], case["start"], case["stop"], idx["icol_point"], idx["icol_elem"], case["pd"], case["pl"], case["bc"][0],
                       case["force_scaling"])
    for ip, xi in enumerate(case["points"][:, 0]):
        assert abs(a.point_u[ip, 2] - case["deflection"](xi)) <= 1e-8


@pytest.mark.parametrize("two_d", [False, True])
def test_gravity_pointmass_follower_solve_parity(two_d):
    O = _oracle()
    case = cases.random_anisotropic(5, nelem=6)
    case["gravity"] = np.array([0.0, 0.0, -9.81])
    if two_d:   # keep the planar problem well posed: loads in the x-y plane only
        case = cases.random_anisotropic(5, nelem=6, followers=False, pmass=False, dloads=False, gravity=False)
    ref = O.static_solve(cases.oracle_packed(case, two_dimensional=two_d))
    res, st, idx = _solve_case(case, two_dimensional=two_d)
    assert bool(res["converged"][0]) == ref["converged"]
    if ref["converged"]:
        assert int(res["iters"][0]) == ref["iters"]
        _check_states(case, res["x"][0], ref["x"], idx)


def test_tipforce_batch_with_backtracking_parity():
    """A mixed batch (different loads -> different iteration counts and line-search paths) against per-instance oracle solves."""
    O = _oracle()
    lams = [0.5, 2.0, 4.0, 8.0, 12.0, 16.0]
    cs = [cases.tipforce(l) for l in lams]
    case = dict(cs[0])
    case["elems"] = np.concatenate([c["elems"] for c in cs])
    case["bc"] = np.concatenate([c["bc"] for c in cs])
    case["force_scaling"] = np.array([c["force_scaling"] for c in cs])
    res, st, idx = _solve_case(case)
    for b, c in enumerate(cs):
        ref = O.static_solve(cases.oracle_packed(c))
        assert bool(res["converged"][b]) == ref["converged"]
        assert int(res["iters"][b]) == ref["iters"], (lams[b], res["iters"][b], ref["iters"])
        _check_states(case, res["x"][b], ref["x"], idx, b)
    assert st["newton_iters"] == int(res["iters"].sum())


def test_c2_subset_parity_against_oracle():
    """BASELINE config 2 at reduced batch: every instance checked against the oracle."""
    O = _oracle()
    w = workloads.cantilever_tipmoment_batch(48, 100)
    ctx = _ctx_for(w)
    res = ctx.static_solve()
    idx = ctx.indices
    ctx.close()
    pk = O.Packed(w["start"], w["stop"], w["elems"], w["points"], w["pd"], w["pl"], w["bc"])
    ref = O.static_solve_batch(pk, 48, force_scaling=w["force_scaling"])
    assert np.array_equal(res["converged"], ref["converged"].astype(bool)) and res["converged"].all()
    assert np.array_equal(res["iters"], ref["iters"])
    for b in range(48):
        _check_states(w, res["x"][b], ref["x"][b], idx, b)


def test_c2_full_size_properties():
    """BASELINE config 2 at full size (4096 x 100 elements): size-independent properties + a sampled oracle check."""
    O = _oracle()
    w = workloads.cantilever_tipmoment_batch(4096, 100)
    ctx = _ctx_for(w)
    res = ctx.static_solve()
    st = ctx.stats()
    idx = ctx.indices
    # (1) everything converged, residual norms under ftol
    assert res["converged"].all() and (res["resid_inf"] <= 1e-9).all()
    assert st["newton_iters"] == int(res["iters"].sum())
    # (2) idempotence: restarting from the solution takes zero iterations and returns the same state bit for bit
    res2 = ctx.static_solve(x0=res["x"])
    assert (res2["iters"] == 0).all() and np.array_equal(res2["x"], res["x"])
    # (3) physics: pure bending -> tip rotation = λ π about z (Wiener-Milenkovic parameter 4 tan(φ/4)), |error| from discretisation only
    ic = idx["icol_point"][-1] - 1
    thz = res["x"][:, ic + 5]
    assert np.allclose(thz, 4 * np.tan(w["lam"] * np.pi / 4), rtol=1e-4)
    # (4) the oracle's residual at the GPU solution is below ftol on a sample, and the sample agrees with oracle solves
    pick = np.random.default_rng(0).choice(4096, 24, replace=False)
    for b in pick:
        pk = cases.oracle_packed(w, int(b))
        assert np.abs(O.static_residual(pk, res["x"][b])).max() <= 1e-9
        ref = O.static_solve(pk)
        assert ref["iters"] == res["iters"][b]
        _check_states(w, res["x"][b], ref["x"], idx, int(b))
    ctx.close()


def test_multi_device_entry_point_single_gpu():
    w = workloads.cantilever_tipmoment_batch(32, 20)
    out = api.static_analysis_batch(w["start"], w["stop"], w["elems"], w["pd"], w["pl"], w["bc"], force_scaling=w["force_scaling"], devices=(0,))
    assert out.converged.all() and out.x.shape == (32, 12 * 20 + 6)


def test_no_gravity_hint_uploads_less_and_changes_nothing():
    """gxb_hint_no_gravity: only L, x, compliance, Cab of every element record cross PCIe; the solve is bit-identical, and a later
    non-zero gravity is refused loudly."""
    w = workloads.cantilever_tipmoment_batch(64, 30)
    rng = np.random.default_rng(5)
    elems = w["elems"].copy()
    elems[:, :, 40:76] = rng.random(elems[:, :, 40:76].shape)          # junk in the mass fields must not matter without gravity
    out = []
    for hint in (False, True):
        ctx = api.Context(0)
        ctx.topology(w["start"], w["stop"], w["pd"], w["pl"])
        ctx.set_batch(64)
        if hint:
            ctx.hint_no_gravity()
        ctx.set_elements(elems); ctx.set_bc(w["bc"]); ctx.set_body(None, w["force_scaling"])
        out.append(ctx.static_solve())
        if hint:
            with pytest.raises(api.GXBError, match="gravity"):
                ctx.set_body(np.ones((64, 3)), w["force_scaling"])
        ctx.close()
    assert out[0]["converged"].all() and np.array_equal(out[0]["iters"], out[1]["iters"]) and np.array_equal(out[0]["x"], out[1]["x"])


def _ctx_case(case, nb=1):
    ctx = api.Context(0)
    ctx.topology(case["start"], case["stop"], case["pd"], case["pl"])
    ctx.set_batch(nb)
    return ctx


def test_static_joined_wing_linear_sweep_as_one_batch():
    """test/examples/static-joined-wing.jl:103-121: 141 load levels of the joined wing (a chain clamped at BOTH ends), linear = true --
    independent solves, so they go out as ONE batch."""
    O = _oracle()
    Fz = np.linspace(0, 70e3, 141)
    cs = [cases.static_joined_wing(f) for f in Fz]
    ctx = _ctx_case(cs[0], 141)
    ctx.set_elements(np.concatenate([c["elems"] for c in cs])); ctx.set_bc(np.concatenate([c["bc"] for c in cs]))
    ctx.set_body(None, np.array([c["force_scaling"] for c in cs]))
    res = ctx.static_solve(api.make_opts(linear=True))
    ctx.close()
    assert res["converged"].all()
    for k in (1, 70, 140):
        ref = O.static_solve(cases.oracle_packed(cs[k]), linear=True)
        assert state_error(res["x"][k], ref["x"], 0) <= 1e-9
    # linear: the response scales with the load
    assert np.abs(res["x"][140] - 2 * res["x"][70]).max() <= 1e-9 * np.abs(res["x"][140]).max()


@pytest.mark.parametrize("follower", [False, True])
def test_static_joined_wing_nonlinear_continuation(follower):
    """static-joined-wing.jl:123-150: the nonlinear sweeps with reset_state = false -- every load level starts from the previous
    solution (from zero the full load does not converge).  Dead and follower load; iteration counts must match the oracle's."""
    O = _oracle()
    Fz = np.linspace(0, 70e3, 141)
    case0 = cases.static_joined_wing(0.0, follower)
    ctx = _ctx_case(case0, 1)
    ctx.set_elements(case0["elems"]); ctx.set_body(None, np.array([case0["force_scaling"]]))
    x = None; xo = None
    its, its_ref = [], []
    for f in Fz:
        c = cases.static_joined_wing(f, follower)
        ctx.set_bc(c["bc"])
        r = ctx.static_solve(x0=x)
        ref = O.static_solve(cases.oracle_packed(c), x0=xo)
        assert r["converged"][0] and ref["converged"]
        x, xo = r["x"], ref["x"]
        its.append(int(r["iters"][0])); its_ref.append(ref["iters"])
    ctx.close()
    assert its == its_ref
    assert state_error(x[0], xo, 0) <= 1e-9


def _dyadic_free_free_beam(nelem=8):
    """A free-free beam (no displacement prescribed anywhere) whose Jacobian at x = 0 has only dyadic entries (lengths 1/4, compliances
    2^-k, Cab = I): Gaussian elimination is exact, so the six rigid-body modes show up as EXACTLY zero pivots in any pivot order --
    LAPACK-style `info != 0` in the oracle, the zero-pivot flag on the device."""
    from gxbeam_b200.assembly import PrescribedConditions, pack_prescribed_conditions
    w = workloads.cantilever_tipmoment_batch(1, nelem)
    el = w["elems"].copy()
    el[:, :, 0] = 0.25
    for k, v in enumerate([2.0 ** -10, 2.0 ** -8, 2.0 ** -8, 2.0 ** -6, 2.0 ** -5, 2.0 ** -5]):
        el[:, :, 4 + 7 * k] = v
    pd, pl, bc1 = pack_prescribed_conditions(nelem + 1, {nelem + 1: PrescribedConditions(Fz=1.0), 3: PrescribedConditions(Fy=-0.5)})
    return dict(w, elems=el, pd=pd, pl=pl, bc=bc1[None], force_scaling=np.ones(1))


@pytest.mark.parametrize("parts", [1, 2])
def test_singular_jacobian_fallback_matches_nlsolve_restated(parts):
    """NLsolve's singular-Jacobian fallback p = -(J'J + λI) \\ J'F, λ = 1e6 sqrt(n eps) ||J'J||_1 (oracle/gxb_solve.hpp:135-160) on
    the device (k_singular_fallback): one Newton iteration of the exactly singular free-free beam must land on the oracle's iterate
    (x1 = x0 + α p with the same α from the line search, whose slope is the true g'p here), and the iteration count must be 1 on both."""
    O = _oracle()
    case = _dyadic_free_free_beam()
    pk = O.Packed(case["start"], case["stop"], case["elems"], case["points"], case["pd"], case["pl"], case["bc"])
    n = 12 * 8 + 6
    J = O.static_jacobian(pk, np.zeros(n))
    assert np.linalg.matrix_rank(J) == n - 6
    ref = O.static_solve(pk, iterations=1)
    # the oracle did take the fallback: its iterate is the regularised step
    A = J.T @ J
    lam = 1e6 * np.sqrt(n * np.finfo(float).eps) * np.abs(A).sum(0).max()
    p = -np.linalg.solve(A + lam * np.eye(n), J.T @ O.static_residual(pk, np.zeros(n)))
    assert np.abs(ref["x"] - p).max() <= 1e-12 * np.abs(p).max()
    ctx = _ctx_for(case)
    res = ctx.static_solve(api.make_opts(iterations=1, lu_parts=parts))
    assert int(res["iters"][0]) == ref["iters"] == 1 and not res["converged"][0] and not ref["converged"]
    assert state_error(res["x"][0], ref["x"], 0) <= 1e-9
    # a regular instance in the same batch is not disturbed, and the batch keeps going after the fallback
    res3 = ctx.static_solve(api.make_opts(iterations=3, lu_parts=parts))
    ref3 = O.static_solve(pk, iterations=3)
    ctx.close()
    assert not res3["converged"][0] and not ref3["converged"] and int(res3["iters"][0]) == ref3["iters"] == 3


def test_exact_dphi0_option_agrees_on_iteration_counts():
    """opts.exact_dphi0: dφ0 = (J'F)'p as NLsolve evaluates it instead of -F'F.  Same Newton iteration counts and states on a BASELINE
    config-2 sample (the two differ by the residual of the linear solve, ~1e-13 relative)."""
    O = _oracle()
    w = workloads.cantilever_tipmoment_batch(64, 100, tip_force=True)
    ctx = _ctx_for(w)
    a = ctx.static_solve(api.make_opts())
    b = ctx.static_solve(api.make_opts(exact_dphi0=True))
    ctx.close()
    pk = O.Packed(w["start"], w["stop"], w["elems"], w["points"], w["pd"], w["pl"], w["bc"])
    ref = O.static_solve_batch(pk, 64, force_scaling=w["force_scaling"])
    assert a["converged"].all() and b["converged"].all()
    assert np.array_equal(a["iters"], b["iters"]) and np.array_equal(b["iters"], ref["iters"])
    assert state_error(b["x"], ref["x"], 0) <= 1e-9 and state_error(a["x"], b["x"], 0) <= 1e-9


def test_bc_compact_upload_is_equivalent_to_the_full_records():
    """gxb_set_bc_compact: one shared np x 18 record + K per-instance values must leave exactly the bytes gxb_set_bc leaves (NaN pattern
    included): the solves are bit-identical."""
    w = workloads.cantilever_tipmoment_batch(40, 24, tip_force=True)
    ctx = _ctx_for(w)
    a = ctx.static_solve()
    ctx.set_bc_compact(w["bc"][0], [25, 25], [11, 8], np.ascontiguousarray(w["bc"][:, 24, [11, 8]]))
    b = ctx.static_solve()
    with pytest.raises(api.GXBError, match="1-based"):
        ctx.set_bc_compact(w["bc"][0], [0], [11], np.zeros((40, 1)))
    ctx.close()
    assert a["converged"].all() and np.array_equal(a["iters"], b["iters"]) and np.array_equal(a["x"], b["x"])


def test_multi_gpu_entry_is_bit_identical_to_one_gpu():
    """gxb_multi_* (one context + worker thread per GPU inside the library): the sharded solve returns exactly the single-context
    result.  With one visible GPU the two shards run as two contexts of the same device (same code path); under the 2-GPU lease they
    run on devices 0 and 1."""
    import ctypes
    ndev = ctypes.c_int(0)
    ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(ndev))
    devices = (0, 1) if ndev.value >= 2 else (0, 0)
    w = workloads.cantilever_tipmoment_batch(50, 40, tip_force=True)
    # lu_parts fixed: the automatic choice depends on the shard size, and different partitionings agree to round-off only
    one = api.static_analysis_batch(w["start"], w["stop"], w["elems"], w["pd"], w["pl"], w["bc"], force_scaling=w["force_scaling"], devices=(0,), lu_parts=4)
    two = api.static_analysis_batch(w["start"], w["stop"], w["elems"], w["pd"], w["pl"], w["bc"], force_scaling=w["force_scaling"], devices=devices, lu_parts=4)
    assert two.converged.all() and np.array_equal(one.iters, two.iters) and np.array_equal(one.x, two.x)
    assert len(two.stats["shard_ms"]) == 2 and all(t > 0 for t in two.stats["shard_ms"])
    # steady_state_analysis through the same entry
    wb = workloads.rotating_blade_batch(6, rpm_jitter=0.3)
    mc = api.MultiContext(devices)
    mc.topology(wb["start"], wb["stop"], wb["pd"], wb["pl"], kind=1)
    mc.set_batch(6)
    r2 = mc.solve(wb["elems"], wb["bc"], points=wb["points_b"], force_scaling=wb["force_scaling"], motion=wb["motion"])
    mc.close()
    ctx = api.Context(0)
    ctx.topology(wb["start"], wb["stop"], wb["pd"], wb["pl"], kind=1)
    ctx.set_batch(6)
    ctx.set_elements(wb["elems"]); ctx.set_points(wb["points_b"]); ctx.set_bc(wb["bc"]); ctx.set_body(None, wb["force_scaling"]); ctx.set_motion(wb["motion"])
    r1 = ctx.steady_solve()
    ctx.close()
    assert r2["converged"].all() and np.array_equal(r1["iters"], r2["iters"]) and np.array_equal(r1["x"], r2["x"])


def test_multi_solve_compact_pipeline_lanes_match_one_context():
    """gxb_multi_solve_compact: three shards that share GPU 0 (pipeline lanes with ordered uploads), no-gravity hint, prescribed conditions
    as one shared record + 2 varying entries per instance -- bit-identical to one context fed the full arrays (lu_parts fixed: the
    automatic kernel choice depends on the shard size)."""
    w = workloads.cantilever_tipmoment_batch(90, 24, tip_force=True)
    opts = api.make_opts(lu_parts=1)
    ctx = _ctx_for(w)
    ref = ctx.static_solve(opts)
    ctx.close()
    mc = api.MultiContext((0, 0, 0))
    mc.topology(w["start"], w["stop"], w["pd"], w["pl"])
    mc.set_batch(90)
    mc.hint_no_gravity()
    vals = np.ascontiguousarray(w["bc"][:, 24, [11, 8]])
    for _ in range(2):          # the second call reuses the shards' device buffers
        r = mc.solve_compact(w["elems"], w["bc"][0], [25, 25], [11, 8], vals, opts, force_scaling=w["force_scaling"])
        assert r["converged"].all() and np.array_equal(r["iters"], ref["iters"]) and np.array_equal(r["x"], ref["x"])
        assert len(r["shard_ms"]) == 3 and all(t > 0 for t in r["shard_ms"])
    with pytest.raises(api.GXBError, match="1-based"):
        mc.solve_compact(w["elems"], w["bc"][0], [0, 25], [11, 8], vals, opts, force_scaling=w["force_scaling"])
    mc.close()


def test_device_assembly_state_matches_host_restatement():
    """gxb_assembly_state: AssemblyState(x, system, assembly; prescribed_conditions) on the device (src/output.jl:146-203, 317-403) --
    PointState / ElementState records of every instance against the numpy restatement (gxbeam_b200/state.py), follower loads, prescribed
    displacements and reaction unknowns included.  Same formulas on both sides: agreement to round-off (1e-13)."""
    from gxbeam_b200.state import assembly_state_records
    case = cases.random_anisotropic(3, nelem=9)
    ctx = _ctx_for(case)
    res = ctx.static_solve()
    pts, els = ctx.assembly_state()
    ctx.close()
    # (the extraction is a function of the state vector alone: this stiff random case need not have converged)
    rp, re_ = assembly_state_records(res["x"][0], 0, case["pd"], case["pl"], case["bc"][0], float(np.ravel(case["force_scaling"])[0]))
    assert np.abs(pts[0] - rp).max() <= 1e-13 * np.abs(rp).max() and np.abs(els[0] - re_).max() <= 1e-13 * np.abs(re_).max()
    # the same observables as the dataclass the solve tests use
    a = assembly_state(res["x"][0], case["start"], case["stop"], ctx.indices["icol_point"], ctx.indices["icol_elem"], case["pd"], case["pl"],
                       case["bc"][0], float(np.ravel(case["force_scaling"])[0]))
    assert np.allclose(pts[0][:, 0:3], a.point_u, rtol=0, atol=1e-13 * np.abs(a.point_u).max())
    assert np.allclose(pts[0][:, 24:27], a.point_F, rtol=1e-13, atol=1e-13 * np.abs(a.point_F).max())
    assert np.allclose(els[0][:, 27:30], a.elem_Mi, rtol=1e-13, atol=1e-13 * np.abs(a.elem_Mi).max())


@pytest.mark.parametrize("parts", [0, 1, 2, 4, 8])
def test_nan_member_is_isolated_under_every_factorisation_kernel(parts):
    """A NaN compliance entry in one member of a static batch: that member is reported unconverged (NLsolve would throw on the non-finite
    residual), every other member takes exactly the oracle's iterations and lands on its state -- with one warp per instance, the
    two-sided kernel and the partitioned kernel alike (the NaN must not leak through shared memory or the interface systems)."""
    O = _oracle()
    w = workloads.cantilever_tipmoment_batch(12, 40, tip_force=True)
    elems = w["elems"].copy()
    elems[5, 17, 4] = np.nan
    ctx = _ctx_for({**w, "elems": elems})
    res = ctx.static_solve(api.make_opts(lu_parts=parts))
    ctx.close()
    ok = np.ones(12, bool); ok[5] = False
    assert np.array_equal(res["converged"], ok)
    pk = O.Packed(w["start"], w["stop"], w["elems"], w["points"], w["pd"], w["pl"], w["bc"])
    ref = O.static_solve_batch(pk, 12, force_scaling=w["force_scaling"])
    assert np.array_equal(res["iters"][ok], ref["iters"][ok])
    assert state_error(res["x"][ok], ref["x"][ok], 0) <= 1e-9
