"""BASELINE configs 3 and 4 at FULL size on the GPU: size-independent properties over the whole batch plus sampled oracle checks
(config 2's full-size test lives in test_gpu_static.py)."""
import numpy as np
import pytest

from gxbeam_b200 import api, workloads
from gxbeam_b200.parity import componentwise_error
from test_gpu_newmark import _ctx, _groups, _gust

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import pyoracle as O
    return O


def test_c4_full_size_time_domain_analysis():
    """512-member gust ensemble x 2000 Newmark steps exactly as benchmark/benchmarks.jl:143-216 (initial_condition_analysis of the
    undeformed spinning blade, structural damping on)."""
    O = _oracle()
    nb, nt = 512, 2001
    w, tvec, series = _gust(nb, nt, 0.5)
    opts = api.make_opts(structural_damping=True)
    ctx = _ctx(w, w["motion"], nb, w["points_b"])
    ini = ctx.initial_condition(opts=opts)
    ctx.set_bc_series([21], [7], series)
    res = ctx.newmark_run(tvec, opts=opts, save_every=500)
    ctx.close()
    # (1) every member converges at every step
    assert ini["converged"].all() and res["converged"].all() and (res["steps_done"] == nt).all()
    # (2) linearity of the small-amplitude response in the gust amplitude is NOT assumed; instead: members are independent, so two
    #     members with the same amplitude ordering keep the ordering of their peak tip deflection at t = 0.125 s (during the pulse)
    idx = O.indices(w["start"], w["stop"], False)
    tip_y = idx["icol_point"][-1]                      # 0-based column of u_y at the tip
    amp = series[:, 200, 0]
    defl = res["x"][:, 1, tip_y]                       # saved level 1 = step 500 = 0.125 s
    order = np.argsort(amp)
    assert np.all(np.diff(defl[order]) > 0)
    # (3) long after the pulse (t = 0.5 s, damping on) every member has returned to the same steady rotating state
    final = res["x"][:, -1]
    assert np.abs(final - final[0]).max() <= 1e-6 * np.abs(final[0]).max()
    # (4) sampled members against the oracle, whole saved history, iteration counts included
    for b in (0, 255, 511):
        pk = O.Packed(w["start"], w["stop"], w["elems"][b], w["points"], w["pd"], w["pl"], w["bc"][b], force_scaling=w["force_scaling"][b])
        dyn = O.make_dyn(vb=w["motion"][b, 0:3], wb=w["motion"][b, 3:6], structural_damping=True)
        ri = O.initial_condition(pk, dyn, O.InitialConditions(25, idx["nstates"]))
        bct = np.tile(w["bc"][b][None], (nt, 1, 1)); bct[:, 20, 7] = series[b, :, 0]
        ref = O.newmark_run(pk, dyn, tvec, ri["x"], ri["rates"], bct)
        assert ref["converged"] and int(res["iters_total"][b]) == int(ref["iters"].sum())
        for name, ids in _groups(O, w).items():
            # component-wise (gxbeam_b200/parity.py), 1e-9; the floor scale is the class's maximum over the whole history (at t = 0
            # the undeformed blade carries no internal loads at all: that level alone would compare round-off with round-off)
            err = componentwise_error(res["x"][b][:, ids].ravel(), ref["x"][::500][:, ids].ravel())
            assert err <= 1e-9, (b, name, err)


def test_c3_full_size_steady_and_eigen():
    """1024 rotating wind-turbine blades x 200 elements (n = 3612): steady_state_analysis + eigenvalue_analysis (nev = 20)."""
    O = _oracle()
    nb, nelem, nev, m = 1024, 200, 20, 70
    w = workloads.wind_turbine_blade_batch(nb, nelem)
    ctx = _ctx(w, w["motion"], nb, w["points_b"])
    opts = api.make_opts()                      # the reference's defaults: ftol = 1e-9, iterations = 1000
    res = ctx.steady_solve(opts)
    conv = res["converged"]
    # (1) every design x speed converges (workloads.py: rotor speeds up to 0.9 rad/s), residuals under ftol
    assert conv.all() and (res["resid_inf"] <= 1e-9).all()
    # (2) idempotence on the converged members
    res2 = ctx.steady_solve(opts, x0=res["x"])
    assert (res2["iters"][conv] == 0).all() and np.array_equal(res2["x"][conv], res["x"][conv])
    # (3) eigenpairs of the whole batch: Ritz residuals under ArnoldiMethod's tol for the converged members
    lam, vec, rr = ctx.eigenvalue_analysis(res["x"], nev=nev, m=m)
    assert rr[conv].max() <= 1e-9
    # conjugate pairs: the spectrum is closed under conjugation up to the cut at nev
    assert np.all(np.isfinite(lam[conv].view(float)))
    # (4) sampled members against the oracle: iteration counts, states, and (K + λ M) v = 0 with the ORACLE's K and M
    pick = [0, 7, 400, 1023]
    for b in pick:
        pk = O.Packed(w["start"], w["stop"], w["elems"][b], w["points"], w["pd"], w["pl"], w["bc"][b], force_scaling=w["force_scaling"][b])
        dyn = O.make_dyn(vb=w["motion"][b, 0:3], wb=w["motion"][b, 3:6])
        ref = O.steady_solve(pk, dyn)
        assert ref["converged"] and ref["iters"] == int(res["iters"][b])
        for name, ids in _groups(O, w).items():
            err = componentwise_error(res["x"][b][ids], ref["x"][ids])
            assert err <= 1e-9, (b, name, err)
        K = O.steady_jacobian(pk, dyn, res["x"][b])
        M = O.mass_matrix(pk, res["x"][b])
        for k in range(nev):
            v = vec[b][:, k]
            assert np.abs((K + lam[b, k] * M) @ v).max() <= 1e-6 * np.abs(K @ v).max(), (b, k)
    ctx.close()


def test_c5_full_size_section_properties():
    """BASELINE config 5: 2048 parameterised airfoil layups (840 nodes, 720 quads, KKT system of 2532 unknowns each).  Size-independent
    properties over the whole batch (every elimination succeeded; S and M symmetric, S positive definite; mass = sum of rho * area) plus
    sampled instances against the numpy oracle at 1e-9 of the entry scale."""
    from oracle import section_oracle as SO
    nb = 2048
    w = workloads.airfoil_layup_batch(nb)
    sec = api.Section(w["nn"], w["conn"])
    r = sec.properties(w["nodes"], w["materials"], w["theta"])
    sec.close()
    assert r["ok"].all()
    S, M = r["S"], r["M"]
    d = np.sqrt(np.abs(np.einsum("bii->bi", S)))
    assert (np.abs(S - np.transpose(S, (0, 2, 1))) / (d[:, :, None] * d[:, None, :])).max() <= 1e-9
    assert np.linalg.eigvalsh(0.5 * (S + np.transpose(S, (0, 2, 1)))).min() > 0
    assert np.array_equal(M, np.transpose(M, (0, 2, 1)))
    c = w["conn"] - 1
    xy = w["nodes"][:, c]                                                                     # [nb, ne, 4, 2]
    area = 0.5 * sum(xy[:, :, i, 0] * xy[:, :, (i + 1) % 4, 1] - xy[:, :, (i + 1) % 4, 0] * xy[:, :, i, 1] for i in range(4))
    mass = (w["materials"][:, :, 9] * area).sum(1)
    assert np.abs(M[:, 0, 0] - mass).max() <= 1e-12 * mass.max()
    for b in (0, 1023, 2047):
        Sref, sc, tc = SO.compliance_matrix(w["nodes"][b], c, w["materials"][b], w["theta"][b])
        dd = np.sqrt(np.abs(np.diag(Sref)))
        assert (np.abs(S[b] - Sref) / np.outer(dd, dd)).max() <= 1e-9, b
        assert np.abs(r["sc"][b] - sc).max() <= 1e-9 and np.abs(r["tc"][b] - tc).max() <= 1e-9
