"""GPU parity tests for the eigenvalue analysis pieces: mass matrix M = ∂r/∂ẋ (system_mass_matrix!) and the Krylov part of
solve_eigensystem (K^{-1} M Arnoldi on the device, small dense eig on the host), against the oracle + a dense solver and against
the reference's recorded frequencies (test/issues/pointmass.jl:54-58)."""
import numpy as np
import pytest

import cases
from gxbeam_b200 import api, workloads

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("seed,nelem,two_d", [(0, 2, False), (1, 5, False), (2, 4, True)])
def test_mass_matrix_parity(seed, nelem, two_d):
    O = _oracle()
    case = cases.random_anisotropic(seed, nelem=nelem, pmass=True)
    pk = cases.oracle_packed(case, two_dimensional=two_d)
    ctx = _ctx(case, np.zeros((1, 12)))
    n = ctx.n
    brow, bcol = ctx.pattern()
    x = np.random.default_rng(seed).random((1, n))
    Mb = ctx.mass_matrix(x, api.make_opts(two_dimensional=two_d))
    M = api.blocks_to_dense(n, brow, bcol, Mb[0])
    M_ref = O.mass_matrix(pk, x[0])
    assert np.abs(M - M_ref).max() <= 1e-13 * np.abs(M_ref).max()
    mask = api.blocks_to_dense(n, brow, bcol, np.ones((len(brow), 9))) != 0
    assert np.all(M_ref[~mask] == 0.0)
    ctx.close()


def test_pointmass_eigenfrequencies_match_recorded_reference_output():
    case = cases.pointmass_jl()
    ctx = _ctx(case, np.zeros((1, 12)))
    st = ctx.steady_solve()
    assert st["converged"][0]
    lam, vec, res = ctx.eigenvalue_analysis(st["x"], nev=14, m=60)
    ctx.close()
    freq = np.sort(np.abs(lam[0].imag)) / (2 * np.pi)
    for f in case["ref_freq"]:
        assert np.min(np.abs(freq - f)) <= 1.5e-8 * f          # pointmass.jl:54-58 (default isapprox)
    assert res[0].max() <= 1e-9                                 # ArnoldiMethod tol = 1e-9 (analyses.jl:954)


def test_rotating_blade_eigenvalues_parity():
    """Linearised eigen sweep over rotor speed (BASELINE config 3 style, on the config-1 topology): every instance against a dense
    eigen-solve of the oracle's K and M."""
    O = _oracle()
    nb, nev = 4, 8
    w = workloads.rotating_blade_batch(nb, rpm_jitter=0.5)
    ctx = _ctx(w, w["motion"], nb, w["points_b"])
    st = ctx.steady_solve()
    lam, vec, res = ctx.eigenvalue_analysis(st["x"], nev=nev, m=70)
    ctx.close()
    assert res.max() <= 1e-9
    for b in range(nb):
        pk = O.Packed(w["start"], w["stop"], w["elems"][b], w["points"], w["pd"], w["pl"], w["bc"][b], force_scaling=w["force_scaling"][b])
        dyn = O.make_dyn(vb=w["motion"][b, 0:3], wb=w["motion"][b, 3:6])
        K = O.steady_jacobian(pk, dyn, st["x"][b])
        M = O.mass_matrix(pk, st["x"][b])
        lam_ref, V_ref = O.eigenvalues_dense(K, M, nev)
        # compare as sets (conjugate pairs may come in either order)
        for l in lam_ref:
            assert np.min(np.abs(lam[b] - l)) <= 1e-9 * abs(l), (b, l)
        # eigenvectors: (K + λ M) v = 0
        for k in range(nev):
            v = vec[b][:, k]
            assert np.abs((K + lam[b, k] * M) @ v).max() <= 1e-7 * np.abs(K @ v).max()
