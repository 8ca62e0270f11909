"""CPU tests: pin the oracle's steady Jacobian (with point masses) and mass matrix against the reference's recorded eigen-
frequencies (test/issues/pointmass.jl:54-58, 16 digits, default isapprox rtol ~1.5e-8), and check M = ∂r/∂ẋ by differentiation."""
import numpy as np

import cases
from oracle import pyoracle as O


def test_pointmass_frequencies_match_recorded_reference_output():
    case = cases.pointmass_jl()
    pk = cases.oracle_packed(case)
    dyn = O.make_dyn()
    st = O.steady_solve(pk, dyn)                       # eigenvalue_analysis runs steady_state_analysis first (analyses.jl:1299)
    assert st["converged"]
    K = O.steady_jacobian(pk, dyn, st["x"])
    M = O.mass_matrix(pk, st["x"])
    lam, V = O.eigenvalues_dense(K, M, 14)             # nev = 14 as in the test
    freq = np.sort(np.abs(lam.imag)) / (2 * np.pi)
    for f in case["ref_freq"]:
        assert np.min(np.abs(freq - f)) <= 1.5e-8 * f  # pointmass.jl:54-58
    # eigenpairs satisfy (K + λ M) v = 0
    for k in range(len(lam)):
        res = (K + lam[k] * M) @ V[:, k]
        assert np.abs(res).max() <= 1e-6 * np.abs(K @ V[:, k]).max()


def test_mass_matrix_is_derivative_of_newmark_residual_wrt_rates():
    """With xdot = 2/dt x - xdot_init, ∂r/∂x of the Newmark residual = K_steady-like part + (2/dt) M at fixed rates:
    M equals -∂r/∂(xdot_init) (analyses.jl:3749; jacobians.jl:287-300 checks the same identity with ForwardDiff)."""
    case = cases.random_anisotropic(2, nelem=3)
    pk = cases.oracle_packed(case)
    rng = np.random.default_rng(5)
    dyn = O.make_dyn(vb=rng.random(3), wb=rng.random(3))
    n = O.nstates(pk, static=False)
    npnt = len(case["points"])
    x = rng.random(n)
    init = rng.random((npnt, 12))
    dt = 0.37
    M = O.mass_matrix(pk, x)
    idx = O.indices(case["start"], case["stop"], False)
    h = 1e-6
    r0 = O.newmark_residual(pk, dyn, init, dt, x)
    for ip in range(npnt):
        for k in range(12):
            if k < 6 and case["pd"][ip, k]:
                continue                                  # prescribed displacement: the rate of that slot is not a state rate
            ini2 = init.copy(); ini2[ip, k] += h
            dr = (O.newmark_residual(pk, dyn, ini2, dt, x) - r0) / h      # ∂r/∂init = -∂r/∂xdot
            col = idx["icol_point"][ip] - 1 + k
            assert np.abs(-dr - M[:, col]).max() <= 1e-5 * max(1.0, np.abs(M[:, col]).max())
