"""CPU tests: pin the oracle's steady-state (DynamicSystem) path against the reference's own self-consistency test
(test/jacobians.jl:177-205: analytic Jacobian == derivative of the residual with structural damping, random body-frame
velocities/accelerations, a body-acceleration state) and against the static path it must reduce to."""
import numpy as np
import pytest

import cases
from oracle import pyoracle as O


def _rand_dyn(rng, damping):
    return O.make_dyn(vb=1e2 * rng.random(3), wb=1e2 * rng.random(3), ab=1e2 * rng.random(3), alb=1e2 * rng.random(3),
                      structural_damping=damping)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("damping", [False, True])
def test_steady_jacobian_is_derivative_of_residual_reference_system(seed, damping):
    case = cases.jacobians_jl(seed)
    pk = cases.oracle_packed(case)
    kl, ku, icb = O.steady_bandwidth(pk)
    assert list(icb) == [0, 0, 0, 0, 5, 0]        # root theta_y has pd & pl: αb[2] lives in x[icol_point[1]+4]
    rng = np.random.default_rng(seed + 10)
    dyn = _rand_dyn(rng, damping)
    n = O.nstates(pk, static=False)
    assert n == 18 * 2 + 12
    x = 1e2 * rng.random(n)                       # test/jacobians.jl:181
    J = O.steady_jacobian(pk, dyn, x)
    Jc = O.steady_jacobian(pk, dyn, x, complex_step=True)
    assert np.abs(J - Jc).max() <= 1e-10 * max(1.0, np.abs(Jc).max())


@pytest.mark.parametrize("two_d", [False, True])
def test_steady_jacobian_general_loads(two_d):
    case = cases.random_anisotropic(3, nelem=4)
    pk = cases.oracle_packed(case, two_dimensional=two_d)
    assert O.steady_bandwidth(pk)[:2] == (23, 29)
    rng = np.random.default_rng(7)
    dyn = _rand_dyn(rng, True)
    x = rng.random(O.nstates(pk, static=False))
    J = O.steady_jacobian(pk, dyn, x)
    Jc = O.steady_jacobian(pk, dyn, x, complex_step=True)
    assert np.abs(J - Jc).max() <= 1e-10 * max(1.0, np.abs(Jc).max())


def test_steady_reduces_to_static_at_rest():
    """No body motion, zero velocities: the steady solution's (u, θ, F, M) equal the static solution (rotating.jl uses this at 0 rpm)."""
    case = cases.tipmoment(0.8)
    pk = cases.oracle_packed(case)
    rs = O.static_solve(pk)
    rd = O.steady_solve(pk, O.make_dyn())
    assert rs["converged"] and rd["converged"]
    ids, idd = O.indices(case["start"], case["stop"], True), O.indices(case["start"], case["stop"], False)
    for ip in range(len(ids["icol_point"])):
        a, b = ids["icol_point"][ip] - 1, idd["icol_point"][ip] - 1
        assert np.allclose(rd["x"][b:b + 6], rs["x"][a:a + 6], rtol=1e-9, atol=1e-12)
        assert np.allclose(rd["x"][b + 6:b + 12], 0.0, atol=1e-9)
    for ie in range(len(ids["icol_elem"])):
        a, b = ids["icol_elem"][ie] - 1, idd["icol_elem"][ie] - 1
        assert np.allclose(rd["x"][b:b + 6], rs["x"][a:a + 6], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("builder", ["zero_mass_jl", "zero_length_element_jl"])
def test_zeros_jl_steady_cases_converge(builder):
    """test/issues/zeros.jl:3-168: zero mass matrix, zero-length element -- the reference asserts `converged`."""
    case = getattr(cases, builder)()
    pk = cases.oracle_packed(case)
    r = O.steady_solve(pk, O.make_dyn(wb=case["angular_velocity"]))
    assert r["converged"] and r["error"] == 0
