"""CPU tests: pin the oracle's Newmark (time-domain) path against the reference's self-consistency test
(test/jacobians.jl:238-261: analytic Newmark Jacobian == derivative of the residual, structural damping on, random initialisation
terms and dt) and against physical invariants of time marching from a steady state."""
import numpy as np
import pytest

import cases
from gxbeam_b200 import workloads
from oracle import pyoracle as O


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("damping", [False, True])
def test_newmark_jacobian_is_derivative_of_residual(seed, damping):
    case = cases.random_anisotropic(seed, nelem=2 + seed)
    pk = cases.oracle_packed(case)
    rng = np.random.default_rng(seed + 20)
    dyn = O.make_dyn(vb=1e2 * rng.random(3), wb=1e2 * rng.random(3), structural_damping=damping)
    n = O.nstates(pk, static=False)
    init = rng.random((len(case["points"]), 12))          # udot_init .. Ωdot_init = rand  (jacobians.jl:243-246)
    dt = rng.random()
    x = 1e2 * rng.random(n)
    J = O.newmark_jacobian(pk, dyn, init, dt, x)
    Jc = O.newmark_jacobian(pk, dyn, init, dt, x, complex_step=True)
    assert np.abs(J - Jc).max() <= 1e-10 * max(1.0, np.abs(Jc).max())


def _blade():
    w = workloads.rotating_blade_batch(1)
    pk = O.Packed(w["start"], w["stop"], w["elems"][0], w["points"], w["pd"], w["pl"], w["bc"][0], force_scaling=w["force_scaling"][0])
    return w, pk


@pytest.mark.parametrize("damping", [False, True])
def test_time_marching_from_steady_state_stays_put(damping):
    """A steady rotating equilibrium with zero rates is a fixed point of the Newmark recurrence (analyses.jl:2693-2738): every step
    must converge in zero iterations and leave x unchanged."""
    w, pk = _blade()
    dyn = O.make_dyn(vb=w["motion"][0, 0:3], wb=w["motion"][0, 3:6], structural_damping=damping)
    st = O.steady_solve(pk, dyn, ftol=1e-12)
    tvec = np.linspace(0, 0.01, 21)
    r = O.newmark_run(pk, dyn, tvec, st["x"], np.zeros((25, 12)))
    assert r["converged"] and r["steps_done"] == 21
    assert np.abs(r["x"] - st["x"]).max() <= 1e-8 * np.abs(st["x"]).max()


def test_gust_response_config4_topology():
    """BASELINE config 4 topology (benchmark/benchmarks.jl:143-216): force pulse 100·tf1(t) at point 21, structural damping on."""
    w, pk = _blade()
    dyn = O.make_dyn(vb=w["motion"][0, 0:3], wb=w["motion"][0, 3:6], structural_damping=True)
    st = O.steady_solve(pk, dyn)
    nt = 101
    tvec = np.linspace(0, 0.025, nt)      # dt = 0.25 ms as in the benchmark
    tf1 = lambda t: 0.1 * (1 - np.sin(2 * np.pi * (t / 0.4 + 0.25))) if t < 0.2 else 0.0
    bct = np.tile(w["bc"][0][None], (nt, 1, 1))
    for k, t in enumerate(tvec):
        bct[k, 20, 7] = 100 * tf1(t)
    r = O.newmark_run(pk, dyn, tvec, st["x"], np.zeros((25, 12)), bct)
    assert r["converged"] and r["steps_done"] == nt
    assert 1 <= r["iters"][1:].max() <= 4
    idx = O.indices(w["start"], w["stop"], False)
    tip = idx["icol_point"][-1] - 1
    assert r["x"][-1, tip + 1] > r["x"][0, tip + 1]      # the blade tip moves in +y under the +y pulse
