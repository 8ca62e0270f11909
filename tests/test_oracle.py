"""CPU tests: pin the oracle (oracle/) against the reference's own tests, fixtures and recorded outputs.

Fixtures restated: test/jacobians.jl:3-39 (rotation derivative identities, θ = 1e3·rand),
test/jacobians.jl:158-175 (analytic Jacobian == derivative of the residual, atol 1e-10), test/examples/tipmoment.jl,
cantilever.jl, overdetermined.jl, tipforce.jl (analytic solutions), curved.jl:56-58 (16-digit reference output).
"""
import numpy as np
import pytest

import cases
from oracle import pyoracle as O
from gxbeam_b200.state import assembly_state


@pytest.mark.parametrize("scale", [1.0, 3.9, 5.0, 13.0, 1e3])
def test_rotation_derivatives_match_complex_step(scale):
    rng = np.random.default_rng(int(scale * 10))
    for _ in range(5):
        th = scale * rng.random(3)
        dth = scale * rng.random(3)
        a = O.rotation(th, dth)
        c = O.rotation_cs(th, dth)
        for k in ("C_th", "Q_th", "Qinv_th", "dQ_th"):
            assert np.allclose(a[k], c[k], rtol=1e-10, atol=1e-10 * max(1.0, np.abs(c[k]).max())), k
        assert np.allclose(a["C"] @ a["C"].T, np.eye(3), atol=1e-13)
        assert np.allclose(a["Q"] @ a["Qinv"], np.eye(3), atol=1e-12)
        # get_ΔQ(θ,Δθ) == mul3(Q_θ1,Q_θ2,Q_θ3, Δθ)   (test/jacobians.jl:30-31)
        mul3 = np.stack([a["Q_th"][k] @ dth for k in range(3)], axis=1)
        assert np.allclose(a["dQ"], mul3, rtol=1e-10, atol=1e-12)


def test_rotation_scaling_branches():
    # math.jl:52-67: scaling = 1 inside |θ| <= 4, wraps by 8m/|θ| outside
    assert O.rotation([1.0, 2.0, 3.0], [0, 0, 0])["scaling"] == 1.0
    assert O.rotation([4.0, 0.0, 0.0], [0, 0, 0])["scaling"] == 1.0
    assert np.isclose(O.rotation([5.0, 0.0, 0.0], [0, 0, 0])["scaling"], 1 - 8 / 5)
    assert np.isclose(O.rotation([0.0, 13.0, 0.0], [0, 0, 0])["scaling"], 1 - 16 / 13)


def test_indices_chain_formula():
    # chain => icol_point[p] = s(p-1)+1, icol_elem[e] = s(e-1)+(s-5); s = 12 static / 18 dynamic
    n = 7
    start, stop = np.arange(1, n + 1), np.arange(2, n + 2)
    for static, s in ((True, 12), (False, 18)):
        idx = O.indices(start, stop, static)
        assert idx["nstates"] == s * n + (s - 6)
        assert list(idx["icol_point"]) == [s * p + 1 for p in range(n + 1)]
        assert list(idx["icol_elem"]) == [s * e + (s - 5) for e in range(n)]
        assert list(idx["irow_point"]) == list(idx["icol_point"]) and list(idx["irow_elem"]) == list(idx["icol_elem"])


def test_indices_general_connectivity():
    # a Y-joint: elements (1->2), (2->3), (2->4): points are numbered in order of first appearance (system.jl:55-116)
    idx = O.indices([1, 2, 2], [2, 3, 4], True)
    assert idx["nstates"] == 6 * 4 + 6 * 3
    assert list(idx["icol_point"]) == [1, 13, 25, 37]
    assert list(idx["icol_elem"]) == [7, 19, 31]


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("two_d", [False, True])
def test_static_jacobian_is_derivative_of_residual(seed, two_d):
    case = cases.random_anisotropic(seed, nelem=3)
    pk = cases.oracle_packed(case, two_dimensional=two_d)
    n = O.nstates(pk)
    rng = np.random.default_rng(seed + 100)
    x = rng.random(n)       # jacobians.jl uses x = rand(n) scaled; θ entries stay in the scaling == 1 branch ...
    J = O.static_jacobian(pk, x)
    Jc = O.static_jacobian(pk, x, complex_step=True)
    assert np.abs(J - Jc).max() <= 1e-10 * max(1.0, np.abs(Jc).max())
    x = 1e2 * rng.random(n)  # ... and this one exercises scaling != 1 (test/jacobians.jl:162)
    J = O.static_jacobian(pk, x)
    Jc = O.static_jacobian(pk, x, complex_step=True)
    assert np.abs(J - Jc).max() <= 1e-9 * max(1.0, np.abs(Jc).max())
    assert O.static_bandwidth(pk) == (14, 17)


@pytest.mark.parametrize("lam", [0.0, 0.4, 0.8, 1.2, 1.6, 1.8, 2.0])
def test_tipmoment_analytic(lam):
    case = cases.tipmoment(lam)
    pk = cases.oracle_packed(case)
    r = O.static_solve(pk)
    assert r["converged"] and r["error"] == 0
    idx = O.indices(case["start"], case["stop"])
    st = assembly_state(r["x"], case["start"], case["stop"], idx["icol_point"], idx["icol_elem"], case["pd"], case["pl"], case["bc"][0],
                        case["force_scaling"])
    rho = case["rho"]
    for ip, xi in enumerate(case["points"][:, 0]):
        ua = np.zeros(2) if np.isinf(rho) else np.array([rho * np.sin(xi / rho) - xi, rho * (1 - np.cos(xi / rho))])
        assert np.allclose(st.point_u[ip, :2], ua, atol=5e-2)      # tipmoment.jl:76-77
    for ie, el in enumerate(case["asm"].elements):
        xi = el.x[0]
        ua = np.zeros(2) if np.isinf(rho) else np.array([rho * np.sin(xi / rho) - xi, rho * (1 - np.cos(xi / rho))])
        assert np.allclose(st.elem_u[ie, :2], ua, atol=5e-2)       # tipmoment.jl:67-68


def test_curved_beam_matches_recorded_reference_output():
    case = cases.curved()
    r = O.static_solve(cases.oracle_packed(case))
    assert r["converged"]
    idx = O.indices(case["start"], case["stop"])
    ic = idx["icol_point"][-1] - 1
    u = r["x"][ic:ic + 3]
    # the reference's own output, recorded in test/examples/curved.jl:56-58
    assert np.allclose(u, case["ref_tip_u"], rtol=1e-9, atol=0)


def test_cantilever_partial_load_linear():
    case = cases.cantilever_partial_load()
    r = O.static_solve(cases.oracle_packed(case), linear=True)
    assert r["converged"] and r["iters"] == 1
    idx = O.indices(case["start"], case["stop"])
    st = assembly_state(r["x"], case["start"], case["stop"], idx["icol_point"], idx["icol_elem"], case["pd"], case["pl"], case["bc"][0],
                        case["force_scaling"])
    for ip, xi in enumerate(case["points"][:, 0]):
        assert abs(st.point_u[ip, 2] - case["deflection"](xi)) <= 1e-8     # cantilever.jl:93
    for ie, el in enumerate(case["asm"].elements):
        assert abs(st.elem_u[ie, 2] - case["deflection"](el.x[0])) <= 1e-9  # cantilever.jl:85


def test_overdetermined_linear():
    case = cases.overdetermined()
    r = O.static_solve(cases.oracle_packed(case), linear=True)
    idx = O.indices(case["start"], case["stop"])
    st = assembly_state(r["x"], case["start"], case["stop"], idx["icol_point"], idx["icol_elem"], case["pd"], case["pl"], case["bc"][0],
                        case["force_scaling"])
    for ip, xi in enumerate(case["points"][:, 0]):
        assert abs(st.point_u[ip, 2] - case["deflection"](xi)) <= 1e-8     # overdetermined.jl:66


def test_tipforce_converges_and_batch_matches_single():
    cs = [cases.tipforce(lam) for lam in (0.5, 2.0, 8.0, 16.0)]
    singles = [O.static_solve(cases.oracle_packed(c)) for c in cs]
    assert all(s["converged"] for s in singles)
    # same four instances through the threaded batch entry
    elems = np.concatenate([c["elems"] for c in cs])
    bc = np.concatenate([c["bc"] for c in cs])
    pk = O.Packed(cs[0]["start"], cs[0]["stop"], elems, cs[0]["points"], cs[0]["pd"], cs[0]["pl"], bc)
    fs = np.array([c["force_scaling"] for c in cs])
    rb = O.static_solve_batch(pk, 4, force_scaling=fs, nthreads=2)
    for i, s in enumerate(singles):
        assert np.array_equal(rb["x"][i], s["x"]) and rb["iters"][i] == s["iters"] and rb["converged"][i] == 1


def test_opcount_matches_bench_constants():
    import os
    import sys
    """SURVEY §8(d): f_e, f_p are frozen by an operation-counting run of the oracle; bench.py's roofline uses exactly those numbers."""
    import re
    import subprocess
    root0 = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root0 not in sys.path:
        sys.path.insert(0, root0)
    import bench
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run(["make", "-s", "-C", os.path.join(root, "oracle"), "opcount"], capture_output=True, text=True, check=True).stdout
    fe = int(re.search(r"f_e \(fused: properties once\)\s+= (\d+)", out).group(1))
    fp = int(re.search(r"f_p \(residual \+ Jacobian, free interior point\)\s+= (\d+)", out).group(1))
    assert (fe, fp) == (bench.F_ELEMENT_STATIC, bench.F_POINT_STATIC)
    assert bench.algorithmic_flops_per_iter(100) == 100 * fe + 101 * fp + 1155348
