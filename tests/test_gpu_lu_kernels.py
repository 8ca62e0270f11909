"""The two factorisation kernels behind `nlsolve!`'s linear solve (src/analyses.jl:3518, 3585): the tensor-core one (gxb_lu_mma.cuh, default)
and the row-per-lane one (gxb_lu.cuh, GXB_LU_MMA=0).  Both are held to a dense LAPACK solve of the oracle's Jacobian on chain lengths that
exercise every window situation of the 8-column elimination (fewer rows than slots, virtual identity rows padding n to a multiple of 8,
6-row blocks entering one or two at a time), and to each other."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from gxbeam_b200 import api

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import cases
import test_gpu_static as T
out = {{}}
for nelem in {sizes!r}:
    case = cases.random_anisotropic(5, nelem=nelem)
    ctx = T._ctx_for(case)
    x = np.random.default_rng(nelem).random((1, ctx.n))
    out[str(nelem)] = ctx.static_newton_direction(x)[0].tolist()
    ctx.close()
print("RESULT" + json.dumps(out))
"""

SIZES = [1, 2, 3, 5, 7, 12, 33, 100, 255]


def _directions(mma):
    env = dict(os.environ, GXB_LU_MMA=str(mma))
    code = _CHILD.format(root=ROOT, tests=os.path.join(ROOT, "tests"), sizes=SIZES)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout
    line = [l for l in out.splitlines() if l.startswith("RESULT")][-1]
    return {int(k): np.array(v) for k, v in json.loads(line[6:]).items()}


def test_both_factorisations_match_dense_solve_and_each_other():
    from oracle import pyoracle as O
    mma, rowlane = _directions(3), _directions(0)
    for nelem in SIZES:
        case = cases.random_anisotropic(5, nelem=nelem)
        pk = cases.oracle_packed(case)
        n = 12 * nelem + 6
        x = np.random.default_rng(nelem).random(n)
        p_ref = -np.linalg.solve(O.static_jacobian(pk, x), O.static_residual(pk, x))
        scale = np.abs(p_ref).max()
        # tolerance: 1e-9 relative (BASELINE north_star), the dense solve itself carries cond(J) * eps
        assert np.abs(mma[nelem] - p_ref).max() < 1e-9 * scale, nelem
        assert np.abs(rowlane[nelem] - p_ref).max() < 1e-9 * scale, nelem
        assert np.abs(mma[nelem] - rowlane[nelem]).max() < 1e-9 * scale, nelem


_CHILD_DYN = r"""
import json, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import cases
import test_gpu_steady as T
out = {{}}
for nelem in {sizes!r}:
    case = cases.random_anisotropic(4, nelem=nelem, pmass=True)
    rng = np.random.default_rng(100 + nelem)
    mo = 10 * rng.random(12)
    ctx = T._ctx(case, mo[None])
    x = rng.random((1, ctx.n))
    out[str(nelem)] = ctx.newton_direction(x)[0].tolist()
    ctx.close()
print("RESULT" + json.dumps(out))
"""

SIZES_DYN = [1, 2, 3, 4, 8, 24, 61]


def _directions_dyn(mma, two=0):
    env = dict(os.environ, GXB_LU_MMA=str(mma), GXB_LU_TWO=str(two))
    code = _CHILD_DYN.format(root=ROOT, tests=os.path.join(ROOT, "tests"), sizes=SIZES_DYN)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout
    line = [l for l in out.splitlines() if l.startswith("RESULT")][-1]
    return {int(k): np.array(v) for k, v in json.loads(line[6:]).items()}


def test_both_factorisations_dynamic_systems():
    """Same for DynamicSystem Jacobians (18 N + 12 states, 32 x 56 window, two pending slots per lane in the back substitution)."""
    from oracle import pyoracle as O
    mma, rowlane, two = _directions_dyn(3), _directions_dyn(0), _directions_dyn(3, two=1)
    for nelem in SIZES_DYN:
        case = cases.random_anisotropic(4, nelem=nelem, pmass=True)
        pk = cases.oracle_packed(case)
        rng = np.random.default_rng(100 + nelem)
        mo = 10 * rng.random(12)
        dyn = O.make_dyn(vb=mo[0:3], wb=mo[3:6], ab=mo[6:9], alb=mo[9:12])
        x = rng.random(18 * nelem + 12)
        p_ref = -np.linalg.solve(O.steady_jacobian(pk, dyn, x), O.steady_residual(pk, dyn, x))
        scale = np.abs(p_ref).max()
        assert np.abs(mma[nelem] - p_ref).max() < 1e-9 * scale, nelem
        assert np.abs(rowlane[nelem] - p_ref).max() < 1e-9 * scale, nelem
        assert np.abs(mma[nelem] - rowlane[nelem]).max() < 1e-9 * scale, nelem
        # the two-sided factorisation (gxb_lu_two.cuh; chains of fewer than 5 elements fall back to one warp per instance)
        assert np.abs(two[nelem] - p_ref).max() < 1e-9 * scale, nelem


# ---- partitioned factorisation (gxb_lu_part.cuh): P warps per instance, static systems ----
PART_SIZES = [15, 16, 33, 64, 100, 255]


@pytest.mark.parametrize("parts", [2, 3, 4, 6, 8])
def test_partitioned_factorisation_matches_dense_solve(parts):
    """Newton direction of the partitioned (substructured) factorisation vs a dense LAPACK solve of the oracle's Jacobian and vs the
    one-warp-per-instance kernel, on chains whose cuts fall at every position relative to the 6-row blocks / TMA ring, including the
    shortest chain a given P accepts (4 stations per partition; shorter chains fall back to fewer partitions)."""
    from oracle import pyoracle as O
    import test_gpu_static as T
    for nelem in PART_SIZES:
        case = cases.random_anisotropic(5, nelem=nelem)
        ctx = T._ctx_for(case)
        n = 12 * nelem + 6
        x = np.random.default_rng(nelem).random((1, n))
        p_part = ctx.static_newton_direction(x, api.make_opts(lu_parts=parts))[0]
        p_one = ctx.static_newton_direction(x, api.make_opts(lu_parts=1))[0]
        ctx.close()
        pk = cases.oracle_packed(case)
        p_ref = -np.linalg.solve(O.static_jacobian(pk, x[0]), O.static_residual(pk, x[0]))
        scale = np.abs(p_ref).max()
        # 1e-9 relative (BASELINE north_star); both are backward-stable pivoted eliminations of the same matrix
        assert np.abs(p_part - p_ref).max() < 1e-9 * scale, (parts, nelem)
        assert np.abs(p_part - p_one).max() < 1e-9 * scale, (parts, nelem)


@pytest.mark.parametrize("parts", [2, 3, 4, 6, 8])
def test_partitioned_solve_iteration_parity(parts):
    """Whole static_analysis of a sample of BASELINE config 2 with the partitioned factorisation: converged, Newton iteration counts
    equal to the oracle's, states within 1e-9 (component-wise against the group scale)."""
    from oracle import pyoracle as O
    from gxbeam_b200 import workloads
    nb = 48
    w = workloads.cantilever_tipmoment_batch(nb, 100, tip_force=True)
    ctx = api.Context(0)
    ctx.topology(w["start"], w["stop"], w["pd"], w["pl"])
    ctx.set_batch(nb)
    ctx.set_elements(w["elems"]); ctx.set_bc(w["bc"]); ctx.set_body(None, w["force_scaling"])
    res = ctx.static_solve(api.make_opts(lu_parts=parts))
    ctx.close()
    pk = O.Packed(w["start"], w["stop"], w["elems"], w["points"], w["pd"], w["pl"], w["bc"])
    ref = O.static_solve_batch(pk, nb, force_scaling=w["force_scaling"])
    assert res["converged"].all() and ref["converged"].all()
    assert np.array_equal(res["iters"], ref["iters"])
    assert np.abs(res["x"] - ref["x"]).max() < 1e-9 * np.abs(ref["x"]).max()
