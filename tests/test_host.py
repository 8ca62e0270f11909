"""CPU tests of the host-side mirror of the reference's data model (gxbeam_b200/assembly.py, state.py, workloads.py)."""
import numpy as np

from gxbeam_b200.assembly import (Assembly, PrescribedConditions, default_force_scaling, discretize_beam, distributed_loads,
                                  pack_prescribed_conditions)
from gxbeam_b200 import workloads


def test_prescribed_conditions_flags_and_nan_fill():
    pc = PrescribedConditions(ux=0.0, theta_z=0.1, Fz=5.0, Mz=2.0)
    # ux: displacement only -> pd, !pl, load NaN ; theta_z with Mz: both (body-acceleration DOF) ; uz free with Fz
    assert list(pc.pd) == [1, 0, 0, 0, 0, 1] and list(pc.pl) == [0, 1, 1, 1, 1, 1]
    assert np.isnan(pc.values[6]) and pc.values[8] == 5.0 and pc.values[5] == 0.1 and pc.values[11] == 2.0
    assert np.isnan(pc.values[1]) and np.isnan(pc.values[2])


def test_prescribed_conditions_dof2_quirk():
    # loads.jl:121: the second DOF decides "displacement only" by looking at Fz, not Fy
    pc = PrescribedConditions(uy=0.0, Fy=3.0)
    assert pc.pd[1] == 1 and pc.pl[1] == 0 and np.isnan(pc.values[7])
    pc = PrescribedConditions(uy=0.0, Fz=3.0)
    assert pc.pd[1] == 1 and pc.pl[1] == 1 and pc.values[7] == 0.0


def test_default_packing_of_points_without_conditions():
    pd, pl, bc = pack_prescribed_conditions(3, {2: PrescribedConditions(Fx=1.0)})
    assert pd.sum() == 0 and pl.all() and bc[1, 6] == 1.0 and np.all(bc[0, 6:] == 0)


def test_default_force_scaling_is_power_of_two():
    asm = Assembly([[0, 0, 0], [1, 0, 0]], [1], [2], compliance=[np.diag([1e-6, 0, 0, 0, 1e-4, 1e-4])])
    fs = default_force_scaling(asm)
    assert fs == 2.0 ** np.ceil(np.log2(3 / (2e-4 + 1e-6) / 100))


def test_distributed_load_integration():
    asm = Assembly([[0, 0, 0], [2, 0, 0]], [1], [2])
    d = distributed_loads(asm, 1, fz=lambda s: 6.0 * s)   # s in [0,1] over a length-2 element
    # f1 = ∫(1-ξ) f ΔL dξ = 2*6*(1/2-1/3) = 2 ; f2 = 2*6/3 = 4
    assert np.isclose(d[2], 2.0) and np.isclose(d[5], 4.0) and np.count_nonzero(d) == 2


def test_discretize_curved_beam_endpoints():
    R = 100.0
    dL, xp, xm, Cab = discretize_beam(R * np.pi / 4, [0, 0, 0], 16, frame=np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]]),
                                     curvature=[0, 0, -1 / R])
    assert np.allclose(np.linalg.norm(xp - np.array([R, 0, 0]), axis=1), R, atol=1e-9)   # points lie on a circle of radius R
    assert all(np.allclose(c.T @ c, np.eye(3), atol=1e-12) for c in Cab)


def test_c2_workload_is_seeded_and_shaped():
    a = workloads.cantilever_tipmoment_batch(8, 10)
    b = workloads.cantilever_tipmoment_batch(8, 10)
    assert np.array_equal(a["elems"], b["elems"]) and np.array_equal(a["bc"], b["bc"], equal_nan=True)
    assert a["elems"].shape == (8, 10, 91) and a["bc"].shape == (8, 11, 18)
    assert np.all(a["force_scaling"] == 2.0 ** np.round(np.log2(a["force_scaling"])))


def test_compact_conditions_round_trip():
    """api.compact_conditions: the shared record + the varying entries reproduce the packed prescribed conditions exactly (NaN pattern
    included) -- the expansion gxb_set_bc_compact performs on the device, restated on the host."""
    from gxbeam_b200 import api, workloads
    w = workloads.cantilever_tipmoment_batch(7, 12, tip_force=True)
    bc = w["bc"]
    base, pts, fld, vals = api.compact_conditions(bc)
    assert base.shape == bc.shape[1:] and vals.shape == (7, len(pts)) and len(pts) == len(fld) == 2
    assert set(zip(pts.tolist(), fld.tolist())) == {(13, 11), (13, 8)}            # tip point: Mz and Fz vary
    rebuilt = np.tile(base[None], (7, 1, 1))
    rebuilt[:, pts - 1, fld] = vals
    assert np.array_equal(np.isnan(rebuilt), np.isnan(bc)) and np.array_equal(np.nan_to_num(rebuilt), np.nan_to_num(bc))
    # identical instances: nothing varies
    b2, p2, f2, v2 = api.compact_conditions(np.tile(bc[:1], (3, 1, 1)))
    assert len(p2) == 0 and v2.shape == (3, 0)
