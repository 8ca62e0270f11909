"""GPU parity tests of the batched cross-section analysis (gxb_section_*, gxbeam_b200/csrc/gxb_section.cu) against the numpy oracle
(oracle/section_oracle.py, pinned on the BECAS tables in tests/test_oracle_section.py) and against the reference's own fixtures."""
import numpy as np
import pytest

from gxbeam_b200 import api
from oracle import section_oracle as SO

pytestmark = pytest.mark.gpu

ISO1 = (100.0, 100.0, 100.0, 41.667, 41.667, 41.667, 0.2, 0.2, 0.2, 1.0)
ORTHO = (480.0, 120.0, 120.0, 60.0, 50.0, 60.0, 0.19, 0.26, 0.19, 1.0)


def _scaled_err(S, Sref):
    """|ΔS_ij| relative to sqrt(S_ii S_jj): the natural entry scale of a compliance / mass matrix (entries that vanish by symmetry are
    compared against the diagonal scale, not against their own round-off)."""
    d = np.sqrt(np.abs(np.diag(Sref)))
    return float((np.abs(S - Sref) / np.outer(d, d)).max())


def _random_layups(nb, ne, seed):
    rng = np.random.default_rng(seed)
    mats = np.tile(np.array(ORTHO)[None, None, :], (nb, ne, 1))
    mats[:, :, 0:6] *= rng.uniform(0.7, 1.3, (nb, ne, 1))          # stiffness scale per element (plies of different material)
    mats[:, :, 9] = rng.uniform(0.5, 2.0, (nb, ne))                  # density
    theta = rng.uniform(-np.pi / 3, np.pi / 3, (nb, ne))
    return mats, theta


def test_element_matrices_match_oracle():
    """element_submatrix (src/section.jl:486-501) of every element of a distorted mesh with random orthotropic plies: 1e-12."""
    nodes, conn = SO.square_mesh(4, 0.05)
    rng = np.random.default_rng(1)
    nodes = nodes + 0.003 * rng.standard_normal(nodes.shape)        # skewed quads: non-trivial Jacobians and element orientations
    ne = len(conn)
    mats, theta = _random_layups(3, ne, 2)
    sec = api.Section(len(nodes), conn + 1)
    out = sec.element_matrices(np.tile(nodes[None], (3, 1, 1)), mats, theta)
    sec.close()
    for b in range(3):
        for e in range(ne):
            ref = np.concatenate([m.ravel() for m in SO.element_submatrix(mats[b, e], theta[b, e], nodes[conn[e]])])
            assert np.abs(out[b, e] - ref).max() <= 1e-12 * np.abs(ref).max(), (b, e)


def test_becas_square_fixtures_on_the_device():
    """test/section.jl:45-84 and :140-160 on the DEVICE's output: isotropic square S1 and the orthotropic square at 22.5 degrees."""
    nodes, conn = SO.square_mesh()
    sec = api.Section(len(nodes), conn + 1)
    mats = np.stack([np.tile(np.array(ISO1)[None], (100, 1)), np.tile(np.array(ORTHO)[None], (100, 1))])
    theta = np.stack([np.zeros(100), np.full(100, 22.5 * np.pi / 180)])
    r = sec.properties(np.tile(nodes[None], (2, 1, 1)), mats, theta, gxbeam_order=False, shear_center=False)
    sec.close()
    assert r["ok"].all()
    K = np.linalg.inv(r["S"][0])
    for (i, v, tol) in ((0, 3.4899e-1, 1e-5), (1, 3.4899e-1, 1e-5), (2, 1.0, 1e-4), (3, 8.3384e-4, 1e-8), (4, 8.3384e-4, 1e-8), (5, 5.9084e-4, 1e-8)):
        assert abs(K[i, i] - v) <= tol
    assert np.abs(r["sc"][0]).max() <= 1e-8 and np.abs(r["tc"][0]).max() <= 1e-8
    K = np.linalg.inv(r["S"][1])
    for (i, j, v, tol) in ((0, 0, 7.598e-1, 1e-4), (2, 2, 3.435, 1e-3), (5, 5, 9.499e-4, 1e-7), (0, 2, 7.387e-1, 1e-4), (3, 5, -4.613e-4, 1e-7)):
        assert abs(K[i, j] - v) <= tol


@pytest.mark.parametrize("mesh", ["square", "ring", "half_ring"])
def test_section_properties_match_oracle(mesh):
    """compliance_matrix + mass_matrix of a batch of random layups (per-element material scale, density and ply angle, per-instance node
    perturbation) against the oracle, instance by instance: S and M to 1e-9 of their entry scale, centres to 1e-9 of the section size.
    The closed ring exercises the node renumbering (its natural numbering has the full bandwidth)."""
    if mesh == "square":
        nodes, conn = SO.square_mesh(8, 0.05)
    elif mesh == "ring":
        nodes, conn = SO.ring_mesh(0.1, 0.01, 4, 60)
    else:
        nodes, conn = SO.ring_mesh(0.1, 0.01, 4, 40, np.pi / 2, 3 * np.pi / 2, closed=False)
    nb, ne, nn = 6, len(conn), len(nodes)
    mats, theta = _random_layups(nb, ne, 7)
    rng = np.random.default_rng(11)
    nb_nodes = np.tile(nodes[None], (nb, 1, 1)) * rng.uniform(0.9, 1.1, (nb, 1, 1))       # per-instance scaling of the geometry
    sec = api.Section(nn, conn + 1)
    assert sec.half_bandwidth < 3 * nn // 2
    for go, scf in ((True, True), (False, False)):
        r = sec.properties(nb_nodes, mats, theta, gxbeam_order=go, shear_center=scf)
        assert r["ok"].all()
        size = np.abs(nodes).max()
        for b in range(nb):
            Sref, sc, tc = SO.compliance_matrix(nb_nodes[b], conn, mats[b], theta[b], gxbeam_order=go, shear_center=scf)
            Mref, mc = SO.mass_matrix(nb_nodes[b], conn, mats[b])
            assert _scaled_err(r["S"][b], Sref) <= 1e-9, (mesh, b, _scaled_err(r["S"][b], Sref))
            assert np.abs(r["sc"][b] - sc).max() <= 1e-9 * size and np.abs(r["tc"][b] - tc).max() <= 1e-9 * size
            assert _scaled_err(r["M"][b], Mref) <= 1e-12 and np.abs(r["mc"][b] - mc).max() <= 1e-12 * size
    sec.close()


def test_section_argument_errors():
    nodes, conn = SO.square_mesh(2, 0.05)
    with pytest.raises(api.GXBError, match="1-based"):
        api.Section(len(nodes), conn)                  # 0-based connectivity is refused


@pytest.mark.parametrize("gxbeam_order", [True, False])
def test_strain_recovery_and_tsai_wu_match_oracle(gxbeam_order):
    """strain_recovery (src/section.jl:952-1040) and tsai_wu (:1108-1131) of random layups of a closed multi-ply ring under random section loads,
    element by element against the oracle: every strain / stress within 1e-9 of the largest entry of its array (per layup), the Tsai-Wu index to
    1e-9 absolute-or-relative; S comes back with it (shear-centre form) and equals gxb_section_properties'."""
    nodes, conn = SO.ring_mesh(0.1, 0.012, 4, 48)
    nb, ne, nn = 5, len(conn), len(nodes)
    mats, theta = _random_layups(nb, ne, 21)
    rng = np.random.default_rng(22)
    nb_nodes = np.tile(nodes[None], (nb, 1, 1)) * rng.uniform(0.9, 1.1, (nb, 1, 1))
    F = rng.standard_normal((nb, 3)) * 50.0
    M = rng.standard_normal((nb, 3)) * 5.0
    strengths = np.tile(np.array([1500.0, 1200.0, 50.0, 250.0, 50.0, 250.0, 70.0, 70.0, 40.0])[None, None], (nb, ne, 1)) * rng.uniform(0.8, 1.2, (nb, ne, 1))
    sec = api.Section(nn, conn + 1)
    r = sec.strain_recovery(nb_nodes, mats, theta, F, M, strengths=strengths, gxbeam_order=gxbeam_order)
    p = sec.properties(nb_nodes, mats, theta, gxbeam_order=gxbeam_order)
    sec.close()
    # the global matrices are summed with floating-point atomics: two runs agree to round-off, not bit for bit
    assert r["ok"].all() and max(_scaled_err(r["S"][b], p["S"][b]) for b in range(nb)) <= 1e-11
    for b in range(nb):
        _, _, _, parts = SO.compliance_matrix(nb_nodes[b], conn, mats[b], theta[b], gxbeam_order=False, shear_center=False, return_parts=True)
        ref = SO.strain_recovery(F[b], M[b], nb_nodes[b], conn, mats[b], theta[b], parts, gxbeam_order=gxbeam_order)
        for name, want in zip(("strain_b", "stress_b", "strain_p", "stress_p"), ref):
            got = r[name][b].T
            assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max(), (b, name, np.abs(got - want).max() / np.abs(want).max())
        fw = SO.tsai_wu(ref[3], strengths[b])
        assert np.abs(r["failure"][b] - fw).max() <= 1e-9 * max(1.0, np.abs(fw).max()), b


def test_strain_recovery_beam_theory_on_the_device():
    """The isotropic square under axial force and bending: the DEVICE's axial stress is Fz / A + Mx y / Ixx - My x / Iyy (the same check that pins
    the oracle, tests/test_oracle_section.py); without strengths no Tsai-Wu index is returned, and asking for one without strengths is refused."""
    nodes, conn = SO.square_mesh()
    mats = np.tile(np.array(ISO1)[None, None], (2, 100, 1)); theta = np.zeros((2, 100))
    F = np.array([[0, 0, 3.0], [0, 0, 1.0]]); M = np.array([[0, 0, 0], [0.7, 0.4, 0]])
    sec = api.Section(len(nodes), conn + 1)
    r = sec.strain_recovery(np.tile(nodes[None], (2, 1, 1)), mats, theta, F, M, gxbeam_order=False)
    assert r["failure"] is None
    cen = nodes[conn].mean(axis=1)
    for b, tol in ((0, 1e-9), (1, 1e-3)):
        want = F[b, 2] / 0.01 + M[b, 0] * cen[:, 1] / (1e-4 / 12) - M[b, 1] * cen[:, 0] / (1e-4 / 12)
        assert np.abs(r["stress_b"][b][:, 2] - want).max() <= tol * np.abs(want).max()
    import ctypes as C
    out = np.empty((2, 100))
    rc = sec.lib.gxb_section_strain_recovery(sec.h, C.c_int64(2), api._ptr(np.tile(nodes[None], (2, 1, 1))), api._ptr(mats), api._ptr(theta), None,
                                             api._ptr(F), api._ptr(M), 0, None, None, None, None, api._ptr(out), None, None, None)
    assert rc != 0 and b"strengths" in sec.lib.gxb_last_error()
    sec.close()


def test_wide_band_section_matches_oracle():
    """A thick ring (20 layers through the wall): the renumbered half bandwidth exceeds 115, so the two columns of a factorisation step are
    handed over by more items than the CTA has threads and the border lanes sit above the widest column -- same parity bar as the thin
    meshes."""
    nodes, conn = SO.ring_mesh(0.1, 0.03, 20, 40)
    nb, ne, nn = 2, len(conn), len(nodes)
    mats, theta = _random_layups(nb, ne, 31)
    sec = api.Section(nn, conn + 1)
    assert sec.half_bandwidth > 115
    r = sec.properties(np.tile(nodes[None], (nb, 1, 1)), mats, theta, gxbeam_order=False, shear_center=False)
    sec.close()
    assert r["ok"].all()
    for b in range(nb):
        Sref, sc, tc = SO.compliance_matrix(nodes, conn, mats[b], theta[b], gxbeam_order=False, shear_center=False)
        assert _scaled_err(r["S"][b], Sref) <= 1e-9, (b, _scaled_err(r["S"][b], Sref))
        assert np.abs(r["sc"][b] - sc).max() <= 1e-9 * 0.1 and np.abs(r["tc"][b] - tc).max() <= 1e-9 * 0.1
