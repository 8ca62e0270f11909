"""The section-analysis oracle (oracle/section_oracle.py, a numpy restatement of src/section.jl) against the BECAS user-guide tables that the
reference's own tests hold: test/section.jl:40-340 (same meshes, same materials, same tolerances)."""
import numpy as np
import pytest

from oracle import section_oracle as S

ISO1 = (100.0, 100.0, 100.0, 41.667, 41.667, 41.667, 0.2, 0.2, 0.2, 1.0)
ISO2 = tuple(v / 10 for v in ISO1[:6]) + (0.2, 0.2, 0.2, 1.0)
ORTHO = (480.0, 120.0, 120.0, 60.0, 50.0, 60.0, 0.19, 0.26, 0.19, 1.0)


def _K(nodes, conn, mats, thetas):
    Sm, sc, tc = S.compliance_matrix(nodes, conn, mats, thetas, gxbeam_order=False, shear_center=False)
    return np.linalg.inv(Sm), sc, tc


def test_material_stiffness_is_symmetric_and_isotropic_limit():
    """test/section.jl:6-36 in spirit: the constitutive matrix of an isotropic material has the textbook entries."""
    E, nu = 100.0, 0.2
    G = E / (2 * (1 + nu))
    Q = S.stiffness((E, E, E, G, G, G, nu, nu, nu, 1.0))
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    assert np.allclose(Q, Q.T)
    assert np.isclose(Q[0, 0], lam + 2 * G) and np.isclose(Q[5, 5], lam + 2 * G) and np.isclose(Q[0, 1], lam) and np.isclose(Q[2, 2], G)


def test_square_isotropic_S1():
    """test/section.jl:45-84"""
    nodes, conn = S.square_mesh()
    K, sc, tc = _K(nodes, conn, [ISO1] * 100, np.zeros(100))
    for (i, j, v, tol) in ((0, 0, 3.4899e-1, 1e-5), (1, 1, 3.4899e-1, 1e-5), (2, 2, 1.0, 1e-4), (3, 3, 8.3384e-4, 1e-8), (4, 4, 8.3384e-4, 1e-8),
                           (5, 5, 5.9084e-4, 1e-8)):
        assert abs(K[i, j] - v) <= tol, (i, j, K[i, j])
    assert np.abs(sc).max() <= 1e-8 and np.abs(tc).max() <= 1e-8


def test_square_two_materials_S2():
    """test/section.jl:86-115"""
    nodes, conn = S.square_mesh()
    mats = [ISO1 if e // 10 < 5 else ISO2 for e in range(100)]
    K, _, _ = _K(nodes, conn, mats, np.zeros(100))
    for (i, j, v, tol) in ((0, 0, 1.28e-1, 1e-3), (1, 1, 1.92e-1, 1e-3), (2, 2, 5.50e-1, 1e-3), (3, 3, 4.59e-4, 1e-6), (4, 4, 4.59e-4, 1e-6),
                           (5, 5, 2.77e-4, 1e-6), (1, 5, -3.93e-3, 1e-5), (2, 4, 1.13e-2, 1e-4)):
        assert abs(K[i, j] - v) <= tol, (i, j, K[i, j])


@pytest.mark.parametrize("deg,table", [
    (0.0, ((0, 0, 5.039e-1, 1e-4), (1, 1, 4.201e-1, 1e-4), (2, 2, 4.800, 1e-3), (3, 3, 4.001e-3, 1e-6), (4, 4, 4.001e-3, 1e-6), (5, 5, 7.737e-4, 1e-7))),
    (22.5, ((0, 0, 7.598e-1, 1e-4), (1, 1, 4.129e-1, 1e-4), (2, 2, 3.435, 1e-3), (3, 3, 2.489e-3, 1e-6), (4, 4, 2.274e-3, 1e-6), (5, 5, 9.499e-4, 1e-7),
            (0, 2, 7.387e-1, 1e-4), (3, 5, -4.613e-4, 1e-7))),
    (90.0, ((0, 0, 5.0202e-1, 1e-5), (1, 1, 5.0406e-1, 1e-5), (2, 2, 1.2, 1e-3), (3, 3, 1.0004e-3, 1e-7), (4, 4, 1.0002e-3, 1e-7), (5, 5, 8.5081e-4, 1e-8))),
])
def test_square_orthotropic_S3(deg, table):
    """test/section.jl:117-181: orthotropic material at ply angles 0, 22.5 and 90 degrees"""
    nodes, conn = S.square_mesh()
    K, _, _ = _K(nodes, conn, [ORTHO] * 100, np.full(100, deg * np.pi / 180))
    for (i, j, v, tol) in table:
        assert abs(K[i, j] - v) <= tol, (deg, i, j, K[i, j])


def test_cylinder_C1_and_half_cylinder_C2():
    """test/section.jl:184-290"""
    nodes, conn = S.ring_mesh(0.1, 0.01, 4, 120)
    K, _, _ = _K(nodes, conn, [ISO1] * len(conn), np.zeros(len(conn)))
    for (i, j, v, tol) in ((0, 0, 1.249e-1, 0.5e-4), (1, 1, 1.249e-1, 0.5e-4), (2, 2, 5.965e-1, 1.5e-4), (3, 3, 2.697e-3, 2e-6), (4, 4, 2.697e-3, 2e-6),
                           (5, 5, 2.248e-3, 1e-6)):
        assert abs(K[i, j] - v) <= tol, (i, j, K[i, j])
    nodes, conn = S.ring_mesh(0.1, 0.01, 4, 60, np.pi / 2, 3 * np.pi / 2, closed=False)
    K, sc, tc = _K(nodes, conn, [ISO1] * len(conn), np.zeros(len(conn)))
    for (i, j, v, tol) in ((0, 0, 4.964e-2, 2e-5), (1, 1, 6.244e-2, 2e-5), (2, 2, 2.982e-1, 2e-4), (3, 3, 1.349e-3, 1e-6), (4, 4, 1.349e-3, 1e-6),
                           (5, 5, 9.120e-4, 3e-7), (2, 4, 1.805e-2, 1e-5), (1, 5, -7.529e-3, 3e-6)):
        assert abs(K[i, j] - v) <= tol, (i, j, K[i, j])
    assert abs(sc[0] + 1.206e-1) <= 1e-4 and abs(sc[1]) <= 1e-6 and abs(tc[0] + 6.051e-2) <= 2e-5 and abs(tc[1]) <= 1e-6


def test_circular_tube_gxbeam_order():
    """test/section.jl:292-346 (Chen, Yu, Capellaro): default ordering and shear-centre transformation"""
    E, nu, R = 73e9, 0.33, 0.3
    G = E / (2 * (1 + nu))
    t = (1.0 / 3) * 2 * R
    mat = (E, E, E, G, G, G, nu, nu, nu, 2800.0)
    nodes, conn = S.ring_mesh(R, t, 20, 100)
    Sm, _, _ = S.compliance_matrix(nodes, conn, [mat] * len(conn), np.zeros(len(conn)))
    K = np.linalg.inv(Sm)
    for (i, v, rt) in ((0, 1.835e10, 1e-3), (1, 4.682e9, 1e-3), (2, 4.682e9, 1e-3), (3, 3.519e8, 3e-2), (4, 4.587e8, 2e-3), (5, 4.587e8, 2e-3)):
        assert abs(K[i, i] - v) <= rt * v, (i, K[i, i])
    # mass matrix (section.jl:840-890): mass per length and polar inertia of the annulus, up to the polygonal discretisation
    M, mc = S.mass_matrix(nodes, conn, [mat] * len(conn))
    area = np.pi * (R ** 2 - (R - t) ** 2)
    assert abs(M[0, 0] - 2800.0 * area) <= 2e-3 * 2800.0 * area and np.abs(mc).max() <= 1e-12
    assert np.allclose(M, M.T)


def test_strain_recovery_isotropic_square_beam_theory():
    """strain_recovery (section.jl:952-1040) on the homogeneous isotropic square: axial force and pure bending reproduce the Euler-Bernoulli
    stress field szz = Fz / A + Mx y / Ixx - My x / Iyy at the element centres, torsion-free loads leave the shear stresses at zero, and the ply
    frame of an unrotated element on a horizontal edge is the beam frame with the fibre (1) axis along z."""
    nodes, conn = S.square_mesh()
    mats = np.array([ISO1] * 100); th = np.zeros(100)
    _, _, _, parts = S.compliance_matrix(nodes, conn, mats, th, gxbeam_order=False, shear_center=False, return_parts=True)
    side = 0.1; area = side ** 2; inertia = side ** 4 / 12
    cen = nodes[conn].mean(axis=1)
    for F, M in (([0, 0, 3.0], [0, 0, 0]), ([0, 0, 0], [2.0, 0, 0]), ([0, 0, 0], [0, -1.5, 0]), ([0, 0, 1.0], [0.7, 0.4, 0])):
        eb, sb, ep, sp_ = S.strain_recovery(F, M, nodes, conn, mats, th, parts, gxbeam_order=False)
        want = F[2] / area + M[0] * cen[:, 1] / inertia - M[1] * cen[:, 0] / inertia
        scale = np.abs(want).max()
        # exact for the axial load; bending carries the 10 x 10 bilinear mesh's discretisation error (anticlastic in-plane warping), 5.5e-4
        tol = 1e-9 if not any(M) else 1e-3
        assert np.abs(sb[2] - want).max() <= tol * scale, np.abs(sb[2] - want).max() / scale
        assert np.abs(sb[[0, 1, 3, 4, 5]]).max() <= tol * scale                  # uniaxial stress state
        assert np.allclose(eb[2], want / 100.0, rtol=0, atol=tol * scale / 100)  # ezz = szz / E
        assert np.allclose(eb[0], -0.2 * eb[2], rtol=0, atol=tol * scale / 100)  # Poisson contraction
        assert np.allclose(sp_[0], sb[2], rtol=0, atol=1e-9 * scale)              # ply 11 = beam zz for theta = 0
    # the GXBeam axis order permutes loads (x along the beam) and outputs: the same physical load gives the same field, xx first
    eb2, sb2, _, _ = S.strain_recovery([1.0, 0, 0], [0, 0.7, 0.4], nodes, conn, mats, th, parts, gxbeam_order=True)
    assert np.allclose(sb2[0], sb[2], rtol=0, atol=1e-12 * scale) and np.allclose(eb2[0], eb[2], rtol=0, atol=1e-14)


def test_tsai_wu_known_values():
    """tsai_wu (section.jl:1108-1131): uniaxial stress at the tensile / compressive strength sits exactly on the failure surface."""
    st = np.array([[1500.0, 1200.0, 50.0, 250.0, 50.0, 250.0, 70.0, 70.0, 40.0]] * 4)
    s = np.zeros((6, 4))
    s[0, 0] = 1500.0; s[0, 1] = -1200.0; s[1, 2] = 50.0; s[3, 3] = 70.0
    f = S.tsai_wu(s, st)
    assert np.allclose(f, 1.0, rtol=1e-13)
