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
