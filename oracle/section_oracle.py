"""CPU oracle for the cross-section analysis of src/section.jl (BASELINE config 5).  TEST INFRASTRUCTURE ONLY: imported
by tests/, never by the product package.

numpy restatement, function by function, of /root/reference/src/section.jl (citations are relative to that file):
    stiffness :216-249, get_Ttheta :260-278, get_Tbeta :293-309, elementQ :317-322, element_orientation :331-345,
    element_matrices :347-462, element_integrand :464-484, element_submatrix :486-501 (2 x 2 Gauss), node2idx :508-538,
    strain_recovery :952-1040, tsai_wu :1108-1131,
    compliance_matrix :641-783 (scatter, KKT system [E R D; R' A 0; D' 0 0], two solves with 6 right-hand sides, the nine products
    that form S, shear / tension centre, optional transformation to the shear centre, reorder :540-551), mass_matrix :840-890.
The reference factorises the sparse KKT matrix with UMFPACK (`factorize`, :553); here scipy.sparse.linalg.splu (SuperLU) -- any
backward-stable LU agrees to ~cond * eps.  Parity is PINNED on the BECAS user-guide tables the reference's own tests hold
(test/section.jl:40-340: square S1-S3 with three ply angles, cylinder C1, half cylinder C2 with shear / tension centre, circular tube).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

GP = 1.0 / np.sqrt(3.0)


def stiffness(mat):
    """section.jl:216-249.  mat = (E1, E2, E3, G12, G13, G23, nu12, nu13, nu23, rho)."""
    E1, E2, E3, G12, G13, G23, nu12, nu13, nu23 = mat[:9]
    nu21 = nu12 * E2 / E1
    nu31 = nu13 * E3 / E1
    nu32 = nu23 * E3 / E2
    delta = 1.0 / (1 - nu12 * nu21 - nu23 * nu32 - nu13 * nu31 - 2 * nu21 * nu32 * nu13)
    Q66 = E1 * (1 - nu23 * nu32) * delta
    Q11 = E2 * (1 - nu13 * nu31) * delta
    Q22 = E3 * (1 - nu12 * nu21) * delta
    Q16 = E1 * (nu21 + nu31 * nu23) * delta
    Q12 = E2 * (nu32 + nu12 * nu31) * delta
    Q26 = E1 * (nu31 + nu21 * nu32) * delta
    Q = np.zeros((6, 6))
    Q[0, 0], Q[0, 1], Q[0, 5] = Q11, Q12, Q16
    Q[1, 0], Q[1, 1], Q[1, 5] = Q12, Q22, Q26
    Q[2, 2], Q[3, 3], Q[4, 4] = G23, G12, G13
    Q[5, 0], Q[5, 1], Q[5, 5] = Q16, Q26, Q66
    return Q


def get_Ttheta(theta):
    """section.jl:260-278"""
    s, c = np.sin(theta), np.cos(theta)
    c2, s2, sc = c * c, s * s, s * c
    return np.array([[c2, 0, 0, 2 * sc, 0, s2], [0, 1, 0, 0, 0, 0], [0, 0, c, 0, s, 0], [-sc, 0, 0, c2 - s2, 0, sc], [0, 0, -s, 0, c, 0],
                     [s2, 0, 0, -2 * sc, 0, c2]])


def get_Tbeta(c, s):
    """section.jl:293-309"""
    c2, s2, sc = c * c, s * s, s * c
    return np.array([[c2, s2, -2 * sc, 0, 0, 0], [s2, c2, 2 * sc, 0, 0, 0], [sc, -sc, c2 - s2, 0, 0, 0], [0, 0, 0, c, -s, 0], [0, 0, 0, s, c, 0],
                     [0, 0, 0, 0, 0, 1.0]])


def elementQ(mat, theta, cbeta, sbeta):
    """section.jl:317-322"""
    Q = stiffness(mat)
    T = get_Ttheta(theta)
    Q = T @ Q @ T.T
    T = get_Tbeta(cbeta, sbeta)
    return T @ Q @ T.T


def element_orientation(xy):
    """section.jl:331-345; xy: [4, 2] node coordinates of the element"""
    xl, yl = 0.5 * (xy[0, 0] + xy[3, 0]), 0.5 * (xy[0, 1] + xy[3, 1])
    xr, yr = 0.5 * (xy[1, 0] + xy[2, 0]), 0.5 * (xy[1, 1] + xy[2, 1])
    dx, dy = xr - xl, yr - yl
    ds = np.sqrt(dx * dx + dy * dy)
    return dx / ds, dy / ds


def element_matrices(ksi, eta, mat, theta, xy):
    """section.jl:347-462: Q, J, SZ, SN, BN at one point of the element."""
    mk, pk, me, pe = 1 - ksi, 1 + ksi, 1 - eta, 1 + eta
    N = np.array([mk * me, pk * me, pk * pe, mk * pe]) / 4
    x, y = N @ xy[:, 0], N @ xy[:, 1]
    cb, sb = element_orientation(xy)
    Q = elementQ(mat, theta, cb, sb)
    SZ = np.zeros((6, 6))
    SZ[3, 0] = 1; SZ[3, 5] = -y
    SZ[4, 1] = 1; SZ[4, 5] = x
    SZ[5, 2] = 1; SZ[5, 3] = y; SZ[5, 4] = -x
    SN = np.zeros((6, 12))
    for i in range(4):
        SN[3, 3 * i] = SN[4, 3 * i + 1] = SN[5, 3 * i + 2] = N[i]
    N_ksi = np.array([-me, me, pe, -pe]) / 4
    N_eta = np.array([-mk, -pk, pk, mk]) / 4
    J = np.array([[N_ksi @ xy[:, 0], N_ksi @ xy[:, 1]], [N_eta @ xy[:, 0], N_eta @ xy[:, 1]]])
    Ji = np.linalg.inv(J)
    Bksi = np.zeros((6, 3)); Beta = np.zeros((6, 3))
    Bksi[0, 0] = Ji[0, 0]; Bksi[1, 1] = Ji[1, 0]; Bksi[2, 0] = Ji[1, 0]; Bksi[2, 1] = Ji[0, 0]; Bksi[3, 2] = Ji[0, 0]; Bksi[4, 2] = Ji[1, 0]
    Beta[0, 0] = Ji[0, 1]; Beta[1, 1] = Ji[1, 1]; Beta[2, 0] = Ji[1, 1]; Beta[2, 1] = Ji[0, 1]; Beta[3, 2] = Ji[0, 1]; Beta[4, 2] = Ji[1, 1]
    NMk = np.zeros((3, 12)); NMe = np.zeros((3, 12))
    for i in range(4):
        for d in range(3):
            NMk[d, 3 * i + d] = N_ksi[i]; NMe[d, 3 * i + d] = N_eta[i]
    BN = Bksi @ NMk + Beta @ NMe
    return Q, J, SZ, SN, BN


def element_integrand(ksi, eta, mat, theta, xy):
    """section.jl:464-484: (Ae 6x6, Re 12x6, Ee 12x12, Ce 12x12, Le 12x6, Me 12x12) at one Gauss point."""
    Q, J, SZ, SN, BN = element_matrices(ksi, eta, mat, theta, xy)
    dJ = np.linalg.det(J)
    return (SZ.T @ Q @ SZ * dJ, BN.T @ Q @ SZ * dJ, BN.T @ Q @ BN * dJ, BN.T @ Q @ SN * dJ, SN.T @ Q @ SZ * dJ, SN.T @ Q @ SN * dJ)


def element_submatrix(mat, theta, xy):
    """section.jl:486-501"""
    acc = None
    for ksi, eta in ((GP, GP), (GP, -GP), (-GP, GP), (-GP, -GP)):
        t = element_integrand(ksi, eta, mat, theta, xy)
        acc = t if acc is None else tuple(a + b for a, b in zip(acc, t))
    return acc


def assemble(nodes, conn, mats, thetas):
    """The scatter loop of compliance_matrix (section.jl:661-677).  nodes [nn, 2]; conn [ne, 4] 0-based; mats [ne, 10]; thetas [ne]."""
    nn = len(nodes); ndof = 3 * nn
    A = np.zeros((6, 6)); R = np.zeros((ndof, 6)); L = np.zeros((ndof, 6))
    E = np.zeros((ndof, ndof)); Cm = np.zeros((ndof, ndof)); M = np.zeros((ndof, ndof))
    for e in range(len(conn)):
        idx = (3 * conn[e][:, None] + np.arange(3)[None, :]).ravel()
        Ae, Re, Ee, Ce, Le, Me = element_submatrix(mats[e], thetas[e], nodes[conn[e]])
        A += Ae; R[idx] += Re; L[idx] += Le
        E[np.ix_(idx, idx)] += Ee; Cm[np.ix_(idx, idx)] += Ce; M[np.ix_(idx, idx)] += Me
    return A, R, E, Cm, L, M


def reorder(K):
    """section.jl:540-551"""
    idx = [2, 0, 1, 5, 3, 4]
    return K[np.ix_(idx, idx)]


def compliance_matrix(nodes, conn, mats, thetas, gxbeam_order=True, shear_center=True, return_parts=False):
    """section.jl:641-783.  Returns S (6x6), sc (2), tc (2)."""
    nodes = np.asarray(nodes, float); conn = np.asarray(conn, int); mats = np.asarray(mats, float); thetas = np.asarray(thetas, float)
    nn = len(nodes); ndof = 3 * nn
    A, R, E, Cm, L, M = assemble(nodes, conn, mats, thetas)
    n = ndof + 12
    Asys = np.zeros((n, n))
    Asys[:ndof, :ndof] = E
    Asys[:ndof, ndof:ndof + 6] = R; Asys[ndof:ndof + 6, :ndof] = R.T
    Asys[ndof:ndof + 6, ndof:ndof + 6] = A
    for i in range(nn):
        s = 3 * i
        for (r, c, v) in ((ndof + 6, s, 1.0), (ndof + 7, s + 1, 1.0), (ndof + 8, s + 2, 1.0), (ndof + 9, s + 2, nodes[i, 1]), (ndof + 10, s + 2, -nodes[i, 0]),
                          (ndof + 11, s, -nodes[i, 1]), (ndof + 11, s + 1, nodes[i, 0])):
            Asys[r, c] = Asys[c, r] = v
    lu = spla.splu(sp.csc_matrix(Asys))
    B2 = np.zeros((n, 6)); B2[ndof + 4, 0] = -1.0; B2[ndof + 3, 1] = 1.0
    X2 = lu.solve(B2)
    dX, dY = X2[:ndof], X2[ndof:ndof + 6]
    B1 = np.zeros((n, 6))
    B1[:ndof] = Cm.T @ dX - Cm @ dX + L @ dY
    B1[ndof:ndof + 6] = np.eye(6) - L.T @ dX
    X1 = lu.solve(B1)
    X, Y = X1[:ndof], X1[ndof:ndof + 6]
    S = (X.T @ (E @ X) + X.T @ (Cm @ dX) + (X.T @ R) @ Y + dX.T @ (Cm.T @ X) + dX.T @ (M @ dX) + (dX.T @ L) @ Y + Y.T @ (R.T @ X)
         + Y.T @ (L.T @ dX) + Y.T @ (A @ Y))
    xs = -S[5, 1] / S[5, 5]; ys = S[5, 0] / S[5, 5]
    den = S[3, 3] * S[4, 4] - S[3, 4] ** 2
    xt = (S[3, 3] * S[4, 2] - S[3, 4] * S[3, 2]) / den
    yt = (-S[3, 2] * S[4, 4] + S[3, 4] * S[4, 2]) / den
    S_raw = S.copy()
    if shear_center:
        P = np.array([[0, 0, ys], [0, 0, -xs], [-ys, xs, 0]])
        Hinv = np.block([[np.eye(3), P.T], [np.zeros((3, 3)), np.eye(3)]])
        HinvT = np.block([[np.eye(3), np.zeros((3, 3))], [P, np.eye(3)]])
        S = Hinv @ S @ HinvT
    if gxbeam_order:
        S = reorder(S)
    if return_parts:
        return S, np.array([xs, ys]), np.array([xt, yt]), dict(A=A, R=R, E=E, C=Cm, L=L, M=M, X=X, Y=Y, dX=dX, dY=dY, S_raw=S_raw)
    return S, np.array([xs, ys]), np.array([xt, yt])


def strain_recovery(F, M, nodes, conn, mats, thetas, parts, gxbeam_order=True):
    """section.jl:952-1040.  parts: the X, dX, Y of compliance_matrix(..., return_parts=True) (the reference's SectionCache).
    Returns strain_b, stress_b, strain_p, stress_p, each [6, ne]."""
    nodes = np.asarray(nodes, float); conn = np.asarray(conn, int); mats = np.asarray(mats, float); thetas = np.asarray(thetas, float)
    F = np.asarray(F, float); M = np.asarray(M, float)
    if gxbeam_order:
        new = [1, 2, 0]
        load = np.concatenate([F[new], M[new]])
        idx_b = [5, 0, 1, 4, 2, 3]
    else:
        load = np.concatenate([F, M])
        idx_b = [0, 1, 5, 2, 3, 4]
    idx_p = [5, 0, 1, 3, 4, 2]
    ne = len(conn)
    eb = np.zeros((6, ne)); sb = np.zeros((6, ne)); ep = np.zeros((6, ne)); sp_ = np.zeros((6, ne))
    X, dX, Y = parts["X"], parts["dX"], parts["Y"]
    for e in range(ne):
        xy = nodes[conn[e]]
        Q, _, SZ, SN, BN = element_matrices(0.0, 0.0, mats[e], thetas[e], xy)
        idx = (3 * conn[e][:, None] + np.arange(3)[None, :]).ravel()
        eps = SZ @ (Y @ load) + BN @ (X[idx] @ load) + SN @ (dX[idx] @ load)
        sig = Q @ eps
        cb, sbeta = element_orientation(xy)
        epl = get_Ttheta(thetas[e]).T @ (get_Tbeta(cb, sbeta).T @ eps)
        spl = stiffness(mats[e]) @ epl
        eb[:, e] = eps[idx_b]; sb[:, e] = sig[idx_b]; ep[:, e] = epl[idx_p]; sp_[:, e] = spl[idx_p]
    return eb, sb, ep, sp_


def tsai_wu(stress_p, strengths):
    """section.jl:1108-1131.  stress_p [6, ne]; strengths [ne, 9] = (S1t, S1c, S2t, S2c, S3t, S3c, S12, S13, S23) of each element's material."""
    s = np.asarray(stress_p, float); m = np.asarray(strengths, float)
    S1t, S1c, S2t, S2c, S3t, S3c, S12, S13, S23 = m.T
    return (s[0] ** 2 / (S1t * S1c) + s[1] ** 2 / (S2t * S2c) + s[2] ** 2 / (S3t * S3c) + s[3] ** 2 / S12 ** 2 + s[4] ** 2 / S13 ** 2
            + s[5] ** 2 / S23 ** 2 + s[0] * (1 / S1t - 1 / S1c) + s[1] * (1 / S2t - 1 / S2c) + s[2] * (1 / S3t - 1 / S3c)
            - s[0] * s[1] / np.sqrt(S1t * S1c * S2t * S2c) - s[0] * s[2] / np.sqrt(S1t * S1c * S3t * S3c)
            - s[1] * s[2] / np.sqrt(S2t * S2c * S3t * S3c))


def area_and_centroid(xy):
    """section.jl:790-806"""
    A = 0.5 * sum(xy[i, 0] * xy[(i + 1) % 4, 1] - xy[(i + 1) % 4, 0] * xy[i, 1] for i in range(4))
    return A, xy[:, 0].mean(), xy[:, 1].mean()


def mass_matrix(nodes, conn, mats):
    """section.jl:840-890 (GXBeam ordering).  Returns M (6x6), mass centre (2)."""
    nodes = np.asarray(nodes, float)
    m = xm = ym = 0.0
    geo = []
    for e in range(len(conn)):
        A, xc, yc = area_and_centroid(nodes[conn[e]])
        dm = mats[e][9] * A
        geo.append((dm, xc, yc))
        m += dm; xm += xc * dm; ym += yc * dm
    xm /= m; ym /= m
    Ixx = Iyy = Ixy = 0.0
    for dm, xc, yc in geo:
        Ixx += (yc - ym) ** 2 * dm; Iyy += (xc - xm) ** 2 * dm; Ixy += (xc - xm) * (yc - ym) * dm
    M = np.array([[m, 0, 0, 0, m * ym, -m * xm], [0, m, 0, -m * ym, 0, 0], [0, 0, m, m * xm, 0, 0], [0, -m * ym, m * xm, Ixx + Iyy, 0, 0],
                  [m * ym, 0, 0, 0, Ixx, -Ixy], [-m * xm, 0, 0, 0, -Ixy, Iyy]])
    return M, np.array([xm, ym])


# ---- meshes of the reference's own tests (test/section.jl:40-340), shared by the oracle tests and the GPU parity tests ----
def square_mesh(n=10, half=0.05):
    """test/section.jl:45-66: (n+1)^2 nodes on [-half, half]^2, n^2 elements, node m = (n+1) i + j."""
    x = np.linspace(-half, half, n + 1)
    nodes = np.array([[x[i], x[j]] for i in range(n + 1) for j in range(n + 1)])
    conn = np.array([[(n + 1) * i + j, (n + 1) * (i + 1) + j, (n + 1) * (i + 1) + j + 1, (n + 1) * i + j + 1] for i in range(n) for j in range(n)])
    return nodes, conn


def ring_mesh(R, t, nr, nt, th0=0.0, th1=2 * np.pi, closed=True):
    """test/section.jl:186-226 (closed cylinder, nt-1 sectors) and :239-270 (open half cylinder, nt rows of nodes)."""
    r = np.linspace(R - t, R, nr)
    th = np.linspace(th0, th1, nt)
    rows = nt - 1 if closed else nt
    nodes = np.array([[r[j] * np.cos(th[i]), r[j] * np.sin(th[i])] for i in range(rows) for j in range(nr)])
    conn = []
    for i in range(nt - 1):
        ip = (0 if i == nt - 2 else i + 1) if closed else i + 1
        for j in range(nr - 1):
            conn.append([nr * ip + j, nr * i + j, nr * i + j + 1, nr * ip + j + 1])
    return nodes, np.array(conn)
