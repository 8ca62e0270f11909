"""Host-side data model: Python mirror of the parts of GXBeam's user API that *feed* the batched hot path.

In a real deployment these objects live in Julia (the reference keeps `Assembly`, `PrescribedConditions`,
`DistributedLoads`, `PointMass` untouched) and only their packed arrays cross the C-ABI.  Julia is
not available in this image, so this module restates the packing side in numpy so that tests and benchmarks can be
written the way the reference's own tests are (same names, same keyword arguments, same quirks).

Reference: src/assembly.jl:14-21 (Element layout), :152-195 (Assembly), :216-343 (discretize_beam, curve_*),
src/loads.jl:16-25, :69-244 (PrescribedConditions), :262-271, :375-459 (DistributedLoads), :497-518 (PointMass),
src/system.jl:190-208 (default_force_scaling), src/GXBeam.jl:72-73 (Gauss nodes).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

ELEM_STRIDE = 91  # L, x[3], compliance[36], mass[36], Cab[9], mu[6]   (assembly.jl:14-21; matrices column-major)
BC_STRIDE = 18    # u, theta, F, M, Ff, Mf                              (loads.jl:16-25)
DLOAD_STRIDE = 24  # f1, f2, m1, m2, f1_follower, f2_follower, m1_follower, m2_follower (loads.jl:262-271)

GAUSS_NODES = np.array([-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526])
GAUSS_WEIGHTS = np.array([0.34785484513745385, 0.6521451548625462, 0.6521451548625462, 0.34785484513745385])


def tilde(x):
    return np.array([[0.0, -x[2], x[1]], [x[2], 0.0, -x[0]], [-x[1], x[0], 0.0]])


def gauss_quadrature(f, a, b):
    """math.jl:338-343"""
    h = b - a
    c = (a + b) / 2
    xs = h / 2 * GAUSS_NODES + c
    return h / 2 * float(np.dot(GAUSS_WEIGHTS, [f(s) for s in xs]))


@dataclass
class Element:
    """assembly.jl:14-21"""
    L: float
    x: np.ndarray
    compliance: np.ndarray
    mass: np.ndarray
    Cab: np.ndarray
    mu: np.ndarray

    def pack(self) -> np.ndarray:
        out = np.empty(ELEM_STRIDE)
        out[0] = self.L
        out[1:4] = self.x
        out[4:40] = np.asarray(self.compliance, dtype=float).reshape(6, 6).ravel(order="F")
        out[40:76] = np.asarray(self.mass, dtype=float).reshape(6, 6).ravel(order="F")
        out[76:85] = np.asarray(self.Cab, dtype=float).reshape(3, 3).ravel(order="F")
        out[85:91] = self.mu
        return out


def _compliance_from_stiffness(stiffness):
    """assembly.jl:84-94: invert only the sub-block of non-zero columns."""
    K = np.asarray(stiffness, dtype=float)
    comp = np.zeros((6, 6))
    filled = [j for j in range(6) if np.any(~np.isclose(K[:, j], 0.0))]
    if filled:
        comp[np.ix_(filled, filled)] = np.linalg.inv(K[np.ix_(filled, filled)])
    return comp


@dataclass
class Assembly:
    """assembly.jl:128-195.  `start`/`stop` are 1-based like the reference."""
    points: np.ndarray
    start: np.ndarray
    stop: np.ndarray
    elements: list

    def __init__(self, points, start, stop, stiffness=None, compliance=None, mass=None, frames=None, damping=None,
                 lengths=None, midpoints=None):
        self.points = np.asarray(points, dtype=float).reshape(-1, 3)
        self.start = np.asarray(list(start), dtype=np.int64)
        self.stop = np.asarray(list(stop), dtype=np.int64)
        ne = len(self.start)
        self.elements = []
        for e in range(ne):
            a, b = self.start[e] - 1, self.stop[e] - 1
            L = float(np.linalg.norm(self.points[b] - self.points[a])) if lengths is None else float(lengths[e])
            xm = (self.points[a] + self.points[b]) / 2 if midpoints is None else np.asarray(midpoints[e], dtype=float)
            if compliance is not None:
                comp = np.asarray(compliance[e], dtype=float)
            elif stiffness is not None:
                comp = _compliance_from_stiffness(stiffness[e])
            else:
                comp = np.zeros((6, 6))
            m = np.zeros((6, 6)) if mass is None else np.asarray(mass[e], dtype=float)
            Cab = np.eye(3) if frames is None else np.asarray(frames[e], dtype=float)
            mu = 0.01 * np.ones(6) if damping is None else np.asarray(damping[e], dtype=float)
            self.elements.append(Element(L, xm, comp, m, Cab, mu))

    @property
    def np_(self):
        return self.points.shape[0]

    @property
    def ne(self):
        return len(self.elements)

    def pack_elements(self) -> np.ndarray:
        return np.stack([e.pack() for e in self.elements])


def default_force_scaling(assembly) -> float:
    """system.jl:190-208: nextpow(2, nnz / sum|S_ij| / 100)."""
    nsum = 0
    csum = 0.0
    eps = np.finfo(float).eps
    for elem in assembly.elements:
        a = np.abs(np.asarray(elem.compliance, dtype=float)).ravel()
        csum += float(a.sum())
        nsum += int((a > eps).sum())
    if nsum == 0:
        return 1.0
    v = nsum / csum / 100
    return float(2.0 ** math.ceil(math.log2(v))) if v > 0 else 1.0


@dataclass
class PrescribedConditions:
    """loads.jl:69-244, including the DOF-2 quirk (:121 tests Fz instead of Fy)."""
    pd: np.ndarray = field(default=None)
    pl: np.ndarray = field(default=None)
    values: np.ndarray = field(default=None)  # 18: u, theta, F, M, Ff, Mf

    def __init__(self, ux=None, uy=None, uz=None, theta_x=None, theta_y=None, theta_z=None, Fx=None, Fy=None, Fz=None,
                 Mx=None, My=None, Mz=None, Fx_follower=None, Fy_follower=None, Fz_follower=None, Mx_follower=None,
                 My_follower=None, Mz_follower=None):
        d = [ux, uy, uz, theta_x, theta_y, theta_z]
        l = [Fx, Fy, Fz, Mx, My, Mz]
        f = [Fx_follower, Fy_follower, Fz_follower, Mx_follower, My_follower, Mz_follower]
        # which load decides "displacement only" for each DOF (loads.jl:121 uses Fz/Fz_follower for DOF 2)
        test = [(Fx, Fx_follower), (Fz, Fz_follower), (Fz, Fz_follower), (Mx, Mx_follower), (My, My_follower),
                (Mz, Mz_follower)]
        pd = np.zeros(6, dtype=np.uint8)
        pl = np.zeros(6, dtype=np.uint8)
        dv = np.zeros(6)
        lv = np.zeros(6)
        fv = np.zeros(6)
        for i in range(6):
            if d[i] is None:
                pd[i], pl[i] = 0, 1
                dv[i] = np.nan
                lv[i] = 0.0 if l[i] is None else l[i]
                fv[i] = 0.0 if f[i] is None else f[i]
            elif test[i][0] is None and test[i][1] is None:
                pd[i], pl[i] = 1, 0
                dv[i] = d[i]
                lv[i] = np.nan
                fv[i] = 0.0
            else:
                pd[i], pl[i] = 1, 1
                dv[i] = d[i]
                lv[i] = 0.0 if l[i] is None else l[i]
                fv[i] = 0.0 if f[i] is None else f[i]
        self.pd, self.pl = pd, pl
        self.values = np.concatenate([dv[:3], dv[3:], lv[:3], lv[3:], fv[:3], fv[3:]])


def pack_prescribed_conditions(npoints, prescribed_conditions):
    """Dict{Int,PrescribedConditions} (1-based keys) -> (pd[np,6], pl[np,6], bc[np,18]).

    A point without an entry behaves exactly like `pd = false, pl = true`, zero loads (point.jl:41-47, :169-175, :208-214).
    """
    pd = np.zeros((npoints, 6), dtype=np.uint8)
    pl = np.ones((npoints, 6), dtype=np.uint8)
    bc = np.zeros((npoints, BC_STRIDE))
    bc[:, 0:6] = np.nan
    for ip, pc in prescribed_conditions.items():
        pd[ip - 1] = pc.pd
        pl[ip - 1] = pc.pl
        bc[ip - 1] = pc.values
    return pd, pl, bc


def distributed_loads(assembly, ielem, s1=0.0, s2=1.0, **fn):
    """loads.jl:375-459 (time-independent form :319-357).  `ielem` is 1-based.  Returns the 24 packed doubles."""
    dL = assembly.elements[ielem - 1].L
    xi = lambda s: (s - s1) / (s2 - s1)
    scaling = dL / (s2 - s1)
    zero = lambda s: 0.0
    out = np.zeros(DLOAD_STRIDE)
    names = [("fx", "fy", "fz"), ("mx", "my", "mz"), ("fx_follower", "fy_follower", "fz_follower"),
             ("mx_follower", "my_follower", "mz_follower")]
    # order in memory: f1, f2, m1, m2, f1_follower, f2_follower, m1_follower, m2_follower
    for grp, base in zip(names, (0, 6, 12, 18)):
        for k, nm in enumerate(grp):
            f = fn.get(nm, zero)
            out[base + k] = gauss_quadrature(lambda s: (1 - xi(s)) * f(s) * scaling, s1, s2)
            out[base + 3 + k] = gauss_quadrature(lambda s: xi(s) * f(s) * scaling, s1, s2)
    return out


def pack_distributed_loads(nelem, dloads):
    """Dict{Int, 24 doubles} -> [ne,24] (zeros where absent) or None."""
    if not dloads:
        return None
    out = np.zeros((nelem, DLOAD_STRIDE))
    for ie, d in dloads.items():
        out[ie - 1] = d
    return out


def point_mass(m, p, I):
    """loads.jl:518: PointMass(m, p, I) = [m*I3 -m*tilde(p); m*tilde(p) I-m*tilde(p)*tilde(p)]"""
    p = np.asarray(p, dtype=float)
    pt = tilde(p)
    M = np.zeros((6, 6))
    M[:3, :3] = m * np.eye(3)
    M[:3, 3:] = -m * pt
    M[3:, :3] = m * pt
    M[3:, 3:] = np.asarray(I, dtype=float) - m * pt @ pt
    return M


def pack_point_masses(npoints, pmasses):
    if not pmasses:
        return None
    out = np.zeros((npoints, 36))
    for ip, M in pmasses.items():
        out[ip - 1] = np.asarray(M, dtype=float).reshape(6, 6).ravel(order="F")
    return out


def discretize_beam(L, start, discretization, frame=None, curvature=None):
    """assembly.jl:216-247, curve_triad :331, curve_coordinates :343."""
    r1 = np.asarray(start, dtype=float)
    Cab = np.eye(3) if frame is None else np.asarray(frame, dtype=float)
    k = np.zeros(3) if curvature is None else np.asarray(curvature, dtype=float)
    disc = np.linspace(0.0, 1.0, discretization + 1) if np.isscalar(discretization) else np.asarray(discretization, float)
    sp = L * disc
    sm = (sp[:-1] + sp[1:]) / 2
    dL = sp[1:] - sp[:-1]
    e1 = np.array([1.0, 0.0, 0.0])
    if not np.any(k):
        triads = [Cab.copy() for _ in sm]
        xp = [r1 + s * Cab @ e1 for s in sp]
        xm = [r1 + s * Cab @ e1 for s in sm]
    else:
        kkt = np.outer(k, k)
        kt = tilde(k)
        kn = math.sqrt(k @ k)
        I = np.eye(3)
        triad = lambda s: Cab @ ((I - kkt / kn ** 2) * math.cos(kn * s) + kt / kn * math.sin(kn * s) + kkt / kn ** 2)
        coord = lambda s: r1 + Cab @ ((I / kn - kkt / kn ** 3) * math.sin(kn * s) + kt / kn ** 2 * (1 - math.cos(kn * s))
                                      + kkt / kn ** 2 * s) @ e1
        triads = [triad(s) for s in sm]
        xp = [coord(s) for s in sp]
        xm = [coord(s) for s in sm]
    return dL, np.array(xp), np.array(xm), triads
