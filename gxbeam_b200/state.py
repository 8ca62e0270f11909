"""Post-processing of a solved state vector into the reference's observables (host side, numpy).

Reference: src/output.jl:146-203 (extract_point_state, StaticSystem), :317-352 (extract_element_state), math.jl:52-67.
In a Julia deployment the *existing* `AssemblyState(x, system, assembly; prescribed_conditions)` (output.jl:106) is called
on each returned state vector; this module restates it so that parity can be checked on the same observables.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def rotation_parameter_scaling(theta):
    """math.jl:52-67"""
    n = float(np.sqrt(np.dot(theta, theta)))
    m = np.trunc((np.trunc(n / 4.0) + 1.0) / 2.0)
    return 1.0 - 8.0 * m / n if n > 4.0 else 1.0


def get_C(theta):
    """math.jl:24-36, 94-101"""
    c = rotation_parameter_scaling(theta) * np.asarray(theta, float)
    c0 = 2 - c @ c / 8
    C = np.array([[c0 ** 2 + c[0] ** 2 - c[1] ** 2 - c[2] ** 2, 2 * (c[0] * c[1] + c0 * c[2]), 2 * (c[0] * c[2] - c0 * c[1])],
                  [2 * (c[0] * c[1] - c0 * c[2]), c0 ** 2 - c[0] ** 2 + c[1] ** 2 - c[2] ** 2, 2 * (c[1] * c[2] + c0 * c[0])],
                  [2 * (c[0] * c[2] + c0 * c[1]), 2 * (c[1] * c[2] - c0 * c[0]), c0 ** 2 - c[0] ** 2 - c[1] ** 2 + c[2] ** 2]])
    return C / (4 - c0) ** 2


@dataclass
class AssemblyState:
    """Arrays over points / elements of the fields of PointState / ElementState that a static analysis defines."""
    point_u: np.ndarray      # [np,3]
    point_theta: np.ndarray  # [np,3]  Wiener-Milenkovic parameters (θ·scaling, output.jl:166-167)
    point_F: np.ndarray      # [np,3]  externally applied load / reaction
    point_M: np.ndarray
    elem_u: np.ndarray       # [ne,3]
    elem_theta: np.ndarray
    elem_Fi: np.ndarray      # [ne,3]  internal resultants in the deformed element frame
    elem_Mi: np.ndarray


def assembly_state(x, start, stop, icol_point, icol_elem, pd, pl, bc, force_scaling) -> AssemblyState:
    """x: nstates (static ordering).  icol_*: 1-based (SystemIndices).  pd/pl: [np,6], bc: [np,18]."""
    x = np.asarray(x, float)
    npnt, ne = len(icol_point), len(icol_elem)
    pu = np.zeros((npnt, 3)); pt = np.zeros((npnt, 3)); pF = np.zeros((npnt, 3)); pM = np.zeros((npnt, 3))
    raw_t = np.zeros((npnt, 3))
    for ip in range(npnt):
        ic = icol_point[ip] - 1
        u = np.where(pd[ip, :3] != 0, bc[ip, 0:3], x[ic:ic + 3])
        th = np.where(pd[ip, 3:] != 0, bc[ip, 3:6], x[ic + 3:ic + 6])
        C = get_C(th)
        with np.errstate(invalid="ignore"):
            F = bc[ip, 6:9] + C.T @ bc[ip, 12:15]
            M = bc[ip, 9:12] + C.T @ bc[ip, 15:18]
        pF[ip] = np.where(pl[ip, :3] != 0, F, x[ic:ic + 3] * force_scaling)
        pM[ip] = np.where(pl[ip, 3:] != 0, M, x[ic + 3:ic + 6] * force_scaling)
        pu[ip] = u
        raw_t[ip] = th
        pt[ip] = th * rotation_parameter_scaling(th)
    eu = np.zeros((ne, 3)); et = np.zeros((ne, 3)); eF = np.zeros((ne, 3)); eM = np.zeros((ne, 3))
    for ie in range(ne):
        a, b = start[ie] - 1, stop[ie] - 1
        eu[ie] = (pu[a] + pu[b]) / 2
        th = (raw_t[a] + raw_t[b]) / 2
        et[ie] = th * rotation_parameter_scaling(th)
        ic = icol_elem[ie] - 1
        eF[ie] = x[ic:ic + 3] * force_scaling
        eM[ie] = x[ic + 3:ic + 6] * force_scaling
    return AssemblyState(pu, pt, pF, pM, eu, et, eF, eM)


def assembly_state_records(x, kind, pd, pl, bc, force_scaling, rates=None):
    """The 30-double PointState / ElementState records of a chain (src/output.jl:18-28, 54-64, 146-203, 317-403), host-side numpy
    restatement used by the tests of gxb_assembly_state.  x: nstates; pd/pl: [np, 6]; bc: [np, 18]; rates: [np, 12] or None."""
    x = np.asarray(x, float)
    stride = 12 if kind == 0 else 18
    npnt = pd.shape[0]; ne = npnt - 1
    pts = np.zeros((npnt, 30)); els = np.zeros((ne, 30))
    raw = []
    for ip in range(npnt):
        xp = x[stride * ip:stride * ip + stride]
        u = np.where(pd[ip, :3] != 0, bc[ip, 0:3], xp[0:3]); th = np.where(pd[ip, 3:] != 0, bc[ip, 3:6], xp[3:6])
        C = get_C(th)
        with np.errstate(invalid="ignore"):
            F = bc[ip, 6:9] + C.T @ bc[ip, 12:15]; M = bc[ip, 9:12] + C.T @ bc[ip, 15:18]
        F = np.where(pl[ip, :3] != 0, F, xp[0:3] * force_scaling); M = np.where(pl[ip, 3:] != 0, M, xp[3:6] * force_scaling)
        V = xp[6:9] if kind == 1 else np.zeros(3); Om = xp[9:12] if kind == 1 else np.zeros(3)
        d = np.full(12, 0.0 if kind == 0 else np.nan) if rates is None else np.array(rates[ip], float)
        d[:6] = np.where(pd[ip] != 0, 0.0, d[:6])
        raw.append((u, d[0:3], th, d[3:6], V, d[6:9], Om, d[9:12]))
        pts[ip] = np.concatenate([u, d[0:3], th * rotation_parameter_scaling(th), d[3:6], V, d[6:9], Om, d[9:12], F, M])
    for ie in range(ne):
        a, b = raw[ie], raw[ie + 1]
        avg = [(p + q) / 2 for p, q in zip(a, b)]
        xe = x[stride * ie + stride - 6:stride * ie + stride]
        avg[2] = avg[2] * rotation_parameter_scaling(avg[2])
        els[ie] = np.concatenate(avg + [xe[0:3] * force_scaling, xe[3:6] * force_scaling])
    return pts, els
