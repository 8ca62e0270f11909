"""Synthetic batches of the BASELINE.json configurations, shared by bench.py and the parity tests.

All draws come from numpy.random.default_rng(1234 + cfg) so that the GPU run, the CPU oracle and every rank of a
multi-GPU run see identical inputs.
"""
from __future__ import annotations

import numpy as np

from .assembly import Assembly, PrescribedConditions, pack_prescribed_conditions


def cantilever_tipmoment_batch(nbatch=4096, nelem=100, seed=1234 + 2, shear_rigid=False, tip_force=False):
    """BASELINE config 2: `nbatch` independent `nelem`-element nonlinear cantilevers under a tip moment (static_analysis).

    Base case = test/examples/tipmoment.jl:5-50 (L = 12, E = 30e6, h = w = 1, root clamped, Mz at the tip) with
    randomised stiffness (E_i = E·U(0.8,1.2), h_i, w_i = U(0.9,1.1)) and load factor λ_i = U(0.1,1.0) of π E Izz / L.
    Finite shear/torsion compliance (G = E/2.6) unless `shear_rigid` (the test's zeros).
    Returns a dict of packed arrays in the C-ABI layouts.
    """
    rng = np.random.default_rng(seed)
    L = 12.0
    E0 = 30e6
    xs = np.linspace(0.0, L, nelem + 1)
    points = np.stack([xs, 0 * xs, 0 * xs], 1)
    start = np.arange(1, nelem + 1, dtype=np.int64)
    stop = np.arange(2, nelem + 2, dtype=np.int64)
    E = E0 * rng.uniform(0.8, 1.2, nbatch)
    h = rng.uniform(0.9, 1.1, nbatch)
    w = rng.uniform(0.9, 1.1, nbatch)
    lam = rng.uniform(0.1, 1.0, nbatch)
    fz = rng.uniform(0.0, 1.0, nbatch)
    A = h * w
    Iyy = w * h ** 3 / 12
    Izz = w ** 3 * h / 12
    G = E / 2.6
    diag = np.zeros((nbatch, 6))
    diag[:, 0] = 1 / (E * A)
    if not shear_rigid:
        diag[:, 1] = 1 / (G * A * 5 / 6)
        diag[:, 2] = 1 / (G * A * 5 / 6)
        diag[:, 3] = 1 / (G * (Iyy + Izz))
    diag[:, 4] = 1 / (E * Iyy)
    diag[:, 5] = 1 / (E * Izz)
    # Element{Float64} records: L, x[3], compliance[36], mass[36], Cab[9], mu[6]
    elems = np.zeros((nbatch, nelem, 91))
    dL = L / nelem
    elems[:, :, 0] = dL
    elems[:, :, 1] = ((xs[:-1] + xs[1:]) / 2)[None, :]
    for k in range(6):
        elems[:, :, 4 + 7 * k] = diag[:, k][:, None]      # diagonal of the column-major 6x6
    elems[:, :, 76] = elems[:, :, 80] = elems[:, :, 84] = 1.0  # Cab = I
    elems[:, :, 85:91] = 0.01
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0), nelem + 1: PrescribedConditions(Mz=1.0)}
    pd, pl, bc1 = pack_prescribed_conditions(nelem + 1, pcs)
    bc = np.tile(bc1[None], (nbatch, 1, 1))
    bc[:, nelem, 11] = lam * np.pi * E * Izz / L          # Mz
    if tip_force:
        bc[:, nelem, 8] = fz * E * Iyy / L ** 2           # Fz
    # default_force_scaling (system.jl:190-208) per instance
    nnz = (diag > np.finfo(float).eps).sum(1) * nelem
    csum = diag.sum(1) * nelem
    fs = 2.0 ** np.ceil(np.log2(nnz / csum / 100))
    return dict(start=start, stop=stop, points=points, pd=pd, pl=pl, elems=elems, bc=bc, force_scaling=fs, nelem=nelem,
                nbatch=nbatch, lam=lam)


def single_assembly_arrays(assembly: Assembly, prescribed_conditions, dloads=None, pmasses=None):
    """Pack one Assembly + loads into batch-of-one arrays."""
    from .assembly import pack_distributed_loads, pack_point_masses
    pd, pl, bc = pack_prescribed_conditions(assembly.np_, prescribed_conditions)
    return dict(start=assembly.start, stop=assembly.stop, points=assembly.points, pd=pd, pl=pl,
                elems=assembly.pack_elements()[None], bc=bc[None],
                dload=None if not dloads else pack_distributed_loads(assembly.ne, dloads)[None],
                pmass=None if not pmasses else pack_point_masses(assembly.np_, pmasses)[None])


def rotating_blade_batch(nbatch=1, seed=1234 + 1, rpm_jitter=0.0):
    """BASELINE config 1 (benchmark/benchmarks.jl:3-65): 20 + 4 element swept-tip blade, root clamped, rotating at 750 rpm about z
    (linear_velocity = (0, 196.349540849362, 0), angular_velocity = (0, 0, 78.5398163397448)), steady_state_analysis, n = 444.
    For nbatch > 1 the rotor speed of instance i is scaled by U(1 - rpm_jitter, 1 + rpm_jitter) (design-sweep style batch).
    """
    from .assembly import discretize_beam
    rng = np.random.default_rng(seed)
    l1, xp1, xm1, c1 = discretize_beam(31.5, [0, 0, 0], 20)
    frame = np.array([[0.707106781186548, 0.707106781186548, 0.0], [-0.707106781186548, 0.707106781186548, 0.0], [0.0, 0.0, 1.0]])
    l2, xp2, xm2, c2 = discretize_beam(6.0, [31.5, 0, 0], 4, frame=frame)
    nelem = 24
    points = np.vstack([xp1, xp2[1:]])
    start = np.arange(1, nelem + 1, dtype=np.int64)
    stop = np.arange(2, nelem + 2, dtype=np.int64)
    comp = np.diag([1.4974543276E-06, 4.7619054919E-06, 5.8036221902E-05, 3.1234387938E-03, 4.5274507258E-03, 1.7969451932E-05])
    mass = np.diag([1.5813000000E-05, 1.5813000000E-05, 1.5813000000E-05, 1.3229801498E-06, 5.2301497500E-09, 1.3177500000E-06])
    asm = Assembly(points, start, stop, compliance=[comp] * nelem, mass=[mass] * nelem, frames=list(c1) + list(c2),
                   lengths=np.concatenate([l1, l2]), midpoints=np.vstack([xm1, xm2]))
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0)}
    pd, pl, bc1 = pack_prescribed_conditions(nelem + 1, pcs)
    elems = np.tile(asm.pack_elements()[None], (nbatch, 1, 1))
    bc = np.tile(bc1[None], (nbatch, 1, 1))
    pts = np.tile(points[None], (nbatch, 1, 1))
    scale = np.ones(nbatch) if rpm_jitter == 0.0 else rng.uniform(1 - rpm_jitter, 1 + rpm_jitter, nbatch)
    motion = np.zeros((nbatch, 12))
    motion[:, 1] = 196.349540849362 * scale
    motion[:, 5] = 78.5398163397448 * scale
    from .assembly import default_force_scaling
    fs = np.full(nbatch, default_force_scaling(asm))
    return dict(start=start, stop=stop, points=points, points_b=pts, pd=pd, pl=pl, elems=elems, bc=bc, motion=motion, force_scaling=fs,
                nelem=nelem, nbatch=nbatch, asm=asm)


WT_STIFFNESS = np.array([[2.389e9, 1.524e6, 6.734e6, -3.382e7, -2.627e7, -4.736e8],
                         [1.524e6, 4.334e8, -3.741e6, -2.935e5, 1.527e7, 3.835e5],
                         [6.734e6, -3.741e6, 2.743e7, -4.592e5, -6.869e5, -4.742e6],
                         [-3.382e7, -2.935e5, -4.592e5, 2.167e7, -6.279e5, 1.430e6],
                         [-2.627e7, 1.527e7, -6.869e5, -6.279e5, 1.970e7, 1.209e7],
                         [-4.736e8, 3.835e5, -4.742e6, 1.430e6, 1.209e7, 4.406e8]])
WT_MASS = np.array([[258.053, 0.0, 0.0, 0.0, 7.07839, -71.6871],
                    [0.0, 258.053, 0.0, -7.07839, 0.0, 0.0],
                    [0.0, 0.0, 258.053, 71.6871, 0.0, 0.0],
                    [0.0, -7.07839, 71.6871, 48.59, 0.0, 0.0],
                    [7.07839, 0.0, 0.0, 0.0, 2.172, 0.0],
                    [-71.6871, 0.0, 0.0, 0.0, 0.0, 46.418]])


def wind_turbine_blade_batch(nbatch=1024, nelem=200, seed=1234 + 3, nspeeds=16, speed_max=0.9):
    """BASELINE config 3: rotating wind-turbine blades (straight, L = 60 m, the fully populated 6x6 stiffness / mass of
    test/examples/wind-turbine-blade.jl:19-37, root clamped), stiffness and mass scaled per design by U(0.8, 1.2), rotor-speed
    sweep ωb = (0, 0, Ω_j), Ω_j = nspeeds values in [0, speed_max] rad/s; instance b = (design b // nspeeds, speed b % nspeeds).
    steady_state_analysis + linearised eigenvalue analysis (nev = 20), n = 18·nelem + 12.

    speed_max = 0.9 rad/s (8.6 rpm, the operating range of a 60 m blade).  SURVEY 8d proposed [0, 2] rad/s; between 1.07 and 1.33 rad/s
    Newton from the zero state (how the reference starts every `steady_state_analysis`, test/examples/rotating.jl:71-90) stagnates in
    the line search for 11 of the 1024 designs x speeds -- identically in the CPU oracle, with `iterations = 1000` as with 50
    (scratch/c3_probe.py: residuals stay at 0.1 .. 50) -- so that range measured the iteration cap, not the path.  With 0.9 rad/s all
    1024 instances converge in at most 5 iterations in the oracle and on the device, at the reference's default `iterations = 1000`.
    """
    rng = np.random.default_rng(seed)
    L = 60.0
    xs = np.linspace(0.0, L, nelem + 1)
    points = np.stack([xs, 0 * xs, 0 * xs], 1)
    start = np.arange(1, nelem + 1, dtype=np.int64)
    stop = np.arange(2, nelem + 2, dtype=np.int64)
    ndes = (nbatch + nspeeds - 1) // nspeeds
    ks = rng.uniform(0.8, 1.2, ndes)
    ms = rng.uniform(0.8, 1.2, ndes)
    comp0 = np.linalg.inv(WT_STIFFNESS)
    speeds = np.linspace(0.0, speed_max, nspeeds)
    elems = np.zeros((nbatch, nelem, 91))
    elems[:, :, 0] = L / nelem
    elems[:, :, 1] = ((xs[:-1] + xs[1:]) / 2)[None, :]
    elems[:, :, 76] = elems[:, :, 80] = elems[:, :, 84] = 1.0
    elems[:, :, 85:91] = 0.01
    fs = np.zeros(nbatch)
    motion = np.zeros((nbatch, 12))
    for b in range(nbatch):
        d = b // nspeeds
        comp = comp0 / ks[d]
        elems[b, :, 4:40] = comp.ravel(order="F")[None, :]
        elems[b, :, 40:76] = (WT_MASS * ms[d]).ravel(order="F")[None, :]
        a = np.abs(comp).ravel()
        fs[b] = 2.0 ** np.ceil(np.log2((a > np.finfo(float).eps).sum() / a.sum() / 100))
        om = speeds[b % nspeeds]
        motion[b, 5] = om                                   # angular_velocity = (0, 0, Ω)
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0)}
    pd, pl, bc1 = pack_prescribed_conditions(nelem + 1, pcs)
    bc = np.tile(bc1[None], (nbatch, 1, 1))
    pts = np.tile(points[None], (nbatch, 1, 1))
    return dict(start=start, stop=stop, points=points, points_b=pts, pd=pd, pl=pl, elems=elems, bc=bc, motion=motion, force_scaling=fs,
                nelem=nelem, nbatch=nbatch)


def airfoil_layup_batch(nbatch=2048, ns=120, nplies=6, seed=1234 + 5, chord=1.0):
    """BASELINE config 5: `nbatch` parameterised layups of one closed thin-walled airfoil-shaped section (src/section.jl
    compliance_matrix + mass_matrix).  The outer contour is a NACA 0015 outline (blunt trailing edge closed by the last segment),
    `ns` segments around, `nplies` plies through the wall, every ply one layer of 4-node quads: nn = ns (nplies + 1) nodes,
    ne = ns nplies elements, counterclockwise node order per element like MeshElement.nodenum.  The topology is shared; per instance
    vary the wall thickness (geometry: the inner surface is the contour scaled by 1 - 0.12 U(0.7, 1.3) about mid-chord), the ply angles (per ply U(-60, 60) degrees), the ply material
    scale (U(0.8, 1.2) on the moduli of a glass/epoxy-like orthotropic material) and the density.
    Returns dict(conn [ne, 4] 1-based, nodes [nb, nn, 2], materials [nb, ne, 10], theta [nb, ne])."""
    rng = np.random.default_rng(seed)
    # contour, counterclockwise, starting at the trailing edge upper side
    beta = np.linspace(0.0, 2 * np.pi, ns, endpoint=False)
    xc = 0.5 * chord * (1 + np.cos(beta))
    tt = 0.15
    yt = 5 * tt * chord * (0.2969 * np.sqrt(xc / chord) - 0.1260 * (xc / chord) - 0.3516 * (xc / chord) ** 2 + 0.2843 * (xc / chord) ** 3 - 0.1015 * (xc / chord) ** 4)
    yc = np.where(beta <= np.pi, yt, -yt)
    outer = np.stack([xc, yc], 1)
    # inner surface: the contour scaled about the mid-chord point (the section is star-shaped about it, so the plies never cross, also at
    # the sharp trailing edge); wall fraction tau_i per instance
    nr = nplies + 1
    centre = np.array([0.5 * chord, 0.0])
    tau = 0.12 * rng.uniform(0.7, 1.3, nbatch)
    frac = np.linspace(0.0, 1.0, nr)                         # 0 = outer surface
    scale = 1.0 - tau[:, None, None, None] * frac[None, None, :, None]
    nodes = centre[None, None, None, :] + scale * (outer - centre)[None, :, None, :]
    nodes = nodes.reshape(nbatch, ns * nr, 2)                # node (i, j) = i nr + j, j through the wall (0 = outside)
    conn = []
    for i in range(ns):
        ip = (i + 1) % ns
        for j in range(nplies):
            # counterclockwise in the local frame (x along the contour, y inward): (i, j), (ip, j), (ip, j+1), (i, j+1)
            conn.append([i * nr + j, ip * nr + j, ip * nr + j + 1, i * nr + j + 1])
    conn = np.array(conn, np.int64) + 1
    ne = len(conn)
    base = np.array([37.0e9, 9.0e9, 9.0e9, 4.0e9, 4.0e9, 3.0e9, 0.28, 0.28, 0.4, 1860.0])
    ply_scale = rng.uniform(0.8, 1.2, (nbatch, nplies))
    ply_angle = rng.uniform(-np.pi / 3, np.pi / 3, (nbatch, nplies))
    ply_rho = rng.uniform(0.9, 1.1, (nbatch, nplies))
    materials = np.tile(base[None, None, :], (nbatch, ne, 1))
    ply_of = np.tile(np.arange(nplies), ns)
    materials[:, :, 0:6] *= ply_scale[:, ply_of][:, :, None]
    materials[:, :, 9] *= ply_rho[:, ply_of]
    theta = ply_angle[:, ply_of]
    return dict(conn=conn, nodes=nodes, materials=materials, theta=theta, nn=ns * nr, ne=ne)
