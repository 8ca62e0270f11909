"""Test cases restated from the reference's own tests (test/examples/*.jl, test/jacobians.jl) as packed C-ABI arrays.

Each builder returns a dict with start, stop, points, pd, pl, elems[1,ne,91], bc[1,np,18], dload, pmass, gravity,
force_scaling plus case-specific extras.  Used for the oracle tests (CPU) and the GPU parity tests alike.
"""
import numpy as np

from gxbeam_b200.assembly import (Assembly, PrescribedConditions, default_force_scaling, discretize_beam, distributed_loads,
                                  pack_distributed_loads, pack_point_masses, pack_prescribed_conditions)


def _pack(asm, pcs, dl=None, pm=None, gravity=(0.0, 0.0, 0.0), **extra):
    pd, pl, bc = pack_prescribed_conditions(asm.np_, pcs)
    d = dict(start=asm.start, stop=asm.stop, points=asm.points, pd=pd, pl=pl, elems=asm.pack_elements()[None], bc=bc[None],
             dload=None if not dl else pack_distributed_loads(asm.ne, dl)[None],
             pmass=None if not pm else pack_point_masses(asm.np_, pm)[None], gravity=np.asarray(gravity, float),
             force_scaling=default_force_scaling(asm), asm=asm)
    d.update(extra)
    return d


def tipmoment(lam, nelem=16):
    """test/examples/tipmoment.jl:5-50"""
    L, h, w, E = 12.0, 1.0, 1.0, 30e6
    A, Iyy, Izz = h * w, w * h ** 3 / 12, w ** 3 * h / 12
    M = lam * np.pi * E * Iyy / L
    x = np.linspace(0, L, nelem + 1)
    points = np.stack([x, 0 * x, 0 * x], 1)
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    comp = [np.diag([1 / (E * A), 0, 0, 0, 1 / (E * Iyy), 1 / (E * Izz)])] * nelem
    asm = Assembly(points, start, stop, compliance=comp)
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0), nelem + 1: PrescribedConditions(Mz=M)}
    rho = E * Iyy / M if M != 0 else np.inf
    return _pack(asm, pcs, rho=rho)


def tipforce(lam, nelem=16):
    """test/examples/tipforce.jl:5-45"""
    L, EI = 1.0, 1e6
    P = lam * EI / L ** 2
    x = np.linspace(0, L, nelem + 1)
    points = np.stack([x, 0 * x, 0 * x], 1)
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    comp = [np.diag([0, 0, 0, 0, 1 / EI, 0])] * nelem
    asm = Assembly(points, start, stop, compliance=comp)
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0), nelem + 1: PrescribedConditions(Fz=P)}
    return _pack(asm, pcs)


def curved(nelem=16, P=600.0):
    """test/examples/curved.jl:5-58 (Bathe-Bolourchi 45-degree bend); reference output recorded to 16 digits at :56-58."""
    R = 100.0
    L = R * np.pi / 4
    h = w = 1.0
    E = 1e7
    G = E / 2
    frame = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    A = h * w
    Iyy, Izz = w * h ** 3 / 12, w ** 3 * h / 12
    Jt = Iyy + Izz
    dL, xp, xm, Cab = discretize_beam(L, [0, 0, 0], nelem, frame=frame, curvature=[0, 0, -1 / R])
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    comp = [np.diag([1 / (E * A), 1 / (G * A), 1 / (G * A), 1 / (G * Jt), 1 / (E * Iyy), 1 / (E * Izz)])] * nelem
    asm = Assembly(xp, start, stop, compliance=comp, frames=Cab, lengths=dL, midpoints=xm)
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0), nelem + 1: PrescribedConditions(Fz=P)}
    return _pack(asm, pcs, ref_tip_u=np.array([-13.577383726758564, -23.545303336988038, 53.45800757548929]))


def cantilever_partial_load(nelem=12):
    """test/examples/cantilever.jl:5-96 (linear, uniform load on the middle third, clamped at the right end)."""
    a, b, L = 0.3, 0.7, 1.0
    n1 = n3 = nelem // 3
    n2 = nelem - n1 - n3
    x = np.concatenate([np.linspace(0, a, n1 + 1), np.linspace(a, b, n2 + 1)[1:], np.linspace(b, L, n3 + 1)[1:]])
    points = np.stack([x, 0 * x, 0 * x], 1)
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    EI = 1e9
    asm = Assembly(points, start, stop, stiffness=[np.diag([0, 0, 0, 0, EI, 0])] * nelem)
    pcs = {nelem + 1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0)}
    q = 1000.0
    dl = {ie: distributed_loads(asm, ie, fz=lambda s: q) for ie in range(n1 + 1, n1 + n2 + 1)}

    def defl(xx):
        s0 = -q / (6 * EI) * ((L - a) ** 3 - (L - b) ** 3)
        d0 = q / (24 * EI) * ((L - a) ** 3 * (3 * L + a) - (L - b) ** 3 * (3 * L + b))
        d = d0 + s0 * xx
        if a < xx <= b:
            d += q / (24 * EI) * (xx - a) ** 4
        elif xx > b:
            d += q / (24 * EI) * ((xx - a) ** 4 - (xx - b) ** 4)
        return d

    return _pack(asm, pcs, dl=dl, deflection=defl)


def overdetermined(nelem=16):
    """test/examples/overdetermined.jl:5-68 (linear, propped cantilever under a linearly varying load)."""
    L, EI, qmax = 1.0, 1e7, 1000.0
    x = np.linspace(0, L, nelem + 1)
    points = np.stack([x, 0 * x, 0 * x], 1)
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    asm = Assembly(points, start, stop, compliance=[np.diag([0, 0, 0, 0, 1 / EI, 0])] * nelem)
    pcs = {1: PrescribedConditions(uz=0), nelem + 1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0)}
    dl = {i: distributed_loads(asm, i, s1=x[i - 1], s2=x[i], fz=lambda s: qmax * s) for i in range(1, nelem + 1)}
    defl = lambda xx: qmax * (1 - xx) ** 2 / (120 * EI) * (4 - 8 * (1 - xx) + 5 * (1 - xx) ** 2 - (1 - xx) ** 3)
    return _pack(asm, pcs, dl=dl, deflection=defl)


def random_anisotropic(seed=0, nelem=2, followers=True, gravity=True, pmass=True, dloads=True):
    """The system of test/jacobians.jl:98-175 (fully populated 6x6 stiffness/mass, tip point mass, gravity, follower loads,
    a DOF with theta_y = 0 prescribed) with our own seeded draws of the same magnitudes (Julia's RNG is unavailable)."""
    rng = np.random.default_rng(seed)
    L = 60.0
    x = np.linspace(0, L, nelem + 1)
    points = np.stack([x, 0 * x, 0 * x], 1)
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    A = rng.random((6, 6))
    stiff = (A @ A.T + 6 * np.eye(6)) * 1e5
    B = rng.random((6, 6))
    mass = B @ B.T + np.eye(6)
    frames = []
    for _ in range(nelem):
        q, _r = np.linalg.qr(rng.standard_normal((3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        frames.append(q)
    asm = Assembly(points, start, stop, stiffness=[stiff] * nelem, mass=[mass] * nelem, frames=frames)
    tip = dict(Fx=1e3 * rng.random(), Mz=1e3 * rng.random(), theta_y=0.0)
    if followers:
        tip.update(Fy_follower=1e3 * rng.random(), Mx_follower=1e2 * rng.random())
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0), nelem + 1: PrescribedConditions(**tip)}
    if nelem > 2 and followers:
        pcs[2] = PrescribedConditions(Fz_follower=50.0 * rng.random(), My=10.0)
    dl = None
    if dloads:
        dl = {1: distributed_loads(asm, 1, fx=lambda s: 10 * s, fz_follower=lambda s: 3.0, my=lambda s: 2.0, mz_follower=lambda s: s)}
    pm = None
    if pmass:
        M = rng.random((6, 6))
        pm = {nelem + 1: 1e3 * (M + M.T) / 2}
    g = 1e3 * rng.random(3) if gravity else (0, 0, 0)
    return _pack(asm, pcs, dl=dl, pm=pm, gravity=g)


def oracle_packed(case, b=0, two_dimensional=False):
    """Build the oracle's Problem for instance b of a (possibly batched) case."""
    from oracle import pyoracle as O
    g = np.asarray(case.get("gravity", np.zeros(3)), float)
    g = g[b] if g.ndim == 2 else g
    fs = case["force_scaling"]
    fs = float(fs[b]) if np.ndim(fs) else float(fs)
    dl = case.get("dload")
    pm = case.get("pmass")
    return O.Packed(case["start"], case["stop"], case["elems"][b], case["points"], case["pd"], case["pl"], case["bc"][b],
                    None if dl is None else dl[b], None if pm is None else pm[b], gravity=g, force_scaling=fs,
                    two_dimensional=two_dimensional)


def jacobians_jl(seed=0, nelem=2):
    """The exact system of test/jacobians.jl:98-150: L = 60, 2 elements, the test's own 6x6 stiffness and mass matrices, root with
    `theta_y = 0, My = 0` (a DOF with pd & pl -> body-frame angular acceleration becomes a state, system.jl:138-150), tip point
    mass, gravity; random draws (damping, point mass, gravity) from our own seeded generator with the test's magnitudes."""
    rng = np.random.default_rng(seed)
    L = 60.0
    x = np.linspace(0, L, nelem + 1)
    points = np.stack([x, 0 * x, 0 * x], 1)
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    stiff = np.array([[2.389e9, 1.524e6, 6.734e6, -3.382e7, -2.627e7, -4.736e8],
                      [1.524e6, 4.334e8, -3.741e6, -2.935e5, 1.527e7, 3.835e5],
                      [6.734e6, -3.741e6, 2.743e7, -4.592e5, -6.869e5, -4.742e6],
                      [-3.382e7, -2.935e5, -4.592e5, 2.167e7, -6.279e5, 1.430e6],
                      [-2.627e7, 1.527e7, -6.869e5, -6.279e5, 1.970e7, 1.209e7],
                      [-4.736e8, 3.835e5, -4.742e6, 1.430e6, 1.209e7, 4.406e8]])
    mass = np.array([[258.053, 0.0, 0.0, 0.0, 7.07839, -71.6871],
                     [0.0, 258.053, 0.0, -7.07839, 0.0, 0.0],
                     [0.0, 0.0, 258.053, 71.6871, 0.0, 0.0],
                     [0.0, -7.07839, 71.6871, 48.59, 0.0, 0.0],
                     [7.07839, 0.0, 0.0, 0.0, 2.172, 0.0],
                     [-71.6871, 0.0, 0.0, 0.0, 0.0, 46.418]])
    damping = [0.1 * rng.random(6)] * nelem
    asm = Assembly(points, start, stop, stiffness=[stiff] * nelem, mass=[mass] * nelem, damping=damping)
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, My=0, theta_z=0)}
    M = rng.random((6, 6))
    pm = {nelem + 1: 1e3 * (np.triu(M) + np.triu(M, 1).T)}     # Symmetric(1e3*rand(6,6))
    return _pack(asm, pcs, pm=pm, gravity=1e3 * rng.random(3))


def pointmass_jl():
    """test/issues/pointmass.jl:5-45: 10 massless anisotropic elements along y, point masses at every point (transform_properties of
    the test's 6x6 mass matrix), root clamped.  The reference records five eigen-frequencies to 16 digits (:54-58)."""
    nodes = np.array([[0, i / 10, 0] for i in range(11)], float)
    nelem = 10
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    stiff = np.array([[1.0e6, 0, 0, -0.5, -1.0, -50000.0], [0, 3.0e6, 0, 0, 0, 0], [0, 0, 3.0e6, 0, 0, 0],
                      [-0.5, 0, 0, 7.0, 0.1, -0.02], [-1.0, 0, 0, 0.1, 5.0, 0.1], [-50000.0, 0, 0, -0.02, 0.1, 3000.0]])
    mass = np.array([[0.02, 0, 0, 0, -5.0e-7, -1.0e-7], [0, 0.02, 0, 5.0e-7, 0, 0.0001], [0, 0, 0.02, 1.0e-7, -0.0001, 0],
                     [0, 5.0e-7, 1.0e-7, 1.0e-5, 1.0e-8, 2.0e-10], [-5.0e-7, 0, -0.0001, 1.0e-8, 6.0e-7, 9.0e-9],
                     [-1.0e-7, 0.0001, 0, 2.0e-10, 9.0e-9, 1.0e-5]])
    T = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    asm = Assembly(nodes, start, stop, frames=[T] * nelem, stiffness=[stiff] * nelem)
    Z = np.zeros((3, 3))
    TT = np.block([[T, Z], [Z, T]])
    pm = TT.T @ mass @ TT                                    # transform_properties (math.jl:43)
    pms = {1: pm / 2, nelem + 1: pm / 2}
    for i in range(2, nelem + 1):
        pms[i] = pm
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0)}
    return _pack(asm, pcs, pm=pms, ref_freq=np.array([2.8182004347800804, 17.66611982731975, 27.978670985969078, 49.93431945680836,
                                                     66.07594270678581]))


def initial_jl():
    """test/initial.jl:5-70: straight 31.5 in section (20 elements) + 6 in tip swept 45 degrees (4 elements), diagonal compliance / mass,
    root clamped, body frame spinning (angular_velocity) and translating (linear_velocity); structural damping off."""
    L1, n1, L2, n2 = 31.5, 20, 6.0, 4
    len1, xp1, xm1, Cab1 = discretize_beam(L1, [0, 0, 0], n1)
    c = 0.707106781186548
    frame = np.array([[c, c, 0.0], [-c, c, 0.0], [0.0, 0.0, 1.0]])
    len2, xp2, xm2, Cab2 = discretize_beam(L2, [31.5, 0, 0], n2, frame=frame)
    nelem = n1 + n2
    points = np.array(list(xp1) + list(xp2[1:]))
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    comp = np.diag([1.4974543276E-06, 4.7619054919E-06, 5.8036221902E-05, 3.1234387938E-03, 4.5274507258E-03, 1.7969451932E-05])
    mass = np.diag([1.5813000000E-05, 1.5813000000E-05, 1.5813000000E-05, 1.3229801498E-06, 5.2301497500E-09, 1.3177500000E-06])
    asm = Assembly(points, start, stop, compliance=[comp] * nelem, mass=[mass] * nelem, frames=list(Cab1) + list(Cab2),
                   lengths=list(len1) + list(len2), midpoints=list(xm1) + list(xm2))
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0),
           n1 + 1: PrescribedConditions(Fx=0, Fy=0, Fz=0, Mx=0, My=0, Mz=0)}
    return _pack(asm, pcs, linear_velocity=np.array([0, 196.349540849362, 0]), angular_velocity=np.array([0, 0, 78.5398163397448]))


def massless_initial_jl():
    """test/issues/zeros.jl:170-222 "Massless Element Time Domain Initialization": 10 massless elements along y, point masses at points
    2, 4, 6, 8, 10 (the last one halved), tip loads Fz = 30, My = -0.2.  The reference asserts `converged`."""
    comp = np.array([[6.00001e-6, 0.0, 0.0, 7.25923e-7, -8.1452e-7, 0.0001], [0.0, 3.33333e-7, 0.0, 0.0, 0.0, 0.0],
                     [0.0, 0.0, 3.33333e-7, 0.0, 0.0, 0.0], [7.25923e-7, 0.0, 0.0, 0.142898, -0.00285808, 1.31466e-5],
                     [-8.1452e-7, 0.0, 0.0, -0.00285808, 0.200057, -2.0263e-5], [0.0001, 0.0, 0.0, 1.31466e-5, -2.0263e-5, 0.002]])
    comp = np.triu(comp) + np.triu(comp, 1).T                      # Symmetric(...)
    mass = np.array([[0.02, 0.0, 0.0, 0.0, -5.0e-7, -1.0e-7], [0.0, 0.02, 0.0, 5.0e-7, 0.0, 0.0001], [0.0, 0.0, 0.02, 1.0e-7, -0.0001, 0.0],
                     [0.0, 5.0e-7, 1.0e-7, 1.0e-5, 1.0e-8, 2.0e-10], [-5.0e-7, 0.0, -0.0001, 1.0e-8, 6.0e-7, 9.0e-9],
                     [-1.0e-7, 0.0001, 0.0, 2.0e-10, 9.0e-9, 1.0e-5]])
    mass = np.triu(mass) + np.triu(mass, 1).T
    nodes = np.array([[0, i / 10, 0] for i in range(11)], float)
    nelem = 10
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    T = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    asm = Assembly(nodes, start, stop, frames=[T] * nelem, compliance=[comp] * nelem)
    Z = np.zeros((3, 3))
    TT = np.block([[T, Z], [Z, T]])
    pm = TT @ mass @ TT.T                                          # transform_properties(mass, T') = [T 0; 0 T] X [T' 0; 0 T'] (math.jl:43)
    pms = {2: pm}
    for i in range(4, nelem + 1, 2):
        pms[i] = pm
    pms[nelem] = pm / 2
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0), nelem + 1: PrescribedConditions(Fz=30, My=-0.2)}
    return _pack(asm, pcs, pm=pms)


def free_flying_beam(nelem=8, nstate_dofs=6):
    """A free-flying beam: the root has displacements AND loads prescribed (all zero) on `nstate_dofs` DOFs, so the corresponding
    body-frame accelerations become state variables (system.jl:138-150); a tip force accelerates the body.  Exact rigid-body check:
    linear acceleration = F / (rho A L)."""
    L = 2.0
    x = np.linspace(0, L, nelem + 1)
    points = np.stack([x, 0 * x, 0 * x], 1)
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    E, G, A_, I_, J_ = 70e9, 26e9, 1e-3, 1e-7, 2e-7
    comp = np.diag([1 / (E * A_), 1 / (G * A_ * 0.8), 1 / (G * A_ * 0.8), 1 / (G * J_), 1 / (E * I_), 1 / (E * I_)])
    rho = 2700.0
    mass = np.diag([rho * A_] * 3 + [rho * J_, rho * I_, rho * I_])
    asm = Assembly(points, start, stop, compliance=[comp] * nelem, mass=[mass] * nelem)
    names_d = ["ux", "uy", "uz", "theta_x", "theta_y", "theta_z"]
    names_l = ["Fx", "Fy", "Fz", "Mx", "My", "Mz"]
    root = {k: 0 for k in names_d}
    root.update({names_l[i]: 0 for i in range(nstate_dofs)})
    # the DOF-2 quirk of loads.jl:121 (Fz decides for uy): keep the flags consistent by prescribing all six loads when nstate_dofs == 6
    pcs = {1: PrescribedConditions(**root), nelem + 1: PrescribedConditions(Fz=5.0, Fx=20.0)}
    return _pack(asm, pcs, total_mass=rho * A_ * L)


def _zeros_jl_section():
    """Cross-section of test/issues/zeros.jl:31-56 (also test/examples/rotating.jl)."""
    w, h = 1.0, 0.063
    E, nu, rho = 1.06e7, 0.325, 2.51e-4
    ky, kz, kt = 1.2000001839588001, 14.625127919304001, 65.85255016982444
    A = h * w
    Iyy, Izz = w * h ** 3 / 12, w ** 3 * h / 12
    J = Iyy + Izz
    Ay, Az, Jx = A / ky, A / kz, J / kt
    G = E / (2 * (1 + nu))
    comp = np.diag([1 / (E * A), 1 / (G * Ay), 1 / (G * Az), 1 / (G * Jx), 1 / (E * Iyy), 1 / (E * Izz)])
    mass = np.diag([rho * A, rho * A, rho * A, rho * J, rho * Iyy, rho * Izz])
    return comp, mass


def zero_mass_jl():
    """test/issues/zeros.jl:3-77 "Zero Mass Matrix": swept-tip blade at 750 rpm with an all-zero mass matrix; steady_state_analysis
    must converge."""
    sweep, rpm = 45 * np.pi / 180, 750
    len1, xp1, xm1, Cab1 = discretize_beam(31.5, [2.5, 0, 0], 13)
    cs, ss = np.cos(sweep), np.sin(sweep)
    frame = np.array([[cs, ss, 0], [-ss, cs, 0], [0, 0, 1.0]])
    len2, xp2, xm2, Cab2 = discretize_beam(6.0, [34, 0, 0], 3, frame=frame)
    nelem = 16
    points = np.array(list(xp1) + list(xp2[1:]))
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    comp, _ = _zeros_jl_section()
    asm = Assembly(points, start, stop, compliance=[comp] * nelem, mass=[np.zeros((6, 6))] * nelem, frames=list(Cab1) + list(Cab2),
                   lengths=list(len1) + list(len2), midpoints=list(xm1) + list(xm2))
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0)}
    return _pack(asm, pcs, angular_velocity=np.array([0, 0, rpm * 2 * np.pi / 60]))


def zero_length_element_jl():
    """test/issues/zeros.jl:79-168 "Zero Length Element": the same blade with its real mass matrix and a zero-length element between
    the straight and the swept section (the joint point is duplicated); steady_state_analysis must converge."""
    sweep, rpm = 45 * np.pi / 180, 750
    len1, xp1, xm1, Cab1 = discretize_beam(31.5, [2.5, 0, 0], 13)
    len12, xp12, xm12, Cab12 = discretize_beam(0.0, [34, 0, 0], 1)
    cs, ss = np.cos(sweep), np.sin(sweep)
    frame = np.array([[cs, ss, 0], [-ss, cs, 0], [0, 0, 1.0]])
    len2, xp2, xm2, Cab2 = discretize_beam(6.0, [34, 0, 0], 3, frame=frame)
    nelem = 17
    points = np.array(list(xp1) + list(xp2))                      # vcat(xp_b1, xp_b2): the joint appears twice
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    comp, mass = _zeros_jl_section()
    asm = Assembly(points, start, stop, compliance=[comp] * nelem, mass=[mass] * nelem, frames=list(Cab1) + list(Cab12) + list(Cab2),
                   lengths=list(len1) + list(len12) + list(len2), midpoints=list(xm1) + list(xm12) + list(xm2))
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0)}
    return _pack(asm, pcs, angular_velocity=np.array([0, 0, rpm * 2 * np.pi / 60]))


def static_joined_wing(Fz=0.0, follower=False):
    """test/examples/static-joined-wing.jl:5-100: five beams (5+5+5+5+20 elements) forming a chain clamped at BOTH ends, load Fz (dead
    or follower) at the joint."""
    p = [np.array(v, float) for v in ([-7.1726, -12, -3.21539], [-5.37945, -9, -2.41154], [-3.5863, -6, -1.6077],
                                      [-1.79315, -3, -0.803848], [0, 0, 0], [7.1726, -12, 3.21539])]

    def frame(q, sgn):
        t1 = np.hypot(q[0], q[1]); c1, s1 = sgn * q[0] / t1, sgn * q[1] / t1
        rot1 = np.array([[c1, -s1, 0], [s1, c1, 0], [0, 0, 1.0]])
        t2 = np.linalg.norm(q); c2, s2 = t1 / t2, sgn * q[2] / t2
        rot2 = np.array([[c2, 0, -s2], [0, 1, 0], [s2, 0, c2]])
        return rot1 @ rot2
    Cab_1, Cab_2 = frame(p[0], -1.0), frame(p[5], 1.0)
    comps = [[1.05204e-9, 3.19659e-9, 2.13106e-8, 1.15475e-7, 1.52885e-7, 7.1672e-9],
             [1.24467e-9, 3.77682e-9, 2.51788e-8, 1.90461e-7, 2.55034e-7, 1.18646e-8],
             [1.60806e-9, 4.86724e-9, 3.24482e-8, 4.07637e-7, 5.57611e-7, 2.55684e-8],
             [2.56482e-9, 7.60456e-9, 5.67609e-8, 1.92171e-6, 2.8757e-6, 1.02718e-7],
             [2.77393e-9, 7.60456e-9, 1.52091e-7, 1.27757e-5, 2.7835e-5, 1.26026e-7]]
    nel = [5, 5, 5, 5, 20]
    lengths, points, mids, frames, comp = [], [], [], [], []
    for k in range(5):
        L = np.linalg.norm(p[k + 1] - p[k])
        ln, xp, xm, cab = discretize_beam(L, p[k], nel[k], frame=Cab_1 if k < 4 else Cab_2)
        lengths += list(ln); mids += list(xm); frames += list(cab); comp += [np.diag(comps[k])] * nel[k]
        points += list(xp) if k == 0 else list(xp[1:])
    nelem = sum(nel)
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    asm = Assembly(np.array(points), start, stop, compliance=comp, frames=frames, lengths=lengths, midpoints=mids)
    clamp = dict(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0)
    load = dict(Fz_follower=Fz) if follower else dict(Fz=Fz)
    pcs = {1: PrescribedConditions(**clamp), 21: PrescribedConditions(**load), nelem + 1: PrescribedConditions(**clamp)}
    return _pack(asm, pcs, joint=21)


def dynamic_joined_wing():
    """test/examples/dynamic-joined-wing.jl:5-83: two 4-element beams joined at the origin, both outer ends clamped, time-varying loads
    Fx = Fy = F_L(t), Fz = F_S(t) at the joint; tvec = range(0, 0.01, 201).  Returns the case at t = 0 plus the load series
    (fields 6, 7, 8 of point 5)."""
    p1, p2, p3 = np.zeros(3), np.array([-7.1726, -12, -3.21539]), np.array([7.1726, -12, 3.21539])
    Cab_1 = np.array([[0.5, 0.866025, 0.0], [0.836516, -0.482963, 0.258819], [0.224144, -0.12941, -0.965926]])
    Cab_2 = np.array([[0.5, 0.866025, 0.0], [-0.836516, 0.482963, 0.258819], [0.224144, -0.12941, 0.965926]])
    l1, xp1, xm1, c1 = discretize_beam(np.linalg.norm(p1 - p2), p2, 4, frame=Cab_1)
    l2, xp2, xm2, c2 = discretize_beam(np.linalg.norm(p3 - p1), p1, 4, frame=Cab_2)
    nelem = 8
    points = np.array(list(xp1) + list(xp2[1:]))
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    comp = np.diag([2.93944738387698e-10, 8.42991725049126e-10, 3.38313996669689e-08, 4.69246721094557e-08, 6.79584100559513e-08,
                    1.37068861370898e-09])
    mass = np.diag([4.86e-2, 4.86e-2, 4.86e-2, 1.0632465e-2, 2.10195e-4, 1.042227e-2])
    asm = Assembly(points, start, stop, compliance=[comp] * nelem, mass=[mass] * nelem, frames=list(c1) + list(c2),
                   lengths=list(l1) + list(l2), midpoints=list(xm1) + list(xm2))
    tvec = np.linspace(0, 0.01, 201)
    F_L = np.where(tvec < 0.01, 1e6 * tvec, np.where(tvec < 0.02, -1e6 * (tvec - 0.02), 0.0))
    F_S = np.where(tvec < 0.02, 5e3 * (1 - np.cos(np.pi * tvec / 0.02)), 1e4)
    clamp = dict(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0)
    pcs = {1: PrescribedConditions(**clamp), 5: PrescribedConditions(Fx=F_L[0], Fy=F_L[0], Fz=F_S[0]), nelem + 1: PrescribedConditions(**clamp)}
    series = np.stack([F_L, F_L, F_S], 1)                      # [nt, 3] -> fields 6, 7, 8 of point 5
    return _pack(asm, pcs, tvec=tvec, series=series)


def rotating_swept_tip(sweep_deg, nelem_b1=20, nelem_b2=20):
    """test/examples/rotating.jl:68-100: 31.5 in straight blade + 6 in tip swept by `sweep_deg`, 20 + 20 elements, root clamped;
    the reference compares its eigen-frequencies with the experiment of Epps & Chandra (table at rotating.jl:139-181)."""
    sweep = sweep_deg * np.pi / 180
    len1, xp1, xm1, Cab1 = discretize_beam(31.5, [2.5, 0, 0], nelem_b1)
    cs, ss = np.cos(sweep), np.sin(sweep)
    frame = np.array([[cs, ss, 0], [-ss, cs, 0], [0, 0, 1.0]])
    len2, xp2, xm2, Cab2 = discretize_beam(6.0, [34, 0, 0], nelem_b2, frame=frame)
    nelem = nelem_b1 + nelem_b2
    points = np.array(list(xp1) + list(xp2[1:]))
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    comp, mass = _zeros_jl_section()
    asm = Assembly(points, start, stop, compliance=[comp] * nelem, mass=[mass] * nelem, frames=list(Cab1) + list(Cab2),
                   lengths=list(len1) + list(len2), midpoints=list(xm1) + list(xm2))
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0)}
    return _pack(asm, pcs)


def wind_turbine_blade_jl():
    """test/examples/wind-turbine-blade.jl:5-52: 60 m blade, 5 elements, fully populated 6x6 stiffness and mass, tip load
    Fz = 1e5 sin(20 t), tvec = 0:0.001:0.1."""
    from gxbeam_b200.workloads import WT_STIFFNESS, WT_MASS
    nelem = 5
    x = np.linspace(0, 60.0, nelem + 1)
    points = np.stack([x, 0 * x, 0 * x], 1)
    start, stop = np.arange(1, nelem + 1), np.arange(2, nelem + 2)
    asm = Assembly(points, start, stop, stiffness=[WT_STIFFNESS] * nelem, mass=[WT_MASS] * nelem)
    tvec = np.arange(0, 0.1 + 1e-12, 0.001)
    pcs = {1: PrescribedConditions(ux=0, uy=0, uz=0, theta_x=0, theta_y=0, theta_z=0), nelem + 1: PrescribedConditions(Fz=0.0)}
    return _pack(asm, pcs, tvec=tvec, series=(1e5 * np.sin(20 * tvec))[:, None])
