"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_build/libgxb_oracle.so.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU legs; the product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libgxb_oracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp"))]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "CXX=g++"], stdout=subprocess.DEVNULL)
    return _LIB


class Problem(C.Structure):
    _fields_ = [("np", C.c_int64), ("ne", C.c_int64), ("start", C.c_void_p), ("stop", C.c_void_p), ("elems", C.c_void_p),
                ("points", C.c_void_p), ("pd", C.c_void_p), ("pl", C.c_void_p), ("bc", C.c_void_p), ("dload", C.c_void_p),
                ("pmass", C.c_void_p), ("gravity", C.c_double * 3), ("force_scaling", C.c_double),
                ("two_dimensional", C.c_int32)]


class Opts(C.Structure):
    _fields_ = [("linear", C.c_int32), ("ftol", C.c_double), ("iterations", C.c_int64), ("maxstep", C.c_double),
                ("c1", C.c_double), ("rho_hi", C.c_double), ("rho_lo", C.c_double)]


class Dyn(C.Structure):
    _fields_ = [("vb", C.c_double * 3), ("wb", C.c_double * 3), ("ab", C.c_double * 3), ("alb", C.c_double * 3),
                ("structural_damping", C.c_int32)]


def make_dyn(vb=(0, 0, 0), wb=(0, 0, 0), ab=(0, 0, 0), alb=(0, 0, 0), structural_damping=False):
    return Dyn((C.c_double * 3)(*vb), (C.c_double * 3)(*wb), (C.c_double * 3)(*ab), (C.c_double * 3)(*alb), int(structural_damping))


def make_opts(linear=False, ftol=1e-9, iterations=1000, maxstep=1e6, c1=1e-4, rho_hi=0.5, rho_lo=0.1):
    return Opts(int(linear), ftol, iterations, maxstep, c1, rho_hi, rho_lo)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Packed:
    """Keeps the numpy arrays alive and exposes the C struct."""

    def __init__(self, start, stop, elems, points, pd, pl, bc, dload=None, pmass=None, gravity=(0, 0, 0), force_scaling=1.0,
                 two_dimensional=False):
        self.start = np.ascontiguousarray(start, dtype=np.int64)
        self.stop = np.ascontiguousarray(stop, dtype=np.int64)
        self.elems = np.ascontiguousarray(elems, dtype=np.float64)
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.pd = np.ascontiguousarray(pd, dtype=np.uint8)
        self.pl = np.ascontiguousarray(pl, dtype=np.uint8)
        self.bc = np.ascontiguousarray(bc, dtype=np.float64)
        self.dload = None if dload is None else np.ascontiguousarray(dload, dtype=np.float64)
        self.pmass = None if pmass is None else np.ascontiguousarray(pmass, dtype=np.float64)
        self.ne = len(self.start)
        self.np = int(max(self.start.max(), self.stop.max()))
        self.c = Problem(self.np, self.ne, _p(self.start), _p(self.stop), _p(self.elems), _p(self.points), _p(self.pd),
                         _p(self.pl), _p(self.bc), _p(self.dload), _p(self.pmass), (C.c_double * 3)(*gravity),
                         float(force_scaling), int(two_dimensional))


def indices(start, stop, static=True):
    start = np.ascontiguousarray(start, dtype=np.int64)
    stop = np.ascontiguousarray(stop, dtype=np.int64)
    ne = len(start)
    npnt = int(max(start.max(), stop.max()))
    n = C.c_int64(0)
    irp, icp = np.zeros(npnt, np.int64), np.zeros(npnt, np.int64)
    ire, ice = np.zeros(ne, np.int64), np.zeros(ne, np.int64)
    lib().gxo_indices(C.c_int64(ne), _p(start), _p(stop), int(static), C.byref(n), _p(irp), _p(ire), _p(icp), _p(ice))
    return dict(nstates=n.value, irow_point=irp, irow_elem=ire, icol_point=icp, icol_elem=ice)


def nstates(pk: Packed, static=True):
    return indices(pk.start, pk.stop, static)["nstates"]


def rotation(theta, dtheta):
    out = np.zeros(145)
    lib().gxo_rotation(_p(np.ascontiguousarray(theta, float)), _p(np.ascontiguousarray(dtheta, float)), _p(out))
    r = {}
    names = ["C", "Q", "Qinv", "dQ"]
    for i, nm in enumerate(names):
        r[nm] = out[36 * i:36 * i + 9].reshape(3, 3)
        r[nm + "_th"] = out[36 * i + 9:36 * i + 36].reshape(3, 3, 3)  # [k] = d/dθ_k
    r["scaling"] = out[144]
    return r


def rotation_cs(theta, dtheta):
    out = np.zeros(108)
    lib().gxo_rotation_cs(_p(np.ascontiguousarray(theta, float)), _p(np.ascontiguousarray(dtheta, float)), _p(out))
    return {nm + "_th": out[27 * i:27 * i + 27].reshape(3, 3, 3) for i, nm in enumerate(["C", "Q", "Qinv", "dQ"])}


def static_residual(pk: Packed, x):
    x = np.ascontiguousarray(x, float)
    r = np.zeros_like(x)
    lib().gxo_static_residual(C.byref(pk.c), _p(x), _p(r))
    return r


def static_jacobian(pk: Packed, x, complex_step=False):
    x = np.ascontiguousarray(x, float)
    n = len(x)
    J = np.zeros((n, n), order="F")
    fn = lib().gxo_static_jacobian_cs if complex_step else lib().gxo_static_jacobian
    fn(C.byref(pk.c), _p(x), _p(J))
    return J


def static_bandwidth(pk: Packed):
    kl, ku = C.c_int64(0), C.c_int64(0)
    lib().gxo_static_bandwidth(C.byref(pk.c), C.byref(kl), C.byref(ku))
    return kl.value, ku.value


def static_solve(pk: Packed, x0=None, **kw):
    n = nstates(pk)
    x = np.zeros(n)
    r = np.zeros(n)
    stats = np.zeros(4, np.int64)
    rinf = C.c_double(0)
    o = make_opts(**kw)
    x0a = None if x0 is None else np.ascontiguousarray(x0, float)
    err = lib().gxo_static_solve(C.byref(pk.c), C.byref(o), _p(x0a), _p(x), _p(r), _p(stats), C.byref(rinf))
    return dict(x=x, resid=r, converged=bool(stats[0]), iters=int(stats[1]), f_calls=int(stats[2]),
                ls_backtracks=int(stats[3]), resid_inf=rinf.value, error=err)


def static_solve_batch(pk0: Packed, nbatch, gravity=None, force_scaling=None, x0=None, nthreads=0, **kw):
    """pk0's per-instance arrays (elems, points, bc, dload, pmass) hold the whole batch contiguously."""
    n = nstates(pk0)
    x = np.zeros((nbatch, n))
    conv = np.zeros(nbatch, np.int32)
    its = np.zeros(nbatch, np.int32)
    rinf = np.zeros(nbatch)
    o = make_opts(**kw)
    g = None if gravity is None else np.ascontiguousarray(gravity, float)
    fs = None if force_scaling is None else np.ascontiguousarray(force_scaling, float)
    x0a = None if x0 is None else np.ascontiguousarray(x0, float)
    lib().gxo_static_solve_batch(C.c_int64(nbatch), C.byref(pk0.c), _p(g), _p(fs), C.byref(o), _p(x0a), _p(x), _p(conv),
                                 _p(its), _p(rinf), int(nthreads))
    return dict(x=x, converged=conv, iters=its, resid_inf=rinf)


def num_threads():
    return lib().gxo_num_threads()


# ---------------------------------------------------------------- steady state (DynamicSystem) ----
def steady_residual(pk: Packed, dyn: Dyn, x):
    x = np.ascontiguousarray(x, float)
    r = np.zeros_like(x)
    lib().gxo_steady_residual(C.byref(pk.c), C.byref(dyn), _p(x), _p(r))
    return r


def steady_jacobian(pk: Packed, dyn: Dyn, x, complex_step=False):
    x = np.ascontiguousarray(x, float)
    n = len(x)
    J = np.zeros((n, n), order="F")
    lib().gxo_steady_jacobian(C.byref(pk.c), C.byref(dyn), _p(x), _p(J), int(complex_step))
    return J


def steady_bandwidth(pk: Packed):
    kl, ku = C.c_int64(0), C.c_int64(0)
    icb = np.zeros(6, np.int64)
    lib().gxo_steady_bandwidth(C.byref(pk.c), C.byref(kl), C.byref(ku), _p(icb))
    return kl.value, ku.value, icb


def steady_solve(pk: Packed, dyn: Dyn, x0=None, **kw):
    n = nstates(pk, static=False)
    x = np.zeros(n)
    r = np.zeros(n)
    stats = np.zeros(4, np.int64)
    rinf = C.c_double(0)
    o = make_opts(**kw)
    x0a = None if x0 is None else np.ascontiguousarray(x0, float)
    err = lib().gxo_steady_solve(C.byref(pk.c), C.byref(dyn), C.byref(o), _p(x0a), _p(x), _p(r), _p(stats), C.byref(rinf))
    return dict(x=x, resid=r, converged=bool(stats[0]), iters=int(stats[1]), f_calls=int(stats[2]),
                ls_backtracks=int(stats[3]), resid_inf=rinf.value, error=err)


def steady_solve_batch(pk0: Packed, nbatch, dyns, gravity=None, force_scaling=None, x0=None, nthreads=0, **kw):
    """dyns: list of Dyn (one per instance).  pk0's per-instance arrays (elems, points, bc, ...) hold the whole batch."""
    n = nstates(pk0, static=False)
    x = np.zeros((nbatch, n))
    conv = np.zeros(nbatch, np.int32)
    its = np.zeros(nbatch, np.int32)
    rinf = np.zeros(nbatch)
    o = make_opts(**kw)
    arr = (Dyn * nbatch)(*dyns)
    g = None if gravity is None else np.ascontiguousarray(gravity, float)
    fs = None if force_scaling is None else np.ascontiguousarray(force_scaling, float)
    x0a = None if x0 is None else np.ascontiguousarray(x0, float)
    lib().gxo_steady_solve_batch(C.c_int64(nbatch), C.byref(pk0.c), arr, _p(g), _p(fs), C.byref(o), _p(x0a), _p(x), _p(conv),
                                 _p(its), _p(rinf), int(nthreads))
    return dict(x=x, converged=conv, iters=its, resid_inf=rinf)


# ---------------------------------------------------------------- Newmark time marching ----
def newmark_residual(pk: Packed, dyn: Dyn, init, dt, x):
    x = np.ascontiguousarray(x, float)
    init = np.ascontiguousarray(init, float)
    r = np.zeros_like(x)
    lib().gxo_newmark_residual(C.byref(pk.c), C.byref(dyn), _p(init), C.c_double(dt), _p(x), _p(r))
    return r


def newmark_jacobian(pk: Packed, dyn: Dyn, init, dt, x, complex_step=False):
    x = np.ascontiguousarray(x, float)
    init = np.ascontiguousarray(init, float)
    n = len(x)
    J = np.zeros((n, n), order="F")
    lib().gxo_newmark_jacobian(C.byref(pk.c), C.byref(dyn), _p(init), C.c_double(dt), _p(x), _p(J), int(complex_step))
    return J


def newmark_run(pk: Packed, dyn: Dyn, tvec, x0, rates0, bc_t=None, update_linearization=True, **kw):
    """time_domain_analysis! from a given initial state.  rates0: [np,12] = udot, θdot, Vdot, Ωdot at t0.  bc_t: [nt,np,18] or None."""
    tvec = np.ascontiguousarray(tvec, float)
    nt = len(tvec)
    n = nstates(pk, static=False)
    xh = np.zeros((nt, n))
    its = np.zeros(nt, np.int32)
    conv = C.c_int32(0)
    o = make_opts(**kw)
    bct = None if bc_t is None else np.ascontiguousarray(bc_t, float)
    lib().gxo_newmark_run2.restype = C.c_int64
    done = lib().gxo_newmark_run2(C.byref(pk.c), C.byref(dyn), C.byref(o), C.c_int64(nt), _p(tvec), _p(bct), _p(np.ascontiguousarray(x0, float)),
                                  _p(np.ascontiguousarray(rates0, float)), _p(xh), _p(its), C.byref(conv), int(update_linearization))
    return dict(x=xh, iters=its, converged=bool(conv.value), steps_done=int(done))


# ---------------------------------------------------------------- initial condition analysis ----
class Initial(C.Structure):
    _fields_ = [("rv1", C.c_void_p), ("rv2", C.c_void_p), ("u0", C.c_void_p), ("th0", C.c_void_p), ("V0", C.c_void_p),
                ("Om0", C.c_void_p), ("Vdot0", C.c_void_p), ("Omdot0", C.c_void_p)]


class InitialConditions:
    """u0, theta0, V0, Omega0, Vdot0, Omegadot0 ([np,3] each, default zeros) + rate_vars1/2 ([nstates] bool)."""

    def __init__(self, npnt, n, u0=None, theta0=None, V0=None, Omega0=None, Vdot0=None, Omegadot0=None, rv1=None, rv2=None):
        def arr(a):
            return np.zeros((npnt, 3)) if a is None else np.ascontiguousarray(a, float).reshape(npnt, 3)
        self.u0, self.th0, self.V0, self.Om0, self.Vdot0, self.Omdot0 = map(arr, (u0, theta0, V0, Omega0, Vdot0, Omegadot0))
        self.rv1 = np.zeros(n, np.uint8) if rv1 is None else np.ascontiguousarray(rv1, np.uint8)
        self.rv2 = np.zeros(n, np.uint8) if rv2 is None else np.ascontiguousarray(rv2, np.uint8)
        self.c = Initial(_p(self.rv1), _p(self.rv2), _p(self.u0), _p(self.th0), _p(self.V0), _p(self.Om0), _p(self.Vdot0),
                         _p(self.Omdot0))


def initial_rate_vars(pk: Packed):
    n = nstates(pk, static=False)
    rv1, rv2 = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    lib().gxo_initial_rate_vars(C.byref(pk.c), _p(rv1), _p(rv2))
    return rv1, rv2


def initial_residual(pk: Packed, dyn: Dyn, ic: InitialConditions, x):
    x = np.ascontiguousarray(x, float)
    r = np.zeros_like(x)
    lib().gxo_initial_residual(C.byref(pk.c), C.byref(dyn), C.byref(ic.c), _p(x), _p(r))
    return r


def initial_jacobian(pk: Packed, dyn: Dyn, ic: InitialConditions, x, compressed=True):
    x = np.ascontiguousarray(x, float)
    n = len(x)
    J = np.zeros((n, n), order="F")
    lib().gxo_initial_jacobian(C.byref(pk.c), C.byref(dyn), C.byref(ic.c), _p(x), _p(J), int(compressed))
    return J


def initial_condition(pk: Packed, dyn: Dyn, ic: InitialConditions, x0=None, **kw):
    """initial_condition_analysis!: returns the solved initial-condition vector, the DynamicSystem state vector `x`, the point
    rates [np,12] that seed newmark_run, and the rate_vars."""
    n = nstates(pk, static=False)
    x_ic, xx, rates = np.zeros(n), np.zeros(n), np.zeros((pk.np, 12))
    stats = np.zeros(4, np.int64)
    rinf = C.c_double(0)
    o = make_opts(**kw)
    x0a = None if x0 is None else np.ascontiguousarray(x0, float)
    err = lib().gxo_initial_condition(C.byref(pk.c), C.byref(dyn), C.byref(o), C.byref(ic.c), _p(x0a), _p(x_ic), _p(xx), _p(rates),
                                      _p(ic.rv1), _p(ic.rv2), _p(stats), C.byref(rinf))
    return dict(x_ic=x_ic, x=xx, rates=rates, rv1=ic.rv1.copy(), rv2=ic.rv2.copy(), converged=bool(stats[0]), iters=int(stats[1]),
                f_calls=int(stats[2]), ls_backtracks=int(stats[3]), resid_inf=rinf.value, error=err)


# ---------------------------------------------------------------- mass matrix / eigen ----
def mass_matrix(pk: Packed, x):
    x = np.ascontiguousarray(x, float)
    n = len(x)
    M = np.zeros((n, n), order="F")
    lib().gxo_mass_matrix(C.byref(pk.c), _p(x), _p(M))
    return M


def eigenvalues_dense(K, M, nev):
    """solve_eigensystem (src/analyses.jl:943-965) with a dense solver: eigenvalues of K^{-1} M of largest magnitude, sorted by
    (abs, imag) descending, returned as lambda = -1/mu together with the eigenvectors."""
    import scipy.linalg as sla
    mu, V = sla.eig(np.linalg.solve(K, M))
    order = sorted(range(len(mu)), key=lambda i: (abs(mu[i]), mu[i].imag), reverse=True)[:nev]
    mu = mu[order]
    return -1.0 / mu, V[:, order]
