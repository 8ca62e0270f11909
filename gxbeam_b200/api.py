"""ctypes binding of libgxbeam_b200.so (include/gxbeam_b200.h) and the batched analysis entry point.

`static_analysis_batch` mirrors the reference's `static_analysis(assembly; prescribed_conditions, distributed_loads,
point_masses, gravity, linear, two_dimensional, ftol, iterations, ...)` (src/analyses.jl:66-179) for a batch of
independent assemblies that share one topology and one pd/pl pattern.  The product path is the CUDA library only: if the
library is missing or no GPU is usable this module raises, it never falls back to a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

from . import assembly as asm_mod

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path():
    # GXBEAM_B200_LIB lets experiments point at an alternative build of the same CUDA library (never a CPU implementation)
    return os.environ.get("GXBEAM_B200_LIB") or os.path.join(_HERE, "libgxbeam_b200.so")


class GXBError(RuntimeError):
    pass


class Opts(C.Structure):
    _fields_ = [("linear", C.c_int32), ("two_dimensional", C.c_int32), ("ftol", C.c_double), ("iterations", C.c_int64),
                ("maxstep", C.c_double), ("c1", C.c_double), ("rho_hi", C.c_double), ("rho_lo", C.c_double),
                ("profile", C.c_int32), ("structural_damping", C.c_int32), ("persistent", C.c_int32), ("update_linearization", C.c_int32),
                ("lu_parts", C.c_int32), ("exact_dphi0", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("rounds", C.c_int64), ("launches", C.c_int64), ("newton_iters", C.c_int64), ("resid_evals", C.c_int64),
                ("n_assemble", C.c_int64), ("n_factor", C.c_int64), ("n_update", C.c_int64), ("ms_assemble", C.c_double),
                ("ms_factor", C.c_double), ("ms_update", C.c_double), ("ms_total", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ["gxb_default_opts", "gxb_create", "gxb_destroy", "gxb_last_error", "gxb_topology", "gxb_pattern", "gxb_set_batch",
           "gxb_set_elements", "gxb_set_points", "gxb_set_bc", "gxb_set_dloads", "gxb_set_pmass", "gxb_set_body",
           "gxb_set_motion", "gxb_static_solve", "gxb_steady_solve", "gxb_steady_solve_resident", "gxb_upload_x0",
           "gxb_static_solve_resident", "gxb_fetch", "gxb_get_stats", "gxb_residual_jacobian", "gxb_newton_direction",
           "gxb_set_bc_series", "gxb_newmark_run", "gxb_newmark_residual_jacobian", "gxb_mass_matrix", "gxb_eigen_arnoldi",
           "gxb_initial_condition", "gxb_initial_residual_jacobian", "gxb_body_columns", "gxb_eigen_ritz_vectors", "gxb_hint_no_gravity", "gxb_eigen_status", "gxb_set_bc_compact", "gxb_assembly_state", "gxb_eigen_arnoldi_left",
           "gxb_multi_create", "gxb_multi_destroy", "gxb_multi_topology", "gxb_multi_set_batch", "gxb_multi_solve",
           "gxb_section_create", "gxb_section_destroy", "gxb_section_info", "gxb_section_properties", "gxb_section_element_matrices",
           "gxb_section_strain_recovery", "gxb_multi_hint_no_gravity", "gxb_multi_solve_compact"]

_lib = None


def load_library():
    """Load the CUDA library.  Fails loudly when it has not been built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise GXBError(f"{path} not found: build it with gxbeam_b200/csrc/build.sh (there is no CPU fallback)")
        lib = C.CDLL(path)
        lib.gxb_last_error.restype = C.c_char_p
        for name in EXPORTS:
            getattr(lib, name)  # AttributeError if a declared symbol is missing
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data) if isinstance(a, np.ndarray) else C.c_void_p(int(a))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def make_opts(linear=False, two_dimensional=False, ftol=1e-9, iterations=1000, maxstep=1e6, c1=1e-4, rho_hi=0.5, rho_lo=0.1,
              profile=False, structural_damping=False, persistent=0, update_linearization=False, lu_parts=0, exact_dphi0=False):
    """persistent: 0 auto, 1 force, 2 forbid the persistent per-instance kernel (kind=1 contexts).
    lu_parts: warps per instance of the partitioned factorisation (kind=0): 0 auto, 1 off, 2 / 3 / 4 / 6 / 8."""
    return Opts(int(linear), int(two_dimensional), ftol, iterations, maxstep, c1, rho_hi, rho_lo, int(profile), int(structural_damping),
                int(persistent), int(update_linearization), int(lu_parts), int(exact_dphi0))


class Context:
    """One gxb_ctx = one CUDA device.  Thin, literal wrapper over the C-ABI (host pointers in, host pointers out)."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        self._ck(self.lib.gxb_create(C.byref(self.h), int(device)))
        self.device = device
        self.n = 0
        self.nbatch = 0

    def _ck(self, rc):
        if rc != 0:
            raise GXBError(f"gxbeam_b200 error {rc}: {self.lib.gxb_last_error().decode()}")

    def close(self):
        if self.h:
            self.lib.gxb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- topology / pattern ---
    def topology(self, start, stop, pd=None, pl=None, kind=0):
        start = np.ascontiguousarray(start, dtype=np.int64)
        stop = np.ascontiguousarray(stop, dtype=np.int64)
        ne = len(start)
        npnt = int(max(start.max(), stop.max()))
        pd = None if pd is None else np.ascontiguousarray(pd, dtype=np.uint8)
        pl = None if pl is None else np.ascontiguousarray(pl, dtype=np.uint8)
        n = C.c_int64(0)
        irp, icp = np.zeros(npnt, np.int64), np.zeros(npnt, np.int64)
        ire, ice = np.zeros(ne, np.int64), np.zeros(ne, np.int64)
        self._ck(self.lib.gxb_topology(self.h, int(kind), C.c_int64(npnt), C.c_int64(ne), _ptr(start), _ptr(stop), _ptr(pd), _ptr(pl),
                                       C.byref(n), _ptr(irp), _ptr(ire), _ptr(icp), _ptr(ice)))
        self.n, self.np, self.ne, self.kind = n.value, npnt, ne, int(kind)
        self.start, self.stop = start, stop
        self.indices = dict(nstates=n.value, irow_point=irp, irow_elem=ire, icol_point=icp, icol_elem=ice)
        return self.indices

    def pattern(self):
        nb = C.c_int64(0)
        self._ck(self.lib.gxb_pattern(self.h, C.byref(nb), None, None))
        brow, bcol = np.zeros(nb.value, np.int64), np.zeros(nb.value, np.int64)
        self._ck(self.lib.gxb_pattern(self.h, C.byref(nb), _ptr(brow), _ptr(bcol)))
        return brow, bcol

    # --- inputs ---
    def set_batch(self, nbatch):
        self._ck(self.lib.gxb_set_batch(self.h, C.c_int64(nbatch)))
        self.nbatch = nbatch

    def hint_no_gravity(self, on=True):
        """Static contexts: gravity is zero, upload only the 49 doubles per element that the path reads (54 % of the PCIe bytes)."""
        self._ck(self.lib.gxb_hint_no_gravity(self.h, int(on)))

    def set_elements(self, elems):
        e = _f64(elems)
        assert e.size == self.nbatch * self.ne * asm_mod.ELEM_STRIDE, "elements must be nbatch x ne x 91"
        self._ck(self.lib.gxb_set_elements(self.h, _ptr(e)))

    def set_points(self, xyz):
        x = _f64(xyz)
        assert x.size == self.nbatch * self.np * 3, "points must be nbatch x np x 3"
        self._ck(self.lib.gxb_set_points(self.h, _ptr(x)))

    def set_motion(self, motion):
        """nbatch x 12: linear_velocity, angular_velocity, linear_acceleration, angular_acceleration (kind=1)."""
        m = _f64(motion)
        assert m is None or m.size == self.nbatch * 12
        self._ck(self.lib.gxb_set_motion(self.h, _ptr(m)))

    def set_bc(self, bc):
        b = _f64(bc)
        assert b.size == self.nbatch * self.np * asm_mod.BC_STRIDE, "bc must be nbatch x np x 18"
        self._ck(self.lib.gxb_set_bc(self.h, _ptr(b)))

    def set_bc_compact(self, base, point, field, values):
        """base [np, 18] shared by the batch; values[b, k] -> field[k] (0..17) of point[k] (1-based): K doubles per instance cross PCIe."""
        base = _f64(base); v = _f64(values)
        assert base.size == self.np * asm_mod.BC_STRIDE and v.shape[0] == self.nbatch
        pt = np.ascontiguousarray(point, dtype=np.int32); fd = np.ascontiguousarray(field, dtype=np.int32)
        K = len(pt)
        assert v.size == self.nbatch * K and len(fd) == K
        self._ck(self.lib.gxb_set_bc_compact(self.h, _ptr(base), C.c_int32(K), _ptr(pt), _ptr(fd), _ptr(v)))

    def set_dloads(self, dl):
        d = _f64(dl)
        assert d is None or d.size == self.nbatch * self.ne * asm_mod.DLOAD_STRIDE
        self._ck(self.lib.gxb_set_dloads(self.h, _ptr(d)))

    def set_pmass(self, pm):
        p = _f64(pm)
        assert p is None or p.size == self.nbatch * self.np * 36
        self._ck(self.lib.gxb_set_pmass(self.h, _ptr(p)))

    def set_body(self, gravity=None, force_scaling=None):
        g = _f64(gravity)
        fs = _f64(force_scaling)
        assert g is None or g.size == self.nbatch * 3
        assert fs is None or fs.size == self.nbatch
        self._ck(self.lib.gxb_set_body(self.h, _ptr(g), _ptr(fs)))

    # --- solves ---
    def static_solve(self, opts=None, x0=None, out=None):
        """Host buffers in / out (the e2e path).  `out` may carry preallocated (pinned) arrays x, converged, iters, resid_inf."""
        o = opts or make_opts()
        nb, n = self.nbatch, self.n
        out = out or {}
        x = out.get("x") if out.get("x") is not None else np.empty((nb, n))
        conv = out.get("converged") if out.get("converged") is not None else np.empty(nb, np.int32)
        its = out.get("iters") if out.get("iters") is not None else np.empty(nb, np.int32)
        rinf = out.get("resid_inf") if out.get("resid_inf") is not None else np.empty(nb)
        x0a = _f64(x0)
        fn = self.lib.gxb_static_solve if self.kind == 0 else self.lib.gxb_steady_solve
        self._ck(fn(self.h, C.byref(o), _ptr(x0a), _ptr(x), _ptr(conv), _ptr(its), _ptr(rinf)))
        return dict(x=x, converged=conv.astype(bool), iters=its, resid_inf=rinf)

    steady_solve = static_solve   # same call shape; dispatches on the context's kind

    def upload_x0(self, x0=None):
        self._ck(self.lib.gxb_upload_x0(self.h, _ptr(_f64(x0))))

    def static_solve_resident(self, opts=None):
        o = opts or make_opts()
        fn = self.lib.gxb_static_solve_resident if self.kind == 0 else self.lib.gxb_steady_solve_resident
        self._ck(fn(self.h, C.byref(o)))

    steady_solve_resident = static_solve_resident

    def fetch(self):
        nb, n = self.nbatch, self.n
        x = np.empty((nb, n)); conv = np.empty(nb, np.int32); its = np.empty(nb, np.int32); rinf = np.empty(nb)
        self._ck(self.lib.gxb_fetch(self.h, _ptr(x), _ptr(conv), _ptr(its), _ptr(rinf)))
        return dict(x=x, converged=conv.astype(bool), iters=its, resid_inf=rinf)

    def assembly_state(self, rates=None):
        """PointState / ElementState records of the resident states: (points [nb, np, 30], elements [nb, ne, 30]); field order u, udot,
        theta, thetadot, V, Vdot, Omega, Omegadot, F, M (the reference's struct layout)."""
        pts = np.empty((self.nbatch, self.np, 30)); els = np.empty((self.nbatch, self.ne, 30))
        r = None if rates is None else _f64(rates).reshape(self.nbatch, self.np, 12)
        self._ck(self.lib.gxb_assembly_state(self.h, _ptr(r), _ptr(pts), _ptr(els)))
        return pts, els

    def stats(self):
        st = Stats()
        self._ck(self.lib.gxb_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def static_residual_jacobian(self, x, opts=None, want_jacobian=True):
        o = opts or make_opts()
        x = _f64(x).reshape(self.nbatch, self.n)
        r = np.empty_like(x)
        J = None
        if want_jacobian:
            brow, _ = self.pattern()
            J = np.empty((self.nbatch, len(brow), 9))
        self._ck(self.lib.gxb_residual_jacobian(self.h, C.byref(o), _ptr(x), _ptr(r), _ptr(J)))
        return r, J

    residual_jacobian = static_residual_jacobian

    def newmark_residual_jacobian(self, init, dt, x, opts=None):
        o = opts or make_opts()
        x = _f64(x).reshape(self.nbatch, self.n)
        init = _f64(init).reshape(self.nbatch, self.np, 12)
        r = np.empty_like(x)
        brow, _ = self.pattern()
        J = np.empty((self.nbatch, len(brow), 9))
        self._ck(self.lib.gxb_newmark_residual_jacobian(self.h, C.byref(o), _ptr(init), C.c_double(dt), _ptr(x), _ptr(r), _ptr(J)))
        return r, J

    # --- initial condition analysis ---
    def _pack_ic(self, u0=None, theta0=None, V0=None, Omega0=None, Vdot0=None, Omegadot0=None):
        """[nb, np, 18] = u0, theta0, V0, Omega0, Vdot0, Omegadot0 per point ([nb,np,3] or [np,3] each; None = zeros)."""
        parts = (u0, theta0, V0, Omega0, Vdot0, Omegadot0)
        if all(a is None for a in parts):
            return None
        ic = np.zeros((self.nbatch, self.np, 18))
        for k, a in enumerate(parts):
            if a is not None:
                ic[:, :, 3 * k:3 * k + 3] = np.broadcast_to(np.asarray(a, float).reshape(-1, self.np, 3), (self.nbatch, self.np, 3))
        return ic

    def initial_condition(self, x0=None, opts=None, want_rate_vars=False, **ic):
        """initial_condition_analysis! (src/analyses.jl:1695-1945) for the whole batch.  Returns the DynamicSystem states x [nb,n], the
        point rates [nb,np,12] (udot, θdot, Vdot, Ωdot) and the convergence data; both stay resident for newmark_run(x0=None)."""
        o = opts or make_opts(structural_damping=True)
        nb, n = self.nbatch, self.n
        icv = self._pack_ic(**ic)
        x = np.empty((nb, n)); rates = np.empty((nb, self.np, 12))
        conv = np.empty(nb, np.int32); its = np.empty(nb, np.int32); rinf = np.empty(nb)
        rv = np.empty((nb, self.np, 12), np.uint8) if want_rate_vars else None
        x0a = None if x0 is None else _f64(x0).reshape(nb, n)
        self._ck(self.lib.gxb_initial_condition(self.h, C.byref(o), _ptr(icv), _ptr(x0a), _ptr(x), _ptr(rates), _ptr(conv), _ptr(its),
                                                _ptr(rinf), _ptr(rv)))
        out = dict(x=x, rates=rates, converged=conv.astype(bool), iters=its, resid_inf=rinf)
        if want_rate_vars:
            out["rate_vars1"], out["rate_vars2"] = rv[:, :, :6], rv[:, :, 6:]
        return out

    def initial_residual_jacobian(self, x, opts=None, rate_vars1=None, rate_vars2=None, **ic):
        o = opts or make_opts()
        x = _f64(x).reshape(self.nbatch, self.n)
        icv = self._pack_ic(**ic)
        r = np.empty_like(x)
        brow, _ = self.pattern()
        J = np.empty((self.nbatch, len(brow), 9))
        rv1 = None if rate_vars1 is None else np.ascontiguousarray(rate_vars1, np.uint8)
        rv2 = None if rate_vars2 is None else np.ascontiguousarray(rate_vars2, np.uint8)
        self._ck(self.lib.gxb_initial_residual_jacobian(self.h, C.byref(o), _ptr(icv), _ptr(rv1), _ptr(rv2), _ptr(x), _ptr(r), _ptr(J)))
        return r, J

    def body_columns(self, want_columns=True):
        """icol_body (1-based, 0 = none) and the dense Jacobian columns [6, nb, n] of the body-acceleration states."""
        icb = np.zeros(6, np.int64)
        D = np.zeros((6, self.nbatch, self.n)) if want_columns else None
        self._ck(self.lib.gxb_body_columns(self.h, _ptr(icb), _ptr(D)))
        return icb, D

    # --- eigenvalue analysis ---
    def mass_matrix(self, x, opts=None):
        o = opts or make_opts()
        x = _f64(x).reshape(self.nbatch, self.n)
        brow, _ = self.pattern()
        M = np.empty((self.nbatch, len(brow), 9))
        self._ck(self.lib.gxb_mass_matrix(self.h, C.byref(o), _ptr(x), _ptr(M)))
        return M

    def eigen_arnoldi(self, x, m, opts=None, want_basis=True):
        o = opts or make_opts()
        x = _f64(x).reshape(self.nbatch, self.n)
        H = np.empty((self.nbatch, m + 1, m))
        V = np.empty((self.nbatch, m + 1, self.n)) if want_basis else None
        self._ck(self.lib.gxb_eigen_arnoldi(self.h, C.byref(o), _ptr(x), C.c_int32(m), _ptr(H), _ptr(V)))
        return H, V

    def eigen_status(self):
        s = np.zeros(self.nbatch, np.int32)
        self._ck(self.lib.gxb_eigen_status(self.h, _ptr(s)))
        return s.astype(bool)

    def eigenvalue_analysis(self, x, nev=6, m=None, opts=None, tol=1e-9, want_vectors=True, max_m=None, strict=True):
        """solve_eigensystem (src/analyses.jl:943-965) at the linearisation states x: returns (lam [nb, nev] complex,
        vec [nb, n, nev] complex or None, resid [nb, nev] Ritz residual estimates relative to |mu|).
        The Krylov basis never leaves the GPU: the m x m Hessenberg eigenproblems are solved on the host (stacked LAPACK call), the
        chosen eigenvectors go back and the Ritz vectors V y are formed on the device (gxb_eigen_ritz_vectors).

        `tol` is enforced like `partialschur(...; tol = 1e-9)` (analyses.jl:954): if some instance's nev leading Ritz pairs have a
        residual estimate above tol, the Krylov space is rebuilt twice as large (up to max_m, default min(n, 8 m)) until every
        instance is under tol.  If that cannot be reached, or K is singular for some instance (`lu(K)` would throw), `strict` raises;
        with strict = False the per-instance flags are left in `self.eigen_converged` / `self.eigen_singular`."""
        m = m or min(self.n, max(3 * nev + 10, 30))
        max_m = min(self.n, max_m or 8 * m)
        while True:
            lam, vec, res, H = self._eigen_once(x, nev, m, opts, want_vectors)
            sing = self.eigen_status()
            ok = (res <= tol).all(axis=1) & ~sing
            if ok.all() or sing.any() or m >= max_m:
                break
            m = min(2 * m, max_m)
        self.eigen_converged, self.eigen_singular, self.eigen_m = ok, sing, m
        if strict and sing.any():
            raise GXBError(f"eigenvalue_analysis: K is singular for instances {np.nonzero(sing)[0][:8].tolist()} (lu(K) throws in the reference)")
        if strict and not ok.all():
            raise GXBError(f"eigenvalue_analysis: {int((~ok).sum())} instance(s) did not reach tol = {tol} with a Krylov space of {m} "
                           f"(worst Ritz residual {float(res.max()):.2e}); pass a larger max_m or strict=False")
        return lam, vec, res

    def mass_times(self, x, vec, opts=None, chunk=64):
        """M v for complex vectors vec [nb, n, k], M = system_mass_matrix! at the states x (device assembly, host product)."""
        brow, bcol = self.pattern()
        Mb = self.mass_matrix(x, opts).reshape(self.nbatch, -1, 3, 3)
        out = np.zeros_like(vec)
        cols = bcol[:, None] + np.arange(3)[None, :]
        rows = (brow[:, None] + np.arange(3)[None, :]).ravel()
        for b0 in range(0, self.nbatch, chunk):
            sl = slice(b0, min(self.nbatch, b0 + chunk))
            prod = np.einsum("bkij,bkjn->bkin", Mb[sl], vec[sl][:, cols, :])
            acc = np.zeros((prod.shape[0], self.n, vec.shape[2]), complex)
            np.add.at(acc, (slice(None), rows), prod.reshape(prod.shape[0], -1, vec.shape[2]))
            out[sl] = acc
        return out

    def left_eigenvectors(self, x, lam, V, m=None, opts=None):
        """left_eigenvectors(system, λ, V) (src/analyses.jl:966-1079): U [nb, nev, n] complex, row k the left eigenvector of λ_k, normalised
        like the reference so that U M V = I.  The reference runs three passes of inverse iteration with a complex shifted factorisation
        per mode; here the left eigenvectors come out of ONE Arnoldi process on the transposed pencil (gxb_eigen_arnoldi_left), are matched
        to the given eigenvalues and scaled by 1 / (u_k' M v_k).  Returns (U, match) with match [nb, nev] = |λ_left - λ| / |λ| of the pairing."""
        nb, nev = lam.shape
        m = m or min(self.n, max(3 * nev + 10, 30))
        o = opts or make_opts()
        xs = _f64(x).reshape(nb, self.n)
        H = np.empty((nb, m + 1, m))
        self._ck(self.lib.gxb_eigen_arnoldi_left(self.h, C.byref(o), _ptr(xs), C.c_int32(m), _ptr(H), None))
        mu, Y = np.linalg.eig(np.ascontiguousarray(H[:, :m, :m]))
        lam_l = -1.0 / mu
        order = np.zeros((nb, nev), int); match = np.zeros((nb, nev))
        for b in range(nb):
            free = np.ones(m, bool)
            for k in range(nev):
                d = np.where(free, np.abs(lam_l[b] - lam[b, k]), np.inf)
                j = int(np.argmin(d))
                order[b, k] = j; free[j] = False; match[b, k] = d[j] / abs(lam[b, k])
        Yc = np.ascontiguousarray(np.take_along_axis(Y.astype(complex), order[:, None, :], 2))
        Ul = np.empty((nb, self.n, nev), complex)
        self._ck(self.lib.gxb_eigen_ritz_vectors(self.h, C.c_int32(nev), _ptr(Yc.view(np.float64)), _ptr(Ul.view(np.float64))))
        MV = self.mass_times(xs, V, o)                                        # [nb, n, nev]
        unorm = np.einsum("bnk,bnk->bk", Ul, MV)                              # u_k' M v_k (plain transpose: rows of U are conj(u) in the reference)
        U = np.transpose(Ul / unorm[:, None, :], (0, 2, 1))
        return np.ascontiguousarray(U), match

    def _eigen_once(self, x, nev, m, opts, want_vectors):
        H, _ = self.eigen_arnoldi(x, m, opts, want_basis=False)
        # stacked LAPACK dgeev, spread over host threads (numpy releases the GIL inside the gufunc loop)
        nth = max(1, min(os.cpu_count() or 1, 32, self.nbatch // 8))
        bounds = np.linspace(0, self.nbatch, nth + 1).astype(int)
        Hm = np.ascontiguousarray(H[:, :m, :m])
        with ThreadPoolExecutor(nth) as ex:
            parts = list(ex.map(lambda k: np.linalg.eig(Hm[bounds[k]:bounds[k + 1]]), range(nth)))
        mu = np.concatenate([q[0].astype(complex) for q in parts])            # [nb, m]
        Y = np.concatenate([q[1].astype(complex) for q in parts])             # [nb, m, m]
        # sort by (abs, imag) descending like `sortperm(..., by = mu -> (abs(mu), imag(mu)), rev = true)` (analyses.jl:956)
        order = np.lexsort((-mu.imag, -np.abs(mu)), axis=1)[:, :nev]
        mu = np.take_along_axis(mu, order, 1)
        Y = np.take_along_axis(Y, order[:, None, :], 2)                       # [nb, m, nev]
        lam = -1.0 / mu
        res = np.abs(H[:, m, m - 1][:, None] * Y[:, m - 1, :]) / np.abs(mu)
        vec = None
        if want_vectors:
            Yc = np.ascontiguousarray(Y.astype(complex))
            vec = np.empty((self.nbatch, self.n, nev), complex)
            self._ck(self.lib.gxb_eigen_ritz_vectors(self.h, C.c_int32(nev), _ptr(Yc.view(np.float64)), _ptr(vec.view(np.float64))))
        return lam, vec, res, H

    # --- time marching ---
    def set_bc_series(self, point, field, values):
        """Time-varying prescribed values: values[b, step, k] goes to field[k] (0..17: u,theta,F,M,Ff,Mf) of point[k] (1-based)."""
        v = _f64(values)
        nb, nt, K = v.shape
        assert nb == self.nbatch
        pt = np.ascontiguousarray(point, dtype=np.int32)
        fd = np.ascontiguousarray(field, dtype=np.int32)
        self._ck(self.lib.gxb_set_bc_series(self.h, C.c_int64(nt), C.c_int32(K), _ptr(pt), _ptr(fd), _ptr(v)))

    def newmark_run(self, tvec, x0=None, rates0=None, opts=None, save_every=1, want_history=True, want_rates=False):
        """time_domain_analysis! from the initial state (x0 [nb,n], rates0 [nb,np,12] = udot, θdot, Vdot, Ωdot)."""
        o = opts or make_opts()
        tvec = _f64(tvec)
        nt = len(tvec)
        nb, n = self.nbatch, self.n
        nsave = (nt - 1) // save_every + 1
        xh = np.empty((nb, nsave, n)) if want_history else None
        steps = np.empty(nb, np.int32); itot = np.empty(nb, np.int32); conv = np.empty(nb, np.int32)
        x0a = None if x0 is None else _f64(x0).reshape(nb, n)              # None: march from the state gxb_initial_condition left in HBM
        r0a = None if rates0 is None else _f64(rates0).reshape(nb, self.np, 12)
        rh = np.empty((nb, nsave, self.np, 12)) if want_rates else None      # udot, θdot, Vdot, Ωdot of every saved level (newmark_output)
        self._ck(self.lib.gxb_newmark_run(self.h, C.byref(o), C.c_int64(nt), _ptr(tvec), _ptr(x0a), _ptr(r0a), C.c_int64(save_every), _ptr(xh),
                                          _ptr(steps), _ptr(itot), _ptr(conv), _ptr(rh)))
        return dict(x=xh, rates=rh, steps_done=steps, iters_total=itot, converged=conv.astype(bool))

    def static_newton_direction(self, x, opts=None):
        o = opts or make_opts()
        x = _f64(x).reshape(self.nbatch, self.n)
        p = np.empty_like(x)
        self._ck(self.lib.gxb_newton_direction(self.h, C.byref(o), _ptr(x), _ptr(p)))
        return p

    newton_direction = static_newton_direction


class Section:
    """Batched cross-section analysis (src/section.jl: compliance_matrix :641-783, mass_matrix :840-890) for layups sharing one quad mesh.
    conn: [ne, 4] node numbers, 1-based, counterclockwise from the bottom-left node (MeshElement.nodenum)."""

    def __init__(self, nn, conn, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        conn = np.ascontiguousarray(conn, dtype=np.int64)
        self.nn, self.ne = int(nn), conn.shape[0]
        self._ck(self.lib.gxb_section_create(C.byref(self.h), int(device), C.c_int64(self.nn), C.c_int64(self.ne), _ptr(conn)))
        hb, sz = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.gxb_section_info(self.h, C.byref(hb), C.byref(sz)))
        self.half_bandwidth, self.scratch_doubles = hb.value, sz.value

    def _ck(self, rc):
        if rc != 0:
            raise GXBError(f"gxbeam_b200 error {rc}: {self.lib.gxb_last_error().decode()}")

    def _inputs(self, nodes, materials, theta):
        nodes = _f64(nodes); materials = _f64(materials); theta = _f64(theta)
        nb = theta.shape[0]
        assert nodes.shape == (nb, self.nn, 2) and materials.shape == (nb, self.ne, 10) and theta.shape == (nb, self.ne)
        return nb, nodes, materials, theta

    def properties(self, nodes, materials, theta, gxbeam_order=True, shear_center=True):
        """nodes [nb, nn, 2], materials [nb, ne, 10], theta [nb, ne] -> dict(S [nb,6,6], sc, tc [nb,2], M [nb,6,6], mc [nb,2], ok, device_ms)."""
        nb, nodes, materials, theta = self._inputs(nodes, materials, theta)
        S = np.empty((nb, 6, 6)); sc = np.empty((nb, 2)); tc = np.empty((nb, 2)); M = np.empty((nb, 6, 6)); mc = np.empty((nb, 2))
        ok = np.empty(nb, np.int32); ms = C.c_double(0)
        self._ck(self.lib.gxb_section_properties(self.h, C.c_int64(nb), _ptr(nodes), _ptr(materials), _ptr(theta), int(gxbeam_order), int(shear_center),
                                                 _ptr(S), _ptr(sc), _ptr(tc), _ptr(M), _ptr(mc), _ptr(ok), C.byref(ms)))
        return dict(S=S, sc=sc, tc=tc, M=M, mc=mc, ok=ok.astype(bool), device_ms=ms.value)

    def strain_recovery(self, nodes, materials, theta, F, M, strengths=None, gxbeam_order=True):
        """strain_recovery (src/section.jl:952-1040) and, with `strengths` [nb, ne, 9], tsai_wu (:1108-1131) of every layup under its section loads
        F, M [nb, 3] -> dict(strain_b, stress_b, strain_p, stress_p [nb, ne, 6], failure [nb, ne] or None, S [nb, 6, 6], ok, device_ms)."""
        nb, nodes, materials, theta = self._inputs(nodes, materials, theta)
        F = _f64(F); M = _f64(M)
        assert F.shape == (nb, 3) and M.shape == (nb, 3)
        st = None if strengths is None else _f64(strengths)
        assert st is None or st.shape == (nb, self.ne, 9)
        out = {k: np.empty((nb, self.ne, 6)) for k in ("strain_b", "stress_b", "strain_p", "stress_p")}
        fail = None if st is None else np.empty((nb, self.ne))
        S = np.empty((nb, 6, 6)); ok = np.empty(nb, np.int32); ms = C.c_double(0)
        self._ck(self.lib.gxb_section_strain_recovery(self.h, C.c_int64(nb), _ptr(nodes), _ptr(materials), _ptr(theta), _ptr(st), _ptr(F), _ptr(M),
                                                      int(gxbeam_order), _ptr(out["strain_b"]), _ptr(out["stress_b"]), _ptr(out["strain_p"]),
                                                      _ptr(out["stress_p"]), _ptr(fail), _ptr(S), _ptr(ok), C.byref(ms)))
        out.update(failure=fail, S=S, ok=ok.astype(bool), device_ms=ms.value)
        return out

    def element_matrices(self, nodes, materials, theta):
        nb, nodes, materials, theta = self._inputs(nodes, materials, theta)
        out = np.empty((nb, self.ne, 612))
        self._ck(self.lib.gxb_section_element_matrices(self.h, C.c_int64(nb), _ptr(nodes), _ptr(materials), _ptr(theta), _ptr(out)))
        return out

    def close(self):
        if self.h:
            self.lib.gxb_section_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiContext:
    """gxb_multi: one context + one host worker thread per GPU inside the library, contiguous shards, no collective."""

    def __init__(self, devices):
        self.lib = load_library()
        self.h = C.c_void_p()
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        self.ngpu = len(dev)
        rc = self.lib.gxb_multi_create(C.byref(self.h), C.c_int32(len(dev)), _ptr(dev))
        if rc:
            raise GXBError(f"gxbeam_b200 error {rc}: {self.lib.gxb_last_error().decode()}")

    def _ck(self, rc):
        if rc != 0:
            raise GXBError(f"gxbeam_b200 error {rc}: {self.lib.gxb_last_error().decode()}")

    def topology(self, start, stop, pd=None, pl=None, kind=0):
        start = np.ascontiguousarray(start, dtype=np.int64); stop = np.ascontiguousarray(stop, dtype=np.int64)
        ne = len(start); npnt = int(max(start.max(), stop.max()))
        pd = None if pd is None else np.ascontiguousarray(pd, dtype=np.uint8)
        pl = None if pl is None else np.ascontiguousarray(pl, dtype=np.uint8)
        n = C.c_int64(0)
        irp, icp, ire, ice = np.zeros(npnt, np.int64), np.zeros(npnt, np.int64), np.zeros(ne, np.int64), np.zeros(ne, np.int64)
        self._ck(self.lib.gxb_multi_topology(self.h, int(kind), C.c_int64(npnt), C.c_int64(ne), _ptr(start), _ptr(stop), _ptr(pd), _ptr(pl),
                                             C.byref(n), _ptr(irp), _ptr(ire), _ptr(icp), _ptr(ice)))
        self.n, self.np, self.ne, self.kind = n.value, npnt, ne, int(kind)
        self.indices = dict(nstates=n.value, irow_point=irp, irow_elem=ire, icol_point=icp, icol_elem=ice)
        return self.indices

    def set_batch(self, nbatch):
        self._ck(self.lib.gxb_multi_set_batch(self.h, C.c_int64(nbatch)))
        self.nbatch = nbatch

    def solve(self, elems, bc, opts=None, points=None, dloads=None, pmass=None, gravity=None, force_scaling=None, motion=None, x0=None, out=None):
        o = opts or make_opts()
        nb, n = self.nbatch, self.n
        out = out or {}
        x = out.get("x") if out.get("x") is not None else np.empty((nb, n))
        conv = out.get("converged") if out.get("converged") is not None else np.empty(nb, np.int32)
        its = out.get("iters") if out.get("iters") is not None else np.empty(nb, np.int32)
        rinf = out.get("resid_inf") if out.get("resid_inf") is not None else np.empty(nb)
        ms = np.zeros(self.ngpu)
        a = [_f64(v) for v in (elems, points, bc, dloads, pmass, gravity, force_scaling, motion, x0)]
        self._ck(self.lib.gxb_multi_solve(self.h, C.byref(o), *[_ptr(v) for v in a], _ptr(x), _ptr(conv), _ptr(its), _ptr(rinf), _ptr(ms)))
        return dict(x=x, converged=conv.astype(bool), iters=its, resid_inf=rinf, shard_ms=ms)

    def hint_no_gravity(self, on=True):
        self._ck(self.lib.gxb_multi_hint_no_gravity(self.h, int(on)))

    def solve_compact(self, elems, bc_base, bc_point, bc_field, bc_values, opts=None, gravity=None, force_scaling=None, x0=None, out=None):
        """gxb_multi_solve_compact: static_analysis of the whole batch from host buffers, prescribed conditions as one shared [np, 18] record
        plus K varying entries per instance; shards that share a GPU act as pipeline lanes (ordered uploads)."""
        o = opts or make_opts()
        nb, n = self.nbatch, self.n
        out = out or {}
        x = out.get("x") if out.get("x") is not None else np.empty((nb, n))
        conv = out.get("converged") if out.get("converged") is not None else np.empty(nb, np.int32)
        its = out.get("iters") if out.get("iters") is not None else np.empty(nb, np.int32)
        rinf = out.get("resid_inf") if out.get("resid_inf") is not None else np.empty(nb)
        ms = np.zeros(self.ngpu)
        pt = np.ascontiguousarray(bc_point, np.int32); fd = np.ascontiguousarray(bc_field, np.int32)
        vals = _f64(bc_values)
        assert vals.shape == (nb, len(pt))
        self._ck(self.lib.gxb_multi_solve_compact(self.h, C.byref(o), _ptr(_f64(elems)), _ptr(_f64(bc_base)), C.c_int32(len(pt)), _ptr(pt), _ptr(fd),
                                                  _ptr(vals), _ptr(_f64(gravity)), _ptr(_f64(force_scaling)), _ptr(_f64(x0)), _ptr(x), _ptr(conv),
                                                  _ptr(its), _ptr(rinf), _ptr(ms)))
        return dict(x=x, converged=conv.astype(bool) if conv.dtype != bool else conv, iters=its, resid_inf=rinf, shard_ms=ms)

    def close(self):
        if self.h:
            self.lib.gxb_multi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def compact_conditions(bc):
    """Split packed prescribed conditions bc [nb, np, 18] into what gxb_set_bc_compact / gxb_multi_solve_compact take: the record all
    instances share (base [np, 18] = instance 0), the 1-based point numbers and 0-based field numbers of the entries that differ between
    instances, and their per-instance values [nb, K].  NaN (unused) entries compare equal."""
    bc = np.asarray(bc, float)
    same = (bc == bc[:1]) | (np.isnan(bc) & np.isnan(bc[:1]))
    p, f = np.nonzero(~same.all(axis=0))
    return bc[0].copy(), (p + 1).astype(np.int32), f.astype(np.int32), np.ascontiguousarray(bc[:, p, f])


def correlate_eigenmodes(Cm):
    """correlate_eigenmodes(C) (src/analyses.jl:1114-1142), restated: for every row of the correlation matrix (C = U_p M V etc.) the
    best-fitting mode that is not assigned yet, and the corruption index = largest magnitude not chosen / magnitude chosen.
    Returns (perm, corruption), perm 0-based."""
    Cm = np.asarray(Cm)
    nev = Cm.shape[0]
    perm = -np.ones(nev, int)
    corruption = np.zeros(nev)
    for i in range(nev):
        tmp = np.abs(Cm[i])
        ranked = np.argsort(-tmp, kind="stable")
        i1 = next(j for j in ranked if j not in perm)
        i2 = ranked[1] if i1 == ranked[0] else ranked[0]
        perm[i] = i1
        corruption[i] = tmp[i2] / tmp[i1]
    return perm, corruption


def blocks_to_dense(n, brow, bcol, Jblk):
    """Scatter one instance's 3x3 pattern blocks into a dense n x n matrix (tests only)."""
    J = np.zeros((n, n))
    for k in range(len(brow)):
        J[brow[k]:brow[k] + 3, bcol[k]:bcol[k] + 3] = Jblk[k].reshape(3, 3)
    return J


@dataclass
class BatchResult:
    x: np.ndarray          # [nbatch, nstates]
    converged: np.ndarray  # [nbatch] bool
    iters: np.ndarray      # [nbatch] Newton iterations
    resid_inf: np.ndarray  # [nbatch]
    indices: dict
    stats: dict


class PipelinedSolver:
    """`static_analysis_batch` / `steady_state_analysis_batch` as a persistent service with host buffers in and out.

    `lanes` contexts live on one GPU, each with its own stream and chunk-sized device buffers.  A batch is cut into chunks that the
    lanes take round-robin from host threads (ctypes releases the GIL), so the host->device copy and packing of chunk k+1 overlap the
    Newton solve of chunk k and the device->host copy of chunk k-1: the PCIe transfer of the reference-layout inputs (728 B per
    element) is hidden behind the solve instead of preceding it."""

    def __init__(self, start, stop, pd, pl, chunk, lanes=2, device=0, kind=0, no_gravity=False, chunk_sizes=None):
        """chunk_sizes: optional list, one lane per entry, each lane sized to its own chunk (unequal cuts: a small LAST chunk shortens
        the tail after the last upload); otherwise `lanes` lanes take chunks of `chunk` instances round-robin."""
        self.chunk, self.kind = int(chunk), kind
        self.chunk_sizes = None if chunk_sizes is None else [int(v) for v in chunk_sizes]
        sizes = self.chunk_sizes if self.chunk_sizes is not None else [self.chunk] * lanes
        self.ctxs = []
        for sz in sizes:
            c = Context(device)
            self.indices = c.topology(start, stop, pd, pl, kind=kind)
            c.set_batch(sz)
            if kind == 0 and no_gravity:
                c.hint_no_gravity()
            self.ctxs.append(c)
        self.n = self.ctxs[0].n
        self.pool = ThreadPoolExecutor(len(sizes))
        self.upload_lock = threading.Lock()

    def solve(self, elements, prescribed_values, force_scaling, out, opts=None, gravity=None, points=None, motion=None, bc_compact=None):
        """elements [nb, ne, 91], prescribed_values [nb, np, 18], force_scaling [nb]; out: dict of preallocated host arrays x [nb, n],
        converged, iters (int32 [nb]), resid_inf [nb] (pinned memory makes the copies asynchronous).  nb must be a multiple of `chunk`.
        bc_compact = (base [np, 18], point [K], field [K], values [nb, K]) replaces prescribed_values (gxb_set_bc_compact)."""
        nb = elements.shape[0]
        o = opts or make_opts()
        if self.chunk_sizes is not None:
            if sum(self.chunk_sizes) != nb:
                raise GXBError("PipelinedSolver.solve: the chunk sizes must add up to the batch")
            bounds = np.concatenate([[0], np.cumsum(self.chunk_sizes)])
            slices = [[slice(int(bounds[k]), int(bounds[k + 1]))] for k in range(len(self.ctxs))]
        else:
            if nb % self.chunk:
                raise GXBError("PipelinedSolver.solve: the batch must be a multiple of the chunk size")
            nchunks = nb // self.chunk
            slices = [[slice(k * self.chunk, (k + 1) * self.chunk) for k in range(lane, nchunks, len(self.ctxs))] for lane in range(len(self.ctxs))]
        stats = [None] * len(self.ctxs)

        def run(lane):
            c = self.ctxs[lane]
            acc = {}
            for sl in slices[lane]:
                # one upload at a time: the PCIe link is the shared resource; serialising the uploads staggers the lanes so that one
                # lane's Newton solve runs under the other lane's transfer instead of both alternating in lock-step
                with self.upload_lock:
                    c.set_elements(elements[sl])
                    if bc_compact is not None:
                        c.set_bc_compact(bc_compact[0], bc_compact[1], bc_compact[2], bc_compact[3][sl])
                    else:
                        c.set_bc(prescribed_values[sl])
                    if points is not None:
                        c.set_points(points[sl])
                    if motion is not None:
                        c.set_motion(motion[sl])
                    c.set_body(None if gravity is None else gravity[sl], force_scaling[sl])
                c.static_solve(o, None, {key: out[key][sl] for key in ("x", "converged", "iters", "resid_inf")})
                st = c.stats()
                for key, v in st.items():
                    acc[key] = acc.get(key, 0) + v
            stats[lane] = acc
        list(self.pool.map(run, range(len(self.ctxs))))
        tot = {}
        for a in stats:
            for key, v in (a or {}).items():
                tot[key] = tot.get(key, 0) + v
        return tot

    def close(self):
        self.pool.shutdown()
        for c in self.ctxs:
            c.close()


def _solve_shard(device, start, stop, pd, pl, elems, bc, dl, pm, grav, fs, x0, opts):
    ctx = Context(device)
    try:
        idx = ctx.topology(start, stop, pd, pl)
        ctx.set_batch(elems.shape[0])
        ctx.set_elements(elems)
        ctx.set_bc(bc)
        if dl is not None:
            ctx.set_dloads(dl)
        if pm is not None:
            ctx.set_pmass(pm)
        ctx.set_body(grav, fs)
        res = ctx.static_solve(opts, x0)
        return res, idx, ctx.stats()
    finally:
        ctx.close()


def static_analysis_batch(start, stop, elements, prescribed_pd, prescribed_pl, prescribed_values, distributed_loads=None,
                          point_masses=None, gravity=None, force_scaling=None, x0=None, linear=False, two_dimensional=False,
                          ftol=1e-9, iterations=1000, devices=(0,), profile=False, lu_parts=0) -> BatchResult:
    """Batched `static_analysis` (src/analyses.jl:66-179) over nbatch independent assemblies sharing a topology.

    elements: [nbatch, ne, 91]; prescribed_values: [nbatch, np, 18]; pd/pl: [np, 6]; distributed_loads: [nbatch, ne, 24] | None;
    point_masses: [nbatch, np, 36] | None; gravity: [nbatch, 3] | None; force_scaling: [nbatch] | None (-> 1.0).
    The batch is sharded contiguously over `devices` (one context and one host thread per GPU, no collective); results
    are gathered into single host arrays.
    """
    elements = _f64(elements)
    nbatch = elements.shape[0]
    bc = _f64(prescribed_values).reshape(nbatch, -1, 18)
    opts = make_opts(linear=linear, two_dimensional=two_dimensional, ftol=ftol, iterations=iterations, profile=profile, lu_parts=lu_parts)
    fs = np.ones(nbatch) if force_scaling is None else _f64(force_scaling)
    bounds = np.linspace(0, nbatch, len(devices) + 1).astype(int)
    sl = [slice(bounds[i], bounds[i + 1]) for i in range(len(devices)) if bounds[i + 1] > bounds[i]]
    cut = lambda a, s: None if a is None else _f64(a)[s]
    args = [(devices[i], start, stop, prescribed_pd, prescribed_pl, elements[s], bc[s], cut(distributed_loads, s), cut(point_masses, s),
             cut(gravity, s), fs[s], cut(x0, s), opts) for i, s in enumerate(sl)]
    if len(args) == 1:
        outs = [_solve_shard(*args[0])]
    else:
        # several GPUs: the library's own multi-GPU entry (one worker thread + context per GPU inside the .so, gxb_multi_*)
        mc = MultiContext([devices[i] for i in range(len(sl))])
        try:
            idx = mc.topology(start, stop, prescribed_pd, prescribed_pl)
            mc.set_batch(nbatch)
            r = mc.solve(elements, bc, opts, dloads=distributed_loads, pmass=point_masses, gravity=gravity, force_scaling=fs, x0=x0)
        finally:
            mc.close()
        return BatchResult(r["x"], r["converged"], r["iters"], r["resid_inf"], idx, dict(shard_ms=r["shard_ms"].tolist()))
    x = np.concatenate([o[0]["x"] for o in outs])
    conv = np.concatenate([o[0]["converged"] for o in outs])
    its = np.concatenate([o[0]["iters"] for o in outs])
    rinf = np.concatenate([o[0]["resid_inf"] for o in outs])
    return BatchResult(x, conv, its, rinf, outs[0][1], outs[0][2])
