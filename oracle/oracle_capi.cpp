// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points used through ctypes by tests/, smoke() and bench.py's CPU legs.
// The product library (gxbeam_b200/csrc) never includes, links or calls anything in this directory.
#include <complex>
#include <cstring>
#include <vector>
#include <atomic>
#include <thread>
#include "gxb_math.hpp"
#include "gxb_static.hpp"
#include "gxb_dynamic.hpp"
#include "gxb_solve.hpp"

using namespace gxo;
typedef std::complex<double> cplx;

extern "C" {

// mirrors gxo::Problem field by field (ctypes Structure in tests/oracle.py)
struct gxo_problem {
    int64_t np, ne;
    const int64_t* start;
    const int64_t* stop;
    const double* elems;
    const double* points;
    const uint8_t* pd;
    const uint8_t* pl;
    const double* bc;
    const double* dload;
    const double* pmass;
    double gravity[3];
    double force_scaling;
    int32_t two_dimensional;
};

struct gxo_dyn {       // mirrors gxo::DynParams
    double vb[3], wb[3], ab[3], alb[3];
    int32_t structural_damping;
};

struct gxo_opts {
    int32_t linear;
    double ftol;
    int64_t iterations;
    double maxstep, c1, rho_hi, rho_lo;
};
}

static Problem to_problem(const gxo_problem* q) {
    Problem P;
    P.np = q->np; P.ne = q->ne; P.start = q->start; P.stop = q->stop; P.elems = q->elems; P.points = q->points;
    P.pd = q->pd; P.pl = q->pl; P.bc = q->bc; P.dload = q->dload; P.pmass = q->pmass;
    for (int i = 0; i < 3; i++) P.gravity[i] = q->gravity[i];
    P.force_scaling = q->force_scaling; P.two_dimensional = q->two_dimensional;
    return P;
}
static NewtonOpts to_opts(const gxo_opts* o) {
    NewtonOpts r;
    if (o) { r.linear = o->linear; r.ftol = o->ftol; r.iterations = o->iterations; r.maxstep = o->maxstep; r.c1 = o->c1; r.rho_hi = o->rho_hi; r.rho_lo = o->rho_lo; }
    return r;
}

// scalar bandwidths of the static Jacobian from the 3x3-block insertion pattern (element.jl:2487-2552, point.jl:1795-1797)
static void static_bandwidth(const Problem& P, const Indices& idx, int64_t& kl, int64_t& ku) {
    kl = 0; ku = 0;
    auto blk = [&](int64_t r0, int64_t nr, int64_t c0, int64_t nc) {
        kl = std::max(kl, (r0 + nr - 1) - c0);
        ku = std::max(ku, (c0 + nc - 1) - r0);
    };
    for (int64_t ip = 0; ip < P.np; ip++) blk(idx.irow_point[ip], 6, idx.icol_point[ip], 6);
    for (int64_t ie = 0; ie < P.ne; ie++) {
        int64_t ic = idx.icol_elem[ie], ic1 = idx.icol_point[P.start[ie] - 1], ic2 = idx.icol_point[P.stop[ie] - 1];
        int64_t ir = idx.irow_elem[ie], ir1 = idx.irow_point[P.start[ie] - 1], ir2 = idx.irow_point[P.stop[ie] - 1];
        blk(ir, 3, ic1, 6); blk(ir, 3, ic2, 6); blk(ir, 6, ic, 6); blk(ir + 3, 3, ic1 + 3, 3); blk(ir + 3, 3, ic2 + 3, 3);
        for (int64_t irp : {ir1, ir2}) {
            blk(irp, 6, ic1 + 3, 3); blk(irp, 6, ic2 + 3, 3); blk(irp, 3, ic, 3); blk(irp + 3, 3, ic, 6);
        }
    }
    if (P.two_dimensional) { kl = std::max<int64_t>(kl, 0); ku = std::max<int64_t>(ku, 0); }
}

extern "C" {

int gxo_indices(int64_t ne, const int64_t* start, const int64_t* stop, int is_static, int64_t* nstates, int64_t* irow_point,
                int64_t* irow_elem, int64_t* icol_point, int64_t* icol_elem) {
    Indices idx = system_indices(ne, start, stop, is_static != 0);
    *nstates = idx.nstates;
    std::memcpy(irow_point, idx.irow_point.data(), idx.irow_point.size() * 8);
    std::memcpy(icol_point, idx.icol_point.data(), idx.icol_point.size() * 8);
    std::memcpy(irow_elem, idx.irow_elem.data(), idx.irow_elem.size() * 8);
    std::memcpy(icol_elem, idx.icol_elem.data(), idx.icol_elem.size() * 8);
    return 0;
}

static void put33(double* out, const M3<double>& A) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[3 * i + j] = A.m[i][j]; }

// Rotation kernels and their analytic derivatives at (θ, Δθ); every 3x3 is written row-major.
//   out: C[9] C_θ[27] Q[9] Q_θ[27] Qinv[9] Qinv_θ[27] ΔQ[9] ΔQ_θ[27] scaling[1]   (145 doubles)
void gxo_rotation(const double* th_, const double* dth_, double* out) {
    V3<double> th = cvec<double>(th_), dth = cvec<double>(dth_);
    M3<double> C = get_C(th), A1, A2, A3;
    put33(out, C); get_C_th(C, th, A1, A2, A3); put33(out + 9, A1); put33(out + 18, A2); put33(out + 27, A3);
    M3<double> Q = get_Q(th), Q1, Q2, Q3;
    put33(out + 36, Q); get_Q_th(Q, th, Q1, Q2, Q3); put33(out + 45, Q1); put33(out + 54, Q2); put33(out + 63, Q3);
    put33(out + 72, get_Qinv(th)); get_Qinv_th(th, A1, A2, A3); put33(out + 81, A1); put33(out + 90, A2); put33(out + 99, A3);
    put33(out + 108, get_dQ(th, dth, Q)); get_dQ_th(th, dth, Q, Q1, Q2, Q3, A1, A2, A3);
    put33(out + 117, A1); put33(out + 126, A2); put33(out + 135, A3);
    out[144] = rotation_parameter_scaling(th);
}
// Same derivatives by complex step (the oracle's stand-in for ForwardDiff in test/jacobians.jl:3-39):
//   out: C_θ[27] Q_θ[27] Qinv_θ[27] ΔQ_θ[27]
void gxo_rotation_cs(const double* th_, const double* dth_, double* out) {
    const double h = 1e-30;
    for (int k = 0; k < 3; k++) {
        V3<cplx> th = cvec<cplx>(th_), dth = cvec<cplx>(dth_);
        th.v[k] += cplx(0, h);
        M3<cplx> C = get_C(th), Q = get_Q(th), Qi = get_Qinv(th), dQ = get_dQ(th, dth, Q);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                out[9 * k + 3 * i + j] = C.m[i][j].imag() / h;
                out[27 + 9 * k + 3 * i + j] = Q.m[i][j].imag() / h;
                out[54 + 9 * k + 3 * i + j] = Qi.m[i][j].imag() / h;
                out[81 + 9 * k + 3 * i + j] = dQ.m[i][j].imag() / h;
            }
    }
}

int gxo_static_residual(const gxo_problem* q, const double* x, double* r) {
    Problem P = to_problem(q);
    Indices idx = system_indices(P.ne, P.start, P.stop, true);
    static_system_residual<double>(P, idx, x, r);
    return 0;
}
// dense column-major n x n
int gxo_static_jacobian(const gxo_problem* q, const double* x, double* J) {
    Problem P = to_problem(q);
    Indices idx = system_indices(P.ne, P.start, P.stop, true);
    DenseMat<double> D{idx.nstates, J};
    static_system_jacobian<double>(P, idx, x, D);
    return 0;
}
// complex-step derivative of the residual, dense column-major
int gxo_static_jacobian_cs(const gxo_problem* q, const double* x, double* J) {
    Problem P = to_problem(q);
    Indices idx = system_indices(P.ne, P.start, P.stop, true);
    int64_t n = idx.nstates;
    const double h = 1e-30;
    std::vector<cplx> xc(n), rc(n);
    for (int64_t j = 0; j < n; j++) {
        for (int64_t i = 0; i < n; i++) xc[i] = cplx(x[i], 0);
        xc[j] += cplx(0, h);
        for (int64_t i = 0; i < n; i++) rc[i] = cplx(0, 0);
        static_system_residual<cplx>(P, idx, xc.data(), rc.data());
        for (int64_t i = 0; i < n; i++) J[i + n * j] = rc[i].imag() / h;
    }
    return 0;
}
int gxo_static_bandwidth(const gxo_problem* q, int64_t* kl, int64_t* ku) {
    Problem P = to_problem(q);
    Indices idx = system_indices(P.ne, P.start, P.stop, true);
    static_bandwidth(P, idx, *kl, *ku);
    return 0;
}

static NewtonResult static_solve_one(const Problem& P, const Indices& idx, int64_t kl, int64_t ku, const NewtonOpts& o, const double* x0,
                                     double* x, double* F) {
    int64_t n = idx.nstates;
    BandMat J; J.init(n, kl, ku);
    for (int64_t i = 0; i < n; i++) x[i] = x0 ? x0[i] : 0.0;
    auto rf = [&](const double* xx, double* FF) { static_system_residual<double>(P, idx, xx, FF); };
    auto jf = [&](const double* xx, BandMat& JJ) { static_system_jacobian<double>(P, idx, xx, JJ); };
    return newton_solve(n, x, F, J, rf, jf, o);
}

// static_analysis for one instance (analyses.jl:78-179 with a single time step)
//   stats[4] = {converged, iters, f_calls, ls_backtracks}
int gxo_static_solve(const gxo_problem* q, const gxo_opts* opts, const double* x0, double* x, double* resid, int64_t* stats,
                     double* resid_inf) {
    Problem P = to_problem(q);
    Indices idx = system_indices(P.ne, P.start, P.stop, true);
    int64_t kl, ku; static_bandwidth(P, idx, kl, ku);
    NewtonResult r = static_solve_one(P, idx, kl, ku, to_opts(opts), x0, x, resid);
    stats[0] = r.converged; stats[1] = r.iters; stats[2] = r.f_calls; stats[3] = r.ls_backtracks;
    *resid_inf = r.resid_inf;
    return r.error;
}

// Batched static_analysis, instances independent, a std::thread pool with a dynamic schedule over instances (the analogue of
// Threads.@threads over reference GXBeam instances).  Topology + pd/pl shared; per-instance arrays are contiguous
// [nbatch][...] in the layouts of include/gxbeam_b200.h.  gravity: nbatch x 3 ; force_scaling: nbatch.
int gxo_static_solve_batch(int64_t nbatch, const gxo_problem* q0, const double* gravity, const double* force_scaling, const gxo_opts* opts,
                           const double* x0, double* x, int32_t* converged, int32_t* iters, double* resid_inf, int nthreads) {
    Problem P0 = to_problem(q0);
    Indices idx = system_indices(P0.ne, P0.start, P0.stop, true);
    int64_t kl, ku; static_bandwidth(P0, idx, kl, ku);
    int64_t n = idx.nstates;
    NewtonOpts o = to_opts(opts);
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
      for (;;) {
        int64_t b = next.fetch_add(1);   // dynamic schedule, one instance at a time
        if (b >= nbatch) break;
        Problem P = P0;
        P.elems = P0.elems + b * P0.ne * 91;
        P.points = P0.points ? P0.points + b * P0.np * 3 : nullptr;
        P.bc = P0.bc + b * P0.np * 18;
        P.dload = P0.dload ? P0.dload + b * P0.ne * 24 : nullptr;
        P.pmass = P0.pmass ? P0.pmass + b * P0.np * 36 : nullptr;
        if (gravity) for (int i = 0; i < 3; i++) P.gravity[i] = gravity[3 * b + i];
        if (force_scaling) P.force_scaling = force_scaling[b];
        std::vector<double> F(n);
        NewtonResult r = static_solve_one(P, idx, kl, ku, o, x0 ? x0 + b * n : nullptr, x + b * n, F.data());
        converged[b] = r.converged; iters[b] = (int32_t)r.iters; resid_inf[b] = r.resid_inf;
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    return 0;
}

int gxo_num_threads() { int n = (int)std::thread::hardware_concurrency(); return n > 0 ? n : 1; }

}  // extern "C"

// ---------------------------------------------------------------- steady state (DynamicSystem) ----
static DynParams to_dyn(const gxo_dyn* d) {
    DynParams D;
    if (d) { for (int i = 0; i < 3; i++) { D.vb[i] = d->vb[i]; D.wb[i] = d->wb[i]; D.ab[i] = d->ab[i]; D.alb[i] = d->alb[i]; } D.structural_damping = d->structural_damping; }
    return D;
}
static void dynamic_bandwidth(const Problem& P, const Indices& idx, int64_t& kl, int64_t& ku) {
    kl = 0; ku = 0;
    auto blk = [&](int64_t r0, int64_t nr, int64_t c0, int64_t nc) { kl = std::max(kl, (r0 + nr - 1) - c0); ku = std::max(ku, (c0 + nc - 1) - r0); };
    for (int64_t ip = 0; ip < P.np; ip++) blk(idx.irow_point[ip], 12, idx.icol_point[ip], 12);
    for (int64_t ie = 0; ie < P.ne; ie++) {
        int64_t ic = idx.icol_elem[ie], ic1 = idx.icol_point[P.start[ie] - 1], ic2 = idx.icol_point[P.stop[ie] - 1];
        int64_t ir = idx.irow_elem[ie], ir1 = idx.irow_point[P.start[ie] - 1], ir2 = idx.irow_point[P.stop[ie] - 1];
        blk(ir, 6, ic1, 12); blk(ir, 6, ic2, 12); blk(ir, 6, ic, 6);
        for (int64_t irp : {ir1, ir2}) { blk(irp, 6, ic1, 12); blk(irp, 6, ic2, 12); blk(irp, 6, ic, 6); }
    }
    for (int i = 0; i < 6; i++)
        if (idx.icol_body[i]) for (int64_t ip = 0; ip < P.np; ip++) blk(idx.irow_point[ip], 6, idx.icol_body[i], 1);
}

extern "C" {

int gxo_steady_residual(const gxo_problem* q, const gxo_dyn* d, const double* x, double* r) {
    Problem P = to_problem(q); DynParams D = to_dyn(d);
    Indices idx = system_indices(P.ne, P.start, P.stop, false);
    update_body_acceleration_indices(P, idx);
    static_assert(sizeof(double) == 8, "");
    steady_system_residual<double>(P, idx, D, x, r);
    return 0;
}
int gxo_steady_jacobian(const gxo_problem* q, const gxo_dyn* d, const double* x, double* J, int complex_step) {
    Problem P = to_problem(q); DynParams D = to_dyn(d);
    Indices idx = system_indices(P.ne, P.start, P.stop, false);
    update_body_acceleration_indices(P, idx);
    int64_t n = idx.nstates;
    if (!complex_step) {
        DenseMat<double> M{n, J};
        steady_system_jacobian<double>(P, idx, D, x, M);
        return 0;
    }
    const double h = 1e-30;
    std::vector<cplx> xc(n), rc(n);
    for (int64_t j = 0; j < n; j++) {
        for (int64_t i = 0; i < n; i++) xc[i] = cplx(x[i], 0);
        xc[j] += cplx(0, h);
        for (int64_t i = 0; i < n; i++) rc[i] = cplx(0, 0);
        steady_system_residual<cplx>(P, idx, D, xc.data(), rc.data());
        for (int64_t i = 0; i < n; i++) J[i + n * j] = rc[i].imag() / h;
    }
    return 0;
}
int gxo_steady_bandwidth(const gxo_problem* q, int64_t* kl, int64_t* ku, int64_t* icol_body) {
    Problem P = to_problem(q);
    Indices idx = system_indices(P.ne, P.start, P.stop, false);
    update_body_acceleration_indices(P, idx);
    dynamic_bandwidth(P, idx, *kl, *ku);
    if (icol_body) for (int i = 0; i < 6; i++) icol_body[i] = idx.icol_body[i];
    return 0;
}
static NewtonResult steady_solve_one(const Problem& P, const Indices& idx, const DynParams& D, int64_t kl, int64_t ku, const NewtonOpts& o,
                                     const double* x0, double* x, double* F) {
    int64_t n = idx.nstates;
    BandMat J; J.init(n, kl, ku);
    for (int64_t i = 0; i < n; i++) x[i] = x0 ? x0[i] : 0.0;
    auto rf = [&](const double* xx, double* FF) { steady_system_residual<double>(P, idx, D, xx, FF); };
    auto jf = [&](const double* xx, BandMat& JJ) { steady_system_jacobian<double>(P, idx, D, xx, JJ); };
    return newton_solve(n, x, F, J, rf, jf, o);
}
// steady_state_analysis for one instance (analyses.jl:383-527 with a single time step)
int gxo_steady_solve(const gxo_problem* q, const gxo_dyn* d, const gxo_opts* opts, const double* x0, double* x, double* resid, int64_t* stats,
                     double* resid_inf) {
    Problem P = to_problem(q); DynParams D = to_dyn(d);
    Indices idx = system_indices(P.ne, P.start, P.stop, false);
    update_body_acceleration_indices(P, idx);
    int64_t kl, ku; dynamic_bandwidth(P, idx, kl, ku);
    NewtonResult r = steady_solve_one(P, idx, D, kl, ku, to_opts(opts), x0, x, resid);
    stats[0] = r.converged; stats[1] = r.iters; stats[2] = r.f_calls; stats[3] = r.ls_backtracks;
    *resid_inf = r.resid_inf;
    return r.error;
}
// batched; dyn: nbatch records (vb, wb, ab, alb may differ per instance)
int gxo_steady_solve_batch(int64_t nbatch, const gxo_problem* q0, const gxo_dyn* dyn, const double* gravity, const double* force_scaling,
                           const gxo_opts* opts, const double* x0, double* x, int32_t* converged, int32_t* iters, double* resid_inf, int nthreads) {
    Problem P0 = to_problem(q0);
    Indices idx = system_indices(P0.ne, P0.start, P0.stop, false);
    update_body_acceleration_indices(P0, idx);
    int64_t kl, ku; dynamic_bandwidth(P0, idx, kl, ku);
    int64_t n = idx.nstates;
    NewtonOpts o = to_opts(opts);
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
      for (;;) {
        int64_t b = next.fetch_add(1);
        if (b >= nbatch) break;
        Problem P = P0;
        P.elems = P0.elems + b * P0.ne * 91;
        P.points = P0.points + b * P0.np * 3;
        P.bc = P0.bc + b * P0.np * 18;
        P.dload = P0.dload ? P0.dload + b * P0.ne * 24 : nullptr;
        P.pmass = P0.pmass ? P0.pmass + b * P0.np * 36 : nullptr;
        if (gravity) for (int i = 0; i < 3; i++) P.gravity[i] = gravity[3 * b + i];
        if (force_scaling) P.force_scaling = force_scaling[b];
        DynParams D = to_dyn(dyn + b);
        std::vector<double> F(n);
        NewtonResult r = steady_solve_one(P, idx, D, kl, ku, o, x0 ? x0 + b * n : nullptr, x + b * n, F.data());
        converged[b] = r.converged; iters[b] = (int32_t)r.iters; resid_inf[b] = r.resid_inf;
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------- Newmark time marching ----
extern "C" {

int gxo_newmark_residual(const gxo_problem* q, const gxo_dyn* d, const double* init, double dt, const double* x, double* r) {
    Problem P = to_problem(q); DynParams D = to_dyn(d);
    Indices idx = system_indices(P.ne, P.start, P.stop, false);
    NewmarkParams N; N.init = init; N.dt = dt;
    newmark_system_residual<double>(P, idx, D, N, x, r);
    return 0;
}
int gxo_newmark_jacobian(const gxo_problem* q, const gxo_dyn* d, const double* init, double dt, const double* x, double* J, int complex_step) {
    Problem P = to_problem(q); DynParams D = to_dyn(d);
    Indices idx = system_indices(P.ne, P.start, P.stop, false);
    NewmarkParams N; N.init = init; N.dt = dt;
    int64_t n = idx.nstates;
    if (!complex_step) { DenseMat<double> M{n, J}; newmark_system_jacobian<double>(P, idx, D, N, x, M); return 0; }
    const double h = 1e-30;
    std::vector<cplx> xc(n), rc(n);
    for (int64_t j = 0; j < n; j++) {
        for (int64_t i = 0; i < n; i++) xc[i] = cplx(x[i], 0);
        xc[j] += cplx(0, h);
        for (int64_t i = 0; i < n; i++) rc[i] = cplx(0, 0);
        newmark_system_residual<cplx>(P, idx, D, N, xc.data(), rc.data());
        for (int64_t i = 0; i < n; i++) J[i + n * j] = rc[i].imag() / h;
    }
    return 0;
}

// time_domain_analysis! with a given initial state (src/analyses.jl:2600-2790), one instance.
//   tvec: nt times.  bc_t: nt x np x 18 prescribed values per time (NULL: constant = q->bc).  x0: nstates.  rates0: np x 12
//   (udot, θdot, Vdot, Ωdot of the initial state).  x_hist: nt x nstates (row 0 = x0).  iters: nt (iters[0] = 0).
//   Returns the number of time levels completed (stops early when a step does not converge).
int64_t gxo_newmark_run2(const gxo_problem* q, const gxo_dyn* d, const gxo_opts* opts, int64_t nt, const double* tvec, const double* bc_t,
                         const double* x0, const double* rates0, double* x_hist, int32_t* iters, int32_t* converged, int update_linearization);
int64_t gxo_newmark_run(const gxo_problem* q, const gxo_dyn* d, const gxo_opts* opts, int64_t nt, const double* tvec, const double* bc_t,
                        const double* x0, const double* rates0, double* x_hist, int32_t* iters, int32_t* converged) {
    return gxo_newmark_run2(q, d, opts, nt, tvec, bc_t, x0, rates0, x_hist, iters, converged, 1);
}
// update_linearization only matters with opts->linear: false = every step is linearised about the INITIAL state x0 (analyses.jl:2741-2755)
int64_t gxo_newmark_run2(const gxo_problem* q, const gxo_dyn* d, const gxo_opts* opts, int64_t nt, const double* tvec, const double* bc_t,
                         const double* x0, const double* rates0, double* x_hist, int32_t* iters, int32_t* converged, int update_linearization) {
    Problem P = to_problem(q); DynParams D = to_dyn(d);
    Indices idx = system_indices(P.ne, P.start, P.stop, false);
    int64_t kl, ku; dynamic_bandwidth(P, idx, kl, ku);
    int64_t n = idx.nstates;
    NewtonOpts o = to_opts(opts);
    std::vector<double> x(x0, x0 + n), F(n), paug(12 * P.np, 0.0);
    std::copy(x.begin(), x.end(), x_hist);
    iters[0] = 0; *converged = 1;
    BandMat J; J.init(n, kl, ku);
    int64_t done = 1;
    for (int64_t it = 1; it < nt; it++) {
        double dt = tvec[it] - tvec[it - 1];
        if (bc_t) P.bc = bc_t + it * P.np * 18;
        for (int64_t ip = 0; ip < P.np; ip++) {
            V3<double> u, th;
            int64_t icol = idx.icol_point[ip] - 1;
            point_displacement(P, x.data(), ip, icol, u, th);
            V3<double> V = ld3(x.data(), icol + 6), Om = ld3(x.data(), icol + 9);
            double* pa = &paug[12 * ip];
            V3<double> r[4];
            if (it <= 1) { for (int k = 0; k < 4; k++) r[k] = cvec<double>(rates0 + 12 * ip + 3 * k); }
            else {
                double c = 2.0 / (tvec[it - 1] - tvec[it - 2]);
                r[0] = c * u - cvec<double>(pa); r[1] = c * th - cvec<double>(pa + 3); r[2] = c * V - cvec<double>(pa + 6); r[3] = c * Om - cvec<double>(pa + 9);
            }
            double c = 2.0 / dt;
            V3<double> s[4] = {c * u + r[0], c * th + r[1], c * V + r[2], c * Om + r[3]};
            for (int k = 0; k < 4; k++) for (int i = 0; i < 3; i++) pa[3 * k + i] = s[k][i];
        }
        NewmarkParams N; N.init = paug.data(); N.dt = dt;
        auto rf = [&](const double* xx, double* FF) { newmark_system_residual<double>(P, idx, D, N, xx, FF); };
        auto jf = [&](const double* xx, BandMat& JJ) { newmark_system_jacobian<double>(P, idx, D, N, xx, JJ); };
        if (o.linear && !update_linearization) std::copy(x0, x0 + n, x.begin());     // x = newmark_lsolve!(x0, paug, constants)
        NewtonResult res = newton_solve(n, x.data(), F.data(), J, rf, jf, o);
        std::copy(x.begin(), x.end(), x_hist + it * n);
        iters[it] = (int32_t)res.iters;
        done = it + 1;
        if (!res.converged) { *converged = 0; break; }
    }
    return done;
}

}  // extern "C"

// ---------------------------------------------------------------- mass matrix ----
extern "C" {
// system_mass_matrix! (src/system.jl:961-995) as a dense column-major n x n matrix
int gxo_mass_matrix(const gxo_problem* q, const double* x, double* M) {
    Problem P = to_problem(q);
    Indices idx = system_indices(P.ne, P.start, P.stop, false);
    int64_t n = idx.nstates;
    DenseMat<double> D{n, M};
    D.zero();
    system_mass_matrix<double>(P, idx, x, 1.0, D);
    return 0;
}
}  // extern "C"

// ---------------------------------------------------------------- initial condition analysis ----
extern "C" {
struct gxo_initial {   // mirrors gxo::InitialParams
    const uint8_t* rv1;
    const uint8_t* rv2;
    const double* u0; const double* th0; const double* V0; const double* Om0; const double* Vdot0; const double* Omdot0;
};
}
static InitialParams to_initial(const gxo_initial* c) {
    InitialParams I;
    I.rv1 = c->rv1; I.rv2 = c->rv2; I.u0 = c->u0; I.th0 = c->th0; I.V0 = c->V0; I.Om0 = c->Om0; I.Vdot0 = c->Vdot0; I.Omdot0 = c->Omdot0;
    return I;
}
extern "C" {
int gxo_initial_rate_vars(const gxo_problem* q, uint8_t* rv1, uint8_t* rv2) {
    Problem P = to_problem(q);
    Indices idx = system_indices(P.ne, P.start, P.stop, false);
    initial_rate_vars(P, idx, rv1, rv2);
    return 0;
}
int gxo_initial_residual(const gxo_problem* q, const gxo_dyn* d, const gxo_initial* c, const double* x, double* r) {
    Problem P = to_problem(q); DynParams D = to_dyn(d); InitialParams I = to_initial(c);
    Indices idx = system_indices(P.ne, P.start, P.stop, false);
    update_body_acceleration_indices(P, idx);
    initial_system_residual<double>(P, idx, D, I, x, r);
    return 0;
}
// dense column-major Jacobian; compressed = 1 uses the banded column compression the solver uses, 0 perturbs one column at a time
int gxo_initial_jacobian(const gxo_problem* q, const gxo_dyn* d, const gxo_initial* c, const double* x, double* J, int compressed) {
    Problem P = to_problem(q); DynParams D = to_dyn(d); InitialParams I = to_initial(c);
    Indices idx = system_indices(P.ne, P.start, P.stop, false);
    update_body_acceleration_indices(P, idx);
    int64_t n = idx.nstates;
    DenseMat<double> M{n, J};
    if (compressed) {
        int64_t kl, ku; dynamic_bandwidth(P, idx, kl, ku);
        initial_system_jacobian(P, idx, D, I, x, kl, ku, M);
    } else {
        initial_system_jacobian(P, idx, D, I, x, n, n, M);
    }
    return 0;
}
// initial_condition_analysis! (src/analyses.jl:1695-1945) for one instance.  c->rv1 / c->rv2 are ignored; the rate_vars
// are computed here and returned in rv1 / rv2 (nstates each).  x0: initial-condition-layout guess or NULL (reset_state).
// Outputs: x_ic (the solved initial-condition vector), xx (DynamicSystem state vector), rates (np x 12).
int gxo_initial_condition(const gxo_problem* q, const gxo_dyn* d, const gxo_opts* opts, const gxo_initial* c, const double* x0, double* x_ic,
                          double* xx, double* rates, uint8_t* rv1, uint8_t* rv2, int64_t* stats, double* resid_inf) {
    Problem P = to_problem(q); DynParams D = to_dyn(d); InitialParams I = to_initial(c);
    Indices idx = system_indices(P.ne, P.start, P.stop, false);
    update_body_acceleration_indices(P, idx);
    int64_t kl, ku; dynamic_bandwidth(P, idx, kl, ku);
    int64_t n = idx.nstates;
    initial_rate_vars(P, idx, rv1, rv2);
    I.rv1 = rv1; I.rv2 = rv2;
    BandMat J; J.init(n, kl, ku);
    std::vector<double> F(n);
    for (int64_t i = 0; i < n; i++) x_ic[i] = x0 ? x0[i] : 0.0;
    auto rf = [&](const double* xv, double* FF) { initial_system_residual<double>(P, idx, D, I, xv, FF); };
    auto jf = [&](const double* xv, BandMat& JJ) { initial_system_jacobian(P, idx, D, I, xv, kl, ku, JJ); };
    NewtonResult r = newton_solve(n, x_ic, F.data(), J, rf, jf, to_opts(opts));
    initial_output(P, idx, D, I, x_ic, xx, rates);
    stats[0] = r.converged; stats[1] = r.iters; stats[2] = r.f_calls; stats[3] = r.ls_backtracks;
    *resid_inf = r.resid_inf;
    return r.error;
}
}  // extern "C"
