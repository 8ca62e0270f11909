// ORACLE — TEST INFRASTRUCTURE ONLY (see gxb_math.hpp).
//
// CPU restatement of GXBeam.jl's steady-state residual / Jacobian assembly on a DynamicSystem (12 states per point):
//   update_body_acceleration_indices!, body_accelerations     src/system.jl:138-182
//   point_velocities                                          src/point.jl:228-234
//   steady_point_properties / _jacobian_properties            src/point.jl:630-672, 964-1023
//   dynamic/steady_point_resultants (+ jacobians)             src/point.jl:1438-1449, 1495-1540
//   point_velocity_residuals / dynamic_point_velocity_jacobians  src/point.jl:1658-1666, 1728-1739
//   insert_dynamic/steady_point_jacobians!                    src/point.jl:1907-1935, 1809-1844
//   steady_point_residual! / _jacobian!                       src/point.jl:2137-2158, 2346-2367
//   steady_element_properties / _jacobian_properties          src/element.jl:117-217, 649-840   (incl. structural damping)
//   dynamic_compatibility_jacobians                           src/element.jl:1550-1578
//   dynamic_element_resultants (+ jacobians)                  src/element.jl:1901-1919, 2035-2064, 2219-2328
//   insert_dynamic/steady_element_jacobians!                  src/element.jl:2685-2759, 2554-2583
//   steady_element_residual! / _jacobian!                     src/element.jl:3019-3035, 3209-3229
//   steady_system_residual! / _jacobian!                      src/system.jl:427-455, 693-720
#pragma once
#include <complex>
#include "gxb_static.hpp"

namespace gxo {

struct DynParams {
    double vb[3] = {0, 0, 0}, wb[3] = {0, 0, 0}, ab[3] = {0, 0, 0}, alb[3] = {0, 0, 0};
    int structural_damping = 0;
};

// system.jl:138-150 (first point, in ascending index order, that has pd & pl on DOF i)
inline void update_body_acceleration_indices(const Problem& P, Indices& idx) {
    for (int i = 0; i < 6; i++) {
        idx.icol_body[i] = 0;
        for (int64_t ip = 0; ip < P.np; ip++)
            if (P.pd[6 * ip + i] && P.pl[6 * ip + i]) { idx.icol_body[i] = idx.icol_point[ip] + i; break; }
    }
}
// system.jl:167-182
template <class T> void body_accelerations(const Indices& idx, const T* x, const DynParams& D, V3<T>& ab, V3<T>& alb) {
    for (int i = 0; i < 3; i++) {
        ab.v[i] = idx.icol_body[i] ? x[idx.icol_body[i] - 1] : T(D.ab[i]);
        alb.v[i] = idx.icol_body[3 + i] ? x[idx.icol_body[3 + i] - 1] : T(D.alb[i]);
    }
}

template <class T> struct DynPoint : PointProps<T> {
    M3<T> Qinv, Qi_t1, Qi_t2, Qi_t3;
    V3<T> dx, vb, wb, ab, alb, V, Om, Pm, H, udot, thdot, Vdot, Omdot, Pdot, Hdot;
    M3<T> P_th, P_V, P_Om, H_th, H_V, H_Om;
    M3<T> Pdot_ab, Pdot_alb, Pdot_u, Pdot_th, Pdot_V, Pdot_Om, Hdot_ab, Hdot_alb, Hdot_u, Hdot_th, Hdot_V, Hdot_Om;
};

// point.jl:630-672
template <class T> DynPoint<T> steady_point_properties(const Problem& P, const Indices& idx, const T* x, int64_t ip, const DynParams& D,
                                                       const V3<T>& ab, const V3<T>& alb) {
    DynPoint<T> p;
    static_cast<PointProps<T>&>(p) = static_point_properties(P, idx, x, ip);
    p.Qinv = get_Qinv(p.th);
    p.dx = cvec<T>(P.points + 3 * ip);
    p.vb = cvec<T>(D.vb); p.wb = cvec<T>(D.wb); p.ab = ab; p.alb = alb;
    int64_t icol = idx.icol_point[ip] - 1;
    p.V = ld3(x, icol + 6); p.Om = ld3(x, icol + 9);
    M3<T> Ct = tr(p.C);
    p.Pm = Ct * (p.mass11 * (p.C * p.V)) + Ct * (p.mass12 * (p.C * p.Om));
    p.H = Ct * (p.mass21 * (p.C * p.V)) + Ct * (p.mass22 * (p.C * p.Om));
    p.udot = zero3<T>(); p.thdot = zero3<T>();
    p.Vdot = ab + cross(alb, p.dx + p.u) + cross(p.wb, p.udot);
    p.Omdot = alb;
    p.Pdot = Ct * (p.mass11 * (p.C * p.Vdot)) + Ct * (p.mass12 * (p.C * p.Omdot));
    p.Hdot = Ct * (p.mass21 * (p.C * p.Vdot)) + Ct * (p.mass22 * (p.C * p.Omdot));
    return p;
}
// point.jl:964-1023
template <class T> void steady_point_jacobian_properties(const Problem& P, const Indices& idx, const T* x, int64_t ip, DynPoint<T>& p) {
    static_point_jacobian_properties(P, idx, x, ip, p);
    get_Qinv_th(p.th, p.Qi_t1, p.Qi_t2, p.Qi_t3);
    M3<T> Ct = tr(p.C), Ct1 = tr(p.C_t1), Ct2 = tr(p.C_t2), Ct3 = tr(p.C_t3);
    auto m3c = [&](const V3<T>& w) { return mul3(p.C_t1, p.C_t2, p.C_t3, w); };
    auto m3t = [&](const V3<T>& w) { return mul3(Ct1, Ct2, Ct3, w); };
    p.P_th = m3t(p.mass11 * (p.C * p.V) + p.mass12 * (p.C * p.Om)) + Ct * (p.mass11 * m3c(p.V)) + Ct * (p.mass12 * m3c(p.Om));
    p.P_V = Ct * p.mass11 * p.C; p.P_Om = Ct * p.mass12 * p.C;
    p.H_th = m3t(p.mass21 * (p.C * p.V) + p.mass22 * (p.C * p.Om)) + Ct * (p.mass21 * m3c(p.V)) + Ct * (p.mass22 * m3c(p.Om));
    p.H_V = Ct * p.mass21 * p.C; p.H_Om = Ct * p.mass22 * p.C;
    M3<T> Vdot_alb = -tilde(p.dx + p.u), Vdot_u = tilde(p.alb);
    M3<T> Pdot_Vdot = Ct * p.mass11 * p.C, Pdot_Omdot = Ct * p.mass12 * p.C, Hdot_Vdot = Ct * p.mass21 * p.C, Hdot_Omdot = Ct * p.mass22 * p.C;
    p.Pdot_ab = Pdot_Vdot; p.Pdot_alb = Pdot_Vdot * Vdot_alb + Pdot_Omdot;
    p.Hdot_ab = Hdot_Vdot; p.Hdot_alb = Hdot_Vdot * Vdot_alb + Hdot_Omdot;
    p.Pdot_u = Pdot_Vdot * Vdot_u; p.Hdot_u = Hdot_Vdot * Vdot_u;
    p.Pdot_th = m3t(p.mass11 * (p.C * p.Vdot)) + Ct * (p.mass11 * m3c(p.Vdot)) + m3t(p.mass12 * (p.C * p.Omdot)) + Ct * (p.mass12 * m3c(p.Omdot));
    p.Hdot_th = m3t(p.mass21 * (p.C * p.Vdot)) + Ct * (p.mass21 * m3c(p.Vdot)) + m3t(p.mass22 * (p.C * p.Omdot)) + Ct * (p.mass22 * m3c(p.Omdot));
    p.Pdot_V = zero33<T>(); p.Hdot_V = zero33<T>(); p.Pdot_Om = zero33<T>(); p.Hdot_Om = zero33<T>();
}
// point.jl:1438-1449
template <class T> void dynamic_point_resultants(const DynPoint<T>& p, V3<T>& F, V3<T>& M) {
    static_point_resultants(p, F, M);
    F = F - (cross(p.wb, p.Pm) + p.Pdot);
    M = M - (cross(p.wb, p.H) + cross(p.V, p.Pm) + p.Hdot);
}
template <class T> struct PointResJac { M3<T> F_th, M_th, F_u, F_V, F_Om, M_u, M_V, M_Om, F_ab, F_alb, M_ab, M_alb; };
// point.jl:1495-1540
template <class T> PointResJac<T> steady_point_resultant_jacobians(const DynPoint<T>& p) {
    PointResJac<T> j;
    static_point_resultant_jacobians(p, j.F_th, j.M_th);
    M3<T> wt = tilde(p.wb), Vt = tilde(p.V);
    j.F_u = -p.Pdot_u;
    j.F_th = j.F_th - (wt * p.P_th + p.Pdot_th);
    j.F_V = -(wt * p.P_V) - p.Pdot_V;
    j.F_Om = -(wt * p.P_Om) - p.Pdot_Om;
    j.M_u = -p.Hdot_u;
    j.M_th = j.M_th - (wt * p.H_th + Vt * p.P_th + p.Hdot_th);
    j.M_V = -(wt * p.H_V) - Vt * p.P_V + tilde(p.Pm) - p.Hdot_V;
    j.M_Om = -(wt * p.H_Om) - Vt * p.P_Om - p.Hdot_Om;
    j.F_ab = -p.Pdot_ab; j.F_alb = -p.Pdot_alb; j.M_ab = -p.Hdot_ab; j.M_alb = -p.Hdot_alb;
    return j;
}
// point.jl:1658-1666
template <class T> void point_velocity_residuals(const DynPoint<T>& p, V3<T>& rV, V3<T>& rOm) {
    rV = p.V - p.vb - cross(p.wb, p.dx) - cross(p.wb, p.u) - p.udot;
    rOm = p.Qinv * (p.C * (p.Om - p.wb)) - p.thdot;
}
// point.jl:2137-2158
template <class T> void steady_point_residual(const Problem& P, const Indices& idx, const T* x, int64_t ip, const DynParams& D, const V3<T>& ab,
                                              const V3<T>& alb, T* resid) {
    int64_t irow = idx.irow_point[ip] - 1;
    DynPoint<T> p = steady_point_properties(P, idx, x, ip, D, ab, alb);
    V3<T> F, M, rV, rOm;
    dynamic_point_resultants(p, F, M);
    point_velocity_residuals(p, rV, rOm);
    T fs = T(P.force_scaling);
    for (int i = 0; i < 3; i++) { resid[irow + i] = -F[i] / fs; resid[irow + 3 + i] = -M[i] / fs; resid[irow + 6 + i] = rV[i]; resid[irow + 9 + i] = rOm[i]; }
}
template <class J, class T> void setcol3(J& jac, int64_t r0, int64_t c, const V3<T>& v) { for (int i = 0; i < 3; i++) jac.set(r0 + i, c, v[i]); }
template <class J, class T> void addcol3(J& jac, int64_t r0, int64_t c, const V3<T>& v) { for (int i = 0; i < 3; i++) jac.add(r0 + i, c, v[i]); }
// point.jl:2346-2367, 1785-1800, 1907-1935, 1809-1844
template <class T, class J> void steady_point_jacobian(const Problem& P, const Indices& idx, const T* x, int64_t ip, const DynParams& D, const V3<T>& ab,
                                                       const V3<T>& alb, J& jac) {
    DynPoint<T> p = steady_point_properties(P, idx, x, ip, D, ab, alb);
    steady_point_jacobian_properties(P, idx, x, ip, p);
    PointResJac<T> r = steady_point_resultant_jacobians(p);
    // dynamic_point_velocity_jacobians (point.jl:1728-1739)
    M3<T> rV_u = -tilde(p.wb), rV_V = eye3<T>();
    M3<T> rOm_th = mul3(p.Qi_t1, p.Qi_t2, p.Qi_t3, p.C * (p.Om - p.wb)) + p.Qinv * mul3(p.C_t1, p.C_t2, p.C_t3, p.Om - p.wb);
    M3<T> rOm_Om = p.Qinv * p.C;
    int64_t irow = idx.irow_point[ip] - 1, icol = idx.icol_point[ip] - 1;
    T fs = T(P.force_scaling);
    // insert_static_point_jacobians!
    set33(jac, irow, icol, -p.F_F);
    set33(jac, irow, icol + 3, (-(r.F_th * p.th_th)) / fs);
    set33(jac, irow + 3, icol + 3, -p.M_M - (r.M_th * p.th_th) / fs);
    // insert_dynamic_point_jacobians!
    set33(jac, irow, icol + 6, (-r.F_V) / fs);
    set33(jac, irow, icol + 9, (-r.F_Om) / fs);
    set33(jac, irow + 3, icol + 6, (-r.M_V) / fs);
    set33(jac, irow + 3, icol + 9, (-r.M_Om) / fs);
    set33(jac, irow + 6, icol, rV_u * p.u_u);
    set33(jac, irow + 6, icol + 6, rV_V);
    set33(jac, irow + 9, icol + 3, rOm_th * p.th_th);
    set33(jac, irow + 9, icol + 9, rOm_Om);
    // insert_steady_point_jacobians!
    sub33m(jac, irow, icol, (r.F_u * p.u_u) / fs);
    sub33m(jac, irow + 3, icol, (r.M_u * p.u_u) / fs);
    for (int i = 0; i < 3; i++) {
        if (idx.icol_body[i]) { setcol3(jac, irow, idx.icol_body[i] - 1, (-col(r.F_ab, i)) / fs); setcol3(jac, irow + 3, idx.icol_body[i] - 1, (-col(r.M_ab, i)) / fs); }
        if (idx.icol_body[3 + i]) { setcol3(jac, irow, idx.icol_body[3 + i] - 1, (-col(r.F_alb, i)) / fs); setcol3(jac, irow + 3, idx.icol_body[3 + i] - 1, (-col(r.M_alb, i)) / fs); }
    }
}

// ---------------------------------------------------------------- elements ----
template <class T> struct DynElem : ElemProps<T> {
    M3<T> Q;
    V3<T> dx, vb, wb, ab, alb, V1, V2, Om1, Om2, V, Om, Pm, H, udot, thdot, Vdot, Omdot, Pdot, Hdot;
    // structural damping
    int damping = 0;
    M3<T> mu11, mu22, C1, C2, Qinv1, Qinv2, dQ;
    V3<T> dx1, dx2, udot1, udot2, uedot, thdot1, thdot2, thedot, du, dth, dudot, dthdot, gamdot, kapdot;
    // jacobian properties
    M3<T> gam_u1, gam_u2, gam_th1, gam_th2, gam_V1, gam_V2, gam_Om1, gam_Om2, kap_th1, kap_th2, kap_Om1, kap_Om2;
    M3<T> P_th, P_V, P_Om, H_th, H_V, H_Om;
    M3<T> Pdot_ab, Pdot_alb, Pdot_u, Pdot_th, Pdot_V, Pdot_Om, Hdot_ab, Hdot_alb, Hdot_u, Hdot_th, Hdot_V, Hdot_Om;
};
// element.jl:117-217
template <class T> DynElem<T> steady_element_properties(const Problem& P, const Indices& idx, const T* x, int64_t ie, const DynParams& D,
                                                        const V3<T>& ab, const V3<T>& alb) {
    DynElem<T> p;
    static_cast<ElemProps<T>&>(p) = static_element_properties(P, idx, x, ie);
    const double* e = P.elems + 91 * ie;
    V3<T> e1 = unit<T>(0);
    p.Q = get_Q(p.th);
    p.dx = cvec<T>(e + 1);
    p.vb = cvec<T>(D.vb); p.wb = cvec<T>(D.wb); p.ab = ab; p.alb = alb;
    int64_t ia = P.start[ie] - 1, ib = P.stop[ie] - 1;
    int64_t ic1 = idx.icol_point[ia] - 1, ic2 = idx.icol_point[ib] - 1;
    p.V1 = ld3(x, ic1 + 6); p.Om1 = ld3(x, ic1 + 9); p.V2 = ld3(x, ic2 + 6); p.Om2 = ld3(x, ic2 + 9);
    p.V = (p.V1 + p.V2) / T(2); p.Om = (p.Om1 + p.Om2) / T(2);
    M3<T> CtCabT = tr(p.CtCab);
    p.Pm = p.CtCab * (p.mass11 * (CtCabT * p.V)) + p.CtCab * (p.mass12 * (CtCabT * p.Om));
    p.H = p.CtCab * (p.mass21 * (CtCabT * p.V)) + p.CtCab * (p.mass22 * (CtCabT * p.Om));
    p.udot = zero3<T>(); p.thdot = zero3<T>();
    p.Vdot = ab + cross(alb, p.dx + p.u) + cross(p.wb, p.udot);
    p.Omdot = alb;
    p.Pdot = p.CtCab * (p.mass11 * (CtCabT * p.Vdot)) + p.CtCab * (p.mass12 * (CtCabT * p.Omdot));
    p.Hdot = p.CtCab * (p.mass21 * (CtCabT * p.Vdot)) + p.CtCab * (p.mass22 * (CtCabT * p.Omdot));
    p.damping = D.structural_damping;
    if (p.damping) {
        const double* mu = e + 85;
        p.mu11 = zero33<T>(); p.mu22 = zero33<T>();
        for (int i = 0; i < 3; i++) { p.mu11.m[i][i] = T(mu[i]); p.mu22.m[i][i] = T(mu[3 + i]); }
        p.C1 = get_C(p.th1); p.C2 = get_C(p.th2); p.Qinv1 = get_Qinv(p.th1); p.Qinv2 = get_Qinv(p.th2);
        p.dx1 = cvec<T>(P.points + 3 * ia); p.dx2 = cvec<T>(P.points + 3 * ib);
        p.udot1 = p.V1 - p.vb - cross(p.wb, p.dx1 + p.u1);
        p.udot2 = p.V2 - p.vb - cross(p.wb, p.dx2 + p.u2);
        p.uedot = (p.udot1 + p.udot2) / T(2);
        p.thdot1 = p.Qinv1 * (p.C1 * (p.Om1 - p.wb));
        p.thdot2 = p.Qinv2 * (p.C2 * (p.Om2 - p.wb));
        p.thedot = (p.thdot1 + p.thdot2) / T(2);
        p.du = p.u2 - p.u1; p.dth = p.th2 - p.th1;
        p.dudot = p.udot2 - p.udot1; p.dthdot = p.thdot2 - p.thdot1;
        p.dQ = get_dQ(p.th, p.dth, p.Q);
        M3<T> Wt = tilde(p.Om - p.wb);
        p.gamdot = -(CtCabT * (Wt * p.du)) + CtCabT * p.dudot - p.L * (CtCabT * (Wt * (p.Cab * e1)));
        M3<T> CabT = tr(p.Cab);
        p.kapdot = CabT * (p.Q * p.dthdot) + CabT * (p.dQ * p.thedot);
        p.gam = p.gam - p.mu11 * p.gamdot;
        p.kap = p.kap - p.mu22 * p.kapdot;
    }
    return p;
}
// element.jl:649-840
template <class T> void steady_element_jacobian_properties(const Problem& P, int64_t ie, DynElem<T>& p) {
    static_element_jacobian_properties(P, ie, p);
    V3<T> e1 = unit<T>(0);
    M3<T> Q_t1, Q_t2, Q_t3;
    get_Q_th(p.Q, p.th, Q_t1, Q_t2, Q_t3);
    M3<T> Ct1 = tr(p.C_t1), Ct2 = tr(p.C_t2), Ct3 = tr(p.C_t3), CtCabT = tr(p.CtCab), CabT = tr(p.Cab);
    auto m3c = [&](const V3<T>& w) { return mul3(p.C_t1, p.C_t2, p.C_t3, w); };
    auto m3t = [&](const V3<T>& w) { return mul3(Ct1, Ct2, Ct3, w); };
    p.P_th = m3t(p.Cab * (p.mass11 * (CtCabT * p.V) + p.mass12 * (CtCabT * p.Om))) + p.CtCab * (p.mass11 * (CabT * m3c(p.V))) +
             p.CtCab * (p.mass12 * (CabT * m3c(p.Om)));
    p.P_V = p.CtCab * p.mass11 * CtCabT; p.P_Om = p.CtCab * p.mass12 * CtCabT;
    p.H_th = m3t(p.Cab * (p.mass21 * (CtCabT * p.V) + p.mass22 * (CtCabT * p.Om))) + p.CtCab * (p.mass21 * (CabT * m3c(p.V))) +
             p.CtCab * (p.mass22 * (CabT * m3c(p.Om)));
    p.H_V = p.CtCab * p.mass21 * CtCabT; p.H_Om = p.CtCab * p.mass22 * CtCabT;
    M3<T> Vdot_alb = -tilde(p.dx + p.u), Vdot_u = tilde(p.alb);
    M3<T> Pdot_Vdot = p.P_V, Pdot_Omdot = p.P_Om, Hdot_Vdot = p.H_V, Hdot_Omdot = p.H_Om;
    p.Pdot_ab = Pdot_Vdot; p.Pdot_alb = Pdot_Vdot * Vdot_alb + Pdot_Omdot;
    p.Hdot_ab = Hdot_Vdot; p.Hdot_alb = Hdot_Vdot * Vdot_alb + Hdot_Omdot;
    p.Pdot_u = Pdot_Vdot * Vdot_u; p.Hdot_u = Hdot_Vdot * Vdot_u;
    p.Pdot_th = m3t(p.Cab * (p.mass11 * (CtCabT * p.Vdot))) + m3t(p.Cab * (p.mass12 * (CtCabT * p.Omdot))) + p.CtCab * (p.mass11 * (CabT * m3c(p.Vdot))) +
                p.CtCab * (p.mass12 * (CabT * m3c(p.Omdot)));
    p.Hdot_th = m3t(p.Cab * (p.mass21 * (CtCabT * p.Vdot))) + m3t(p.Cab * (p.mass22 * (CtCabT * p.Omdot))) + p.CtCab * (p.mass21 * (CabT * m3c(p.Vdot))) +
                p.CtCab * (p.mass22 * (CabT * m3c(p.Omdot)));
    p.Pdot_V = zero33<T>(); p.Pdot_Om = zero33<T>(); p.Hdot_V = zero33<T>(); p.Hdot_Om = zero33<T>();
    M3<T> Z = zero33<T>();
    p.gam_u1 = p.gam_u2 = p.gam_th1 = p.gam_th2 = p.gam_V1 = p.gam_V2 = p.gam_Om1 = p.gam_Om2 = Z;
    p.kap_th1 = p.kap_th2 = p.kap_Om1 = p.kap_Om2 = Z;
    if (p.damping) {
        M3<T> I = eye3<T>();
        M3<T> C1_1, C1_2, C1_3, C2_1, C2_2, C2_3, Qi1_1, Qi1_2, Qi1_3, Qi2_1, Qi2_2, Qi2_3;
        get_C_th(p.C1, p.th1, C1_1, C1_2, C1_3); get_C_th(p.C2, p.th2, C2_1, C2_2, C2_3);
        get_Qinv_th(p.th1, Qi1_1, Qi1_2, Qi1_3); get_Qinv_th(p.th2, Qi2_1, Qi2_2, Qi2_3);
        M3<T> udot1_u1 = -tilde(p.wb), udot2_u2 = -tilde(p.wb), udot1_V1 = I, udot2_V2 = I;
        M3<T> thdot1_th1 = mul3(Qi1_1, Qi1_2, Qi1_3, p.C1 * (p.Om1 - p.wb)) + p.Qinv1 * mul3(C1_1, C1_2, C1_3, p.Om1 - p.wb);
        M3<T> thdot2_th2 = mul3(Qi2_1, Qi2_2, Qi2_3, p.C2 * (p.Om2 - p.wb)) + p.Qinv2 * mul3(C2_1, C2_2, C2_3, p.Om2 - p.wb);
        M3<T> thdot1_Om1 = p.Qinv1 * p.C1, thdot2_Om2 = p.Qinv2 * p.C2;
        M3<T> thdot_th1 = thdot1_th1 / T(2), thdot_th2 = thdot2_th2 / T(2), thdot_Om1 = thdot1_Om1 / T(2), thdot_Om2 = thdot2_Om2 / T(2);
        M3<T> du_u1 = -I, du_u2 = I, dth_th1 = -I, dth_th2 = I;
        M3<T> dudot_u1 = -udot1_u1, dudot_u2 = udot2_u2, dudot_V1 = -udot1_V1, dudot_V2 = udot2_V2;
        M3<T> dthdot_th1 = -thdot1_th1, dthdot_th2 = thdot2_th2, dthdot_Om1 = -thdot1_Om1, dthdot_Om2 = thdot2_Om2;
        M3<T> dQ_t1, dQ_t2, dQ_t3;
        get_dQ_th(p.th, p.dth, p.Q, Q_t1, Q_t2, Q_t3, dQ_t1, dQ_t2, dQ_t3);
        M3<T> dQ_d1 = mul3(Q_t1, Q_t2, Q_t3, unit<T>(0)), dQ_d2 = mul3(Q_t1, Q_t2, Q_t3, unit<T>(1)), dQ_d3 = mul3(Q_t1, Q_t2, Q_t3, unit<T>(2));
        M3<T> Wt = tilde(p.Om - p.wb);
        M3<T> tmp = CtCabT * Wt;
        M3<T> gamdot_u1 = -(tmp * du_u1) + CtCabT * dudot_u1;
        M3<T> gamdot_u2 = -(tmp * du_u2) + CtCabT * dudot_u2;
        tmp = -(CabT * m3c(Wt * p.du)) + CabT * m3c(p.dudot) - p.L * (CabT * m3c(Wt * (p.Cab * e1)));
        M3<T> gamdot_th1 = T(0.5) * tmp, gamdot_th2 = T(0.5) * tmp;
        M3<T> gamdot_V1 = CtCabT * dudot_V1, gamdot_V2 = CtCabT * dudot_V2;
        tmp = CtCabT * tilde(p.du) + p.L * (CtCabT * tilde(p.Cab * e1));
        M3<T> gamdot_Om1 = T(0.5) * tmp, gamdot_Om2 = T(0.5) * tmp;
        M3<T> tmp1 = CabT * mul3(Q_t1, Q_t2, Q_t3, p.dthdot);
        M3<T> tmp2 = CabT * mul3(dQ_t1, dQ_t2, dQ_t3, p.thedot);
        M3<T> tmp3 = CabT * mul3(dQ_d1, dQ_d2, dQ_d3, p.thedot);
        M3<T> CQ = CabT * p.Q, CdQ = CabT * p.dQ;
        M3<T> kapdot_th1 = T(0.5) * tmp1 + T(0.5) * tmp2 + tmp3 * dth_th1 + CQ * dthdot_th1 + CdQ * thdot_th1;
        M3<T> kapdot_th2 = T(0.5) * tmp1 + T(0.5) * tmp2 + tmp3 * dth_th2 + CQ * dthdot_th2 + CdQ * thdot_th2;
        M3<T> kapdot_Om1 = CQ * dthdot_Om1 + CdQ * thdot_Om1;
        M3<T> kapdot_Om2 = CQ * dthdot_Om2 + CdQ * thdot_Om2;
        p.gam_u1 = -(p.mu11 * gamdot_u1); p.gam_u2 = -(p.mu11 * gamdot_u2);
        p.gam_th1 = -(p.mu11 * gamdot_th1); p.gam_th2 = -(p.mu11 * gamdot_th2);
        p.gam_V1 = -(p.mu11 * gamdot_V1); p.gam_V2 = -(p.mu11 * gamdot_V2);
        p.gam_Om1 = -(p.mu11 * gamdot_Om1); p.gam_Om2 = -(p.mu11 * gamdot_Om2);
        p.kap_th1 = -(p.mu22 * kapdot_th1); p.kap_th2 = -(p.mu22 * kapdot_th2);
        p.kap_Om1 = -(p.mu22 * kapdot_Om1); p.kap_Om2 = -(p.mu22 * kapdot_Om2);
    }
}
template <class T> struct DynCompatJac : CompatJac<T> { M3<T> ru_V1, ru_V2, ru_Om1, ru_Om2, rth_Om1, rth_Om2; };
// element.jl:1550-1578
template <class T> DynCompatJac<T> dynamic_compatibility_jacobians(const DynElem<T>& p) {
    DynCompatJac<T> c;
    static_cast<CompatJac<T>&>(c) = static_compatibility_jacobians(p);
    M3<T> QC = p.Qinv * p.Cab;
    c.ru_u1 = c.ru_u1 - p.CtCab * p.gam_u1; c.ru_u2 = c.ru_u2 - p.CtCab * p.gam_u2;
    c.ru_th1 = c.ru_th1 - p.CtCab * p.gam_th1; c.ru_th2 = c.ru_th2 - p.CtCab * p.gam_th2;
    c.ru_V1 = -(p.CtCab * p.gam_V1); c.ru_V2 = -(p.CtCab * p.gam_V2);
    c.ru_Om1 = -(p.CtCab * p.gam_Om1); c.ru_Om2 = -(p.CtCab * p.gam_Om2);
    c.rth_th1 = c.rth_th1 - QC * p.kap_th1; c.rth_th2 = c.rth_th2 - QC * p.kap_th2;
    c.rth_Om1 = -(QC * p.kap_Om1); c.rth_Om2 = -(QC * p.kap_Om2);
    return c;
}
// element.jl:1901-1919
template <class T> Resultants<T> dynamic_element_resultants(const Problem& P, const DynElem<T>& p, int64_t ie) {
    Resultants<T> r = static_element_resultants(P, p, ie);
    V3<T> tmp = cross(p.wb, p.Pm) + p.Pdot;
    r.F1 = r.F1 - T(0.5) * tmp; r.F2 = r.F2 + T(0.5) * tmp;
    tmp = cross(p.wb, p.H) + cross(p.V, p.Pm) + p.Hdot;
    r.M1 = r.M1 - T(0.5) * tmp; r.M2 = r.M2 + T(0.5) * tmp;
    return r;
}
template <class T> struct DynResultantJac : ResultantJac<T> {
    M3<T> F1_u1, F1_u2, F1_V1, F1_V2, F1_Om1, F1_Om2, F2_u1, F2_u2, F2_V1, F2_V2, F2_Om1, F2_Om2;
    M3<T> M1_u1, M1_u2, M1_V1, M1_V2, M1_Om1, M1_Om2, M2_u1, M2_u2, M2_V1, M2_V2, M2_Om1, M2_Om2;
    M3<T> F1_ab, F1_alb, F2_ab, F2_alb, M1_ab, M1_alb, M2_ab, M2_alb;
};
// element.jl:2219-2328, 2035-2064
template <class T> DynResultantJac<T> steady_element_resultant_jacobians(const Problem& P, const DynElem<T>& p, int64_t ie) {
    DynResultantJac<T> r;
    static_cast<ResultantJac<T>&>(r) = static_element_resultant_jacobians(P, p, ie);
    M3<T> wt = tilde(p.wb), Vt = tilde(p.V);
    M3<T> tmp = T(0.5) * (p.CtCab * tilde(p.F));
    r.M1_u1 = -(tmp * p.gam_u1); r.M2_u1 = tmp * p.gam_u1;
    r.M1_u2 = -(tmp * p.gam_u2); r.M2_u2 = tmp * p.gam_u2;
    r.M1_th1 = r.M1_th1 - tmp * p.gam_th1; r.M2_th1 = r.M2_th1 + tmp * p.gam_th1;
    r.M1_th2 = r.M1_th2 - tmp * p.gam_th2; r.M2_th2 = r.M2_th2 + tmp * p.gam_th2;
    r.M1_V1 = -(tmp * p.gam_V1); r.M2_V1 = tmp * p.gam_V1;
    r.M1_V2 = -(tmp * p.gam_V2); r.M2_V2 = tmp * p.gam_V2;
    r.M1_Om1 = -(tmp * p.gam_Om1); r.M2_Om1 = tmp * p.gam_Om1;
    r.M1_Om2 = -(tmp * p.gam_Om2); r.M2_Om2 = tmp * p.gam_Om2;
    T h = T(0.5);
    tmp = h * p.Pdot_u;
    r.F1_u1 = -(h * tmp); r.F2_u1 = h * tmp; r.F1_u2 = -(h * tmp); r.F2_u2 = h * tmp;
    tmp = h * (wt * p.P_th + p.Pdot_th);
    r.F1_th1 = r.F1_th1 - h * tmp; r.F2_th1 = r.F2_th1 + h * tmp; r.F1_th2 = r.F1_th2 - h * tmp; r.F2_th2 = r.F2_th2 + h * tmp;
    tmp = h * (wt * p.P_V + p.Pdot_V);
    r.F1_V1 = -(h * tmp); r.F2_V1 = h * tmp; r.F1_V2 = -(h * tmp); r.F2_V2 = h * tmp;
    tmp = h * (wt * p.P_Om + p.Pdot_Om);
    r.F1_Om1 = -(h * tmp); r.F2_Om1 = h * tmp; r.F1_Om2 = -(h * tmp); r.F2_Om2 = h * tmp;
    tmp = h * p.Hdot_u;
    r.M1_u1 = r.M1_u1 - h * tmp; r.M2_u1 = r.M2_u1 + h * tmp; r.M1_u2 = r.M1_u2 - h * tmp; r.M2_u2 = r.M2_u2 + h * tmp;
    tmp = h * (wt * p.H_th + Vt * p.P_th + p.Hdot_th);
    r.M1_th1 = r.M1_th1 - h * tmp; r.M2_th1 = r.M2_th1 + h * tmp; r.M1_th2 = r.M1_th2 - h * tmp; r.M2_th2 = r.M2_th2 + h * tmp;
    tmp = h * (wt * p.H_V + Vt * p.P_V - tilde(p.Pm) + p.Hdot_V);
    r.M1_V1 = r.M1_V1 - h * tmp; r.M2_V1 = r.M2_V1 + h * tmp; r.M1_V2 = r.M1_V2 - h * tmp; r.M2_V2 = r.M2_V2 + h * tmp;
    tmp = h * (wt * p.H_Om + Vt * p.P_Om + p.Hdot_Om);
    r.M1_Om1 = r.M1_Om1 - h * tmp; r.M2_Om1 = r.M2_Om1 + h * tmp; r.M1_Om2 = r.M1_Om2 - h * tmp; r.M2_Om2 = r.M2_Om2 + h * tmp;
    // steady: prescribed / state body accelerations
    tmp = h * p.Pdot_ab; r.F1_ab = -tmp; r.F2_ab = tmp;
    tmp = h * p.Pdot_alb; r.F1_alb = -tmp; r.F2_alb = tmp;
    tmp = h * p.Hdot_ab; r.M1_ab = -tmp; r.M2_ab = tmp;
    tmp = h * p.Hdot_alb; r.M1_alb = -tmp; r.M2_alb = tmp;
    return r;
}
// element.jl:3019-3035
template <class T> void steady_element_residual(const Problem& P, const Indices& idx, const T* x, int64_t ie, const DynParams& D, const V3<T>& ab,
                                                const V3<T>& alb, T* resid) {
    DynElem<T> p = steady_element_properties(P, idx, x, ie, D, ab, alb);
    V3<T> ru, rth;
    compatibility_residuals(p, ru, rth);
    Resultants<T> r = dynamic_element_resultants(P, p, ie);
    insert_element_residuals(P, idx, ie, ru, rth, r, resid);
}
// element.jl:3209-3229, 2685-2759, 2554-2583
template <class T, class J> void steady_element_jacobian(const Problem& P, const Indices& idx, const T* x, int64_t ie, const DynParams& D,
                                                         const V3<T>& ab, const V3<T>& alb, J& jac) {
    DynElem<T> p = steady_element_properties(P, idx, x, ie, D, ab, alb);
    steady_element_jacobian_properties(P, ie, p);
    DynCompatJac<T> c = dynamic_compatibility_jacobians(p);
    DynResultantJac<T> r = steady_element_resultant_jacobians(P, p, ie);
    insert_static_element_jacobians(P, idx, ie, p, c, r, jac);
    T fs = T(P.force_scaling);
    int64_t icol1 = idx.icol_point[P.start[ie] - 1] - 1, icol2 = idx.icol_point[P.stop[ie] - 1] - 1;
    int64_t irow = idx.irow_elem[ie] - 1;
    set33(jac, irow, icol1 + 6, c.ru_V1); set33(jac, irow, icol1 + 9, c.ru_Om1);
    set33(jac, irow, icol2 + 6, c.ru_V2); set33(jac, irow, icol2 + 9, c.ru_Om2);
    set33(jac, irow + 3, icol1 + 9, c.rth_Om1); set33(jac, irow + 3, icol2 + 9, c.rth_Om2);
    int64_t irow1 = idx.irow_point[P.start[ie] - 1] - 1;
    sub33m(jac, irow1, icol1, (r.F1_u1 * p.u1_u1) / fs); sub33m(jac, irow1, icol2, (r.F1_u2 * p.u2_u2) / fs);
    sub33m(jac, irow1, icol1 + 6, r.F1_V1 / fs); sub33m(jac, irow1, icol1 + 9, r.F1_Om1 / fs);
    sub33m(jac, irow1, icol2 + 6, r.F1_V2 / fs); sub33m(jac, irow1, icol2 + 9, r.F1_Om2 / fs);
    sub33m(jac, irow1 + 3, icol1, (r.M1_u1 * p.u1_u1) / fs); sub33m(jac, irow1 + 3, icol2, (r.M1_u2 * p.u2_u2) / fs);
    sub33m(jac, irow1 + 3, icol1 + 6, r.M1_V1 / fs); sub33m(jac, irow1 + 3, icol1 + 9, r.M1_Om1 / fs);
    sub33m(jac, irow1 + 3, icol2 + 6, r.M1_V2 / fs); sub33m(jac, irow1 + 3, icol2 + 9, r.M1_Om2 / fs);
    int64_t irow2 = idx.irow_point[P.stop[ie] - 1] - 1;
    add33(jac, irow2, icol1, (r.F2_u1 * p.u1_u1) / fs); add33(jac, irow2, icol2, (r.F2_u2 * p.u2_u2) / fs);
    add33(jac, irow2, icol1 + 6, r.F2_V1 / fs); add33(jac, irow2, icol1 + 9, r.F2_Om1 / fs);
    add33(jac, irow2, icol2 + 6, r.F2_V2 / fs); add33(jac, irow2, icol2 + 9, r.F2_Om2 / fs);
    add33(jac, irow2 + 3, icol1, (r.M2_u1 * p.u1_u1) / fs); add33(jac, irow2 + 3, icol2, (r.M2_u2 * p.u2_u2) / fs);
    add33(jac, irow2 + 3, icol1 + 6, r.M2_V1 / fs); add33(jac, irow2 + 3, icol1 + 9, r.M2_Om1 / fs);
    add33(jac, irow2 + 3, icol2 + 6, r.M2_V2 / fs); add33(jac, irow2 + 3, icol2 + 9, r.M2_Om2 / fs);
    for (int i = 0; i < 3; i++) {
        if (idx.icol_body[i]) {
            int64_t cb = idx.icol_body[i] - 1;
            addcol3(jac, irow1, cb, (-col(r.F1_ab, i)) / fs); addcol3(jac, irow1 + 3, cb, (-col(r.M1_ab, i)) / fs);
            addcol3(jac, irow2, cb, col(r.F2_ab, i) / fs); addcol3(jac, irow2 + 3, cb, col(r.M2_ab, i) / fs);
        }
        if (idx.icol_body[3 + i]) {
            int64_t cb = idx.icol_body[3 + i] - 1;
            addcol3(jac, irow1, cb, (-col(r.F1_alb, i)) / fs); addcol3(jac, irow1 + 3, cb, (-col(r.M1_alb, i)) / fs);
            addcol3(jac, irow2, cb, col(r.F2_alb, i) / fs); addcol3(jac, irow2 + 3, cb, col(r.M2_alb, i) / fs);
        }
    }
}

// system.jl:427-455
template <class T> void steady_system_residual(const Problem& P, const Indices& idx, const DynParams& D, const T* x, T* resid) {
    V3<T> ab, alb;
    body_accelerations(idx, x, D, ab, alb);
    for (int64_t ip = 0; ip < P.np; ip++) steady_point_residual(P, idx, x, ip, D, ab, alb, resid);
    for (int64_t ie = 0; ie < P.ne; ie++) steady_element_residual(P, idx, x, ie, D, ab, alb, resid);
    if (P.two_dimensional) two_dimensional_residual(idx.nstates, x, resid);
}
// system.jl:693-720
template <class T, class J> void steady_system_jacobian(const Problem& P, const Indices& idx, const DynParams& D, const T* x, J& jac) {
    jac.zero();
    V3<T> ab, alb;
    body_accelerations(idx, x, D, ab, alb);
    for (int64_t ip = 0; ip < P.np; ip++) steady_point_jacobian(P, idx, x, ip, D, ab, alb, jac);
    for (int64_t ie = 0; ie < P.ne; ie++) steady_element_jacobian(P, idx, x, ie, D, ab, alb, jac);
    if (P.two_dimensional) two_dimensional_jacobian<T>(idx.nstates, jac);
}

// ================================================================= Newmark time marching =================================
//   newmark_point_properties / _jacobian_properties          src/point.jl:766-795, 1134-1200
//   newmark_point_velocity_jacobians                         src/point.jl:1708-1720
//   newmark_point_residual! / _jacobian!                     src/point.jl:2200-2221, 2410-2432
//   newmark_element_properties / _jacobian_properties        src/element.jl:381-412, 1049-1112
//   newmark_element_residual! / _jacobian!                   src/element.jl:3074-3090, 3271-3292
//   newmark_system_residual! / _jacobian!                    src/system.jl:530-554, 829-852
struct NewmarkParams {
    const double* init = nullptr;   // [np][12]: udot_init, θdot_init, Vdot_init, Ωdot_init  (`paug`, analyses.jl:2693-2738)
    double dt = 1.0;
};

template <class T> void newmark_point_rates(DynPoint<T>& p, const NewmarkParams& N, int64_t ip) {
    const double* q = N.init + 12 * ip;
    T c = T(2.0 / N.dt);
    p.udot = c * p.u - cvec<T>(q); p.thdot = c * p.th - cvec<T>(q + 3);
    p.Vdot = c * p.V - cvec<T>(q + 6); p.Omdot = c * p.Om - cvec<T>(q + 9);
}
// point.jl:766-795
template <class T> DynPoint<T> newmark_point_properties(const Problem& P, const Indices& idx, const T* x, int64_t ip, const DynParams& D,
                                                        const NewmarkParams& N, M3<T>& Cdot) {
    V3<T> z = zero3<T>();
    DynPoint<T> p = steady_point_properties(P, idx, x, ip, D, z, z);
    newmark_point_rates(p, N, ip);
    Cdot = -(p.C * tilde(p.Om - p.wb));
    M3<T> Ct = tr(p.C), Cdt = tr(Cdot);
    p.Pdot = Ct * (p.mass11 * (p.C * p.Vdot)) + Ct * (p.mass12 * (p.C * p.Omdot)) + Ct * (p.mass11 * (Cdot * p.V)) + Ct * (p.mass12 * (Cdot * p.Om)) +
             Cdt * (p.mass11 * (p.C * p.V)) + Cdt * (p.mass12 * (p.C * p.Om));
    p.Hdot = Ct * (p.mass21 * (p.C * p.Vdot)) + Ct * (p.mass22 * (p.C * p.Omdot)) + Ct * (p.mass21 * (Cdot * p.V)) + Ct * (p.mass22 * (Cdot * p.Om)) +
             Cdt * (p.mass21 * (p.C * p.V)) + Cdt * (p.mass22 * (p.C * p.Om));
    return p;
}
// point.jl:1134-1200
template <class T> void newmark_point_jacobian_properties(const Problem& P, const Indices& idx, const T* x, int64_t ip, DynPoint<T>& p,
                                                          const M3<T>& Cdot, const NewmarkParams& N) {
    steady_point_jacobian_properties(P, idx, x, ip, p);
    M3<T> Ct = tr(p.C), Cdt = tr(Cdot), Ct1 = tr(p.C_t1), Ct2 = tr(p.C_t2), Ct3 = tr(p.C_t3);
    auto m3c = [&](const V3<T>& w) { return mul3(p.C_t1, p.C_t2, p.C_t3, w); };
    auto m3t = [&](const V3<T>& w) { return mul3(Ct1, Ct2, Ct3, w); };
    M3<T> Wt = tilde(p.Om - p.wb);
    T c = T(2.0 / N.dt);
    auto rate_th = [&](const M3<T>& ma, const M3<T>& mb) {
        return m3t(ma * (p.C * p.Vdot)) + m3t(mb * (p.C * p.Omdot)) + Ct * (ma * m3c(p.Vdot)) + Ct * (mb * m3c(p.Omdot)) +
               Wt * m3t(ma * (p.C * p.V)) + Wt * m3t(mb * (p.C * p.Om)) + Cdt * (ma * m3c(p.V)) + Cdt * (mb * m3c(p.Om)) +
               m3t(ma * (Cdot * p.V)) + m3t(mb * (Cdot * p.Om)) - Ct * (ma * m3c(Wt * p.V)) - Ct * (mb * m3c(Wt * p.Om));
    };
    p.Pdot_u = zero33<T>(); p.Hdot_u = zero33<T>();
    p.Pdot_th = rate_th(p.mass11, p.mass12);
    p.Hdot_th = rate_th(p.mass21, p.mass22);
    p.Pdot_V = Ct * p.mass11 * Cdot + Cdt * p.mass11 * p.C + c * (Ct * p.mass11 * p.C);
    p.Pdot_Om = Ct * p.mass12 * Cdot + Cdt * p.mass12 * p.C + Ct * p.mass11 * p.C * tilde(p.V) + Ct * p.mass12 * p.C * tilde(p.Om) -
                tilde(Ct * (p.mass11 * (p.C * p.V))) - tilde(Ct * (p.mass12 * (p.C * p.Om))) + c * (Ct * p.mass12 * p.C);
    p.Hdot_V = Ct * p.mass21 * Cdot + Cdt * p.mass21 * p.C + c * (Ct * p.mass21 * p.C);
    p.Hdot_Om = Ct * p.mass22 * Cdot + Cdt * p.mass22 * p.C + Ct * p.mass21 * p.C * tilde(p.V) + Ct * p.mass22 * p.C * tilde(p.Om) -
                tilde(Ct * (p.mass21 * (p.C * p.V))) - tilde(Ct * (p.mass22 * (p.C * p.Om))) + c * (Ct * p.mass22 * p.C);
}
// point.jl:2200-2221
template <class T> void newmark_point_residual(const Problem& P, const Indices& idx, const T* x, int64_t ip, const DynParams& D, const NewmarkParams& N,
                                               T* resid) {
    int64_t irow = idx.irow_point[ip] - 1;
    M3<T> Cdot;
    DynPoint<T> p = newmark_point_properties(P, idx, x, ip, D, N, Cdot);
    V3<T> F, M, rV, rOm;
    dynamic_point_resultants(p, F, M);
    point_velocity_residuals(p, rV, rOm);
    T fs = T(P.force_scaling);
    for (int i = 0; i < 3; i++) { resid[irow + i] = -F[i] / fs; resid[irow + 3 + i] = -M[i] / fs; resid[irow + 6 + i] = rV[i]; resid[irow + 9 + i] = rOm[i]; }
}
// point.jl:2410-2432 (+ 1708-1720, 1907-1935)
template <class T, class J> void newmark_point_jacobian(const Problem& P, const Indices& idx, const T* x, int64_t ip, const DynParams& D,
                                                        const NewmarkParams& N, J& jac) {
    M3<T> Cdot;
    DynPoint<T> p = newmark_point_properties(P, idx, x, ip, D, N, Cdot);
    newmark_point_jacobian_properties(P, idx, x, ip, p, Cdot, N);
    PointResJac<T> r = steady_point_resultant_jacobians(p);   // the dynamic part of it (F_ab.. unused here)
    T c = T(2.0 / N.dt);
    M3<T> rV_u = -tilde(p.wb) - c * eye3<T>(), rV_V = eye3<T>();
    M3<T> rOm_th = mul3(p.Qi_t1, p.Qi_t2, p.Qi_t3, p.C * (p.Om - p.wb)) + p.Qinv * mul3(p.C_t1, p.C_t2, p.C_t3, p.Om - p.wb) - c * eye3<T>();
    M3<T> rOm_Om = p.Qinv * p.C;
    int64_t irow = idx.irow_point[ip] - 1, icol = idx.icol_point[ip] - 1;
    T fs = T(P.force_scaling);
    set33(jac, irow, icol, -p.F_F);
    set33(jac, irow, icol + 3, (-(r.F_th * p.th_th)) / fs);
    set33(jac, irow + 3, icol + 3, -p.M_M - (r.M_th * p.th_th) / fs);
    set33(jac, irow, icol + 6, (-r.F_V) / fs);
    set33(jac, irow, icol + 9, (-r.F_Om) / fs);
    set33(jac, irow + 3, icol + 6, (-r.M_V) / fs);
    set33(jac, irow + 3, icol + 9, (-r.M_Om) / fs);
    set33(jac, irow + 6, icol, rV_u * p.u_u);
    set33(jac, irow + 6, icol + 6, rV_V);
    set33(jac, irow + 9, icol + 3, rOm_th * p.th_th);
    set33(jac, irow + 9, icol + 9, rOm_Om);
}
// element.jl:381-412
template <class T> DynElem<T> newmark_element_properties(const Problem& P, const Indices& idx, const T* x, int64_t ie, const DynParams& D,
                                                         const NewmarkParams& N, M3<T>& CtCabdot) {
    V3<T> z = zero3<T>();
    DynElem<T> p = steady_element_properties(P, idx, x, ie, D, z, z);
    int64_t ia = P.start[ie] - 1, ib = P.stop[ie] - 1;
    T c = T(2.0 / N.dt);
    V3<T> V1dot = c * p.V1 - cvec<T>(N.init + 12 * ia + 6), Om1dot = c * p.Om1 - cvec<T>(N.init + 12 * ia + 9);
    V3<T> V2dot = c * p.V2 - cvec<T>(N.init + 12 * ib + 6), Om2dot = c * p.Om2 - cvec<T>(N.init + 12 * ib + 9);
    p.Vdot = (V1dot + V2dot) / T(2); p.Omdot = (Om1dot + Om2dot) / T(2);
    CtCabdot = tilde(p.Om - p.wb) * p.CtCab;
    M3<T> A = p.CtCab, At = tr(p.CtCab), B = CtCabdot, Bt = tr(CtCabdot);
    p.Pdot = A * (p.mass11 * (At * p.Vdot)) + A * (p.mass12 * (At * p.Omdot)) + A * (p.mass11 * (Bt * p.V)) + A * (p.mass12 * (Bt * p.Om)) +
             B * (p.mass11 * (At * p.V)) + B * (p.mass12 * (At * p.Om));
    p.Hdot = A * (p.mass21 * (At * p.Vdot)) + A * (p.mass22 * (At * p.Omdot)) + A * (p.mass21 * (Bt * p.V)) + A * (p.mass22 * (Bt * p.Om)) +
             B * (p.mass21 * (At * p.V)) + B * (p.mass22 * (At * p.Om));
    return p;
}
// element.jl:1049-1112
template <class T> void newmark_element_jacobian_properties(const Problem& P, int64_t ie, DynElem<T>& p, const M3<T>& CtCabdot, const NewmarkParams& N) {
    steady_element_jacobian_properties(P, ie, p);
    M3<T> Ct1 = tr(p.C_t1), Ct2 = tr(p.C_t2), Ct3 = tr(p.C_t3), At = tr(p.CtCab), Bt = tr(CtCabdot), CabT = tr(p.Cab);
    const M3<T>&A = p.CtCab, &B = CtCabdot;
    auto m3c = [&](const V3<T>& w) { return mul3(p.C_t1, p.C_t2, p.C_t3, w); };
    auto m3t = [&](const V3<T>& w) { return mul3(Ct1, Ct2, Ct3, w); };
    M3<T> Wt = tilde(p.Om - p.wb);
    T c = T(2.0 / N.dt);
    auto rate_th = [&](const M3<T>& ma, const M3<T>& mb) {
        return m3t(p.Cab * (ma * (At * p.Vdot))) + m3t(p.Cab * (mb * (At * p.Omdot))) + A * (ma * (CabT * m3c(p.Vdot))) + A * (mb * (CabT * m3c(p.Omdot))) +
               Wt * m3t(p.Cab * (ma * (At * p.V))) + Wt * m3t(p.Cab * (mb * (At * p.Om))) + B * (ma * (CabT * m3c(p.V))) + B * (mb * (CabT * m3c(p.Om))) +
               m3t(p.Cab * (ma * (Bt * p.V))) + m3t(p.Cab * (mb * (Bt * p.Om))) - A * (ma * (CabT * m3c(Wt * p.V))) - A * (mb * (CabT * m3c(Wt * p.Om)));
    };
    p.Pdot_u = zero33<T>(); p.Hdot_u = zero33<T>();
    p.Pdot_th = rate_th(p.mass11, p.mass12);
    p.Hdot_th = rate_th(p.mass21, p.mass22);
    p.Pdot_V = A * p.mass11 * Bt + B * p.mass11 * At + c * (A * p.mass11 * At);
    p.Pdot_Om = A * p.mass12 * Bt + B * p.mass12 * At + A * p.mass11 * At * tilde(p.V) + A * p.mass12 * At * tilde(p.Om) -
                tilde(A * (p.mass11 * (At * p.V))) - tilde(A * (p.mass12 * (At * p.Om))) + c * (A * p.mass12 * At);
    p.Hdot_V = A * p.mass21 * Bt + B * p.mass21 * At + c * (A * p.mass21 * At);
    p.Hdot_Om = A * p.mass22 * Bt + B * p.mass22 * At + A * p.mass21 * At * tilde(p.V) + A * p.mass22 * At * tilde(p.Om) -
                tilde(A * (p.mass21 * (At * p.V))) - tilde(A * (p.mass22 * (At * p.Om))) + c * (A * p.mass22 * At);
}
// element.jl:3074-3090
template <class T> void newmark_element_residual(const Problem& P, const Indices& idx, const T* x, int64_t ie, const DynParams& D, const NewmarkParams& N,
                                                 T* resid) {
    M3<T> Bd;
    DynElem<T> p = newmark_element_properties(P, idx, x, ie, D, N, Bd);
    V3<T> ru, rth;
    compatibility_residuals(p, ru, rth);
    Resultants<T> r = dynamic_element_resultants(P, p, ie);
    insert_element_residuals(P, idx, ie, ru, rth, r, resid);
}
// shared tail of steady_element_jacobian / newmark_element_jacobian: insert_dynamic_element_jacobians! (element.jl:2685-2759)
template <class T, class J> void insert_dynamic_element_jacobians(const Problem& P, const Indices& idx, int64_t ie, const DynElem<T>& p,
                                                                  const DynCompatJac<T>& c, const DynResultantJac<T>& r, J& jac) {
    insert_static_element_jacobians(P, idx, ie, p, c, r, jac);
    T fs = T(P.force_scaling);
    int64_t icol1 = idx.icol_point[P.start[ie] - 1] - 1, icol2 = idx.icol_point[P.stop[ie] - 1] - 1;
    int64_t irow = idx.irow_elem[ie] - 1;
    set33(jac, irow, icol1 + 6, c.ru_V1); set33(jac, irow, icol1 + 9, c.ru_Om1);
    set33(jac, irow, icol2 + 6, c.ru_V2); set33(jac, irow, icol2 + 9, c.ru_Om2);
    set33(jac, irow + 3, icol1 + 9, c.rth_Om1); set33(jac, irow + 3, icol2 + 9, c.rth_Om2);
    int64_t irow1 = idx.irow_point[P.start[ie] - 1] - 1;
    sub33m(jac, irow1, icol1, (r.F1_u1 * p.u1_u1) / fs); sub33m(jac, irow1, icol2, (r.F1_u2 * p.u2_u2) / fs);
    sub33m(jac, irow1, icol1 + 6, r.F1_V1 / fs); sub33m(jac, irow1, icol1 + 9, r.F1_Om1 / fs);
    sub33m(jac, irow1, icol2 + 6, r.F1_V2 / fs); sub33m(jac, irow1, icol2 + 9, r.F1_Om2 / fs);
    sub33m(jac, irow1 + 3, icol1, (r.M1_u1 * p.u1_u1) / fs); sub33m(jac, irow1 + 3, icol2, (r.M1_u2 * p.u2_u2) / fs);
    sub33m(jac, irow1 + 3, icol1 + 6, r.M1_V1 / fs); sub33m(jac, irow1 + 3, icol1 + 9, r.M1_Om1 / fs);
    sub33m(jac, irow1 + 3, icol2 + 6, r.M1_V2 / fs); sub33m(jac, irow1 + 3, icol2 + 9, r.M1_Om2 / fs);
    int64_t irow2 = idx.irow_point[P.stop[ie] - 1] - 1;
    add33(jac, irow2, icol1, (r.F2_u1 * p.u1_u1) / fs); add33(jac, irow2, icol2, (r.F2_u2 * p.u2_u2) / fs);
    add33(jac, irow2, icol1 + 6, r.F2_V1 / fs); add33(jac, irow2, icol1 + 9, r.F2_Om1 / fs);
    add33(jac, irow2, icol2 + 6, r.F2_V2 / fs); add33(jac, irow2, icol2 + 9, r.F2_Om2 / fs);
    add33(jac, irow2 + 3, icol1, (r.M2_u1 * p.u1_u1) / fs); add33(jac, irow2 + 3, icol2, (r.M2_u2 * p.u2_u2) / fs);
    add33(jac, irow2 + 3, icol1 + 6, r.M2_V1 / fs); add33(jac, irow2 + 3, icol1 + 9, r.M2_Om1 / fs);
    add33(jac, irow2 + 3, icol2 + 6, r.M2_V2 / fs); add33(jac, irow2 + 3, icol2 + 9, r.M2_Om2 / fs);
}
// element.jl:3271-3292
template <class T, class J> void newmark_element_jacobian(const Problem& P, const Indices& idx, const T* x, int64_t ie, const DynParams& D,
                                                          const NewmarkParams& N, J& jac) {
    M3<T> Bd;
    DynElem<T> p = newmark_element_properties(P, idx, x, ie, D, N, Bd);
    newmark_element_jacobian_properties(P, ie, p, Bd, N);
    DynCompatJac<T> c = dynamic_compatibility_jacobians(p);
    DynResultantJac<T> r = steady_element_resultant_jacobians(P, p, ie);   // = dynamic_element_resultant_jacobians + unused ab/αb parts
    insert_dynamic_element_jacobians(P, idx, ie, p, c, r, jac);
}
// system.jl:530-554
template <class T> void newmark_system_residual(const Problem& P, const Indices& idx, const DynParams& D, const NewmarkParams& N, const T* x, T* resid) {
    for (int64_t ip = 0; ip < P.np; ip++) newmark_point_residual(P, idx, x, ip, D, N, resid);
    for (int64_t ie = 0; ie < P.ne; ie++) newmark_element_residual(P, idx, x, ie, D, N, resid);
    if (P.two_dimensional) two_dimensional_residual(idx.nstates, x, resid);
}
// system.jl:829-852
template <class T, class J> void newmark_system_jacobian(const Problem& P, const Indices& idx, const DynParams& D, const NewmarkParams& N, const T* x, J& jac) {
    jac.zero();
    for (int64_t ip = 0; ip < P.np; ip++) newmark_point_jacobian(P, idx, x, ip, D, N, jac);
    for (int64_t ie = 0; ie < P.ne; ie++) newmark_element_jacobian(P, idx, x, ie, D, N, jac);
    if (P.two_dimensional) two_dimensional_jacobian<T>(idx.nstates, jac);
}

// ================================================================= mass matrix (eigenvalue analysis) ========================
//   mass_matrix_point_jacobian! (point.jl:2522, 1356-1384, 1586-1597, 1746-1755, 2033-2061)
//   mass_matrix_element_jacobian! (element.jl:3395, 1408-1441, 2370-2411, 2903-2953)
//   system_mass_matrix! (system.jl:961-995): M = ∂r/∂ẋ, added to `jac` times gamma
template <class T, class J> void system_mass_matrix(const Problem& P, const Indices& idx, const T* x, T gamma, J& jac) {
    T fs = T(P.force_scaling);
    const bool two_d = P.two_dimensional != 0;
    auto masked = [&](const M3<T>& A, bool angular) {      // lmask = (1,1,0), amask = (0,0,1) on rows when two_dimensional
        M3<T> r = A;
        if (two_d) for (int j = 0; j < 3; j++) { if (angular) { r.m[0][j] = T(0); r.m[1][j] = T(0); } else r.m[2][j] = T(0); }
        return r;
    };
    static const double zeros36[36] = {0};
    for (int64_t ip = 0; ip < P.np; ip++) {
        const double* m = P.pmass ? P.pmass + 36 * ip : zeros36;
        M3<T> m11 = sub33<T>(m, 0, 0), m12 = sub33<T>(m, 0, 1), m21 = sub33<T>(m, 1, 0), m22 = sub33<T>(m, 1, 1);
        int64_t irow = idx.irow_point[ip] - 1, icol = idx.icol_point[ip] - 1;
        V3<T> u, th; point_displacement(P, x, ip, icol, u, th);
        M3<T> u_u, th_th; point_displacement_jacobians(P, ip, u_u, th_th);
        M3<T> C = get_C(th), Ct = tr(C);
        M3<T> PV = Ct * m11 * C, PO = Ct * m12 * C, HV = Ct * m21 * C, HO = Ct * m22 * C;
        // jacob -= F_Vdot gamma / fs with F_Vdot = -Pdot_Vdot
        add33(jac, irow, icol + 6, masked(PV, false) * gamma / fs); add33(jac, irow, icol + 9, masked(PO, false) * gamma / fs);
        add33(jac, irow + 3, icol + 6, masked(HV, true) * gamma / fs); add33(jac, irow + 3, icol + 9, masked(HO, true) * gamma / fs);
        add33(jac, irow + 6, icol, masked(-(u_u * u_u), false) * gamma);
        add33(jac, irow + 9, icol + 3, masked(-(th_th * th_th), true) * gamma);
    }
    for (int64_t ie = 0; ie < P.ne; ie++) {
        const double* e = P.elems + 91 * ie;
        double L = e[0];
        const double* mass = e + 40;
        M3<T> m11 = sub33<T>(mass, 0, 0, L), m12 = sub33<T>(mass, 0, 1, L), m21 = sub33<T>(mass, 1, 0, L), m22 = sub33<T>(mass, 1, 1, L);
        M3<T> Cab = cmat33<T>(e + 76);
        int64_t ia = P.start[ie] - 1, ib = P.stop[ie] - 1;
        V3<T> u1, th1, u2, th2;
        point_displacement(P, x, ia, idx.icol_point[ia] - 1, u1, th1);
        point_displacement(P, x, ib, idx.icol_point[ib] - 1, u2, th2);
        V3<T> th = (th1 + th2) / T(2);
        M3<T> A = tr(get_C(th)) * Cab, At = tr(A);
        T q = T(0.25);
        M3<T> PV = q * (A * m11 * At), PO = q * (A * m12 * At), HV = q * (A * m21 * At), HO = q * (A * m22 * At);
        for (int64_t ipr : {ia, ib}) {
            int64_t irow = idx.irow_point[ipr] - 1;
            for (int64_t ipc : {ia, ib}) {
                int64_t icol = idx.icol_point[ipc] - 1;
                add33(jac, irow, icol + 6, masked(PV, false) * gamma / fs); add33(jac, irow, icol + 9, masked(PO, false) * gamma / fs);
                add33(jac, irow + 3, icol + 6, masked(HV, true) * gamma / fs); add33(jac, irow + 3, icol + 9, masked(HO, true) * gamma / fs);
            }
        }
    }
}

// ================================================================= initial condition analysis ============================
//   initial_point_displacement / _velocity_rates              src/point.jl:305-366, 448-512   (initial_point_displacement_rates = point_velocities :438)
//   initial_point_properties                                  src/point.jl:683-757
//   initial_point_residual!                                   src/point.jl:2168-2190
//   initial_element_properties                                src/element.jl:230-371
//   initial_element_residual!                                 src/element.jl:3046-3064
//   initial_system_residual!                                  src/system.jl:466-521  (incl. the "prescribe velocity rates explicitly" rows)
//   initial_system_jacobian!                                  src/system.jl:731-817
// In this analysis the point columns are re-interpreted: x[icol..+5] holds (per component) a reaction load (displacement
// prescribed), a velocity rate Vdot/Ωdot (rate_vars2 set) or a displacement u/θ; x[icol+6..+11] holds udot, θdot.  V, Ω are
// the given initial values.
//
// JACOBIAN: the oracle does not restate the reference's ~600 lines of analytic initial-condition Jacobian blocks
// (point.jl:1033-1124, 1548-1577, 1690-1701, 1853-1897; element.jl:850-1041, 1523-1548, 2072-2216, 2589-2683).  It takes the
// complex-step derivative of the restated residual instead (exact to round-off); the reference's own test asserts that its
// analytic Jacobian equals the ForwardDiff Jacobian of this residual to atol 1e-10 (test/jacobians.jl:207-236).  One
// deliberate deviation keeps the two identical: follower point loads are rotated with θ taken from x[icol+3..5] even where
// that slot holds Ωdot (point_loads calls the ordinary point_displacement, point.jl:24), while the analytic Jacobian
// multiplies F_θ by the *initial* θ_θ mask (point.jl:1866-1871) — so that dependence is frozen (real part only) below.
struct InitialParams {
    const uint8_t* rv1 = nullptr;   // rate_vars1 [nstates]   (analyses.jl:1835-1846)
    const uint8_t* rv2 = nullptr;   // rate_vars2 [nstates]   (analyses.jl:1848-1873)
    const double* u0 = nullptr;     // [np][3] each
    const double* th0 = nullptr;
    const double* V0 = nullptr;
    const double* Om0 = nullptr;
    const double* Vdot0 = nullptr;
    const double* Omdot0 = nullptr;
};
template <class T> T frozen(const T& v) { return T(realpart(v)); }

// point.jl:305-366
template <class T> void initial_point_displacement(const Problem& P, const T* x, int64_t ip, int64_t icol, const InitialParams& I, V3<T>& u, V3<T>& th) {
    const uint8_t* pd = P.pd + 6 * ip;
    const double* b = P.bc + 18 * ip;
    for (int i = 0; i < 3; i++) {
        u.v[i] = pd[i] ? T(b[i]) : (I.rv2[icol + 6 + i] ? T(I.u0[3 * ip + i]) : x[icol + i]);
        th.v[i] = pd[3 + i] ? T(b[3 + i]) : (I.rv2[icol + 9 + i] ? T(I.th0[3 * ip + i]) : x[icol + 3 + i]);
    }
}
// point.jl:448-512
template <class T> void initial_point_velocity_rates(const Problem& P, const T* x, int64_t ip, int64_t icol, const InitialParams& I, V3<T>& Vdot, V3<T>& Omdot) {
    const uint8_t* pd = P.pd + 6 * ip;
    for (int i = 0; i < 3; i++) {
        Vdot.v[i] = pd[i] ? T(0) : (I.rv2[icol + 6 + i] ? x[icol + i] : T(I.Vdot0[3 * ip + i]));
        Omdot.v[i] = pd[3 + i] ? T(0) : (I.rv2[icol + 9 + i] ? x[icol + 3 + i] : T(I.Omdot0[3 * ip + i]));
    }
}
// point_loads as called from initial_point_properties (point.jl:709, 20-39); see the note on follower loads above
template <class T> void initial_point_loads(const Problem& P, const T* x, int64_t ip, int64_t icol, const InitialParams& I, V3<T>& F, V3<T>& M) {
    const uint8_t* pd = P.pd + 6 * ip;
    const uint8_t* pl = P.pl + 6 * ip;
    const double* b = P.bc + 18 * ip;
    V3<T> th;
    for (int i = 0; i < 3; i++) th.v[i] = pd[3 + i] ? T(b[3 + i]) : (I.rv2[icol + 9 + i] ? frozen(x[icol + 3 + i]) : x[icol + 3 + i]);
    M3<T> Ct = tr(get_C(th));
    V3<T> Fc = cvec<T>(b + 6) + Ct * cvec<T>(b + 12);
    V3<T> Mc = cvec<T>(b + 9) + Ct * cvec<T>(b + 15);
    T fs = T(P.force_scaling);
    for (int i = 0; i < 3; i++) {
        F.v[i] = pl[i] ? Fc[i] : x[icol + i] * fs;
        M.v[i] = pl[3 + i] ? Mc[i] : x[icol + 3 + i] * fs;
    }
}
// point.jl:683-757
template <class T> DynPoint<T> initial_point_properties(const Problem& P, const Indices& idx, const T* x, int64_t ip, const DynParams& D,
                                                        const InitialParams& I, const V3<T>& ab, const V3<T>& alb) {
    DynPoint<T> p;
    static const double zeros36[36] = {0};
    const double* m = P.pmass ? P.pmass + 36 * ip : zeros36;
    p.mass11 = sub33<T>(m, 0, 0); p.mass12 = sub33<T>(m, 0, 1); p.mass21 = sub33<T>(m, 1, 0); p.mass22 = sub33<T>(m, 1, 1);
    int64_t icol = idx.icol_point[ip] - 1;
    initial_point_displacement(P, x, ip, icol, I, p.u, p.th);
    p.C = get_C(p.th);
    p.Qinv = get_Qinv(p.th);
    initial_point_loads(P, x, ip, icol, I, p.F, p.M);
    p.dx = cvec<T>(P.points + 3 * ip);
    p.vb = cvec<T>(D.vb); p.wb = cvec<T>(D.wb); p.ab = ab; p.alb = alb;
    p.gvec = cvec<T>(P.gravity);
    p.V = cvec<T>(I.V0 + 3 * ip) + p.vb + cross(p.wb, p.dx + p.u);
    p.Om = cvec<T>(I.Om0 + 3 * ip) + p.wb;
    M3<T> Ct = tr(p.C);
    p.Pm = Ct * (p.mass11 * (p.C * p.V)) + Ct * (p.mass12 * (p.C * p.Om));
    p.H = Ct * (p.mass21 * (p.C * p.V)) + Ct * (p.mass22 * (p.C * p.Om));
    p.udot = ld3(x, icol + 6); p.thdot = ld3(x, icol + 9);
    initial_point_velocity_rates(P, x, ip, icol, I, p.Vdot, p.Omdot);
    p.Vdot = p.Vdot + ab + cross(alb, p.dx + p.u) + cross(p.wb, p.udot);
    p.Omdot = p.Omdot + alb;
    M3<T> Cdot = -(p.C * tilde(p.Om - p.wb)), Cdt = tr(Cdot);
    p.Pdot = Ct * (p.mass11 * (p.C * p.Vdot)) + Ct * (p.mass12 * (p.C * p.Omdot)) + Ct * (p.mass11 * (Cdot * p.V)) + Ct * (p.mass12 * (Cdot * p.Om)) +
             Cdt * (p.mass11 * (p.C * p.V)) + Cdt * (p.mass12 * (p.C * p.Om));
    p.Hdot = Ct * (p.mass21 * (p.C * p.Vdot)) + Ct * (p.mass22 * (p.C * p.Omdot)) + Ct * (p.mass21 * (Cdot * p.V)) + Ct * (p.mass22 * (Cdot * p.Om)) +
             Cdt * (p.mass21 * (p.C * p.V)) + Cdt * (p.mass22 * (p.C * p.Om));
    return p;
}
// point.jl:2168-2190
template <class T> void initial_point_residual(const Problem& P, const Indices& idx, const T* x, int64_t ip, const DynParams& D, const InitialParams& I,
                                               const V3<T>& ab, const V3<T>& alb, T* resid) {
    int64_t irow = idx.irow_point[ip] - 1;
    DynPoint<T> p = initial_point_properties(P, idx, x, ip, D, I, ab, alb);
    V3<T> F, M, rV, rOm;
    dynamic_point_resultants(p, F, M);
    point_velocity_residuals(p, rV, rOm);
    T fs = T(P.force_scaling);
    for (int i = 0; i < 3; i++) { resid[irow + i] = -F[i] / fs; resid[irow + 3 + i] = -M[i] / fs; resid[irow + 6 + i] = rV[i]; resid[irow + 9 + i] = rOm[i]; }
}
// element.jl:230-371
template <class T> DynElem<T> initial_element_properties(const Problem& P, const Indices& idx, const T* x, int64_t ie, const DynParams& D,
                                                         const InitialParams& I, const V3<T>& ab, const V3<T>& alb) {
    DynElem<T> p;
    static_cast<ElemProps<T>&>(p) = static_element_properties(P, idx, x, ie);     // L, Cab, S, mass, F, M, γ, κ, gvec
    const double* e = P.elems + 91 * ie;
    V3<T> e1 = unit<T>(0);
    int64_t ia = P.start[ie] - 1, ib = P.stop[ie] - 1;
    int64_t ic1 = idx.icol_point[ia] - 1, ic2 = idx.icol_point[ib] - 1;
    initial_point_displacement(P, x, ia, ic1, I, p.u1, p.th1);
    initial_point_displacement(P, x, ib, ic2, I, p.u2, p.th2);
    p.u = (p.u1 + p.u2) / T(2); p.th = (p.th1 + p.th2) / T(2);
    p.C = get_C(p.th); p.CtCab = tr(p.C) * p.Cab; p.Q = get_Q(p.th); p.Qinv = get_Qinv(p.th);
    p.dx = cvec<T>(e + 1);
    p.vb = cvec<T>(D.vb); p.wb = cvec<T>(D.wb); p.ab = ab; p.alb = alb;
    p.V1 = cvec<T>(I.V0 + 3 * ia); p.V2 = cvec<T>(I.V0 + 3 * ib); p.Om1 = cvec<T>(I.Om0 + 3 * ia); p.Om2 = cvec<T>(I.Om0 + 3 * ib);
    p.V = (p.V1 + p.V2) / T(2) + p.vb + cross(p.wb, p.dx + p.u);
    p.Om = (p.Om1 + p.Om2) / T(2) + p.wb;
    M3<T> A = p.CtCab, At = tr(p.CtCab);
    p.Pm = A * (p.mass11 * (At * p.V)) + A * (p.mass12 * (At * p.Om));
    p.H = A * (p.mass21 * (At * p.V)) + A * (p.mass22 * (At * p.Om));
    p.udot1 = ld3(x, ic1 + 6); p.thdot1 = ld3(x, ic1 + 9); p.udot2 = ld3(x, ic2 + 6); p.thdot2 = ld3(x, ic2 + 9);
    p.udot = (p.udot1 + p.udot2) / T(2); p.thdot = (p.thdot1 + p.thdot2) / T(2);
    V3<T> V1d, O1d, V2d, O2d;
    initial_point_velocity_rates(P, x, ia, ic1, I, V1d, O1d);
    initial_point_velocity_rates(P, x, ib, ic2, I, V2d, O2d);
    p.Vdot = (V1d + V2d) / T(2) + ab + cross(alb, p.dx) + cross(alb, p.u) + cross(p.wb, p.udot);
    p.Omdot = (O1d + O2d) / T(2) + alb;
    M3<T> B = tilde(p.Om - p.wb) * p.CtCab, Bt = tr(B);
    p.Pdot = A * (p.mass11 * (At * p.Vdot)) + A * (p.mass12 * (At * p.Omdot)) + A * (p.mass11 * (Bt * p.V)) + A * (p.mass12 * (Bt * p.Om)) +
             B * (p.mass11 * (At * p.V)) + B * (p.mass12 * (At * p.Om));
    p.Hdot = A * (p.mass21 * (At * p.Vdot)) + A * (p.mass22 * (At * p.Omdot)) + A * (p.mass21 * (Bt * p.V)) + A * (p.mass22 * (Bt * p.Om)) +
             B * (p.mass21 * (At * p.V)) + B * (p.mass22 * (At * p.Om));
    p.damping = D.structural_damping;
    if (p.damping) {      // element.jl:333-368
        const double* mu = e + 85;
        p.mu11 = zero33<T>(); p.mu22 = zero33<T>();
        for (int i = 0; i < 3; i++) { p.mu11.m[i][i] = T(mu[i]); p.mu22.m[i][i] = T(mu[3 + i]); }
        p.du = p.u2 - p.u1; p.dth = p.th2 - p.th1;
        p.dudot = p.udot2 - p.udot1; p.dthdot = p.thdot2 - p.thdot1;
        p.dQ = get_dQ(p.th, p.dth, p.Q);
        M3<T> Wt = tilde(p.Om - p.wb);
        p.gamdot = -(At * (Wt * p.du)) + At * p.dudot - p.L * (At * (Wt * (p.Cab * e1)));
        M3<T> CabT = tr(p.Cab);
        p.kapdot = CabT * (p.Q * p.dthdot) + CabT * (p.dQ * p.thdot);
        p.gam = p.gam - p.mu11 * p.gamdot;
        p.kap = p.kap - p.mu22 * p.kapdot;
    }
    return p;
}
// element.jl:3046-3064
template <class T> void initial_element_residual(const Problem& P, const Indices& idx, const T* x, int64_t ie, const DynParams& D, const InitialParams& I,
                                                 const V3<T>& ab, const V3<T>& alb, T* resid) {
    DynElem<T> p = initial_element_properties(P, idx, x, ie, D, I, ab, alb);
    V3<T> ru, rth;
    compatibility_residuals(p, ru, rth);
    Resultants<T> r = dynamic_element_resultants(P, p, ie);
    insert_element_residuals(P, idx, ie, ru, rth, r, resid);
}
// true where the equilibrium row of DOF k of point ip is replaced by `Vdot_k - Vdot0_k = 0`   (system.jl:491-513, 756-813)
inline bool initial_row_replaced(const Problem& P, const Indices& idx, const InitialParams& I, int64_t ip, int k) {
    int64_t icol = idx.icol_point[ip] - 1;
    return !P.pd[6 * ip + k] && !I.rv1[icol + 6 + k] && I.rv2[icol + 6 + k];
}
// system.jl:466-521
template <class T> void initial_system_residual(const Problem& P, const Indices& idx, const DynParams& D, const InitialParams& I, const T* x, T* resid) {
    V3<T> ab, alb;
    body_accelerations(idx, x, D, ab, alb);
    for (int64_t ip = 0; ip < P.np; ip++) initial_point_residual(P, idx, x, ip, D, I, ab, alb, resid);
    for (int64_t ie = 0; ie < P.ne; ie++) initial_element_residual(P, idx, x, ie, D, I, ab, alb, resid);
    for (int64_t ip = 0; ip < P.np; ip++) {
        int64_t irow = idx.irow_point[ip] - 1, icol = idx.icol_point[ip] - 1;
        V3<T> Vdot, Omdot;
        initial_point_velocity_rates(P, x, ip, icol, I, Vdot, Omdot);
        for (int k = 0; k < 6; k++)
            if (initial_row_replaced(P, idx, I, ip, k))
                resid[irow + k] = (k < 3 ? Vdot[k] - T(I.Vdot0[3 * ip + k]) : Omdot[k - 3] - T(I.Omdot0[3 * ip + k - 3]));
    }
    if (P.two_dimensional) two_dimensional_residual(idx.nstates, x, resid);
}
// system.jl:731-817 by complex-step differentiation of initial_system_residual (see the note above).  Column compression: columns
// further apart than kl+ku cannot share a row, so kl+ku+1 residual evaluations recover the whole band.
template <class J> void initial_system_jacobian(const Problem& P, const Indices& idx, const DynParams& D, const InitialParams& I, const double* x,
                                                int64_t kl, int64_t ku, J& jac) {
    typedef std::complex<double> C;
    const int64_t n = idx.nstates, stride = std::min<int64_t>(n, kl + ku + 1);
    const double h = 1e-30;
    std::vector<C> xc(n), rc(n);
    jac.zero();
    for (int64_t c = 0; c < stride; c++) {
        for (int64_t i = 0; i < n; i++) xc[i] = C(x[i], (i % stride == c) ? h : 0.0);
        std::fill(rc.begin(), rc.end(), C(0, 0));
        initial_system_residual<C>(P, idx, D, I, xc.data(), rc.data());
        for (int64_t i = 0; i < n; i++) {
            double v = rc[i].imag() / h;
            if (v == 0.0) continue;
            // the unique column j ≡ c (mod stride) inside [i-kl, i+ku]
            int64_t lo = std::max<int64_t>(0, i - kl);
            int64_t j = lo + ((c - lo % stride) + stride) % stride;
            if (j <= std::min(n - 1, i + ku)) jac.set(i, j, v);
        }
    }
}
// accumulates the column sums of a matrix assembled through add()/set()  (rate_vars = .!iszero.(sum(M, dims=1)), analyses.jl:1846)
struct ColumnSum {
    std::vector<double> s;
    void add(int64_t, int64_t j, double v) { s[j] += v; }
    void set(int64_t, int64_t j, double v) { s[j] += v; }
};
// analyses.jl:1835-1873: rate_vars1 from the mass matrix at x = 0, rate_vars2 with the element mass matrices replaced by the
// compliance matrices and all point masses ignored
inline void initial_rate_vars(const Problem& P, const Indices& idx, uint8_t* rv1, uint8_t* rv2) {
    const int64_t n = idx.nstates;
    std::vector<double> x(n, 0.0);
    ColumnSum cs; cs.s.assign(n, 0.0);
    system_mass_matrix<double>(P, idx, x.data(), 1.0, cs);
    for (int64_t j = 0; j < n; j++) rv1[j] = cs.s[j] != 0.0;
    std::vector<double> el(P.elems, P.elems + 91 * P.ne);
    for (int64_t ie = 0; ie < P.ne; ie++) for (int k = 0; k < 36; k++) el[91 * ie + 40 + k] = el[91 * ie + 4 + k];
    Problem P2 = P; P2.elems = el.data(); P2.pmass = nullptr;
    cs.s.assign(n, 0.0);
    system_mass_matrix<double>(P2, idx, x.data(), 1.0, cs);
    for (int64_t j = 0; j < n; j++) rv2[j] = cs.s[j] != 0.0;
}
// initial_output (analyses.jl:2332-2390): the solved initial-condition vector -> DynamicSystem state vector `xx` and the point
// rates (udot, θdot, Vdot, Ωdot; np x 12) that seed the Newmark march
inline void initial_output(const Problem& P, const Indices& idx, const DynParams& D, const InitialParams& I, const double* x, double* xx, double* rates) {
    const int64_t n = idx.nstates;
    for (int64_t i = 0; i < n; i++) xx[i] = x[i];
    V3<double> ab, alb;
    body_accelerations(idx, x, D, ab, alb);
    V3<double> vb = cvec<double>(D.vb), wb = cvec<double>(D.wb);
    for (int64_t ip = 0; ip < P.np; ip++) {
        int64_t icol = idx.icol_point[ip] - 1;
        V3<double> u, th, F, M, Vdot, Omdot;
        initial_point_displacement(P, x, ip, icol, I, u, th);
        initial_point_loads(P, x, ip, icol, I, F, M);
        V3<double> udot = ld3(x, icol + 6), thdot = ld3(x, icol + 9);
        initial_point_velocity_rates(P, x, ip, icol, I, Vdot, Omdot);
        V3<double> dx = cvec<double>(P.points + 3 * ip);
        V3<double> V = cvec<double>(I.V0 + 3 * ip) + vb + cross(wb, dx + u), Om = cvec<double>(I.Om0 + 3 * ip) + wb;
        Vdot = Vdot + ab + cross(alb, dx + u) + cross(wb, udot);
        Omdot = Omdot + alb;
        const uint8_t* pd = P.pd + 6 * ip; const uint8_t* pl = P.pl + 6 * ip;
        for (int k = 0; k < 3; k++) {
            // set_linear/angular_displacement! then set_external_forces/moments! (input.jl): displacement where !pd, load where !pl
            if (!pd[k]) xx[icol + k] = u[k];
            if (!pd[3 + k]) xx[icol + 3 + k] = th[k];
            if (!pl[k]) xx[icol + k] = F[k] / P.force_scaling;
            if (!pl[3 + k]) xx[icol + 3 + k] = M[k] / P.force_scaling;
            xx[icol + 6 + k] = V[k]; xx[icol + 9 + k] = Om[k];
            rates[12 * ip + k] = pd[k] ? 0.0 : udot[k];
            rates[12 * ip + 3 + k] = pd[3 + k] ? 0.0 : thdot[k];
            rates[12 * ip + 6 + k] = Vdot[k]; rates[12 * ip + 9 + k] = Omdot[k];
        }
    }
}

}  // namespace gxo
