// ORACLE — TEST INFRASTRUCTURE ONLY (see gxb_math.hpp).
//
// CPU restatement of GXBeam.jl's static residual / Jacobian assembly:
//   SystemIndices                      src/system.jl:29-122
//   point helpers                      src/point.jl:9-214
//   static_point_*                     src/point.jl:595-620, 938-954, 1422-1431, 1476-1487, 1785-1800, 2113-2128, 2321-2337
//   static_element_*                   src/element.jl:62-107, 621-640, 1478-1521, 1860-1893, 1945-2027, 2421-2443, 2487-2552,
//                                      2995-3010, 3182-3200
//   static_system_residual!/jacobian!  src/system.jl:400-418, 663-683, 1072-1103
// The structure (properties -> compatibility -> resultants -> insert) deliberately follows the reference,
// residual and Jacobian are evaluated separately exactly like the reference does.
#pragma once
#include <cstdint>
#include <vector>
#include "gxb_math.hpp"

namespace gxo {

// One Assembly instance + its loads, in the memory layout of the C-ABI boundary (include/gxbeam_b200.h).
struct Problem {
    int64_t np = 0, ne = 0;
    const int64_t* start = nullptr;   // ne, 1-based
    const int64_t* stop = nullptr;    // ne, 1-based
    const double* elems = nullptr;    // ne x 91: L, x[3], compliance[36 col-major], mass[36 col-major], Cab[9 col-major], mu[6]
    const double* points = nullptr;   // np x 3
    const uint8_t* pd = nullptr;      // np x 6   (default: all 0)
    const uint8_t* pl = nullptr;      // np x 6   (default: all 1)
    const double* bc = nullptr;       // np x 18: u, theta, F, M, Ff, Mf   (NaN where unused)
    const double* dload = nullptr;    // ne x 24 or null: f1,f2,m1,m2,f1f,f2f,m1f,m2f
    const double* pmass = nullptr;    // np x 36 col-major or null
    double gravity[3] = {0, 0, 0};
    double force_scaling = 1.0;
    int two_dimensional = 0;
};

struct Indices {
    int64_t nstates = 0;
    std::vector<int64_t> irow_point, irow_elem, icol_point, icol_elem;  // 1-based like the reference
    int64_t icol_body[6] = {0, 0, 0, 0, 0, 0};
};

// system.jl:29-122
inline Indices system_indices(int64_t ne, const int64_t* start, const int64_t* stop, bool is_static, bool expanded = false) {
    int64_t np = 0;
    for (int64_t e = 0; e < ne; e++) { if (start[e] > np) np = start[e]; if (stop[e] > np) np = stop[e]; }
    Indices idx;
    std::vector<char> assigned(np, 0);
    idx.irow_point.assign(np, 0); idx.icol_point.assign(np, 0);
    idx.irow_elem.assign(ne, 0); idx.icol_elem.assign(ne, 0);
    int64_t irow = 1, icol = 1;
    auto add_point = [&](int64_t ipt) {
        if (assigned[ipt]) return;
        assigned[ipt] = 1;
        idx.icol_point[ipt] = icol; icol += 6;   // u/F, θ/M
        idx.irow_point[ipt] = irow; irow += 6;   // ΣF = 0, ΣM = 0
        if (!is_static) { icol += 6; irow += 6; }  // V, Ω
    };
    for (int64_t e = 0; e < ne; e++) {
        add_point(start[e] - 1);
        idx.irow_elem[e] = irow; idx.icol_elem[e] = icol;
        int64_t w = expanded ? 18 : 6;
        irow += w; icol += w;
        add_point(stop[e] - 1);
    }
    idx.nstates = icol - 1;
    return idx;
}

template <class T> V3<T> ld3(const T* x, int64_t i0) { return V3<T>{{x[i0], x[i0 + 1], x[i0 + 2]}}; }
template <class T> V3<T> cvec(const double* p) { return V3<T>{{T(p[0]), T(p[1]), T(p[2])}}; }
// 3x3 sub-block (bi,bj) of a 6x6 column-major matrix, optionally scaled
template <class T> M3<T> sub33(const double* m66, int bi, int bj, double scale = 1.0) {
    M3<T> r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = T(m66[(3 * bi + i) + 6 * (3 * bj + j)] * scale);
    return r;
}
template <class T> M3<T> cmat33(const double* m9) {  // column-major 3x3
    M3<T> r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = T(m9[i + 3 * j]);
    return r;
}

// ---- point helpers (point.jl:9-214) ----
// point.jl:147-175
template <class T> void point_displacement(const Problem& P, const T* x, int64_t ip, int64_t icol, V3<T>& u, V3<T>& th) {
    const uint8_t* pd = P.pd + 6 * ip;
    const double* b = P.bc + 18 * ip;
    for (int i = 0; i < 3; i++) {
        u.v[i] = pd[i] ? T(b[i]) : x[icol + i];
        th.v[i] = pd[3 + i] ? T(b[3 + i]) : x[icol + 3 + i];
    }
}
// point.jl:193-206
template <class T> void point_displacement_jacobians(const Problem& P, int64_t ip, M3<T>& u_u, M3<T>& th_th) {
    const uint8_t* pd = P.pd + 6 * ip;
    u_u = zero33<T>(); th_th = zero33<T>();
    for (int i = 0; i < 3; i++) { if (!pd[i]) u_u.m[i][i] = T(1); if (!pd[3 + i]) th_th.m[i][i] = T(1); }
}
// point.jl:20-39
template <class T> void point_loads(const Problem& P, const T* x, int64_t ip, int64_t icol, V3<T>& F, V3<T>& M) {
    const uint8_t* pl = P.pl + 6 * ip;
    const double* b = P.bc + 18 * ip;
    V3<T> u, th;
    point_displacement(P, x, ip, icol, u, th);
    M3<T> Ct = tr(get_C(th));
    V3<T> Fc = cvec<T>(b + 6) + Ct * cvec<T>(b + 12);
    V3<T> Mc = cvec<T>(b + 9) + Ct * cvec<T>(b + 15);
    T fs = T(P.force_scaling);
    for (int i = 0; i < 3; i++) {
        F.v[i] = pl[i] ? Fc[i] : x[icol + i] * fs;
        M.v[i] = pl[3 + i] ? Mc[i] : x[icol + 3 + i] * fs;
    }
}
// point.jl:65-118
template <class T> void point_load_jacobians(const Problem& P, const T* x, int64_t ip, int64_t icol, M3<T>& F_th, M3<T>& F_F,
                                             M3<T>& M_th, M3<T>& M_M) {
    const uint8_t* pl = P.pl + 6 * ip;
    const double* b = P.bc + 18 * ip;
    V3<T> u, th;
    point_displacement(P, x, ip, icol, u, th);
    M3<T> C = get_C(th);
    M3<T> C1, C2, C3;
    get_C_th(C, th, C1, C2, C3);
    V3<T> Ff = cvec<T>(b + 12), Mf = cvec<T>(b + 15);
    F_th = zero33<T>(); M_th = zero33<T>();
    const M3<T>* Ck[3] = {&C1, &C2, &C3};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) {
                // rot_θ[j,k] = C_θk[i,j]
                F_th.m[j][k] = F_th.m[j][k] + Ck[k]->m[i][j] * Ff[i];
                M_th.m[j][k] = M_th.m[j][k] + Ck[k]->m[i][j] * Mf[i];
            }
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) {
            if (!pl[i]) F_th.m[i][k] = T(0);
            if (!pl[3 + i]) M_th.m[i][k] = T(0);
        }
    F_F = zero33<T>(); M_M = zero33<T>();
    for (int i = 0; i < 3; i++) { if (!pl[i]) F_F.m[i][i] = T(1); if (!pl[3 + i]) M_M.m[i][i] = T(1); }
}

// ---- static point (point.jl:595-620, 938-954, 1422-1431, 1476-1487) ----
template <class T> struct PointProps {
    M3<T> mass11, mass12, mass21, mass22;
    V3<T> u, th, F, M, gvec;
    M3<T> C;
    // jacobian properties
    M3<T> u_u, th_th, C_t1, C_t2, C_t3, F_th, M_th, F_F, M_M;
};

template <class T> PointProps<T> static_point_properties(const Problem& P, const Indices& idx, const T* x, int64_t ip) {
    PointProps<T> p;
    static const double zeros36[36] = {0};
    const double* m = P.pmass ? P.pmass + 36 * ip : zeros36;
    p.mass11 = sub33<T>(m, 0, 0); p.mass12 = sub33<T>(m, 0, 1); p.mass21 = sub33<T>(m, 1, 0); p.mass22 = sub33<T>(m, 1, 1);
    int64_t icol = idx.icol_point[ip] - 1;
    point_displacement(P, x, ip, icol, p.u, p.th);
    p.C = get_C(p.th);
    point_loads(P, x, ip, icol, p.F, p.M);
    p.gvec = cvec<T>(P.gravity);
    return p;
}
template <class T> void static_point_jacobian_properties(const Problem& P, const Indices& idx, const T* x, int64_t ip, PointProps<T>& p) {
    int64_t icol = idx.icol_point[ip] - 1;
    point_load_jacobians(P, x, ip, icol, p.F_th, p.F_F, p.M_th, p.M_M);
    point_displacement_jacobians(P, ip, p.u_u, p.th_th);
    get_C_th(p.C, p.th, p.C_t1, p.C_t2, p.C_t3);
}
template <class T> void static_point_resultants(const PointProps<T>& p, V3<T>& F, V3<T>& M) {
    F = p.F + tr(p.C) * (p.mass11 * (p.C * p.gvec));
    M = p.M + tr(p.C) * (p.mass21 * (p.C * p.gvec));
}
template <class T> void static_point_resultant_jacobians(const PointProps<T>& p, M3<T>& F_th, M3<T>& M_th) {
    M3<T> Ct = tr(p.C);
    F_th = p.F_th + mul3(tr(p.C_t1), tr(p.C_t2), tr(p.C_t3), p.mass11 * (p.C * p.gvec)) +
           Ct * (p.mass11 * mul3(p.C_t1, p.C_t2, p.C_t3, p.gvec));
    M_th = p.M_th + mul3(tr(p.C_t1), tr(p.C_t2), tr(p.C_t3), p.mass21 * (p.C * p.gvec)) +
           Ct * (p.mass21 * mul3(p.C_t1, p.C_t2, p.C_t3, p.gvec));
}

// block helpers on an abstract matrix accessor with set(i,j,v)/add(i,j,v), zero-based
template <class J, class T> void set33(J& jac, int64_t r0, int64_t c0, const M3<T>& B) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) jac.set(r0 + i, c0 + j, B.m[i][j]);
}
template <class J, class T> void add33(J& jac, int64_t r0, int64_t c0, const M3<T>& B) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) jac.add(r0 + i, c0 + j, B.m[i][j]);
}
template <class J, class T> void sub33m(J& jac, int64_t r0, int64_t c0, const M3<T>& B) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) jac.add(r0 + i, c0 + j, -B.m[i][j]);
}

// point.jl:2113-2128
template <class T> void static_point_residual(const Problem& P, const Indices& idx, const T* x, int64_t ip, T* resid) {
    int64_t irow = idx.irow_point[ip] - 1;
    PointProps<T> p = static_point_properties(P, idx, x, ip);
    V3<T> F, M;
    static_point_resultants(p, F, M);
    T fs = T(P.force_scaling);
    for (int i = 0; i < 3; i++) { resid[irow + i] = -F[i] / fs; resid[irow + 3 + i] = -M[i] / fs; }
}
// point.jl:2321-2337, 1785-1800
template <class T, class J> void static_point_jacobian(const Problem& P, const Indices& idx, const T* x, int64_t ip, J& jac) {
    PointProps<T> p = static_point_properties(P, idx, x, ip);
    static_point_jacobian_properties(P, idx, x, ip, p);
    M3<T> F_th, M_th;
    static_point_resultant_jacobians(p, F_th, M_th);
    int64_t irow = idx.irow_point[ip] - 1, icol = idx.icol_point[ip] - 1;
    T fs = T(P.force_scaling);
    set33(jac, irow, icol, -p.F_F);
    set33(jac, irow, icol + 3, (-(F_th * p.th_th)) / fs);
    set33(jac, irow + 3, icol + 3, -p.M_M - (M_th * p.th_th) / fs);
}

// ---- static element ----
template <class T> struct ElemProps {
    T L;
    M3<T> C, Cab, CtCab, Qinv, S11, S12, S21, S22, mass11, mass12, mass21, mass22;
    V3<T> u1, u2, th1, th2, u, th, F, M, gam, kap, gvec;
    // jacobian properties
    M3<T> u1_u1, u2_u2, th1_th1, th2_th2, C_t1, C_t2, C_t3, Qi_t1, Qi_t2, Qi_t3;
};

// element.jl:62-107
template <class T> ElemProps<T> static_element_properties(const Problem& P, const Indices& idx, const T* x, int64_t ie) {
    ElemProps<T> p;
    const double* e = P.elems + 91 * ie;
    double L = e[0];
    p.L = T(L);
    const double* comp = e + 4;
    const double* mass = e + 40;
    p.Cab = cmat33<T>(e + 76);
    // compliance *= L ; mass *= L
    p.S11 = sub33<T>(comp, 0, 0, L); p.S12 = sub33<T>(comp, 0, 1, L); p.S21 = sub33<T>(comp, 1, 0, L); p.S22 = sub33<T>(comp, 1, 1, L);
    p.mass11 = sub33<T>(mass, 0, 0, L); p.mass12 = sub33<T>(mass, 0, 1, L); p.mass21 = sub33<T>(mass, 1, 0, L); p.mass22 = sub33<T>(mass, 1, 1, L);
    int64_t ia = P.start[ie] - 1, ib = P.stop[ie] - 1;
    point_displacement(P, x, ia, idx.icol_point[ia] - 1, p.u1, p.th1);
    point_displacement(P, x, ib, idx.icol_point[ib] - 1, p.u2, p.th2);
    p.u = (p.u1 + p.u2) / T(2);
    p.th = (p.th1 + p.th2) / T(2);
    p.C = get_C(p.th);
    p.CtCab = tr(p.C) * p.Cab;
    p.Qinv = get_Qinv(p.th);
    int64_t icol = idx.icol_elem[ie] - 1;
    T fs = T(P.force_scaling);
    p.F = ld3(x, icol) * fs;
    p.M = ld3(x, icol + 3) * fs;
    p.gam = p.S11 * p.F + p.S12 * p.M;
    p.kap = p.S21 * p.F + p.S22 * p.M;
    p.gvec = cvec<T>(P.gravity);
    return p;
}
// element.jl:621-640
template <class T> void static_element_jacobian_properties(const Problem& P, int64_t ie, ElemProps<T>& p) {
    point_displacement_jacobians(P, P.start[ie] - 1, p.u1_u1, p.th1_th1);
    point_displacement_jacobians(P, P.stop[ie] - 1, p.u2_u2, p.th2_th2);
    get_C_th(p.C, p.th, p.C_t1, p.C_t2, p.C_t3);
    get_Qinv_th(p.th, p.Qi_t1, p.Qi_t2, p.Qi_t3);
}
// element.jl:1478-1490
template <class T> void compatibility_residuals(const ElemProps<T>& p, V3<T>& ru, V3<T>& rth) {
    V3<T> e1 = unit<T>(0);
    V3<T> du = p.CtCab * (p.L * e1 + p.gam) - p.L * (p.Cab * e1);
    V3<T> dth = p.Qinv * (p.Cab * p.kap);
    ru = p.u2 - p.u1 - du;
    rth = p.th2 - p.th1 - dth;
}
template <class T> struct CompatJac { M3<T> ru_u1, ru_u2, ru_th1, ru_th2, ru_F, ru_M, rth_th1, rth_th2, rth_F, rth_M; };
// element.jl:1492-1521
template <class T> CompatJac<T> static_compatibility_jacobians(const ElemProps<T>& p) {
    CompatJac<T> c;
    V3<T> e1 = unit<T>(0);
    c.ru_u1 = -p.u1_u1;
    c.ru_u2 = p.u2_u2;
    M3<T> du_th = mul3(tr(p.C_t1), tr(p.C_t2), tr(p.C_t3), p.Cab * (p.L * e1 + p.gam));
    c.ru_th1 = T(-0.5) * du_th;
    c.ru_th2 = T(-0.5) * du_th;
    c.ru_F = -(p.CtCab * p.S11);
    c.ru_M = -(p.CtCab * p.S12);
    M3<T> dth_th = mul3(p.Qi_t1, p.Qi_t2, p.Qi_t3, p.Cab * p.kap);
    c.rth_th1 = -eye3<T>() - T(0.5) * dth_th;
    c.rth_th2 = eye3<T>() - T(0.5) * dth_th;
    c.rth_F = -(p.Qinv * p.Cab * p.S21);
    c.rth_M = -(p.Qinv * p.Cab * p.S22);
    return c;
}
template <class T> struct Resultants { V3<T> F1, F2, M1, M2; };
// element.jl:1860-1893
template <class T> Resultants<T> static_element_resultants(const Problem& P, const ElemProps<T>& p, int64_t ie) {
    Resultants<T> r;
    V3<T> e1 = unit<T>(0);
    V3<T> tmp = p.CtCab * p.F;
    r.F1 = tmp; r.F2 = tmp;
    V3<T> tmp1 = p.CtCab * p.M;
    V3<T> tmp2 = T(0.5) * (p.CtCab * cross(p.L * e1 + p.gam, p.F));
    r.M1 = tmp1 + tmp2; r.M2 = tmp1 - tmp2;
    M3<T> CtCabT = tr(p.CtCab);
    tmp = -(p.CtCab * (p.mass11 * (CtCabT * p.gvec)));
    r.F1 = r.F1 - T(0.5) * tmp; r.F2 = r.F2 + T(0.5) * tmp;
    tmp = -(p.CtCab * (p.mass21 * (CtCabT * p.gvec)));
    r.M1 = r.M1 - T(0.5) * tmp; r.M2 = r.M2 + T(0.5) * tmp;
    if (P.dload) {
        const double* d = P.dload + 24 * ie;
        M3<T> Ct = tr(p.C);
        r.F1 = r.F1 + (cvec<T>(d + 0) + Ct * cvec<T>(d + 12));
        r.F2 = r.F2 - (cvec<T>(d + 3) + Ct * cvec<T>(d + 15));
        r.M1 = r.M1 + (cvec<T>(d + 6) + Ct * cvec<T>(d + 18));
        r.M2 = r.M2 - (cvec<T>(d + 9) + Ct * cvec<T>(d + 21));
    }
    return r;
}
template <class T> struct ResultantJac { M3<T> F1_th1, F1_th2, F1_F, F2_th1, F2_th2, F2_F, M1_th1, M1_th2, M1_F, M1_M, M2_th1, M2_th2, M2_F, M2_M; };
// element.jl:1945-2027
template <class T> ResultantJac<T> static_element_resultant_jacobians(const Problem& P, const ElemProps<T>& p, int64_t ie) {
    ResultantJac<T> r;
    V3<T> e1 = unit<T>(0);
    M3<T> Ct1 = tr(p.C_t1), Ct2 = tr(p.C_t2), Ct3 = tr(p.C_t3);
    M3<T> tmp = mul3(Ct1, Ct2, Ct3, p.Cab * p.F);
    r.F1_th1 = T(0.5) * tmp; r.F2_th1 = T(0.5) * tmp; r.F1_th2 = T(0.5) * tmp; r.F2_th2 = T(0.5) * tmp;
    r.F1_F = p.CtCab; r.F2_F = p.CtCab;
    M3<T> tmp1 = mul3(Ct1, Ct2, Ct3, p.Cab * p.M);
    M3<T> tmp2 = T(0.5) * mul3(Ct1, Ct2, Ct3, p.Cab * cross(p.L * e1 + p.gam, p.F));
    r.M1_th1 = T(0.5) * (tmp1 + tmp2); r.M2_th1 = T(0.5) * (tmp1 - tmp2);
    r.M1_th2 = T(0.5) * (tmp1 + tmp2); r.M2_th2 = T(0.5) * (tmp1 - tmp2);
    tmp = T(0.5) * (p.CtCab * (tilde(p.L * e1 + p.gam) - tilde(p.F) * p.S11));
    r.M1_F = tmp; r.M2_F = -tmp;
    tmp = T(0.5) * (p.CtCab * tilde(p.F) * p.S12);
    r.M1_M = p.CtCab - tmp; r.M2_M = p.CtCab + tmp;
    // gravitational loads
    M3<T> CtCabT = tr(p.CtCab), CabT = tr(p.Cab);
    tmp = T(-0.5) * (mul3(Ct1, Ct2, Ct3, p.Cab * (p.mass11 * (CtCabT * p.gvec))) +
                     p.CtCab * (p.mass11 * (CabT * mul3(p.C_t1, p.C_t2, p.C_t3, p.gvec))));
    r.F1_th1 = r.F1_th1 - T(0.5) * tmp; r.F2_th1 = r.F2_th1 + T(0.5) * tmp;
    r.F1_th2 = r.F1_th2 - T(0.5) * tmp; r.F2_th2 = r.F2_th2 + T(0.5) * tmp;
    tmp = T(-0.5) * (mul3(Ct1, Ct2, Ct3, p.Cab * (p.mass21 * (CtCabT * p.gvec))) +
                     p.CtCab * (p.mass21 * (CabT * mul3(p.C_t1, p.C_t2, p.C_t3, p.gvec))));
    r.M1_th1 = r.M1_th1 - T(0.5) * tmp; r.M2_th1 = r.M2_th1 + T(0.5) * tmp;
    r.M1_th2 = r.M1_th2 - T(0.5) * tmp; r.M2_th2 = r.M2_th2 + T(0.5) * tmp;
    if (P.dload) {
        const double* d = P.dload + 24 * ie;
        tmp = mul3(Ct1, Ct2, Ct3, cvec<T>(d + 12));
        r.F1_th1 = r.F1_th1 + T(0.5) * tmp; r.F1_th2 = r.F1_th2 + T(0.5) * tmp;
        tmp = mul3(Ct1, Ct2, Ct3, cvec<T>(d + 15));
        r.F2_th1 = r.F2_th1 - T(0.5) * tmp; r.F2_th2 = r.F2_th2 - T(0.5) * tmp;
        tmp = mul3(Ct1, Ct2, Ct3, cvec<T>(d + 18));
        r.M1_th1 = r.M1_th1 + T(0.5) * tmp; r.M1_th2 = r.M1_th2 + T(0.5) * tmp;
        tmp = mul3(Ct1, Ct2, Ct3, cvec<T>(d + 21));
        r.M2_th1 = r.M2_th1 - T(0.5) * tmp; r.M2_th2 = r.M2_th2 - T(0.5) * tmp;
    }
    return r;
}
// element.jl:2421-2443
template <class T> void insert_element_residuals(const Problem& P, const Indices& idx, int64_t ie, const V3<T>& ru, const V3<T>& rth,
                                                 const Resultants<T>& r, T* resid) {
    T fs = T(P.force_scaling);
    int64_t irow = idx.irow_elem[ie] - 1;
    for (int i = 0; i < 3; i++) { resid[irow + i] = ru[i]; resid[irow + 3 + i] = rth[i]; }
    irow = idx.irow_point[P.start[ie] - 1] - 1;
    for (int i = 0; i < 3; i++) { resid[irow + i] = resid[irow + i] - r.F1[i] / fs; resid[irow + 3 + i] = resid[irow + 3 + i] - r.M1[i] / fs; }
    irow = idx.irow_point[P.stop[ie] - 1] - 1;
    for (int i = 0; i < 3; i++) { resid[irow + i] = resid[irow + i] + r.F2[i] / fs; resid[irow + 3 + i] = resid[irow + 3 + i] + r.M2[i] / fs; }
}
// element.jl:2995-3010
template <class T> void static_element_residual(const Problem& P, const Indices& idx, const T* x, int64_t ie, T* resid) {
    ElemProps<T> p = static_element_properties(P, idx, x, ie);
    V3<T> ru, rth;
    compatibility_residuals(p, ru, rth);
    Resultants<T> r = static_element_resultants(P, p, ie);
    insert_element_residuals(P, idx, ie, ru, rth, r, resid);
}
// element.jl:2487-2552
template <class T, class J> void insert_static_element_jacobians(const Problem& P, const Indices& idx, int64_t ie, const ElemProps<T>& p,
                                                                 const CompatJac<T>& c, const ResultantJac<T>& r, J& jac) {
    T fs = T(P.force_scaling);
    int64_t icol = idx.icol_elem[ie] - 1;
    int64_t icol1 = idx.icol_point[P.start[ie] - 1] - 1, icol2 = idx.icol_point[P.stop[ie] - 1] - 1;
    int64_t irow = idx.irow_elem[ie] - 1;
    set33(jac, irow, icol1, c.ru_u1 * p.u1_u1);
    set33(jac, irow, icol1 + 3, c.ru_th1 * p.th1_th1);
    set33(jac, irow, icol2, c.ru_u2 * p.u2_u2);
    set33(jac, irow, icol2 + 3, c.ru_th2 * p.th2_th2);
    set33(jac, irow, icol, c.ru_F * fs);
    set33(jac, irow, icol + 3, c.ru_M * fs);
    set33(jac, irow + 3, icol1 + 3, c.rth_th1 * p.th1_th1);
    set33(jac, irow + 3, icol2 + 3, c.rth_th2 * p.th2_th2);
    set33(jac, irow + 3, icol, c.rth_F * fs);
    set33(jac, irow + 3, icol + 3, c.rth_M * fs);
    int64_t irow1 = idx.irow_point[P.start[ie] - 1] - 1;
    sub33m(jac, irow1, icol1 + 3, (r.F1_th1 * p.th1_th1) / fs);
    sub33m(jac, irow1, icol2 + 3, (r.F1_th2 * p.th2_th2) / fs);
    set33(jac, irow1, icol, -r.F1_F);
    sub33m(jac, irow1 + 3, icol1 + 3, (r.M1_th1 * p.th1_th1) / fs);
    sub33m(jac, irow1 + 3, icol2 + 3, (r.M1_th2 * p.th2_th2) / fs);
    set33(jac, irow1 + 3, icol, -r.M1_F);
    set33(jac, irow1 + 3, icol + 3, -r.M1_M);
    int64_t irow2 = idx.irow_point[P.stop[ie] - 1] - 1;
    add33(jac, irow2, icol1 + 3, (r.F2_th1 * p.th1_th1) / fs);
    add33(jac, irow2, icol2 + 3, (r.F2_th2 * p.th2_th2) / fs);
    set33(jac, irow2, icol, r.F2_F);
    add33(jac, irow2 + 3, icol1 + 3, (r.M2_th1 * p.th1_th1) / fs);
    add33(jac, irow2 + 3, icol2 + 3, (r.M2_th2 * p.th2_th2) / fs);
    set33(jac, irow2 + 3, icol, r.M2_F);
    set33(jac, irow2 + 3, icol + 3, r.M2_M);
}
// element.jl:3182-3200
template <class T, class J> void static_element_jacobian(const Problem& P, const Indices& idx, const T* x, int64_t ie, J& jac) {
    ElemProps<T> p = static_element_properties(P, idx, x, ie);
    static_element_jacobian_properties(P, ie, p);
    CompatJac<T> c = static_compatibility_jacobians(p);
    ResultantJac<T> r = static_element_resultant_jacobians(P, p, ie);
    insert_static_element_jacobians(P, idx, ie, p, c, r, jac);
}

// system.jl:1072-1103
template <class T> void two_dimensional_residual(int64_t n, const T* x, T* resid) {
    for (int64_t i = 0; i < n; i += 6) { resid[i + 2] = x[i + 2]; resid[i + 3] = x[i + 3]; resid[i + 4] = x[i + 4]; }
}
template <class T, class J> void two_dimensional_jacobian(int64_t n, J& jac) {
    for (int64_t i = 0; i < n; i += 6)
        for (int k = 2; k <= 4; k++) { jac.zero_row(i + k); jac.set(i + k, i + k, T(1)); }
}

// system.jl:400-418
template <class T> void static_system_residual(const Problem& P, const Indices& idx, const T* x, T* resid) {
    for (int64_t ip = 0; ip < P.np; ip++) static_point_residual(P, idx, x, ip, resid);
    for (int64_t ie = 0; ie < P.ne; ie++) static_element_residual(P, idx, x, ie, resid);
    if (P.two_dimensional) two_dimensional_residual(idx.nstates, x, resid);
}
// system.jl:663-683
template <class T, class J> void static_system_jacobian(const Problem& P, const Indices& idx, const T* x, J& jac) {
    jac.zero();
    for (int64_t ip = 0; ip < P.np; ip++) static_point_jacobian(P, idx, x, ip, jac);
    for (int64_t ie = 0; ie < P.ne; ie++) static_element_jacobian(P, idx, x, ie, jac);
    if (P.two_dimensional) two_dimensional_jacobian<T>(idx.nstates, jac);
}

}  // namespace gxo
