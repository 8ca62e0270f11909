// Fused residual + Jacobian assembly for the initial-condition analysis that starts a time-domain run, one thread per station,
// one CTA per instance (same tiles, passes and row-block layout as gxb_dynamic.cuh).  Replaces
//     initial_system_residual!  src/system.jl:466-521  -> initial_point_residual! (point.jl:2168), initial_element_residual! (element.jl:3046)
//     initial_system_jacobian!  src/system.jl:731-817  -> initial_point_jacobian! (point.jl:2376), initial_element_jacobian! (element.jl:3239)
// (properties point.jl:683-757, 1033-1124; element.jl:230-371, 850-1041; resultants point.jl:1548-1577, element.jl:2072-2216;
//  compatibility element.jl:1523-1548; velocity rows point.jl:1658-1666, 1690-1701; inserts point.jl:1853-1897, element.jl:2589-2683).
//
// The point columns are re-interpreted (point.jl:305-366, 448-512): per component k of point p the slot x[icol+k] holds
//   * the reaction load            where the displacement is prescribed,
//   * the velocity rate Vdot/Ωdot   where rate_vars2[icol+6+k] is set (u/θ then come from u0/θ0),
//   * the displacement u/θ          otherwise (Vdot/Ωdot then come from Vdot0/Ωdot0);
// x[icol+6..11] holds udot, θdot; V, Ω are the given V0, Ω0 plus the body-frame motion.  Where an equilibrium row cannot determine
// its velocity rate (!rate_vars1 & rate_vars2) it is replaced by Vdot_k - Vdot0_k = 0 (system.jl:491-513, 756-813).
//
// With W = Ω - ωb constant, the inertial loads are those of the Newmark step (frame rate tilde(W) A), so inertia_loads<1> and
// inertia_jacobians<1> (with 2/dt = 0) are reused; the extra chain-rule pieces are V_u = tilde(ωb), Vdot_u = tilde(αb),
// Vdot_udot = tilde(ωb), Vdot_Vdot = Ωdot_Ωdot = I (masked).
// Reference quirk kept: follower point loads are rotated with θ read from x[icol+3..5] even where that slot holds Ωdot
// (point_loads -> point_displacement, point.jl:24), while their Jacobian F_θ is masked with the initial θ_θ (point.jl:1866).
#pragma once
#include "gxb_dynamic.cuh"

struct InitArgs {
    DynArgs d;
    const double* icS;        // [b][18][N+1]  u0, θ0, V0, Ω0, Vdot0, Ωdot0
    const unsigned* rvm;      // [b][N+1]      bit k: rate_vars2[icol+6+k];  bit 8+k: equilibrium row k replaced;  bit 16+k: rate_vars1
};

struct InitPoint {
    v3 u, th, Vr, Or, Vd, Od, ud, thd;     // Vr, Or: V0, Ω0 (relative);  Vd, Od: relative velocity rates;  ud, thd: udot, θdot states
    v3 mu, mth, mv, mo;                    // 1 where the slot is u / θ / Vdot / Ωdot
    unsigned fl, rv;
};
GXD void init_point_load(const InitArgs& A, int b, int p, const double* x, InitPoint& q) {
    const StaticArgs& S = A.d.s;
    const int NP = S.N + 1;
    const unsigned fl = S.flags[p], rv = A.rvm[(size_t)b * NP + p];
    const double* bc = S.bcS + ((size_t)b * 18) * NP + p;
    const double* ic = A.icS + ((size_t)b * 18) * NP + p;
    const double* xp = x + 18 * p;
    double t[24];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const bool pd = (fl >> k) & 1u, r = (rv >> k) & 1u;
        t[k] = pd ? bc[(size_t)k * NP] : (r ? ic[(size_t)k * NP] : xp[k]);                 // u, θ
        t[6 + k] = pd ? 0.0 : (r ? xp[k] : ic[(size_t)(12 + k) * NP]);                      // Vdot, Ωdot (relative)
        t[12 + k] = (!pd && !r) ? 1.0 : 0.0;
        t[18 + k] = (!pd && r) ? 1.0 : 0.0;
    }
    q.u = mk(t[0], t[1], t[2]); q.th = mk(t[3], t[4], t[5]); q.Vd = mk(t[6], t[7], t[8]); q.Od = mk(t[9], t[10], t[11]);
    q.mu = mk(t[12], t[13], t[14]); q.mth = mk(t[15], t[16], t[17]); q.mv = mk(t[18], t[19], t[20]); q.mo = mk(t[21], t[22], t[23]);
    q.Vr = ldv(ic + (size_t)6 * NP, NP); q.Or = ldv(ic + (size_t)9 * NP, NP);
    q.ud = mk(xp[6], xp[7], xp[8]); q.thd = mk(xp[9], xp[10], xp[11]);
    q.fl = fl; q.rv = rv;
}

// structural damping with udot, θdot as states (element.jl:333-368, 952-1012)
struct InitDamp {
    v3 gamdot, kapdot;
    m3 g_u1, g_th, g_ud1, k_th1, k_th2, k_thd1, k_thd2;      // coefficients applied; γ_u2 = -γ_u1, γ_u2dot = -γ_u1dot, γ_θ1 = γ_θ2
};
template <bool WANT_J>
__device__ __noinline__ void initial_damping_terms(const m3& Cab, const m3& CtCab, const Rot& R, v3 th, v3 dth, v3 du, v3 dud, v3 dthd, v3 thed,
                                                   v3 W, double L, v3 mua, v3 mub, InitDamp& o) {
    RotQ RQ;
    if (WANT_J) rot_Q_dQ(th, RQ); else RQ.Q = rot_Q(th);
    const m3 dQ = rot_dQ(th, dth, RQ.Q);
    const v3 Cab1 = mk(Cab.a[0], Cab.a[3], Cab.a[6]);
    const v3 gsrc = dud - cross(W, du) - L * cross(W, Cab1);
    o.gamdot = tmul(CtCab, gsrc);
    o.kapdot = tmul(Cab, RQ.Q * dthd + dQ * thed);
    if (!WANT_J) return;
    const m3 At = transpose(CtCab);
    o.g_u1 = -rowscale(At * tilde(W), mua);                   // γdot_u1 = +CtCab' W~
    o.g_th = -0.5 * rowscale(tmul(Cab, mul3c(R, gsrc)), mua);
    o.g_ud1 = rowscale(At, mua);                              // γdot_u1dot = -CtCab'
    m3 D1, D2, D3;
    rot_dQ_dth(th, dth, RQ, D1, D2, D3);
    const m3 t1 = tmul(Cab, cols(RQ.Q1 * dthd, RQ.Q2 * dthd, RQ.Q3 * dthd));
    const m3 t2 = tmul(Cab, cols(D1 * thed, D2 * thed, D3 * thed));
    const m3 t3 = tmul(Cab, thed.x * RQ.Q1 + thed.y * RQ.Q2 + thed.z * RQ.Q3);
    const m3 half = 0.5 * (t1 + t2);
    o.k_th1 = -rowscale(half - t3, mub);
    o.k_th2 = -rowscale(half + t3, mub);
    const m3 CQ = tmul(Cab, RQ.Q), CdQ = tmul(Cab, dQ);
    o.k_thd1 = -rowscale(0.5 * CdQ - CQ, mub);
    o.k_thd2 = -rowscale(0.5 * CdQ + CQ, mub);
}

// derivatives of tf = ωb x P + Pdot and tm = ωb x H + V x P + Hdot of a body carried in frame A with respect to u, θ, udot, Vdot, Ωdot
struct InitInertiaJac { m3 Fu, Fth, Fud, FVd, FOd, Mu, Mth, Mud, MVd, MOd; };
__device__ __noinline__ void initial_inertia_jacobians(const m3& A, const m3& Cab, const Rot& R, const m3& m11, const m3& m12, const m3& m21,
                                                       const m3& m22, v3 V, v3 Om, v3 Vdot, v3 Omdot, v3 wb, v3 alb, v3 Pm, InitInertiaJac& o) {
    InertiaJac ij;
    inertia_jacobians<1>(A, Cab, R, m11, m12, m21, m22, V, Om, Vdot, Omdot, wb, alb, Pm, 0.0, ij);
    const m3 At = transpose(A), wbt = tilde(wb), albt = tilde(alb);
    const m3 A11 = (A * m11) * At, A12 = (A * m12) * At, A21 = (A * m21) * At, A22 = (A * m22) * At;
    o.Fu = ij.FV * wbt + A11 * albt;      // (ω~b P_V + Pdot_V) V_u + Pdot_Vdot Vdot_u
    o.Mu = ij.MV * wbt + A21 * albt;
    o.Fth = ij.Fth; o.Mth = ij.Mth;
    o.Fud = A11 * wbt; o.Mud = A21 * wbt; // Pdot_Vdot Vdot_udot
    o.FVd = A11; o.FOd = A12; o.MVd = A21; o.MOd = A22;
}

template <bool WANT_J>
__device__ void initial_station(const InitArgs& IA, int b, int i, double* smem) {
    const DynArgs& A = IA.d;
    const StaticArgs& S = A.s;
    const int N = S.N, NP = N + 1;
    double* xch = smem;
    double* tiles = smem + (size_t)blockDim.x * GXB_XCH;
    const double fs = S.fs[b];
    const double rfs = 1.0 / fs;
    const double* x = S.x + (size_t)b * S.n;
    double* r = S.r + (size_t)b * S.n;
    double* J = WANT_J ? S.Jrow + (size_t)b * S.n * GXB_DROWW : nullptr;
    const bool has_elem = i < N;
    const bool valid = i <= N;

    v3 g = mk(0, 0, 0);
    if (S.has_grav && S.grav) g = mk(S.grav[3 * b], S.grav[3 * b + 1], S.grav[3 * b + 2]);
    const double* mo = A.motion + 12 * (size_t)b;
    const v3 vb = mk(mo[0], mo[1], mo[2]), wb = mk(mo[3], mo[4], mo[5]);
    v3 ab, alb;
    body_accelerations(A, x, mo, ab, alb);
    const m3 wbt = tilde(wb);

    // ---- point i ----
    InitPoint p1;
    p1.u = p1.th = p1.Vr = p1.Or = p1.Vd = p1.Od = p1.ud = p1.thd = mk(0, 0, 0);
    p1.mu = p1.mth = p1.mv = p1.mo = mk(0, 0, 0); p1.fl = 0; p1.rv = 0;
    v3 Fp = mk(0, 0, 0), Mp = Fp, dxp = Fp, thq = Fp;         // thq: the θ the follower loads are rotated with (see header)
    bool follower = false;
    if (valid) {
        init_point_load(IA, b, i, x, p1);
        const unsigned fl1 = p1.fl;
        const double* bc = S.bcS + ((size_t)b * 18) * NP + i;
        const double* xp = x + 18 * i;
        dxp = ldv(A.pxS + ((size_t)b * 3) * NP + i, NP);
        thq = mk((fl1 & 8) ? bc[3 * NP] : xp[3], (fl1 & 16) ? bc[4 * NP] : xp[4], (fl1 & 32) ? bc[5 * NP] : xp[5]);
        v3 F = ldv(bc + 6 * NP, NP), M = ldv(bc + 9 * NP, NP), Ff = ldv(bc + 12 * NP, NP), Mf = ldv(bc + 15 * NP, NP);
        follower = (Ff.x != 0.0) | (Ff.y != 0.0) | (Ff.z != 0.0) | (Mf.x != 0.0) | (Mf.y != 0.0) | (Mf.z != 0.0);
        if (follower) { m3 C = rot_C(thq); F = F + tmul(C, Ff); M = M + tmul(C, Mf); }
        Fp = mk((fl1 & 64) ? F.x : xp[0] * fs, (fl1 & 128) ? F.y : xp[1] * fs, (fl1 & 256) ? F.z : xp[2] * fs);
        Mp = mk((fl1 & 512) ? M.x : xp[3] * fs, (fl1 & 1024) ? M.y : xp[4] * fs, (fl1 & 2048) ? M.z : xp[5] * fs);
    }
    const v3 u1 = p1.u, th1 = p1.th;
    const unsigned fl1 = p1.fl;

    // ---- element i ----
    const double* es = S.elemS + ((size_t)b * S.elem_nf) * N + (has_elem ? i : 0);
    double L = 0.0;
    m3 Cab = zero33(), CtCab = zero33();
    Rot R;
    InitPoint p2;
    p2.mu = p2.mth = p2.mv = p2.mo = mk(0, 0, 0);
    v3 th = mk(0, 0, 0), F = th, M = th, Le1g = th, kap = th, V = th, Om = th, Vdot = th, Omdot = th, Pm = th, H = th;
    v3 ru = th, rth = th, F1 = th, M1 = th, F2 = th, M2 = th;
    v3 f1f = th, f2f = th, m1f = th, m2f = th, CtCabTg = th;
    InitDamp dmp;
    if (has_elem) {
        L = es[0];
        Cab = ldm(es + N, N);
        init_point_load(IA, b, i + 1, x, p2);
        const double* xe = x + 18 * i + 12;
        F = mk(xe[0] * fs, xe[1] * fs, xe[2] * fs); M = mk(xe[3] * fs, xe[4] * fs, xe[5] * fs);
        th = 0.5 * (th1 + p2.th);
        const v3 u = 0.5 * (u1 + p2.u);
        const v3 W = 0.5 * (p1.Or + p2.Or);                     // Ω - ωb
        const v3 dxe = ldv(A.exS + ((size_t)b * 3) * N + i, N);
        V = 0.5 * (p1.Vr + p2.Vr) + vb + cross(wb, dxe + u);
        Om = W + wb;
        if (WANT_J) rot_C_dC(th, R); else R.C = rot_C(th);
        CtCab = tmul(R.C, Cab);
        v3 gam;
        {
            m3 S11 = ldm(es + 10 * N, N), S12 = ldm(es + 19 * N, N);
            gam = S11 * F + S12 * M;
        }
        {
            m3 S21 = ldm(es + 28 * N, N), S22 = ldm(es + 37 * N, N);
            kap = S21 * F + S22 * M;
        }
        if (A.damping) {
            const v3 mua = ldv(es + 82 * N, N), mub = ldv(es + 85 * N, N);
            initial_damping_terms<WANT_J>(Cab, CtCab, R, th, p2.th - th1, p2.u - u1, p2.ud - p1.ud, p2.thd - p1.thd, 0.5 * (p1.thd + p2.thd),
                                          W, L, mua, mub, dmp);
            gam = gam - mk(mua.x * dmp.gamdot.x, mua.y * dmp.gamdot.y, mua.z * dmp.gamdot.z);
            kap = kap - mk(mub.x * dmp.kapdot.x, mub.y * dmp.kapdot.y, mub.z * dmp.kapdot.z);
        }
        Le1g = mk(L + gam.x, gam.y, gam.z);
        v3 Cab1 = mk(Cab.a[0], Cab.a[3], Cab.a[6]);
        ru = (p2.u - u1) - (CtCab * Le1g - L * Cab1);
        rth = (p2.th - th1) - rot_Qinv_from_c(rot_scaling(th) * th) * (Cab * kap);
        v3 CF = CtCab * F;
        v3 t1 = CtCab * M, t2 = 0.5 * (CtCab * cross(Le1g, F));
        F1 = CF; F2 = CF; M1 = t1 + t2; M2 = t1 - t2;
        if (S.has_grav) {
            CtCabTg = tmul(CtCab, g);
            v3 tg = -(CtCab * (ldm(es + 46 * N, N) * CtCabTg));
            F1 = F1 - 0.5 * tg; F2 = F2 + 0.5 * tg;
            tg = -(CtCab * (ldm(es + 64 * N, N) * CtCabTg));
            M1 = M1 - 0.5 * tg; M2 = M2 + 0.5 * tg;
        }
        if (S.dlS) {
            const double* d = S.dlS + ((size_t)b * 24) * N + i;
            f1f = ldv(d + 12 * N, N); f2f = ldv(d + 15 * N, N); m1f = ldv(d + 18 * N, N); m2f = ldv(d + 21 * N, N);
            F1 = F1 + (ldv(d, N) + tmul(R.C, f1f));
            F2 = F2 - (ldv(d + 3 * N, N) + tmul(R.C, f2f));
            M1 = M1 + (ldv(d + 6 * N, N) + tmul(R.C, m1f));
            M2 = M2 - (ldv(d + 9 * N, N) + tmul(R.C, m2f));
        }
        // inertial loads (element.jl:283-330, 1901-1919)
        Vdot = 0.5 * (p1.Vd + p2.Vd) + ab + cross(alb, dxe) + cross(alb, u) + cross(wb, 0.5 * (p1.ud + p2.ud));
        Omdot = 0.5 * (p1.Od + p2.Od) + alb;
        {
            const m3 m11 = ldm(es + 46 * N, N), m12 = ldm(es + 55 * N, N), m21 = ldm(es + 64 * N, N), m22 = ldm(es + 73 * N, N);
            v3 tf, tm;
            inertia_loads<1>(CtCab, m11, m12, m21, m22, V, Om, Vdot, Omdot, wb, Pm, H, tf, tm);
            F1 = F1 - 0.5 * tf; F2 = F2 + 0.5 * tf;
            M1 = M1 - 0.5 * tm; M2 = M2 + 0.5 * tm;
        }
        double* mine = xch + (size_t)threadIdx.x * GXB_XCH;
        mine[0] = F2.x * rfs; mine[1] = F2.y * rfs; mine[2] = F2.z * rfs;
        mine[3] = M2.x * rfs; mine[4] = M2.y * rfs; mine[5] = M2.z * rfs;
    }
    __syncthreads();

    // ---- residual ----
    m3 Cp = zero33(), Qp = zero33();
    const bool pmass = S.pmS != nullptr;
    v3 V1 = mk(0, 0, 0), Om1 = V1, Vdp = V1, Odp = V1, Pp = V1;
    if (valid) {
        const double* prev = xch + (size_t)(i > 0 ? threadIdx.x - 1 : 0) * GXB_XCH;
        double* rp = r + 18 * i;
        Cp = rot_C(th1); Qp = rot_Qinv_from_c(rot_scaling(th1) * th1);
        V1 = p1.Vr + vb + cross(wb, dxp + u1);
        Om1 = p1.Or + wb;
        if (pmass) {       // point.jl:683-757
            const double* pm = S.pmS + ((size_t)b * 36) * NP + i;
            const m3 m11 = ldm(pm, NP), m12 = ldm(pm + 9 * NP, NP), m21 = ldm(pm + 18 * NP, NP), m22 = ldm(pm + 27 * NP, NP);
            Vdp = p1.Vd + ab + cross(alb, dxp + u1) + cross(wb, p1.ud);
            Odp = p1.Od + alb;
            const m3 Ap = transpose(Cp);
            v3 Hp, tf, tm;
            inertia_loads<1>(Ap, m11, m12, m21, m22, V1, Om1, Vdp, Odp, wb, Pp, Hp, tf, tm);
            Fp = Fp - tf; Mp = Mp - tm;
            if (S.has_grav) { const v3 Cg = Cp * g; Fp = Fp + tmul(Cp, m11 * Cg); Mp = Mp + tmul(Cp, m21 * Cg); }
        }
        double own[6] = {-Fp.x * rfs, -Fp.y * rfs, -Fp.z * rfs, -Mp.x * rfs, -Mp.y * rfs, -Mp.z * rfs};
        double sub[6] = {F1.x * rfs, F1.y * rfs, F1.z * rfs, M1.x * rfs, M1.y * rfs, M1.z * rfs};
        const double* ic = IA.icS + ((size_t)b * 18) * NP + i;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            double v = own[k];
            if (i > 0) v += prev[k];
            if (has_elem) v -= sub[k];
            if ((p1.rv >> (8 + k)) & 1u) v = x[18 * i + k] - ic[(size_t)(12 + k) * NP];      // Vdot_k - Vdot0_k   (system.jl:491-513)
            if (S.two_d && k >= 2 && k <= 4) v = x[18 * i + k];
            rp[k] = v;
        }
        // velocity rows (point.jl:1658-1666) with udot, θdot states
        v3 rV = V1 - vb - cross(wb, dxp) - cross(wb, u1) - p1.ud;
        v3 rO = Qp * (Cp * (Om1 - wb)) - p1.thd;
        double rv[6] = {rV.x, rV.y, rV.z, rO.x, rO.y, rO.z};
#pragma unroll
        for (int k = 0; k < 6; k++) rp[6 + k] = (S.two_d && k >= 2 && k <= 4) ? x[18 * i + 6 + k] : rv[k];
        if (has_elem) {
            double re[6] = {ru.x, ru.y, ru.z, rth.x, rth.y, rth.z};
#pragma unroll
            for (int k = 0; k < 6; k++) rp[12 + k] = (S.two_d && k >= 2 && k <= 4) ? x[18 * i + 12 + k] : re[k];
        }
    }
    if (!WANT_J) return;

    // ---- element inertial Jacobian pieces; "h" = the ¼-weighted terms that enter both end points with the same sign ----
    InitInertiaJac hj;
    if (has_elem) {
        const m3 m11 = ldm(es + 46 * N, N), m12 = ldm(es + 55 * N, N), m21 = ldm(es + 64 * N, N), m22 = ldm(es + 73 * N, N);
        initial_inertia_jacobians(CtCab, Cab, R, m11, m12, m21, m22, V, Om, Vdot, Omdot, wb, alb, Pm, hj);
        hj.Fu = 0.25 * hj.Fu; hj.Fth = 0.25 * hj.Fth; hj.Fud = 0.25 * hj.Fud; hj.FVd = 0.25 * hj.FVd; hj.FOd = 0.25 * hj.FOd;
        hj.Mu = 0.25 * hj.Mu; hj.Mth = 0.25 * hj.Mth; hj.Mud = 0.25 * hj.Mud; hj.MVd = 0.25 * hj.MVd; hj.MOd = 0.25 * hj.MOd;
    }
    // point-mass Jacobian pieces of point i
    InitInertiaJac pj;
    Rot Rp;
    if (valid && (pmass || follower)) rot_C_dC(th1, Rp);
    if (valid && pmass) {
        const double* pm = S.pmS + ((size_t)b * 36) * NP + i;
        const m3 m11 = ldm(pm, NP), m12 = ldm(pm + 9 * NP, NP), m21 = ldm(pm + 18 * NP, NP), m22 = ldm(pm + 27 * NP, NP);
        initial_inertia_jacobians(transpose(Rp.C), eye33(), Rp, m11, m12, m21, m22, V1, Om1, Vdp, Odp, wb, alb, Pp, pj);
    }
    const unsigned repl = (p1.rv >> 8) & 63u;

    // =================== pass P1: force rows of point i ===================
    m3 JFa = zero33(), JFb = zero33();
    if (has_elem) {
        m3 Fth = 0.5 * mul3t(R, Cab * F);
        JFa = Fth; JFb = Fth;
        if (S.has_grav) {
            m3 m11 = ldm(es + 46 * N, N);
            m3 Cthg = mul3c(R, g);
            m3 tg = -0.5 * (mul3t(R, Cab * (m11 * CtCabTg)) + CtCab * (m11 * tmul(Cab, Cthg)));
            JFa = JFa - 0.5 * tg; JFb = JFb + 0.5 * tg;
        }
        if (S.dlS) { JFa = JFa + 0.5 * mul3t(R, f1f); JFb = JFb - 0.5 * mul3t(R, f2f); }
        JFa = JFa - hj.Fth; JFb = JFb + hj.Fth;
    }
    GXB_PASS_BEGIN(c0)
        double* mytile = tiles + (size_t)(i - c0) * GXB_DTILE;
        const bool mine_in = valid && i >= c0 && i < c0 + cnt_;
        const bool next_in = has_elem && i + 1 >= c0 && i + 1 < c0 + cnt_;
        if (mine_in) dtile_clear(mytile);
        __syncthreads();
        if (next_in) {      // rows of the stop point: columns of point i (slots 0..11), of this element (12..17); own columns parked
            double* nt = tiles + (size_t)(i + 1 - c0) * GXB_DTILE;
            dblk(nt, 0, rfs * (colmask(hj.Fu, p1.mu) + colmask(hj.FVd, p1.mv)));
            dblk(nt, 3, rfs * (colmask(JFb, p1.mth) + colmask(hj.FOd, p1.mo)));
            dblk(nt, 6, rfs * hj.Fud);
            dblk(nt, 12, CtCab);                                  // F2_F
            dblk(nt, GXB_DPARK + 0, rfs * (colmask(hj.Fu, p2.mu) + colmask(hj.FVd, p2.mv)));
            dblk(nt, GXB_DPARK + 3, rfs * (colmask(JFb, p2.mth) + colmask(hj.FOd, p2.mo)));
            dblk(nt, GXB_DPARK + 6, rfs * hj.Fud);
        }
        __syncthreads();
        if (mine_in) {
            m3 FF = zero33();
            FF.a[0] = (fl1 & 64) ? 0.0 : -1.0; FF.a[4] = (fl1 & 128) ? 0.0 : -1.0; FF.a[8] = (fl1 & 256) ? 0.0 : -1.0;
            m3 bu = FF + dget(mytile, GXB_DPARK + 0);
            m3 bt = dget(mytile, GXB_DPARK + 3);
            m3 bd = dget(mytile, GXB_DPARK + 6);
            if (follower) {      // follower loads (point.jl:75-105), rotated with thq
                Rot Rq; rot_C_dC(thq, Rq);
                const double* bc = S.bcS + ((size_t)b * 18) * NP + i;
                const m3 Fth = rowmask(mul3t(Rq, ldv(bc + 12 * NP, NP)), fl1 & 64, fl1 & 128, fl1 & 256);
                bt = bt - rfs * colmask(Fth, p1.mth);
            }
            if (pmass) {
                m3 Fth = -pj.Fth;                                 // F_θ -= ω~b P_θ + Pdot_θ   (point.jl:1566)
                if (S.has_grav) {
                    const double* pm = S.pmS + ((size_t)b * 36) * NP + i;
                    const m3 m11 = ldm(pm, NP);
                    Fth = Fth + mul3t(Rp, m11 * (Rp.C * g)) + tmul(Rp.C, m11 * mul3c(Rp, g));
                }
                bt = bt - rfs * colmask(Fth, p1.mth) + rfs * colmask(pj.FOd, p1.mo);
                bu = bu + rfs * (colmask(pj.Fu, p1.mu) + colmask(pj.FVd, p1.mv));
                bd = bd + rfs * pj.Fud;
            }
            if (has_elem) {
                bu = bu + rfs * (colmask(hj.Fu, p1.mu) + colmask(hj.FVd, p1.mv));
                bt = bt + rfs * (colmask(hj.FOd, p1.mo) - colmask(JFa, p1.mth));
                bd = bd + rfs * hj.Fud;
                dblk(mytile, 30, -CtCab);                         // -F1_F
                dblk(mytile, 36, rfs * (colmask(hj.Fu, p2.mu) + colmask(hj.FVd, p2.mv)));
                dblk(mytile, 39, rfs * (colmask(hj.FOd, p2.mo) - colmask(JFa, p2.mth)));
                dblk(mytile, 42, rfs * hj.Fud);
            }
            dblk(mytile, 18, bu); dblk(mytile, 21, bt); dblk(mytile, 24, bd);
#pragma unroll
            for (int k = 0; k < 3; k++)
                if ((repl >> k) & 1u) {
#pragma unroll
                    for (int cc = 0; cc < GXB_DROWW; cc++) mytile[GXB_DTROW * k + cc] = (cc == 18 + k) ? 1.0 : 0.0;
                }
            if (S.two_d) dtile_planar(mytile, 0, 18);
        }
        __syncthreads();
        dtile_flush(tiles, J, c0, cnt_, 0);
        __syncthreads();
    GXB_PASS_END

    // =================== pass P2: moment rows of point i ===================
    m3 JMa = zero33(), JMb = zero33(), M1F = zero33(), hM = zero33();
    m3 dMu1 = zero33(), dMud1 = zero33();       // structural-damping parts (u2, u2dot carry the opposite sign)
    if (has_elem) {
        m3 tm1 = mul3t(R, Cab * M);
        m3 tm2 = 0.5 * mul3t(R, Cab * cross(Le1g, F));
        JMa = 0.5 * (tm1 + tm2);
        JMb = 0.5 * (tm1 - tm2);
        if (S.has_grav) {
            m3 m21 = ldm(es + 64 * N, N);
            m3 Cthg = mul3c(R, g);
            m3 tg = -0.5 * (mul3t(R, Cab * (m21 * CtCabTg)) + CtCab * (m21 * tmul(Cab, Cthg)));
            JMa = JMa - 0.5 * tg; JMb = JMb + 0.5 * tg;
        }
        if (S.dlS) { JMa = JMa + 0.5 * mul3t(R, m1f); JMb = JMb - 0.5 * mul3t(R, m2f); }
        JMa = JMa - hj.Mth; JMb = JMb + hj.Mth;
        m3 CtF = CtCab * tilde(F);
        if (A.damping) {      // M1_x -= ½ CtCab F~ γ_x,  M2_x += ½ CtCab F~ γ_x   (element.jl:2086-2110)
            const m3 tF = 0.5 * CtF;
            const m3 dth_ = tF * dmp.g_th;
            JMa = JMa - dth_; JMb = JMb + dth_;
            dMu1 = tF * dmp.g_u1; dMud1 = tF * dmp.g_ud1;
        }
        M1F = 0.5 * (CtCab * tilde(Le1g) - CtF * ldm(es + 10 * N, N));
        hM = 0.5 * (CtF * ldm(es + 19 * N, N));
    }
    GXB_PASS_BEGIN(c0)
        double* mytile = tiles + (size_t)(i - c0) * GXB_DTILE;
        const bool mine_in = valid && i >= c0 && i < c0 + cnt_;
        const bool next_in = has_elem && i + 1 >= c0 && i + 1 < c0 + cnt_;
        if (mine_in) dtile_clear(mytile);
        __syncthreads();
        if (next_in) {
            double* nt = tiles + (size_t)(i + 1 - c0) * GXB_DTILE;
            dblk(nt, 0, rfs * (colmask(hj.Mu + dMu1, p1.mu) + colmask(hj.MVd, p1.mv)));
            dblk(nt, 3, rfs * (colmask(JMb, p1.mth) + colmask(hj.MOd, p1.mo)));
            dblk(nt, 6, rfs * (hj.Mud + dMud1));
            dblk(nt, 12, -M1F);                                   // M2_F
            dblk(nt, 15, CtCab + hM);                             // M2_M
            dblk(nt, GXB_DPARK + 0, rfs * (colmask(hj.Mu - dMu1, p2.mu) + colmask(hj.MVd, p2.mv)));
            dblk(nt, GXB_DPARK + 3, rfs * (colmask(JMb, p2.mth) + colmask(hj.MOd, p2.mo)));
            dblk(nt, GXB_DPARK + 6, rfs * (hj.Mud - dMud1));
        }
        __syncthreads();
        if (mine_in) {
            m3 MM = zero33();
            MM.a[0] = (fl1 & 512) ? 0.0 : -1.0; MM.a[4] = (fl1 & 1024) ? 0.0 : -1.0; MM.a[8] = (fl1 & 2048) ? 0.0 : -1.0;
            m3 bu = dget(mytile, GXB_DPARK + 0);
            m3 bt = MM + dget(mytile, GXB_DPARK + 3);
            m3 bd = dget(mytile, GXB_DPARK + 6);
            if (follower) {
                Rot Rq; rot_C_dC(thq, Rq);
                const double* bc = S.bcS + ((size_t)b * 18) * NP + i;
                const m3 Mth = rowmask(mul3t(Rq, ldv(bc + 15 * NP, NP)), fl1 & 512, fl1 & 1024, fl1 & 2048);
                bt = bt - rfs * colmask(Mth, p1.mth);
            }
            if (pmass) {
                m3 Mth = -pj.Mth;
                if (S.has_grav) {
                    const double* pm = S.pmS + ((size_t)b * 36) * NP + i;
                    const m3 m21 = ldm(pm + 18 * NP, NP);
                    Mth = Mth + mul3t(Rp, m21 * (Rp.C * g)) + tmul(Rp.C, m21 * mul3c(Rp, g));
                }
                bt = bt - rfs * colmask(Mth, p1.mth) + rfs * colmask(pj.MOd, p1.mo);
                bu = bu + rfs * (colmask(pj.Mu, p1.mu) + colmask(pj.MVd, p1.mv));
                bd = bd + rfs * pj.Mud;
            }
            if (has_elem) {
                bu = bu + rfs * (colmask(hj.Mu + dMu1, p1.mu) + colmask(hj.MVd, p1.mv));
                bt = bt + rfs * (colmask(hj.MOd, p1.mo) - colmask(JMa, p1.mth));
                bd = bd + rfs * (hj.Mud + dMud1);
                dblk(mytile, 30, -M1F);                           // -M1_F
                dblk(mytile, 33, hM - CtCab);                     // -M1_M
                dblk(mytile, 36, rfs * (colmask(hj.Mu - dMu1, p2.mu) + colmask(hj.MVd, p2.mv)));
                dblk(mytile, 39, rfs * (colmask(hj.MOd, p2.mo) - colmask(JMa, p2.mth)));
                dblk(mytile, 42, rfs * (hj.Mud - dMud1));
            }
            dblk(mytile, 18, bu); dblk(mytile, 21, bt); dblk(mytile, 24, bd);
#pragma unroll
            for (int k = 0; k < 3; k++)
                if ((repl >> (3 + k)) & 1u) {
#pragma unroll
                    for (int cc = 0; cc < GXB_DROWW; cc++) mytile[GXB_DTROW * k + cc] = (cc == 21 + k) ? 1.0 : 0.0;
                }
            if (S.two_d) dtile_planar(mytile, 3, 18);
        }
        __syncthreads();
        dtile_flush(tiles, J, c0, cnt_, 3);
        __syncthreads();
    GXB_PASS_END

    // =================== passes V1/V2: velocity rows of point i (block 3i+1): own u/θ at slots 12..17, udot/θdot at 18..23 ===================
    GXB_PASS_BEGIN(c0)
        double* mytile = tiles + (size_t)(i - c0) * GXB_DTILE;
        const bool mine_in = valid && i >= c0 && i < c0 + cnt_;
        if (mine_in) {
            dtile_clear(mytile);
            dblk(mytile, 18, -eye33());                           // rV_udot   (point.jl:1697; u drops out of rV)
            if (S.two_d) dtile_planar(mytile, 0, 18);
        }
        __syncthreads();
        dtile_flush(tiles, J, c0, cnt_, 6);
        __syncthreads();
    GXB_PASS_END
    GXB_PASS_BEGIN(c0)
        double* mytile = tiles + (size_t)(i - c0) * GXB_DTILE;
        const bool mine_in = valid && i >= c0 && i < c0 + cnt_;
        if (mine_in) {
            dtile_clear(mytile);
            Rot Rv; rot_C_dC(th1, Rv);
            const v3 dO = Om1 - wb;
            m3 Qi;
            const m3 rOth = rot_Qinv_dQinv_b(th1, Rv.C * dO, Qi) + Qi * mul3c(Rv, dO);
            dblk(mytile, 15, colmask(rOth, p1.mth));              // rΩ_θ θ_θ
            dblk(mytile, 21, -eye33());                           // rΩ_θdot
            if (S.two_d) dtile_planar(mytile, 3, 18);
        }
        __syncthreads();
        dtile_flush(tiles, J, c0, cnt_, 9);
        __syncthreads();
    GXB_PASS_END

    // =================== passes E1/E2: compatibility rows of element i (block 3i+2) ===================
    // slots: point i u/θ 6..11, udot/θdot 12..17, element 18..23, point i+1 u/θ 24..29, udot/θdot 30..35
    GXB_PASS_BEGIN(c0)
        double* mytile = tiles + (size_t)(i - c0) * GXB_DTILE;
        const bool mine_in = has_elem && i >= c0 && i < c0 + cnt_;
        if (mine_in) {
            dtile_clear(mytile);
            m3 ruth = -0.5 * mul3t(R, Cab * Le1g);
            m3 ruu1 = -eye33(), ruu2 = eye33();
            if (A.damping) {      // initial_compatibility_jacobians (element.jl:1523-1548)
                const m3 Agu = CtCab * dmp.g_u1;
                ruu1 = ruu1 - Agu; ruu2 = ruu2 + Agu;
                ruth = ruth - CtCab * dmp.g_th;
                const m3 Agd = CtCab * dmp.g_ud1;
                dblk(mytile, 12, -Agd); dblk(mytile, 30, Agd);    // ru_u1dot, ru_u2dot
            }
            dblk(mytile, 6, colmask(ruu1, p1.mu));
            dblk(mytile, 9, colmask(ruth, p1.mth));
            dblk(mytile, 18, (-fs) * (CtCab * ldm(es + 10 * N, N)));
            dblk(mytile, 21, (-fs) * (CtCab * ldm(es + 19 * N, N)));
            dblk(mytile, 24, colmask(ruu2, p2.mu));
            dblk(mytile, 27, colmask(ruth, p2.mth));
            if (S.two_d) dtile_planar(mytile, 0, 18);
        }
        __syncthreads();
        dtile_flush(tiles, J, c0, min(cnt_, N - c0), 12);
        __syncthreads();
    GXB_PASS_END
    GXB_PASS_BEGIN(c0)
        double* mytile = tiles + (size_t)(i - c0) * GXB_DTILE;
        const bool mine_in = has_elem && i >= c0 && i < c0 + cnt_;
        if (mine_in) {
            dtile_clear(mytile);
            m3 Qinv;
            m3 hd = -0.5 * rot_Qinv_dQinv_b(th, Cab * kap, Qinv);
            m3 QC = Qinv * Cab;
            m3 rt1 = hd - eye33(), rt2 = hd + eye33();
            if (A.damping) {
                rt1 = rt1 - QC * dmp.k_th1; rt2 = rt2 - QC * dmp.k_th2;
                dblk(mytile, 15, -(QC * dmp.k_thd1)); dblk(mytile, 33, -(QC * dmp.k_thd2));   // rθ_θ1dot, rθ_θ2dot
            }
            dblk(mytile, 9, colmask(rt1, p1.mth));
            dblk(mytile, 18, (-fs) * (QC * ldm(es + 28 * N, N)));
            dblk(mytile, 21, (-fs) * (QC * ldm(es + 37 * N, N)));
            dblk(mytile, 27, colmask(rt2, p2.mth));
            if (S.two_d) dtile_planar(mytile, 3, 18);
        }
        __syncthreads();
        dtile_flush(tiles, J, c0, min(cnt_, N - c0), 15);
        __syncthreads();
    GXB_PASS_END
}
