// Fused static residual + Jacobian assembly, one thread per "station" (point i and element i of a chain), one CTA per
// instance.  Replaces the per-iteration calls
//     static_system_residual!  src/system.jl:400-418   -> static_point_residual! (point.jl:2113), static_element_residual! (element.jl:2995)
//     static_system_jacobian!  src/system.jl:663-683   -> static_point_jacobian! (point.jl:2321), static_element_jacobian! (element.jl:3182)
// The reference evaluates the element/point properties twice (once per call) and inserts 3x3 blocks into a CSC matrix;
// here both are produced in one pass straight into the row-block storage the batched LU consumes (see gxb_lu.cuh):
//
//   Jrow[b][r][30] : row r (natural SystemIndices order) holds the 5 column 6-blocks B-2..B+2 around its own 6-block
//   B = r / 6.  Chain layout [p0 e0 p1 e1 ... pN]: point i is block 2i, element i is block 2i+1.
//
// The rows are produced in four passes of three rows per station (point F rows, point M rows, element ru rows, element
// rθ rows).  In each pass every thread deposits its 3x3 blocks into a shared-memory tile [station][3 rows x 30 slots];
// the tile is then streamed to HBM by the whole CTA with 16-byte, fully coalesced stores (per-thread strided stores cost
// 3x the L2 write sectors and throttled the LSU in the first version of this kernel, profiles/r1_v1_*).
//
// Summation order in point rows is the reference's (point term assigned first, then element i-1 `+=`, then element i
// `-=`; system.jl:668-676, element.jl:2526-2549): element i-1's contribution to the shared θ-block is parked in the
// structurally-zero slots 0..2 of the tile row and folded in by the owning thread, so results are deterministic
// (no atomics).
#pragma once
#include "gxb_math.cuh"

#define GXB_ROWW 30      // doubles per stored Jacobian row (static): 5 blocks x 6
#define GXB_TILE 94      // tile stride per station: 3 rows x 30 slots + 4 pad (16-byte aligned, 2-way bank conflicts at most)
#define GXB_XCH 6        // residual exchange record per thread

struct StaticArgs {
    int N;                 // elements in the chain (points = N+1)
    int n;                 // 12 N + 6
    int two_d;
    int has_grav;          // gravity != 0 for some instance (element mass / point mass terms evaluated)
    const double* elemS;   // [b][82][N]  : L, Cab[9], S11,S12,S21,S22 (=L*compliance, 3x3 row-major each), m11,m12,m21,m22 (=L*mass)
    int elem_nf;           // number of element fields stored
    const double* bcS;     // [b][18][N+1]
    const double* dlS;     // [b][24][N] or null
    const double* pmS;     // [b][36][N+1] (4 3x3 row-major blocks m11,m12,m21,m22) or null
    const double* grav;    // [b][3] or null
    const double* fs;      // [b]
    const unsigned short* flags;  // [N+1] : bits 0-5 pd, bits 6-11 pl
    const double* x;       // [b][n] state to evaluate at
    double* r;             // [b][n] residual out
    double* Jrow;          // [b][n][30] or null (residual only)
    const int* status;     // [b] or null; when given only instances with status == want_status are evaluated
    int want_status;
};

GXD v3 ldv(const double* p, int stride) { return mk(p[0], p[stride], p[2 * stride]); }
GXD m3 ldm(const double* p, int stride) { m3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.a[i] = p[i * stride]; return r; }

// deposit a 3x3 block into a tile (3 rows x 30 slots) at slot s
GXD void tblk(double* t, int s, const m3& B) {
#pragma unroll
    for (int i = 0; i < 3; i++) { t[30 * i + s] = B.a[3 * i]; t[30 * i + s + 1] = B.a[3 * i + 1]; t[30 * i + s + 2] = B.a[3 * i + 2]; }
}
GXD m3 tget(const double* t, int s) { m3 B;
#pragma unroll
    for (int i = 0; i < 3; i++) { B.a[3 * i] = t[30 * i + s]; B.a[3 * i + 1] = t[30 * i + s + 1]; B.a[3 * i + 2] = t[30 * i + s + 2]; }
    return B; }
GXD m3 mul3t(const Rot& R, v3 w) { return cols(tmul(R.C1, w), tmul(R.C2, w), tmul(R.C3, w)); }   // mul3(C_θ1', C_θ2', C_θ3', w)
// zero rows of M whose flag bit is clear (loads that are state variables have no θ-dependence, point.jl:86-105)
GXD m3 rowmask(const m3& M, unsigned f0, unsigned f1, unsigned f2) { m3 r = M;
#pragma unroll
    for (int k = 0; k < 3; k++) { if (!f0) r.a[k] = 0.0; if (!f1) r.a[3 + k] = 0.0; if (!f2) r.a[6 + k] = 0.0; }
    return r; }

// zero this thread's tile, CTA-wide barrier afterwards is the caller's job
GXD void tile_clear(double* t) {
#pragma unroll
    for (int j = 0; j < 90; j += 2) *reinterpret_cast<double2*>(t + j) = make_double2(0.0, 0.0);
}
// stream the tiles of stations [0, nst) to rows 12*st + roff .. +2 of the instance's row-block storage
GXD void tile_flush(const double* tiles, double* J, int nst, int roff) {
    for (int c = threadIdx.x; c < nst * 45; c += blockDim.x) {
        const int st = c / 45, q = c - 45 * st;
        const double2 v = *reinterpret_cast<const double2*>(tiles + (size_t)st * GXB_TILE + 2 * q);
        *reinterpret_cast<double2*>(J + (size_t)(12 * st + roff) * GXB_ROWW + 2 * q) = v;
    }
}
// The same flush by TMA: a station's tile (3 rows x 30 slots) is 720 contiguous bytes in shared memory AND in the row-block storage, so each
// station thread issues ONE bulk store (cp.async.bulk shared -> global) instead of the CTA looping over 16-byte pieces (a quarter of the
// kernel's instructions).  Contains the barriers of the pass: writes -> proxy fence -> barrier -> bulk store -> source read complete -> barrier.
#ifndef GXB_ASM_TMA_FLUSH
#define GXB_ASM_TMA_FLUSH 1
#endif
GXD void tile_flush_bulk(const double* mytile, double* J, int st, int roff, bool active) {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (active) {
        const unsigned src = (unsigned)__cvta_generic_to_shared(mytile);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                     ::"l"(J + (size_t)(12 * st + roff) * GXB_ROWW), "r"(src), "n"(720) : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
    }
    __syncthreads();
}
// planar analysis (system.jl:1072-1103): rows 2,3,4 of every 6-slot become  x_k = 0
GXD void tile_planar(double* t, int first_row_in_slot) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int k = first_row_in_slot + i;
        if (k >= 2 && k <= 4) {
#pragma unroll
            for (int c = 0; c < 30; c++) t[30 * i + c] = (c == 12 + k) ? 1.0 : 0.0;
        }
    }
}

template <bool WANT_J>
__device__ void static_station(const StaticArgs& A, int b, int i, double* smem) {
    const int N = A.N, NP = N + 1;
    double* xch = smem;                                             // [blockDim][6]
    double* tiles = smem + (size_t)blockDim.x * GXB_XCH;            // [N + 1][94]   (WANT_J only)
    double* mytile = tiles + (size_t)threadIdx.x * GXB_TILE;
    double* nexttile = mytile + GXB_TILE;                           // station i+1 (valid when has_elem)
    const double fs = A.fs[b];
    const double rfs = 1.0 / fs;   // the reference divides by fs; fs is a power of two by default so this is exact
    const double* x = A.x + (size_t)b * A.n;
    double* r = A.r + (size_t)b * A.n;
    double* J = WANT_J ? A.Jrow + (size_t)b * A.n * GXB_ROWW : nullptr;
    const bool has_elem = i < N;
    const bool valid = i <= N;

    v3 g = mk(0, 0, 0);
    if (A.has_grav && A.grav) g = mk(A.grav[3 * b], A.grav[3 * b + 1], A.grav[3 * b + 2]);

    // ---- point i: displacement / load selection (point.jl:20-39, 147-167) ----
    v3 u1 = mk(0, 0, 0), th1 = u1, Fp = u1, Mp = u1;
    v3 mu1 = mk(1, 1, 1), mth1 = mu1;                       // u_u, θ_θ diagonal masks (1 = free)
    unsigned fl1 = 0;
    bool point_rot = false;                                 // follower loads or point-mass gravity present at this point
    if (valid) {
        fl1 = A.flags[i];
        const double* bc = A.bcS + ((size_t)b * 18) * NP + i;
        const double* xp = x + 12 * i;
        u1 = mk((fl1 & 1) ? bc[0] : xp[0], (fl1 & 2) ? bc[NP] : xp[1], (fl1 & 4) ? bc[2 * NP] : xp[2]);
        th1 = mk((fl1 & 8) ? bc[3 * NP] : xp[3], (fl1 & 16) ? bc[4 * NP] : xp[4], (fl1 & 32) ? bc[5 * NP] : xp[5]);
        mu1 = mk((fl1 & 1) ? 0.0 : 1.0, (fl1 & 2) ? 0.0 : 1.0, (fl1 & 4) ? 0.0 : 1.0);
        mth1 = mk((fl1 & 8) ? 0.0 : 1.0, (fl1 & 16) ? 0.0 : 1.0, (fl1 & 32) ? 0.0 : 1.0);
        v3 F = ldv(bc + 6 * NP, NP), M = ldv(bc + 9 * NP, NP), Ff = ldv(bc + 12 * NP, NP), Mf = ldv(bc + 15 * NP, NP);
        const bool follower = (Ff.x != 0.0) | (Ff.y != 0.0) | (Ff.z != 0.0) | (Mf.x != 0.0) | (Mf.y != 0.0) | (Mf.z != 0.0);
        const bool pmass = A.has_grav && A.pmS != nullptr;
        point_rot = follower || pmass;
        v3 Fg = mk(0, 0, 0), Mg = Fg;
        if (point_rot) {
            m3 C = rot_C(th1);
            F = F + tmul(C, Ff);                   // F + C' Ff
            M = M + tmul(C, Mf);
            if (pmass) {                           // point.jl:1422-1431
                const double* pm = A.pmS + ((size_t)b * 36) * NP + i;
                v3 Cg = C * g;
                Fg = tmul(C, ldm(pm, NP) * Cg); Mg = tmul(C, ldm(pm + 18 * NP, NP) * Cg);
            }
        }
        // loads: prescribed value where pl, else the state variable times fs  (NaN-safe: selected by flag)
        Fp = mk((fl1 & 64) ? F.x : xp[0] * fs, (fl1 & 128) ? F.y : xp[1] * fs, (fl1 & 256) ? F.z : xp[2] * fs);
        Mp = mk((fl1 & 512) ? M.x : xp[3] * fs, (fl1 & 1024) ? M.y : xp[4] * fs, (fl1 & 2048) ? M.z : xp[5] * fs);
        Fp = Fp + Fg; Mp = Mp + Mg;
    }

    // ---- element i core quantities (element.jl:62-107) ----
    const double* es = A.elemS + ((size_t)b * A.elem_nf) * N + (has_elem ? i : 0);
    double L = 0.0;
    m3 Cab = zero33(), CtCab = zero33();
    Rot R;
    v3 th = mk(0, 0, 0), F = th, M = th, Le1g = th, kap = th, u2 = th, th2 = th;
    v3 mth2 = mk(1, 1, 1), mu2 = mth2;
    v3 ru = mk(0, 0, 0), rth = ru, F1 = ru, M1 = ru, F2 = ru, M2 = ru;
    v3 f1f = ru, f2f = ru, m1f = ru, m2f = ru, CtCabTg = ru;
    if (has_elem) {
        L = es[0];
        Cab = ldm(es + N, N);
        const unsigned fl2 = A.flags[i + 1];
        const double* bc2 = A.bcS + ((size_t)b * 18) * NP + i + 1;
        const double* xq = x + 12 * (i + 1);
        u2 = mk((fl2 & 1) ? bc2[0] : xq[0], (fl2 & 2) ? bc2[NP] : xq[1], (fl2 & 4) ? bc2[2 * NP] : xq[2]);
        th2 = mk((fl2 & 8) ? bc2[3 * NP] : xq[3], (fl2 & 16) ? bc2[4 * NP] : xq[4], (fl2 & 32) ? bc2[5 * NP] : xq[5]);
        mu2 = mk((fl2 & 1) ? 0.0 : 1.0, (fl2 & 2) ? 0.0 : 1.0, (fl2 & 4) ? 0.0 : 1.0);
        mth2 = mk((fl2 & 8) ? 0.0 : 1.0, (fl2 & 16) ? 0.0 : 1.0, (fl2 & 32) ? 0.0 : 1.0);
        const double* xe = x + 12 * i + 6;
        F = mk(xe[0] * fs, xe[1] * fs, xe[2] * fs); M = mk(xe[3] * fs, xe[4] * fs, xe[5] * fs);
        th = 0.5 * (th1 + th2);
        if (WANT_J) rot_C_dC(th, R); else R.C = rot_C(th);
        CtCab = tmul(R.C, Cab);
        {
            m3 S11 = ldm(es + 10 * N, N), S12 = ldm(es + 19 * N, N);
            v3 gam = S11 * F + S12 * M;
            Le1g = mk(L + gam.x, gam.y, gam.z);                // L e1 + γ
        }
        {
            m3 S21 = ldm(es + 28 * N, N), S22 = ldm(es + 37 * N, N);
            kap = S21 * F + S22 * M;
        }
        v3 Cab1 = mk(Cab.a[0], Cab.a[3], Cab.a[6]);            // Cab e1
        // compatibility (element.jl:1478-1490) and resultants (element.jl:1860-1893)
        ru = (u2 - u1) - (CtCab * Le1g - L * Cab1);
        rth = (th2 - th1) - rot_Qinv_from_c(rot_scaling(th) * th) * (Cab * kap);
        v3 CF = CtCab * F;
        v3 t1 = CtCab * M, t2 = 0.5 * (CtCab * cross(Le1g, F));
        F1 = CF; F2 = CF; M1 = t1 + t2; M2 = t1 - t2;
        if (A.has_grav) {
            CtCabTg = tmul(CtCab, g);
            v3 tg = -(CtCab * (ldm(es + 46 * N, N) * CtCabTg));
            F1 = F1 - 0.5 * tg; F2 = F2 + 0.5 * tg;
            tg = -(CtCab * (ldm(es + 64 * N, N) * CtCabTg));
            M1 = M1 - 0.5 * tg; M2 = M2 + 0.5 * tg;
        }
        if (A.dlS) {
            const double* d = A.dlS + ((size_t)b * 24) * N + i;
            f1f = ldv(d + 12 * N, N); f2f = ldv(d + 15 * N, N); m1f = ldv(d + 18 * N, N); m2f = ldv(d + 21 * N, N);
            F1 = F1 + (ldv(d, N) + tmul(R.C, f1f));
            F2 = F2 - (ldv(d + 3 * N, N) + tmul(R.C, f2f));
            M1 = M1 + (ldv(d + 6 * N, N) + tmul(R.C, m1f));
            M2 = M2 - (ldv(d + 9 * N, N) + tmul(R.C, m2f));
        }
        double* mine = xch + (size_t)threadIdx.x * GXB_XCH;
        mine[0] = F2.x * rfs; mine[1] = F2.y * rfs; mine[2] = F2.z * rfs;
        mine[3] = M2.x * rfs; mine[4] = M2.y * rfs; mine[5] = M2.z * rfs;
    }
    if (WANT_J && valid) tile_clear(mytile);      // tiles exist for the N + 1 stations only
    __syncthreads();

    // ---- residual: point rows = own term, += element i-1, -= element i ; element rows = compatibility ----
    if (valid) {
        const double* prev = xch + (size_t)(i > 0 ? threadIdx.x - 1 : 0) * GXB_XCH;
        double* rp = r + 12 * i;
        double own[6] = {-Fp.x * rfs, -Fp.y * rfs, -Fp.z * rfs, -Mp.x * rfs, -Mp.y * rfs, -Mp.z * rfs};
        double sub[6] = {F1.x * rfs, F1.y * rfs, F1.z * rfs, M1.x * rfs, M1.y * rfs, M1.z * rfs};
#pragma unroll
        for (int k = 0; k < 6; k++) {
            double v = own[k];
            if (i > 0) v += prev[k];
            if (has_elem) v -= sub[k];
            if (A.two_d && k >= 2 && k <= 4) v = x[12 * i + k];
            rp[k] = v;
        }
        if (has_elem) {
            double re[6] = {ru.x, ru.y, ru.z, rth.x, rth.y, rth.z};
#pragma unroll
            for (int k = 0; k < 6; k++) rp[6 + k] = (A.two_d && k >= 2 && k <= 4) ? x[12 * i + 6 + k] : re[k];
        }
    }
    if (!WANT_J) return;

    // =================== pass P1: force-equilibrium rows of point i (rows 12i+0..2) ===================
    m3 JFa, JFb;                                             // ∂F1/∂θ, ∂F2/∂θ of element i (element.jl:1954-2000)
    if (has_elem) {
        m3 Fth = 0.5 * mul3t(R, Cab * F);
        JFa = Fth; JFb = Fth;
        if (A.has_grav) {
            m3 m11 = ldm(es + 46 * N, N);
            m3 Cthg = cols(R.C1 * g, R.C2 * g, R.C3 * g);
            m3 tg = -0.5 * (mul3t(R, Cab * (m11 * CtCabTg)) + CtCab * (m11 * tmul(Cab, Cthg)));
            JFa = JFa - 0.5 * tg; JFb = JFb + 0.5 * tg;
        }
        if (A.dlS) { JFa = JFa + 0.5 * mul3t(R, f1f); JFb = JFb - 0.5 * mul3t(R, f2f); }
        // stop-point rows (station i+1): columns of point i (slots 0..5), of this element (6..11); shared θ-block parked in 0..2
        tblk(nexttile, 0, rfs * colmask(JFb, mth2));
        tblk(nexttile, 3, rfs * colmask(JFb, mth1));
        tblk(nexttile, 6, CtCab);                            // F2_F
    }
    __syncthreads();
    if (valid) {
        m3 own = zero33();                                   // -F_θ θ_θ / fs of the point itself
        if (point_rot) {
            Rot Rp; rot_C_dC(th1, Rp);
            const double* bc = A.bcS + ((size_t)b * 18) * NP + i;
            m3 Fth = rowmask(mul3t(Rp, ldv(bc + 12 * NP, NP)), fl1 & 64, fl1 & 128, fl1 & 256);
            if (A.has_grav && A.pmS) {                       // point.jl:1476-1487
                m3 m11 = ldm(A.pmS + ((size_t)b * 36) * NP + i, NP);
                m3 Cthg = cols(Rp.C1 * g, Rp.C2 * g, Rp.C3 * g);
                Fth = Fth + mul3t(Rp, m11 * (Rp.C * g)) + tmul(Rp.C, m11 * Cthg);
            }
            own = (-rfs) * colmask(Fth, mth1);
        }
        m3 blk = own + tget(mytile, 0);                      // += element i-1
        tblk(mytile, 0, zero33());
        if (has_elem) blk = blk - rfs * colmask(JFa, mth1);  // -= element i
        m3 FF = zero33();                                    // -F_F : identity on DOFs whose load is the unknown (point.jl:107-115)
        FF.a[0] = (fl1 & 64) ? 0.0 : -1.0; FF.a[4] = (fl1 & 128) ? 0.0 : -1.0; FF.a[8] = (fl1 & 256) ? 0.0 : -1.0;
        tblk(mytile, 12, FF);
        tblk(mytile, 15, blk);
        if (has_elem) {
            tblk(mytile, 18, -CtCab);                        // -F1_F
            tblk(mytile, 27, (-rfs) * colmask(JFa, mth2));
        }
        if (A.two_d) tile_planar(mytile, 0);
    }
#if GXB_ASM_TMA_FLUSH
    tile_flush_bulk(mytile, J, i, 0, valid);
#else
    __syncthreads();
    tile_flush(tiles, J, NP, 0);
    __syncthreads();
#endif

    // =================== pass P2: moment-equilibrium rows of point i (rows 12i+3..5) ===================
    if (valid) tile_clear(mytile);
    __syncthreads();
    m3 JMa, M1F, hM;
    if (has_elem) {
        m3 tm1 = mul3t(R, Cab * M);
        m3 tm2 = 0.5 * mul3t(R, Cab * cross(Le1g, F));
        JMa = 0.5 * (tm1 + tm2);
        m3 JMb = 0.5 * (tm1 - tm2);
        if (A.has_grav) {
            m3 m21 = ldm(es + 64 * N, N);
            m3 Cthg = cols(R.C1 * g, R.C2 * g, R.C3 * g);
            m3 tg = -0.5 * (mul3t(R, Cab * (m21 * CtCabTg)) + CtCab * (m21 * tmul(Cab, Cthg)));
            JMa = JMa - 0.5 * tg; JMb = JMb + 0.5 * tg;
        }
        if (A.dlS) { JMa = JMa + 0.5 * mul3t(R, m1f); JMb = JMb - 0.5 * mul3t(R, m2f); }
        m3 CtF = CtCab * tilde(F);
        M1F = 0.5 * (CtCab * tilde(Le1g) - CtF * ldm(es + 10 * N, N));     // ½ CtCab (tilde(L e1+γ) - F~ S11)
        hM = 0.5 * (CtF * ldm(es + 19 * N, N));                            // ½ CtCab F~ S12
        tblk(nexttile, 0, rfs * colmask(JMb, mth2));
        tblk(nexttile, 3, rfs * colmask(JMb, mth1));
        tblk(nexttile, 6, -M1F);                             // M2_F
        tblk(nexttile, 9, CtCab + hM);                       // M2_M
    }
    __syncthreads();
    if (valid) {
        m3 MM = zero33();                                    // -M_M
        MM.a[0] = (fl1 & 512) ? 0.0 : -1.0; MM.a[4] = (fl1 & 1024) ? 0.0 : -1.0; MM.a[8] = (fl1 & 2048) ? 0.0 : -1.0;
        m3 own = MM;
        if (point_rot) {
            Rot Rp; rot_C_dC(th1, Rp);
            const double* bc = A.bcS + ((size_t)b * 18) * NP + i;
            m3 Mth = rowmask(mul3t(Rp, ldv(bc + 15 * NP, NP)), fl1 & 512, fl1 & 1024, fl1 & 2048);
            if (A.has_grav && A.pmS) {
                m3 m21 = ldm(A.pmS + ((size_t)b * 36 + 18) * NP + i, NP);
                m3 Cthg = cols(Rp.C1 * g, Rp.C2 * g, Rp.C3 * g);
                Mth = Mth + mul3t(Rp, m21 * (Rp.C * g)) + tmul(Rp.C, m21 * Cthg);
            }
            own = MM - rfs * colmask(Mth, mth1);             // -M_M - M_θ θ_θ / fs
        }
        m3 blk = own + tget(mytile, 0);
        tblk(mytile, 0, zero33());
        if (has_elem) blk = blk - rfs * colmask(JMa, mth1);
        tblk(mytile, 15, blk);
        if (has_elem) {
            tblk(mytile, 18, -M1F);                          // -M1_F
            tblk(mytile, 21, hM - CtCab);                    // -M1_M
            tblk(mytile, 27, (-rfs) * colmask(JMa, mth2));
        }
        if (A.two_d) tile_planar(mytile, 3);
    }
#if GXB_ASM_TMA_FLUSH
    tile_flush_bulk(mytile, J, i, 3, valid);
#else
    __syncthreads();
    tile_flush(tiles, J, NP, 3);
    __syncthreads();
#endif

    // =================== pass E1: compatibility rows ru of element i (rows 12i+6..8) ===================
    if (valid) tile_clear(mytile);
    if (has_elem) {                                          // element.jl:1492-1505, 2507-2515
        m3 ruth = -0.5 * mul3t(R, Cab * Le1g);
        tblk(mytile, 6, colmask(-eye33(), mu1));
        tblk(mytile, 9, colmask(ruth, mth1));
        tblk(mytile, 12, (-fs) * (CtCab * ldm(es + 10 * N, N)));
        tblk(mytile, 15, (-fs) * (CtCab * ldm(es + 19 * N, N)));
        tblk(mytile, 18, colmask(eye33(), mu2));
        tblk(mytile, 21, colmask(ruth, mth2));
        if (A.two_d) tile_planar(mytile, 0);
    }
#if GXB_ASM_TMA_FLUSH
    tile_flush_bulk(mytile, J, i, 6, has_elem);
#else
    __syncthreads();
    tile_flush(tiles, J, N, 6);
    __syncthreads();
#endif

    // =================== pass E2: compatibility rows rθ of element i (rows 12i+9..11) ===================
    if (valid) tile_clear(mytile);
    if (has_elem) {                                          // element.jl:1507-1518, 2517-2524
        m3 Qinv;
        m3 hd = -0.5 * rot_Qinv_dQinv_b(th, Cab * kap, Qinv);
        m3 QC = Qinv * Cab;
        tblk(mytile, 9, colmask(hd - eye33(), mth1));
        tblk(mytile, 12, (-fs) * (QC * ldm(es + 28 * N, N)));
        tblk(mytile, 15, (-fs) * (QC * ldm(es + 37 * N, N)));
        tblk(mytile, 21, colmask(hd + eye33(), mth2));
        if (A.two_d) tile_planar(mytile, 3);
    }
#if GXB_ASM_TMA_FLUSH
    tile_flush_bulk(mytile, J, i, 9, has_elem);
#else
    __syncthreads();
    tile_flush(tiles, J, N, 9);
#endif
}
