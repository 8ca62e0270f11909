// gxbeam_b200 — host side of the C-ABI (include/gxbeam_b200.h) and the kernels that wrap the device code in
// gxb_static.cuh (assembly) and gxb_lu.cuh (batched pivoted block-banded LU).
//
// Newton driver = NLsolve `newton` + LineSearches `BackTracking(order 3)` as reached from nlsolve!
// (reference: src/analyses.jl:3556-3599; algorithm restated in SURVEY.md App. C), run as a per-instance state machine
// that lives entirely in device memory.  One host-visible "round" is three launches:
//     k_factor_solve   instances in state SOLVE : p = -(J \ F), first trial point xt = x + α0 p        -> state TRIAL
//     k_assemble       instances in state TRIAL : Ft = r(xt), J = ∂r/∂x(xt)   (fused, one pass)
//     k_update         instances in state TRIAL : Armijo test; accept (x = xt, F = Ft, J already at xt -> SOLVE / DONE)
//                                                 or backtrack (new α, new xt, stay in TRIAL)
// The host only reads one counter per round (instances still running).  There is no CPU fallback anywhere.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gxbeam_b200.h"
#ifndef GXB_LU_MINBLOCKS
#define GXB_LU_MINBLOCKS 4
#endif
#include "gxb_lu.cuh"
#include "gxb_static.cuh"
#include "gxb_dynamic.cuh"
#include "gxb_initial.cuh"

// ------------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return fail(GXB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

enum { ST_SOLVE = 0, ST_TRIAL = 1, ST_DONE = 2, ST_FAILED = 3 };
enum { FAIL_NONE = 0, FAIL_NONFINITE = 1, FAIL_LINESEARCH = 2, FAIL_MAXITER = 3, FAIL_SINGULAR = 4 };

// per-instance Newton / line-search scalars (structure of arrays)
struct NewtonState {
    int* status; int* failcode; int* iters; int* fcalls; int* ls_k; int* ls_finite;
    double* a1; double* a2; double* phi0; double* dphi0; double* phx0; double* finf;
};

struct gxb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    // topology
    int kind = -1;
    int64_t np = 0, ne = 0, n = 0;
    bool chain = false;
    std::vector<int64_t> start, stop, irow_point, irow_elem, icol_point, icol_elem;
    std::vector<uint8_t> pd, pl;
    std::vector<int64_t> brow, bcol;   // canonical 3x3 pattern
    unsigned short* d_flags = nullptr;
    int64_t* d_brow = nullptr; int64_t* d_bcol = nullptr;
    // batch
    int64_t nbatch = 0;
    double *d_elemS = nullptr, *d_bcS = nullptr, *d_dlS = nullptr, *d_pmS = nullptr, *d_grav = nullptr, *d_fs = nullptr, *d_xyz = nullptr;
    double *d_exS = nullptr, *d_motion = nullptr;     // dynamic only: element midpoints [b][3][N], body motion [b][12]
    double *d_paug = nullptr, *d_rates0 = nullptr, *d_series = nullptr; int *d_ser_point = nullptr, *d_ser_field = nullptr;
    int64_t ser_nt = 0; int ser_K = 0;
    double* d_M = nullptr;   // mass matrix in row-block storage (eigen analysis), allocated on demand
    double* d_Vk = nullptr; int vk_m = 0;   // Krylov basis of the last gxb_eigen_arnoldi ([b][m+1][n]), kept for gxb_eigen_ritz_vectors
    double* d_icS = nullptr; unsigned* d_rvm = nullptr;   // initial-condition analysis: [b][18][NP] initial values, [b][NP] rate_vars masks
    bool have_initial = false;                            // d_x / d_rates0 hold the output of gxb_initial_condition
    int damping = 0;         // structural_damping of the current solve
    bool skip_mass = false;  // gxb_hint_no_gravity: static context, element mass / damping / midpoint fields are not uploaded
    int mode = 0;            // 0: steady residual, 1: Newmark residual (set by the time-marching driver), 2: initial-condition residual
    double c2dt = 0.0;
    bool body_state = false;                          // some DOF has pd & pl (body-frame acceleration is a state, system.jl:138-150)
    int ibody[6] = {-1, -1, -1, -1, -1, -1};          // 0-based column of acceleration component i, or -1 (update_body_acceleration_indices!)
    int nbody = 0;
    double *d_D = nullptr, *d_Z = nullptr;            // [b][6][n] dense Jacobian columns of the body accelerations and B^-1 D (kind = 1)
    int W = 30, UW = 32, KLB = 2, rows_per_station = 12;
    int has_grav = 0;
    double* d_stage = nullptr; size_t stage_bytes = 0;
    // solver state
    double *d_x = nullptr, *d_xt = nullptr, *d_F = nullptr, *d_Ft = nullptr, *d_p = nullptr, *d_J = nullptr, *d_U = nullptr;
    NewtonState ns{};
    int* d_counter = nullptr; int* h_counter = nullptr;
    bool have_x0 = false;
    gxb_stats stats{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr, evC = nullptr;   // evC: end of the last host->device copy
};

// ------------------------------------------------------------------------------------------------------------------
// packing kernels: caller's array-of-structs (the reference's memory layout) -> structure-of-arrays in HBM
// elemS[b][f][e]: f=0 L; 1..9 Cab (row-major); 10..45 S11,S12,S21,S22 = L*compliance; 46..81 m11,m12,m21,m22 = L*mass;
// 82..87 mu (structural damping coefficients)
#define GXB_ELEM_NF 88
// compact variant for static contexts without gravity: raw holds 49 doubles per element (L, x[3], compliance[36], Cab[9]) fetched with
// two strided copies; the mass and damping fields, which that path never reads, are zero-filled
__global__ void k_pack_elements_compact(const double* __restrict__ raw, double* __restrict__ out, int N, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * N) return;
    int64_t b = t / N; int e = (int)(t % N);
    const double* s = raw + t * 49;
    double* o = out + (size_t)b * GXB_ELEM_NF * N + e;
    const double L = s[0];
    o[0] = L;
    for (int k = 0; k < 6; k++) o[(size_t)(82 + k) * N] = 0.0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[(size_t)(1 + 3 * i + j) * N] = s[40 + i + 3 * j];
    for (int bi = 0; bi < 2; bi++) for (int bj = 0; bj < 2; bj++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        int f = 9 * (2 * bi + bj) + 3 * i + j;
        int src = (3 * bi + i) + 6 * (3 * bj + j);
        o[(size_t)(10 + f) * N] = s[4 + src] * L;
        o[(size_t)(46 + f) * N] = 0.0;
    }
}
__global__ void k_pack_elements(const double* __restrict__ raw, double* __restrict__ out, double* __restrict__ exS, int N, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * N) return;
    int64_t b = t / N; int e = (int)(t % N);
    const double* s = raw + t * GXB_ELEM_STRIDE;
    double* o = out + (size_t)b * GXB_ELEM_NF * N + e;
    const double L = s[0];
    o[0] = L;
    if (exS) for (int k = 0; k < 3; k++) exS[((size_t)b * 3 + k) * N + e] = s[1 + k];    // Element.x (midpoint)
    for (int k = 0; k < 6; k++) o[(size_t)(82 + k) * N] = s[85 + k];                       // Element.mu
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[(size_t)(1 + 3 * i + j) * N] = s[76 + i + 3 * j];
    for (int bi = 0; bi < 2; bi++) for (int bj = 0; bj < 2; bj++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        int f = 9 * (2 * bi + bj) + 3 * i + j;
        int src = (3 * bi + i) + 6 * (3 * bj + j);
        o[(size_t)(10 + f) * N] = s[4 + src] * L;     // compliance *= L   (element.jl:69)
        o[(size_t)(46 + f) * N] = s[40 + src] * L;    // mass *= L         (element.jl:70)
    }
}
// out[b][f][p] = raw[b][p][f]
__global__ void k_pack_transpose(const double* __restrict__ raw, double* __restrict__ out, int P, int F, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * P * F) return;
    int f = (int)(t % F); int64_t q = t / F; int p = (int)(q % P); int64_t b = q / P;
    out[((size_t)b * F + f) * P + p] = raw[t];
}
// point masses: raw[b][p][36 col-major] -> out[b][9*(2bi+bj)+3i+j][p]
__global__ void k_pack_pmass(const double* __restrict__ raw, double* __restrict__ out, int P, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * P * 36) return;
    int f = (int)(t % 36); int64_t q = t / 36; int p = (int)(q % P); int64_t b = q / P;
    int blk = f / 9, i = (f % 9) / 3, j = f % 3, bi = blk / 2, bj = blk % 2;
    out[((size_t)b * 36 + f) * P + p] = raw[q * 36 + (3 * bi + i) + 6 * (3 * bj + j)];
}

// ------------------------------------------------------------------------------------------------------------------
template <bool WANT_J> __global__ void __launch_bounds__(256) k_assemble(StaticArgs A) {
    extern __shared__ __align__(16) double asm_smem[];
    const int b = blockIdx.x;
    if (A.status && A.status[b] != A.want_status) return;
    static_station<WANT_J>(A, b, threadIdx.x, asm_smem);
}

template <bool WANT_J, int MODE> __global__ void __launch_bounds__(256) k_assemble_dynamic(DynArgs A) {
    extern __shared__ __align__(16) double asm_smem[];
    const int b = blockIdx.x;
    if (A.s.status && A.s.status[b] != A.s.want_status) return;
    dynamic_station<WANT_J, MODE>(A, b, threadIdx.x, asm_smem);
}

template <bool WANT_J> __global__ void __launch_bounds__(256) k_assemble_initial(InitArgs A) {
    extern __shared__ __align__(16) double asm_smem[];
    const int b = blockIdx.x;
    if (A.d.s.status && A.d.s.status[b] != A.d.s.want_status) return;
    initial_station<WANT_J>(A, b, threadIdx.x, asm_smem);
}

// Dense Jacobian columns of the body-frame accelerations that are state variables (insert_steady_point_jacobians! point.jl:1829-1841,
// insert_steady_element_jacobians! element.jl:2567-2583; the same terms for the initial-condition analysis, point.jl:1879-1895).
// Column i (acceleration component i, stored at column c_i = ibody[i]) has entries in the force/moment rows of EVERY point:
//   + (tf_x, tm_x)[:, i] / fs of the point's own mass and + 1/2 of each adjacent element's, with
//   tf_ab = A m11 A', tf_alb = -A m11 A' tilde(dx + u) + A m12 A', tm likewise with m21, m22   (Vdot_ab = I, Vdot_alb = -tilde(dx + u), Ωdot_alb = I)
// which does not fit the band.  The band matrix B keeps a unit entry at (c_i, c_i) instead (the force-free, displacement-free DOF
// owns that row), and Dn[b][i][:] = column_i - e_{c_i}, so that J = B + sum_i Dn_i e_{c_i}' exactly; the solve undoes the split
// with a Woodbury correction (k_woodbury).  One thread per station, both adjacent elements evaluated by the thread itself.
struct BodyArgs { DynArgs d; const double* icS; const unsigned* rvm; int mode; double* Dn; };   // Dn: [6][nbatch][n]
__global__ void __launch_bounds__(256) k_body_columns(BodyArgs B) {
    const DynArgs& A = B.d; const StaticArgs& S = A.s;
    const int b = blockIdx.x, p = threadIdx.x, N = S.N, NP = N + 1;
    if (S.status && S.status[b] != S.want_status) return;
    if (p > N) return;
    const double rfs = 1.0 / S.fs[b];
    const double* x = S.x + (size_t)b * S.n;
    auto disp = [&](int q, v3& u, v3& th) {
        const unsigned fl = S.flags[q];
        const double* bc = S.bcS + ((size_t)b * 18) * NP + q;
        const double* xp = x + 18 * q;
        double t[6];
        for (int k = 0; k < 6; k++) {
            const bool pd = (fl >> k) & 1u;
            if (B.mode == 2) {
                const bool r = (B.rvm[(size_t)b * NP + q] >> k) & 1u;
                t[k] = pd ? bc[(size_t)k * NP] : (r ? B.icS[((size_t)b * 18 + k) * NP + q] : xp[k]);
            } else t[k] = pd ? bc[(size_t)k * NP] : xp[k];
        }
        u = mk(t[0], t[1], t[2]); th = mk(t[3], t[4], t[5]);
    };
    double G[6][6];                      // rows: F(3), M(3) of this point; columns: ab(3), alb(3)
    for (int r = 0; r < 6; r++) for (int cc = 0; cc < 6; cc++) G[r][cc] = 0.0;
    auto add = [&](const m3& Am, const m3& m11, const m3& m12, const m3& m21, const m3& m22, v3 rr, double w) {
        const m3 At = transpose(Am), rt = tilde(rr);
        const m3 A11 = (Am * m11) * At, A12 = (Am * m12) * At, A21 = (Am * m21) * At, A22 = (Am * m22) * At;
        const m3 Fal = A12 - A11 * rt, Mal = A22 - A21 * rt;
        for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) {
            G[r][cc] += w * A11.a[3 * r + cc]; G[r][3 + cc] += w * Fal.a[3 * r + cc];
            G[3 + r][cc] += w * A21.a[3 * r + cc]; G[3 + r][3 + cc] += w * Mal.a[3 * r + cc];
        }
    };
    v3 up, thp; disp(p, up, thp);
    if (S.pmS) {
        const double* pm = S.pmS + ((size_t)b * 36) * NP + p;
        const v3 dxp = ldv(A.pxS + ((size_t)b * 3) * NP + p, NP);
        add(transpose(rot_C(thp)), ldm(pm, NP), ldm(pm + 9 * NP, NP), ldm(pm + 18 * NP, NP), ldm(pm + 27 * NP, NP), dxp + up, rfs);
    }
    for (int side = 0; side < 2; side++) {
        const int e = p - 1 + side;
        if (e < 0 || e >= N) continue;
        v3 u1, t1, u2, t2; disp(e, u1, t1); disp(e + 1, u2, t2);
        const double* es = S.elemS + ((size_t)b * S.elem_nf) * N + e;
        const m3 Am = tmul(rot_C(0.5 * (t1 + t2)), ldm(es + N, N));
        const v3 dxe = ldv(A.exS + ((size_t)b * 3) * N + e, N);
        add(Am, ldm(es + 46 * N, N), ldm(es + 55 * N, N), ldm(es + 64 * N, N), ldm(es + 73 * N, N), dxe + 0.5 * (u1 + u2), 0.5 * rfs);
    }
    for (int i = 0; i < 6; i++) {
        const int c = A.ibody[i];
        if (c < 0) continue;
        double* Di = B.Dn + ((size_t)i * gridDim.x + b) * S.n + 18 * p;
        for (int k = 0; k < 6; k++) {
            double v = G[k][i];
            if (B.mode == 2 && ((B.rvm[(size_t)b * NP + p] >> (8 + k)) & 1u)) v = 0.0;    // equilibrium row replaced by Vdot_k = Vdot0_k (system.jl:756-813)
            if (S.two_d && k >= 2 && k <= 4) v = 0.0;                                      // planar constraint rows (system.jl:1090-1100)
            if (18 * p + k == c) v -= 1.0;
            Di[k] = v;
        }
        if (c / 18 == p && S.Jrow) S.Jrow[((size_t)b * S.n + c) * GXB_DROWW + 18 + (c % 18)] = 1.0;    // the unit entry B keeps at (c_i, c_i)
    }
}

// rate_vars of the initial-condition analysis (src/analyses.jl:1835-1873): column sums of the mass matrix at x = 0 over the V/Ω
// columns of each point (rate_vars1), and the same with the element mass blocks replaced by the compliance blocks and the point
// masses ignored (rate_vars2).  One thread per (instance, point).  rvm bit k: rate_vars2, bit 8+k: equilibrium row replaced
// (!pd & !rate_vars1 & rate_vars2, system.jl:491-513), bit 16+k: rate_vars1.
__global__ void k_initial_rate_vars(const double* __restrict__ elemS, int elem_nf, const double* __restrict__ bcS, const double* __restrict__ pmS,
                                    const unsigned short* __restrict__ flags, int N, int two_d, unsigned* __restrict__ rvm, int64_t nbatch) {
    const int NP = N + 1;
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * NP) return;
    const int64_t b = t / NP; const int p = (int)(t % NP);
    auto theta = [&](int q) {
        const unsigned fl = flags[q];
        const double* bc = bcS + ((size_t)b * 18) * NP + q;
        return mk((fl & 8) ? bc[3 * NP] : 0.0, (fl & 16) ? bc[4 * NP] : 0.0, (fl & 32) ? bc[5 * NP] : 0.0);
    };
    double s1[6] = {0, 0, 0, 0, 0, 0}, s2[6] = {0, 0, 0, 0, 0, 0};
    auto addcols = [&](double* s, const m3& PV, const m3& PO, const m3& HV, const m3& HO, double w) {
        for (int r = 0; r < 3; r++) {
            const bool lin = !(two_d && r == 2), ang = !(two_d && r < 2);       // lmask = (1,1,0), amask = (0,0,1) (point.jl:2043-2049)
            for (int cc = 0; cc < 3; cc++) {
                if (lin) { s[cc] += w * PV.a[3 * r + cc]; s[3 + cc] += w * PO.a[3 * r + cc]; }
                if (ang) { s[cc] += w * HV.a[3 * r + cc]; s[3 + cc] += w * HO.a[3 * r + cc]; }
            }
        }
    };
    if (pmS) {
        const double* pm = pmS + ((size_t)b * 36) * NP + p;
        const m3 C = rot_C(theta(p));
        addcols(s1, tmul(C, ldm(pm, NP) * C), tmul(C, ldm(pm + 9 * NP, NP) * C), tmul(C, ldm(pm + 18 * NP, NP) * C), tmul(C, ldm(pm + 27 * NP, NP) * C), 1.0);
    }
    for (int side = 0; side < 2; side++) {
        const int e = p - 1 + side;
        if (e < 0 || e >= N) continue;
        const double* es = elemS + ((size_t)b * elem_nf) * N + e;
        const m3 Cab = ldm(es + N, N);
        const m3 Am = tmul(rot_C(0.5 * (theta(e) + theta(e + 1))), Cab), At = transpose(Am);
        // ¼ A m A' enters the rows of both end points of the element
        addcols(s1, (Am * ldm(es + 46 * N, N)) * At, (Am * ldm(es + 55 * N, N)) * At, (Am * ldm(es + 64 * N, N)) * At, (Am * ldm(es + 73 * N, N)) * At, 0.5);
        addcols(s2, (Am * ldm(es + 10 * N, N)) * At, (Am * ldm(es + 19 * N, N)) * At, (Am * ldm(es + 28 * N, N)) * At, (Am * ldm(es + 37 * N, N)) * At, 0.5);
    }
    const unsigned fl = flags[p];
    unsigned m = 0;
    for (int k = 0; k < 6; k++) {
        const bool r1 = s1[k] != 0.0, r2 = s2[k] != 0.0, pd = (fl >> k) & 1u;
        if (r2) m |= 1u << k;
        if (!pd && !r1 && r2) m |= 1u << (8 + k);
        if (r1) m |= 1u << (16 + k);
    }
    rvm[t] = m;
}

// initial_output (src/analyses.jl:2332-2390): the solved initial-condition vector -> DynamicSystem states (in place) and the point
// rates udot, θdot, Vdot, Ωdot ([b][NP][12]) that seed the Newmark march.  One thread per (instance, point).
__global__ void k_initial_output(double* __restrict__ x, const double* __restrict__ bcS, const double* __restrict__ icS, const unsigned* __restrict__ rvm,
                                 const unsigned short* __restrict__ flags, const double* __restrict__ pxS, const double* __restrict__ motion,
                                 double* __restrict__ rates, int NP, int n, int64_t nbatch, DynArgs body) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * NP) return;
    const int64_t b = t / NP; const int p = (int)(t % NP);
    const unsigned fl = flags[p], rv = rvm[t];
    const double* bc = bcS + ((size_t)b * 18) * NP + p;
    const double* ic = icS + ((size_t)b * 18) * NP + p;
    double* xp = x + b * n + 18 * p;
    const double* mo = motion + 12 * b;
    const v3 vb = mk(mo[0], mo[1], mo[2]), wb = mk(mo[3], mo[4], mo[5]);
    v3 ab, alb;
    body_accelerations(body, x + b * n, mo, ab, alb);          // system.jl:167-182 (analyses.jl:2348)
    double ut[6], rt[6];
    for (int k = 0; k < 6; k++) {
        const bool pd = (fl >> k) & 1u, r = (rv >> k) & 1u;
        ut[k] = pd ? bc[(size_t)k * NP] : (r ? ic[(size_t)k * NP] : xp[k]);
        rt[k] = pd ? 0.0 : (r ? xp[k] : ic[(size_t)(12 + k) * NP]);
    }
    const v3 u = mk(ut[0], ut[1], ut[2]);
    const v3 udot = mk(xp[6], xp[7], xp[8]), thdot = mk(xp[9], xp[10], xp[11]);
    const v3 dx = ldv(pxS + ((size_t)b * 3) * NP + p, NP);
    const v3 V = ldv(ic + (size_t)6 * NP, NP) + vb + cross(wb, dx + u), Om = ldv(ic + (size_t)9 * NP, NP) + wb;
    const v3 Vd = mk(rt[0], rt[1], rt[2]) + ab + cross(alb, dx + u) + cross(wb, udot), Od = mk(rt[3], rt[4], rt[5]) + alb;
    double* ra = rates + ((size_t)b * NP + p) * 12;
    for (int k = 0; k < 6; k++) if (!((fl >> k) & 1u)) xp[k] = ut[k];          // displacement where it is a state; loads stay
    ra[0] = (fl & 1) ? 0.0 : udot.x; ra[1] = (fl & 2) ? 0.0 : udot.y; ra[2] = (fl & 4) ? 0.0 : udot.z;
    ra[3] = (fl & 8) ? 0.0 : thdot.x; ra[4] = (fl & 16) ? 0.0 : thdot.y; ra[5] = (fl & 32) ? 0.0 : thdot.z;
    ra[6] = Vd.x; ra[7] = Vd.y; ra[8] = Vd.z; ra[9] = Od.x; ra[10] = Od.y; ra[11] = Od.z;
    xp[6] = V.x; xp[7] = V.y; xp[8] = V.z; xp[9] = Om.x; xp[10] = Om.y; xp[11] = Om.z;
}

// Newmark bookkeeping (src/analyses.jl:2693-2738): per point, rates at the previous time level from the old initialisation
// terms (or the initial state on the first step), then the new terms  2/dt x + xdot.  paug layout [b][12][NP].
__global__ void k_newmark_paug(const double* __restrict__ x, const double* __restrict__ bcS, const unsigned short* __restrict__ flags,
                               double* __restrict__ paug, const double* __restrict__ rates0, int NP, int n, int64_t nbatch,
                               int first, double c_prev, double c_new) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * NP) return;
    int64_t b = t / NP; int p = (int)(t % NP);
    const double* xp = x + b * n + 18 * p;
    const double* bc = bcS + ((size_t)b * 18) * NP + p;
    const unsigned fl = flags[p];
    double* pa = paug + ((size_t)b * 12) * NP + p;
    double s[12];
    for (int k = 0; k < 6; k++) s[k] = (fl & (1u << k)) ? bc[k * NP] : xp[k];      // u, θ honour prescribed displacements at the new time
    for (int k = 0; k < 6; k++) s[6 + k] = xp[6 + k];                              // V, Ω
    for (int k = 0; k < 12; k++) {
        double rate = first ? rates0[((size_t)b * NP + p) * 12 + k] : c_prev * s[k] - pa[(size_t)k * NP];
        pa[(size_t)k * NP] = c_new * s[k] + rate;
    }
}
// time-varying prescribed values: bcS[b][field[k]][point[k]] = series[b][step][k]
__global__ void k_apply_bc_series(double* __restrict__ bcS, const double* __restrict__ series, const int* __restrict__ point,
                                  const int* __restrict__ field, int K, int NP, int64_t nt, int64_t step, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * K) return;
    int64_t b = t / K; int k = (int)(t % K);
    bcS[((size_t)b * 18 + field[k]) * NP + point[k]] = series[((size_t)b * nt + step) * K + k];
}
// xhist[b][slot][:] = x[b][:]
__global__ void k_save_state(const double* __restrict__ x, double* __restrict__ xhist, int n, int64_t nsave, int64_t slot, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * n) return;
    int64_t b = t / n; int c = (int)(t % n);
    xhist[((size_t)b * nsave + slot) * n + c] = x[t];
}

struct SolveArgs {
    int nb, n; int64_t nbatch;
    const double* J; const double* F; const double* x; double* U; double* p; double* xt;
    NewtonState ns;
    double maxstep; int linear;
    int direction_only;   // parity entry point: just p, no state change
    double sign;          // sol = sign * (J \ F): -1 for the Newton direction, +1 for the eigen operator
    // Woodbury-corrected steps (body-acceleration states): F / p above are the right-hand side / solution of ONE band solve,
    // these are the Newton residual and direction the line search works with
    const double* Fnewton; double* pnewton;
};
__device__ void linesearch_start(const SolveArgs& A, int64_t b);

// one warp: factorise + solve instance b, then start the line search (|p|_inf, φ0, dφ0, first trial point)
template <int KLB, int KUB> __device__ void factor_instance(const SolveArgs& A, int64_t b, double* sm) {
    using S = LuShape<KLB, KUB>;
    const int lane = threadIdx.x & 31;
    const size_t n = A.n;
    const double* F = A.F + b * n;
    double* p = A.p + b * n;
    int bad = warp_band_solve<KLB, KUB>(A.J + b * n * S::W, F, A.U + b * n * S::UW, p, A.nb, A.sign, sm);
    __syncwarp();
    if (A.direction_only == 1) return;
    if (bad) { if (lane == 0) { A.ns.status[b] = ST_FAILED; A.ns.failcode[b] = FAIL_SINGULAR; } return; }
    if (A.direction_only) return;        // 2: one of several solves of a Woodbury-corrected step (k_woodbury finishes it)
    linesearch_start(A, b);
}
// |p|_inf, φ0 = ½ F'F, dφ0 = g'p with g = J'F and J p = -F  =>  dφ0 = -F'F; first trial point.  One warp; p is complete.
__device__ void linesearch_start(const SolveArgs& A, int64_t b) {
    const int lane = threadIdx.x & 31;
    const size_t n = A.n;
    const double* F = A.Fnewton ? A.Fnewton + b * n : A.F + b * n;
    double* p = A.pnewton ? A.pnewton + b * n : A.p + b * n;
    {
    double pm = 0.0, ff = 0.0;
    for (int c = lane; c < (int)n; c += 32) { pm = fmax(pm, fabs(p[c])); double f = F[c]; ff = fma(f, f, ff); }
    for (int o = 16; o; o >>= 1) { pm = fmax(pm, __shfl_xor_sync(GXB_FULL, pm, o)); ff += __shfl_xor_sync(GXB_FULL, ff, o); }
    const double a0 = A.linear ? 1.0 : fmin(1.0, A.maxstep / pm);
    const double* x = A.x + b * n; double* xt = A.xt + b * n;
    for (int c = lane; c < (int)n; c += 32) xt[c] = x[c] + a0 * p[c];
    if (lane == 0) {
        A.ns.a1[b] = a0; A.ns.a2[b] = a0; A.ns.phi0[b] = 0.5 * ff; A.ns.dphi0[b] = -ff; A.ns.phx0[b] = 0.5 * ff;
        A.ns.ls_k[b] = 0; A.ns.ls_finite[b] = 0; A.ns.status[b] = ST_TRIAL;
    }
    }
}
template <int KLB, int KUB> __global__ void __launch_bounds__(128, (KLB == 2 ? GXB_LU_MINBLOCKS : 2)) k_factor_solve(SolveArgs A) {
    using S = LuShape<KLB, KUB>;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= A.nbatch) return;
    if (A.direction_only != 1 && A.ns.status[b] != ST_SOLVE) return;
    factor_instance<KLB, KUB>(A, b, smem + (size_t)warp * S::SM_DOUBLES);
}

// Woodbury correction of a Newton step whose Jacobian is J = B + sum_i D_i e_{c_i}' (body-acceleration columns, see k_body_columns):
// with y = B^-1 F (in p) and z_i = B^-1 D_i,   J^-1 F = y - Z (I + E'Z)^-1 E'y.   One warp per instance; the k x k system (k <= 6) is
// solved by lane 0 with partial pivoting.  Then the line search starts exactly as after an ordinary factorisation.
struct WoodArgs { SolveArgs S; const double* Z; int ibody[6]; };
__global__ void __launch_bounds__(128) k_woodbury(WoodArgs W) {
    const SolveArgs& A = W.S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= A.nbatch) return;
    if (A.ns.status[b] != ST_SOLVE) return;
    const size_t n = A.n;
    double* p = A.p + b * n;
    int idx[6], k = 0;
    for (int i = 0; i < 6; i++) if (W.ibody[i] >= 0) idx[k++] = i;
    double a[6] = {0, 0, 0, 0, 0, 0};
    int bad = 0;
    if (lane == 0) {
        double M[6][7];
        for (int j = 0; j < k; j++) {
            const int cj = W.ibody[idx[j]];
            for (int i = 0; i < k; i++) M[j][i] = (i == j ? 1.0 : 0.0) + W.Z[((size_t)idx[i] * A.nbatch + b) * n + cj];
            M[j][k] = p[cj];
        }
        for (int col = 0; col < k && !bad; col++) {
            int piv = col; double best = fabs(M[col][col]);
            for (int r = col + 1; r < k; r++) if (fabs(M[r][col]) > best) { best = fabs(M[r][col]); piv = r; }
            if (!(best > 0.0) || !isfinite(best)) { bad = 1; break; }
            if (piv != col) for (int q = col; q <= k; q++) { double t = M[col][q]; M[col][q] = M[piv][q]; M[piv][q] = t; }
            for (int r = col + 1; r < k; r++) { const double f = M[r][col] / M[col][col]; for (int q = col; q <= k; q++) M[r][q] -= f * M[col][q]; }
        }
        if (!bad) for (int r = k - 1; r >= 0; r--) { double v = M[r][k]; for (int q = r + 1; q < k; q++) v -= M[r][q] * a[q]; a[r] = v / M[r][r]; }
    }
    bad = __shfl_sync(GXB_FULL, bad, 0);
    if (bad) { if (lane == 0) { A.ns.status[b] = ST_FAILED; A.ns.failcode[b] = FAIL_SINGULAR; } return; }
    for (int j = 0; j < 6; j++) a[j] = __shfl_sync(GXB_FULL, a[j], 0);
    for (int c = lane; c < (int)n; c += 32) {
        double v = p[c];
        for (int j = 0; j < k; j++) v = fma(-W.Z[((size_t)idx[j] * A.nbatch + b) * n + c], a[j], v);
        p[c] = -v;                       // Newton direction p = -J^-1 F
    }
    __syncwarp();
    linesearch_start(A, b);
}

struct UpdateArgs {
    int n; int mode;   // mode 0: initial evaluation, 1: line-search step, 2: linear (accept unconditionally, converged)
    double* x; double* xt; double* F; const double* Ft; const double* p;
    NewtonState ns;
    double ftol, c1, rho_hi, rho_lo; int64_t iterations; int ls_iterations;
    int* counter;
};

// Armijo test / accept / convergence for instance b by the whole CTA (any block size up to 256); returns the new status
__device__ int update_instance(const UpdateArgs& A, int b) {
    const size_t n = A.n;
    const double* Ft = A.Ft + b * n;
    // all per-instance scalars are read before the reduction barrier; thread 0 rewrites them only after it
    const double a1 = A.ns.a1[b], a2 = A.ns.a2[b], phi0 = A.ns.phi0[b], dphi0 = A.ns.dphi0[b], phx0 = A.ns.phx0[b];
    int k = A.ns.ls_k[b], nf = A.ns.ls_finite[b];
    int iters = A.ns.iters[b];
    __shared__ double s_sum[8], s_max[8]; __shared__ int s_nan[8];
    double ss = 0.0, mx = 0.0; int nan = 0;
    for (int c = threadIdx.x; c < (int)n; c += blockDim.x) { double f = Ft[c]; ss = fma(f, f, ss); mx = fmax(mx, fabs(f)); nan |= (f != f); }
    for (int o = 16; o; o >>= 1) { ss += __shfl_xor_sync(GXB_FULL, ss, o); mx = fmax(mx, __shfl_xor_sync(GXB_FULL, mx, o)); nan |= __shfl_xor_sync(GXB_FULL, nan, o); }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = ss; s_max[threadIdx.x >> 5] = mx; s_nan[threadIdx.x >> 5] = nan; }
    if (threadIdx.x < 8 && threadIdx.x >= (blockDim.x >> 5)) { s_sum[threadIdx.x] = 0.0; s_max[threadIdx.x] = 0.0; s_nan[threadIdx.x] = 0; }
    __syncthreads();
    ss = (s_sum[0] + s_sum[1]) + (s_sum[2] + s_sum[3]);
    mx = fmax(fmax(s_max[0], s_max[1]), fmax(s_max[2], s_max[3]));
    nan = s_nan[0] | s_nan[1] | s_nan[2] | s_nan[3];
    if (blockDim.x > 128) {      // persistent kernels run with up to 8 warps; the 4-warp order above is kept for bit-compatibility
        ss += (s_sum[4] + s_sum[5]) + (s_sum[6] + s_sum[7]);
        mx = fmax(mx, fmax(fmax(s_max[4], s_max[5]), fmax(s_max[6], s_max[7])));
        nan |= s_nan[4] | s_nan[5] | s_nan[6] | s_nan[7];
    }
    const double phx1 = 0.5 * ss;
    const bool finite = isfinite(ss) && !nan;
    double* x = A.x + b * n; double* xt = A.xt + b * n; double* F = A.F + b * n; const double* p = A.p + b * n;
    // every thread takes the same decision from the same scalars
    bool accept = false; int newstatus = ST_TRIAL; int failcode = FAIL_NONE; double a_new = 0.0;
    if (A.mode == 0) {
        accept = true;
        if (!finite) { newstatus = ST_FAILED; failcode = FAIL_NONFINITE; }     // NLsolve check_isfinite
        else newstatus = (mx <= A.ftol) ? ST_DONE : ST_SOLVE;
        if (newstatus == ST_SOLVE && A.iterations <= 0) { newstatus = ST_FAILED; failcode = FAIL_MAXITER; }
    } else if (A.mode == 2) {
        accept = true; newstatus = ST_DONE; iters = 1;
    } else {
        double n_a1 = a1, n_a2 = a2, n_phx0 = phx0;
        if (k == 0 && !isfinite(phx1) && nf < 52) {                 // back off until finite (LineSearches iterfinite loop)
            nf++; n_a1 = a2; n_a2 = a2 / 2;
        } else if (phx1 > phi0 + A.c1 * a2 * dphi0) {               // Armijo fails -> interpolate
            k++;
            if (k > A.ls_iterations) { newstatus = ST_FAILED; failcode = FAIL_LINESEARCH; }
            double at;
            if (k == 1) at = -(dphi0 * a2 * a2) / (2 * (phx1 - phi0 - dphi0 * a2));
            else {
                double dv = 1.0 / (a1 * a1 * a2 * a2 * (a2 - a1));
                double ca = (a1 * a1 * (phx1 - phi0 - dphi0 * a2) - a2 * a2 * (phx0 - phi0 - dphi0 * a1)) * dv;
                double cb = (-a1 * a1 * a1 * (phx1 - phi0 - dphi0 * a2) + a2 * a2 * a2 * (phx0 - phi0 - dphi0 * a1)) * dv;
                if (fabs(ca) <= 2.220446049250313e-16) at = dphi0 / (2 * cb);
                else { double d = fmax(cb * cb - 3 * ca * dphi0, 0.0); at = (-cb + sqrt(d)) / (3 * ca); }
            }
            n_a1 = a2;
            double hi = a2 * A.rho_hi, lo = a2 * A.rho_lo;
            at = (at != at) ? hi : fmin(at, hi);
            n_a2 = fmax(at, lo);
            n_phx0 = phx1;
        } else {
            accept = true;
        }
        if (!accept && newstatus == ST_TRIAL) {
            a_new = n_a2;
            if (threadIdx.x == 0) { A.ns.a1[b] = n_a1; A.ns.a2[b] = n_a2; A.ns.phx0[b] = n_phx0; A.ns.ls_k[b] = k; A.ns.ls_finite[b] = nf; }
        }
        if (accept) {
            iters++;
            const bool fconv = mx <= A.ftol;                         // maximum(abs, F) <= ftol ; NaN compares false
            bool stopped = nan != 0;
            bool xzero = true;                                       // x-convergence: ||x - xold||_inf <= xtol = 0
            for (int c = threadIdx.x; c < (int)n; c += blockDim.x) { if (xt[c] != x[c]) xzero = false; if (xt[c] != xt[c]) stopped = true; }
            xzero = __syncthreads_and(xzero); stopped = __syncthreads_or(stopped);
            if (fconv) newstatus = ST_DONE;
            else if (stopped) { newstatus = ST_FAILED; failcode = FAIL_NONFINITE; }
            else if (xzero) { newstatus = ST_FAILED; failcode = FAIL_LINESEARCH; }
            else if (iters >= A.iterations) { newstatus = ST_FAILED; failcode = FAIL_MAXITER; }
            else newstatus = ST_SOLVE;
        }
    }
    if (accept) {
        for (int c = threadIdx.x; c < (int)n; c += blockDim.x) { x[c] = xt[c]; F[c] = Ft[c]; }
        if (threadIdx.x == 0) { A.ns.finf[b] = mx; A.ns.iters[b] = iters; }
    } else if (newstatus == ST_TRIAL) {
        for (int c = threadIdx.x; c < (int)n; c += blockDim.x) xt[c] = x[c] + a_new * p[c];
    }
    if (threadIdx.x == 0) {
        A.ns.fcalls[b] += 1;
        A.ns.status[b] = newstatus;
        if (failcode) A.ns.failcode[b] = failcode;
        if (A.counter && (newstatus == ST_SOLVE || newstatus == ST_TRIAL)) atomicAdd(A.counter, 1);
    }
    return newstatus;
}
__global__ void __launch_bounds__(128) k_update(UpdateArgs A) {
    const int b = blockIdx.x;
    if (A.ns.status[b] != ST_TRIAL) return;
    update_instance(A, b);
}

__global__ void k_fill_int(int* p, int v, int64_t n) { int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (t < n) p[t] = v; }

// Jblk[b][k][3i+j] = J[brow[k]+i, bcol[k]+j] read back from the row-block storage
__global__ void k_extract_blocks(const double* __restrict__ Jrow, const int64_t* __restrict__ brow, const int64_t* __restrict__ bcol,
                                 double* __restrict__ out, int64_t nblk, int n, int W, int KLB, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * nblk * 9) return;
    int e = (int)(t % 9); int64_t q = t / 9; int64_t k = q % nblk; int64_t b = q / nblk;
    int64_t r = brow[k] + e / 3, c = bcol[k] + e % 3;
    int64_t slot = c - 6 * (r / 6 - KLB);
    out[t] = (slot >= 0 && slot < W) ? Jrow[((size_t)b * n + r) * W + slot] : 0.0;
}


// ------------------------------------------------------------------------------------------------------------------
// Mass matrix M = ∂r/∂ẋ in the same row-block storage as the Jacobian (system_mass_matrix!, src/system.jl:961-995;
// point.jl:1356-1384, 2033-2061; element.jl:1408-1441, 2370-2411, 2903-2953).  One thread per station; a thread evaluates the
// frames of element i-1 and element i itself (only C(θ) is needed), so no exchange between threads is required.
struct MassArgs { StaticArgs s; double* Mrow; };
__global__ void k_mass_matrix(MassArgs A) {
    const StaticArgs& S = A.s;
    const int b = blockIdx.x, i = threadIdx.x, N = S.N, NP = N + 1;
    if (i > N) return;
    const double rfs = 1.0 / S.fs[b];
    const double* x = S.x + (size_t)b * S.n;
    double* Mr = A.Mrow + ((size_t)b * S.n + 18 * i) * GXB_DROWW;
    auto theta = [&](int p) {
        const unsigned fl = S.flags[p];
        const double* bc = S.bcS + ((size_t)b * 18) * NP + p;
        const double* xp = x + 18 * p;
        return mk((fl & 8) ? bc[3 * NP] : xp[3], (fl & 16) ? bc[4 * NP] : xp[4], (fl & 32) ? bc[5 * NP] : xp[5]);
    };
    const v3 thi = theta(i);
    m3 FV[3], FO[3], MV[3], MO[3];       // column points i-1, i, i+1
    for (int k = 0; k < 3; k++) { FV[k] = zero33(); FO[k] = zero33(); MV[k] = zero33(); MO[k] = zero33(); }
    if (S.pmS) {
        const double* pm = S.pmS + ((size_t)b * 36) * NP + i;
        const m3 C = rot_C(thi);
        FV[1] = rfs * tmul(C, ldm(pm, NP) * C); FO[1] = rfs * tmul(C, ldm(pm + 9 * NP, NP) * C);
        MV[1] = rfs * tmul(C, ldm(pm + 18 * NP, NP) * C); MO[1] = rfs * tmul(C, ldm(pm + 27 * NP, NP) * C);
    }
    for (int side = 0; side < 2; side++) {       // element i-1 (this point is its stop), then element i (this point is its start)
        const int e = i - 1 + side;
        if (e < 0 || e >= N) continue;
        const double* es = S.elemS + ((size_t)b * S.elem_nf) * N + e;
        const m3 Cab = ldm(es + N, N);
        const v3 th = 0.5 * (theta(e) + theta(e + 1));
        const m3 Am = tmul(rot_C(th), Cab), At = transpose(Am);
        const double q = 0.25 * rfs;
        const m3 a11 = q * ((Am * ldm(es + 46 * N, N)) * At), a12 = q * ((Am * ldm(es + 55 * N, N)) * At);
        const m3 a21 = q * ((Am * ldm(es + 64 * N, N)) * At), a22 = q * ((Am * ldm(es + 73 * N, N)) * At);
        const int k0 = side;                     // column points (e, e+1) = (i-1, i) or (i, i+1)
        FV[k0] = FV[k0] + a11; FV[k0 + 1] = FV[k0 + 1] + a11; FO[k0] = FO[k0] + a12; FO[k0 + 1] = FO[k0 + 1] + a12;
        MV[k0] = MV[k0] + a21; MV[k0 + 1] = MV[k0 + 1] + a21; MO[k0] = MO[k0] + a22; MO[k0 + 1] = MO[k0 + 1] + a22;
    }
    const bool two_d = S.two_d != 0;
    auto put = [&](int r0, int slot, const m3& B, bool angular) {
        for (int r = 0; r < 3; r++) {
            const bool zero = two_d && (angular ? r < 2 : r == 2);       // lmask = (1,1,0), amask = (0,0,1)
            for (int cc = 0; cc < 3; cc++) Mr[(size_t)(r0 + r) * GXB_DROWW + slot + cc] = zero ? 0.0 : B.a[3 * r + cc];
        }
    };
    for (int k = 0; k < 3; k++) {
        const int p = i - 1 + k;
        if (p < 0 || p > N) continue;
        const int sV = 6 + 18 * k, sO = 9 + 18 * k;     // slots of (V, Ω) of points i-1, i, i+1 in an equilibrium row
        put(0, sV, FV[k], false); put(0, sO, FO[k], false);
        put(3, sV, MV[k], true); put(3, sO, MO[k], true);
    }
    const unsigned fl = S.flags[i];
    m3 uu = zero33(), tt = zero33();
    uu.a[0] = (fl & 1) ? 0.0 : -1.0; uu.a[4] = (fl & 2) ? 0.0 : -1.0; uu.a[8] = (fl & 4) ? 0.0 : -1.0;
    tt.a[0] = (fl & 8) ? 0.0 : -1.0; tt.a[4] = (fl & 16) ? 0.0 : -1.0; tt.a[8] = (fl & 32) ? 0.0 : -1.0;
    put(6, 12, uu, false);       // rV rows: -u_u
    put(9, 15, tt, true);        // rΩ rows: -θ_θ
}

// y = M v in row-block storage: one thread per row
__global__ void k_rowblock_spmv(const double* __restrict__ Mrow, const double* __restrict__ v, double* __restrict__ y, int n, int W, int KLB,
                                int64_t nbatch, int64_t vstride) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * n) return;
    const int64_t b = t / n; const int r = (int)(t % n);
    const double* row = Mrow + (size_t)t * W;
    const double* vb = v + b * vstride;
    const int c0 = 6 * (r / 6 - KLB);
    double acc = 0.0;
    for (int j = 0; j < W; j++) { const int c = c0 + j; if (c >= 0 && c < n) acc = fma(row[j], vb[c], acc); }
    y[t] = acc;
}

// out[b][c][j] = sum_k Vk[b][k][c] * Y[b][k][j]  (Y complex): thread per state c, four eigenvectors per pass
__global__ void __launch_bounds__(128) k_ritz_vectors(const double* __restrict__ Vk, const double* __restrict__ Y, double* __restrict__ out, int n, int m, int nev) {
    const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    const double* V = Vk + (size_t)b * (m + 1) * n;
    const double* Yb = Y + (size_t)b * m * nev * 2;
    for (int j0 = 0; j0 < nev; j0 += 4) {
        double re[4] = {0, 0, 0, 0}, im[4] = {0, 0, 0, 0};
        if (c < n)
            for (int k = 0; k < m; k++) {
                const double v = V[(size_t)k * n + c];
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (j0 + q < nev) { re[q] = fma(v, Yb[((size_t)k * nev + j0 + q) * 2], re[q]); im[q] = fma(v, Yb[((size_t)k * nev + j0 + q) * 2 + 1], im[q]); }
            }
        if (c < n)
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (j0 + q < nev) { double* o = out + (((size_t)b * n + c) * nev + j0 + q) * 2; o[0] = re[q]; o[1] = im[q]; }
    }
}

// Arnoldi bookkeeping, one CTA per instance.  Vk: [b][m+1][n] Krylov basis, H: [b][m+1][m] (row-major), w: [b][n] = A v_j.
__global__ void __launch_bounds__(256) k_arnoldi_start(double* __restrict__ Vk, int n, int m) {
    const int b = blockIdx.x;
    double* v0 = Vk + (size_t)b * (m + 1) * n;
    __shared__ double red[8];
    double ss = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        // deterministic pseudo-random start vector (the reference's ArnoldiMethod draws a random one)
        unsigned h = (unsigned)c * 2654435761u + 12345u; h ^= h >> 13; h *= 2246822519u; h ^= h >> 16;
        const double val = 0.5 + (double)(h & 0xffffff) / 16777216.0;
        v0[c] = val; ss = fma(val, val, ss);
    }
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(GXB_FULL, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    double tot = 0.0; for (int k = 0; k < (int)(blockDim.x >> 5); k++) tot += red[k];
    const double inv = 1.0 / sqrt(tot);
    for (int c = threadIdx.x; c < n; c += blockDim.x) v0[c] *= inv;
}
__global__ void __launch_bounds__(256) k_arnoldi_orth(double* __restrict__ Vk, double* __restrict__ H, const double* __restrict__ w_in, int n, int m, int j) {
    const int b = blockIdx.x;
    double* V = Vk + (size_t)b * (m + 1) * n;
    double* Hb = H + (size_t)b * (m + 1) * m;
    double* vn = V + (size_t)(j + 1) * n;            // work in place in the slot of v_{j+1}
    const double* w = w_in + (size_t)b * n;
    __shared__ double red[8]; __shared__ double bc;
    const int nw = blockDim.x >> 5;
    auto block_sum = [&](double x) {
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(GXB_FULL, x, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x == 0) { double t = 0.0; for (int k = 0; k < nw; k++) t += red[k]; bc = t; }
        __syncthreads();
        return bc;
    };
    for (int c = threadIdx.x; c < n; c += blockDim.x) vn[c] = w[c];
    // modified Gram-Schmidt, two sweeps (the second one only corrects round-off; its coefficients are added to H)
    for (int sweep = 0; sweep < 2; sweep++)
        for (int i = 0; i <= j; i++) {
            const double* vi = V + (size_t)i * n;
            double d = 0.0;
            for (int c = threadIdx.x; c < n; c += blockDim.x) d = fma(vi[c], vn[c], d);
            const double h = block_sum(d);
            for (int c = threadIdx.x; c < n; c += blockDim.x) vn[c] = fma(-h, vi[c], vn[c]);
            if (threadIdx.x == 0) Hb[(size_t)i * m + j] = (sweep == 0 ? 0.0 : Hb[(size_t)i * m + j]) + h;
        }
    double d = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) d = fma(vn[c], vn[c], d);
    const double nrm = sqrt(block_sum(d));
    if (threadIdx.x == 0) Hb[(size_t)(j + 1) * m + j] = nrm;
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) vn[c] *= inv;
}

// ------------------------------------------------------------------------------------------------------------------
static void free_batch(gxb_ctx* c) {
    cudaFree(c->d_exS); cudaFree(c->d_motion); c->d_exS = c->d_motion = nullptr;
    cudaFree(c->d_M); c->d_M = nullptr;
    cudaFree(c->d_Vk); c->d_Vk = nullptr; c->vk_m = 0;
    cudaFree(c->d_D); cudaFree(c->d_Z); c->d_D = c->d_Z = nullptr;
    cudaFree(c->d_icS); cudaFree(c->d_rvm); c->d_icS = nullptr; c->d_rvm = nullptr; c->have_initial = false;
    cudaFree(c->d_paug); cudaFree(c->d_rates0); cudaFree(c->d_series); cudaFree(c->d_ser_point); cudaFree(c->d_ser_field);
    c->d_paug = c->d_rates0 = c->d_series = nullptr; c->d_ser_point = c->d_ser_field = nullptr; c->ser_nt = 0; c->ser_K = 0; c->mode = 0;
    cudaFree(c->d_elemS); cudaFree(c->d_bcS); cudaFree(c->d_dlS); cudaFree(c->d_pmS); cudaFree(c->d_grav); cudaFree(c->d_fs); cudaFree(c->d_xyz);
    cudaFree(c->d_x); cudaFree(c->d_xt); cudaFree(c->d_F); cudaFree(c->d_Ft); cudaFree(c->d_p); cudaFree(c->d_J); cudaFree(c->d_U);
    cudaFree(c->ns.status); cudaFree(c->ns.failcode); cudaFree(c->ns.iters); cudaFree(c->ns.fcalls); cudaFree(c->ns.ls_k); cudaFree(c->ns.ls_finite);
    cudaFree(c->ns.a1); cudaFree(c->ns.a2); cudaFree(c->ns.phi0); cudaFree(c->ns.dphi0); cudaFree(c->ns.phx0); cudaFree(c->ns.finf);
    cudaFree(c->d_stage);
    c->d_elemS = c->d_bcS = c->d_dlS = c->d_pmS = c->d_grav = c->d_fs = c->d_xyz = nullptr;
    c->d_x = c->d_xt = c->d_F = c->d_Ft = c->d_p = c->d_J = c->d_U = nullptr;
    c->ns = NewtonState{};
    c->d_stage = nullptr; c->stage_bytes = 0; c->nbatch = 0;
}

static int stage_upload(gxb_ctx* c, const void* host, size_t bytes) {
    if (bytes > c->stage_bytes) {
        cudaFree(c->d_stage); c->d_stage = nullptr; c->stage_bytes = 0;
        CK(cudaMalloc(&c->d_stage, bytes));
        c->stage_bytes = bytes;
    }
    CK(cudaMemcpyAsync(c->d_stage, host, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(c->evC, c->stream));      // the caller's buffer is free once this event has fired
    return GXB_OK;
}
// The set_* calls return as soon as the caller's host buffer has been read: the packing kernel that follows is stream-ordered before
// everything that uses its output, so nobody has to wait for it on the host (this keeps a pipelined caller's upload slot short).
static int wait_copy(gxb_ctx* c) { CK(cudaEventSynchronize(c->evC)); return GXB_OK; }

template <class T> static int dalloc(T** p, size_t count) { CK(cudaMalloc((void**)p, count * sizeof(T))); return GXB_OK; }

static inline unsigned grid1(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

extern "C" {

void gxb_default_opts(gxb_opts* o) {
    memset(o, 0, sizeof(*o));
    o->ftol = 1e-9; o->iterations = 1000; o->maxstep = 1e6; o->c1 = 1e-4; o->rho_hi = 0.5; o->rho_lo = 0.1;
}

const char* gxb_last_error(void) { return g_err.c_str(); }

int gxb_create(gxb_ctx** out, int device) {
    if (!out) return fail(GXB_ERR_INVALID, "gxb_create: null ctx pointer");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(GXB_ERR_CUDA, std::string("gxb_create: no usable CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(GXB_ERR_INVALID, "gxb_create: bad device index");
    CK(cudaSetDevice(device));
    gxb_ctx* c = new gxb_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&c->d_counter, sizeof(int)));
    CK(cudaMallocHost(&c->h_counter, sizeof(int)));
    CK(cudaEventCreate(&c->ev0)); CK(cudaEventCreate(&c->ev1)); CK(cudaEventCreate(&c->evA)); CK(cudaEventCreate(&c->evB)); CK(cudaEventCreateWithFlags(&c->evC, cudaEventDisableTiming));
    *out = c;
    return GXB_OK;
}

void gxb_destroy(gxb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_batch(c);
    cudaFree(c->d_flags); cudaFree(c->d_brow); cudaFree(c->d_bcol); cudaFree(c->d_counter); cudaFreeHost(c->h_counter);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->evA); cudaEventDestroy(c->evB); cudaEventDestroy(c->evC);
    cudaStreamDestroy(c->stream);
    delete c;
}

int gxb_topology(gxb_ctx* c, int kind, int64_t np, int64_t ne, const int64_t* start, const int64_t* stop, const uint8_t* pd,
                 const uint8_t* pl, int64_t* nstates, int64_t* irow_point, int64_t* irow_elem, int64_t* icol_point, int64_t* icol_elem) {
    if (!c || !start || !stop || ne < 1 || np < 2) return fail(GXB_ERR_INVALID, "gxb_topology: bad arguments");
    if (kind != 0 && kind != 1) return fail(GXB_ERR_INVALID, "gxb_topology: kind must be 0 (static) or 1 (dynamic)");
    CK(cudaSetDevice(c->device));
    int64_t pmax = 0;
    for (int64_t e = 0; e < ne; e++) {
        if (start[e] < 1 || stop[e] < 1) return fail(GXB_ERR_INVALID, "gxb_topology: point indices are 1-based");
        pmax = std::max(pmax, std::max(start[e], stop[e]));
    }
    if (pmax != np) return fail(GXB_ERR_INVALID, "gxb_topology: np must equal max(maximum(start), maximum(stop)) (system.jl:32)");
    // SystemIndices (system.jl:29-122): walk the elements; a point gets its rows/cols the first time it is seen
    const bool is_static = kind == 0;
    c->kind = kind; c->np = np; c->ne = ne;
    c->start.assign(start, start + ne); c->stop.assign(stop, stop + ne);
    c->irow_point.assign(np, 0); c->icol_point.assign(np, 0); c->irow_elem.assign(ne, 0); c->icol_elem.assign(ne, 0);
    std::vector<char> seen(np, 0);
    int64_t cur = 1;
    for (int64_t e = 0; e < ne; e++) {
        for (int side = 0; side < 2; side++) {
            int64_t ip = (side ? stop[e] : start[e]) - 1;
            if (side == 1) { c->irow_elem[e] = cur; c->icol_elem[e] = cur; cur += 6; }
            if (!seen[ip]) { seen[ip] = 1; c->irow_point[ip] = cur; c->icol_point[ip] = cur; cur += is_static ? 6 : 12; }
        }
    }
    c->n = cur - 1;
    c->pd.assign((size_t)np * 6, 0); c->pl.assign((size_t)np * 6, 1);
    if (pd) c->pd.assign(pd, pd + np * 6);
    if (pl) c->pl.assign(pl, pl + np * 6);
    c->chain = true;
    for (int64_t e = 0; e < ne; e++) if (start[e] != e + 1 || stop[e] != e + 2) c->chain = false;
    if (np != ne + 1) c->chain = false;
    // canonical 3x3-block pattern (static): point.jl:1795-1797, element.jl:2507-2549
    c->brow.clear(); c->bcol.clear();
    if (is_static) {
        auto blk = [&](int64_t r, int64_t cc) { c->brow.push_back(r - 1); c->bcol.push_back(cc - 1); };
        for (int64_t ip = 0; ip < np; ip++) {
            int64_t ir = c->irow_point[ip], ic = c->icol_point[ip];
            blk(ir, ic); blk(ir, ic + 3); blk(ir + 3, ic + 3);
        }
        for (int64_t e = 0; e < ne; e++) {
            int64_t ic = c->icol_elem[e], ic1 = c->icol_point[start[e] - 1], ic2 = c->icol_point[stop[e] - 1];
            int64_t ir = c->irow_elem[e], ir1 = c->irow_point[start[e] - 1], ir2 = c->irow_point[stop[e] - 1];
            blk(ir, ic1); blk(ir, ic1 + 3); blk(ir, ic2); blk(ir, ic2 + 3); blk(ir, ic); blk(ir, ic + 3);
            blk(ir + 3, ic1 + 3); blk(ir + 3, ic2 + 3); blk(ir + 3, ic); blk(ir + 3, ic + 3);
            for (int64_t irp : {ir1, ir2}) {
                blk(irp, ic1 + 3); blk(irp, ic2 + 3); blk(irp, ic);
                blk(irp + 3, ic1 + 3); blk(irp + 3, ic2 + 3); blk(irp + 3, ic); blk(irp + 3, ic + 3);
            }
        }
        // unique (a point's own θ-block is also touched by its elements)
        std::vector<std::pair<int64_t, int64_t>> v;
        for (size_t k = 0; k < c->brow.size(); k++) v.push_back({c->brow[k], c->bcol[k]});
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        c->brow.clear(); c->bcol.clear();
        for (auto& q : v) { c->brow.push_back(q.first); c->bcol.push_back(q.second); }
    }
    if (!is_static) {
        // canonical 3x3-block pattern (dynamic, incl. the structural-damping blocks): point.jl:1795-1797, 1809-1826, 1920-1932;
        // element.jl:2507-2549, 2702-2757
        auto blk = [&](int64_t r, int64_t cc) { c->brow.push_back(r - 1); c->bcol.push_back(cc - 1); };
        for (int64_t ip = 0; ip < np; ip++) {
            int64_t ir = c->irow_point[ip], ic = c->icol_point[ip];
            for (int q = 0; q < 4; q++) { blk(ir, ic + 3 * q); blk(ir + 3, ic + 3 * q); }
            blk(ir + 6, ic); blk(ir + 6, ic + 6); blk(ir + 9, ic + 3); blk(ir + 9, ic + 9);
        }
        for (int64_t e = 0; e < ne; e++) {
            int64_t ic = c->icol_elem[e], ic1 = c->icol_point[start[e] - 1], ic2 = c->icol_point[stop[e] - 1];
            int64_t ir = c->irow_elem[e], ir1 = c->irow_point[start[e] - 1], ir2 = c->irow_point[stop[e] - 1];
            for (int64_t icp : {ic1, ic2}) {
                for (int q = 0; q < 4; q++) blk(ir, icp + 3 * q);
                blk(ir + 3, icp + 3); blk(ir + 3, icp + 9);
            }
            blk(ir, ic); blk(ir, ic + 3); blk(ir + 3, ic); blk(ir + 3, ic + 3);
            for (int64_t irp : {ir1, ir2}) {
                for (int64_t icp : {ic1, ic2}) for (int q = 0; q < 4; q++) { blk(irp, icp + 3 * q); blk(irp + 3, icp + 3 * q); }
                blk(irp, ic); blk(irp + 3, ic); blk(irp + 3, ic + 3);
            }
        }
        std::vector<std::pair<int64_t, int64_t>> v;
        for (size_t k = 0; k < c->brow.size(); k++) v.push_back({c->brow[k], c->bcol[k]});
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        c->brow.clear(); c->bcol.clear();
        for (auto& q : v) { c->brow.push_back(q.first); c->bcol.push_back(q.second); }
    }
    // update_body_acceleration_indices! (system.jl:138-150): the first point (ascending; the reference iterates a Dict) with pd & pl on DOF i
    c->body_state = false; c->nbody = 0;
    for (int i = 0; i < 6; i++) {
        c->ibody[i] = -1;
        if (is_static) continue;
        for (int64_t ip = 0; ip < np; ip++)
            if (c->pd[ip * 6 + i] && c->pl[ip * 6 + i]) { c->ibody[i] = (int)(c->icol_point[ip] - 1 + i); c->nbody++; c->body_state = true; break; }
    }
    if (is_static) { c->W = LuShape<2, 2>::W; c->UW = LuShape<2, 2>::UW; c->KLB = 2; c->rows_per_station = 12; }
    else { c->W = LuShape<3, 4>::W; c->UW = LuShape<3, 4>::UW; c->KLB = 3; c->rows_per_station = 18; }
    std::vector<unsigned short> flags(np);
    for (int64_t ip = 0; ip < np; ip++) {
        unsigned f = 0;
        for (int k = 0; k < 6; k++) { if (c->pd[ip * 6 + k]) f |= 1u << k; if (c->pl[ip * 6 + k]) f |= 64u << k; }
        flags[ip] = (unsigned short)f;
    }
    cudaFree(c->d_flags); cudaFree(c->d_brow); cudaFree(c->d_bcol); c->d_flags = nullptr; c->d_brow = c->d_bcol = nullptr;
    CK(cudaMalloc(&c->d_flags, np * sizeof(unsigned short)));
    CK(cudaMemcpy(c->d_flags, flags.data(), np * sizeof(unsigned short), cudaMemcpyHostToDevice));
    if (!c->brow.empty()) {
        CK(cudaMalloc(&c->d_brow, c->brow.size() * 8)); CK(cudaMalloc(&c->d_bcol, c->bcol.size() * 8));
        CK(cudaMemcpy(c->d_brow, c->brow.data(), c->brow.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_bcol, c->bcol.data(), c->bcol.size() * 8, cudaMemcpyHostToDevice));
    }
    free_batch(c);
    if (nstates) *nstates = c->n;
    if (irow_point) memcpy(irow_point, c->irow_point.data(), np * 8);
    if (icol_point) memcpy(icol_point, c->icol_point.data(), np * 8);
    if (irow_elem) memcpy(irow_elem, c->irow_elem.data(), ne * 8);
    if (icol_elem) memcpy(icol_elem, c->icol_elem.data(), ne * 8);
    return GXB_OK;
}

int gxb_pattern(gxb_ctx* c, int64_t* nblocks, int64_t* brow, int64_t* bcol) {
    if (!c || c->kind < 0) return fail(GXB_ERR_INVALID, "gxb_pattern: call gxb_topology first");
    if (nblocks) *nblocks = (int64_t)c->brow.size();
    if (brow) memcpy(brow, c->brow.data(), c->brow.size() * 8);
    if (bcol) memcpy(bcol, c->bcol.data(), c->bcol.size() * 8);
    return GXB_OK;
}

static int require_chain(gxb_ctx* c, const char* who) {
    if (!c || c->kind < 0) return fail(GXB_ERR_INVALID, std::string(who) + ": call gxb_topology first");
    if (!c->chain) return fail(GXB_ERR_UNSUPPORTED, std::string(who) + ": only chain topologies (start = 1:n, stop = 2:n+1) are implemented on the device");
    if (c->np > 256) return fail(GXB_ERR_UNSUPPORTED, std::string(who) + ": more than 255 elements per assembly not supported yet (one CTA per instance, shared-memory tile per station)");
    return GXB_OK;
}

int gxb_set_batch(gxb_ctx* c, int64_t nbatch) {
    int rc = require_chain(c, "gxb_set_batch"); if (rc) return rc;
    if (nbatch < 1) return fail(GXB_ERR_INVALID, "gxb_set_batch: nbatch < 1");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    free_batch(c);
    c->nbatch = nbatch;
    const size_t nb = (size_t)nbatch, n = (size_t)c->n, N = (size_t)c->ne, P = (size_t)c->np;
    if ((rc = dalloc(&c->d_elemS, nb * GXB_ELEM_NF * N))) return rc;
    if ((rc = dalloc(&c->d_bcS, nb * 18 * P))) return rc;
    if ((rc = dalloc(&c->d_fs, nb))) return rc;
    if ((rc = dalloc(&c->d_grav, nb * 3))) return rc;
    if ((rc = dalloc(&c->d_x, nb * n)) || (rc = dalloc(&c->d_xt, nb * n)) || (rc = dalloc(&c->d_F, nb * n)) || (rc = dalloc(&c->d_Ft, nb * n)) ||
        (rc = dalloc(&c->d_p, nb * n))) return rc;
    if ((rc = dalloc(&c->d_J, nb * n * c->W)) || (rc = dalloc(&c->d_U, nb * n * c->UW))) return rc;
    if (c->kind == 1) {
        if ((rc = dalloc(&c->d_exS, nb * 3 * N)) || (rc = dalloc(&c->d_xyz, nb * 3 * P)) || (rc = dalloc(&c->d_motion, nb * 12))) return rc;
        if ((rc = dalloc(&c->d_paug, nb * 12 * P)) || (rc = dalloc(&c->d_rates0, nb * 12 * P))) return rc;
        CK(cudaMemsetAsync(c->d_paug, 0, nb * 12 * P * 8, c->stream));
        CK(cudaMemsetAsync(c->d_xyz, 0, nb * 3 * P * 8, c->stream));
        CK(cudaMemsetAsync(c->d_motion, 0, nb * 12 * 8, c->stream));
    }
    if ((rc = dalloc(&c->ns.status, nb)) || (rc = dalloc(&c->ns.failcode, nb)) || (rc = dalloc(&c->ns.iters, nb)) || (rc = dalloc(&c->ns.fcalls, nb)) ||
        (rc = dalloc(&c->ns.ls_k, nb)) || (rc = dalloc(&c->ns.ls_finite, nb))) return rc;
    if ((rc = dalloc(&c->ns.a1, nb)) || (rc = dalloc(&c->ns.a2, nb)) || (rc = dalloc(&c->ns.phi0, nb)) || (rc = dalloc(&c->ns.dphi0, nb)) ||
        (rc = dalloc(&c->ns.phx0, nb)) || (rc = dalloc(&c->ns.finf, nb))) return rc;
    // structural zeros of the row-block storage are written once, here
    CK(cudaMemsetAsync(c->d_J, 0, nb * n * c->W * 8, c->stream));
    CK(cudaMemsetAsync(c->d_grav, 0, nb * 3 * 8, c->stream));
    CK(cudaMemsetAsync(c->d_x, 0, nb * n * 8, c->stream));
    std::vector<double> ones(nb, 1.0);
    CK(cudaMemcpyAsync(c->d_fs, ones.data(), nb * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->has_grav = 0; c->have_x0 = false;
    return GXB_OK;
}

#define NEED_BATCH(who) do { if (!c || c->nbatch < 1) return fail(GXB_ERR_INVALID, who ": call gxb_set_batch first"); CK(cudaSetDevice(c->device)); } while (0)

int gxb_set_elements(gxb_ctx* c, const double* elems) {
    NEED_BATCH("gxb_set_elements");
    if (!elems) return fail(GXB_ERR_INVALID, "gxb_set_elements: null");
    if (c->kind == 0 && c->skip_mass) {
        // only what the static path without gravity reads: L, x, compliance (40 doubles) and Cab (9 doubles) of every 91-double record
        const size_t ne = (size_t)c->nbatch * c->ne, bytes = ne * 49 * 8;
        if (bytes > c->stage_bytes) { cudaFree(c->d_stage); c->d_stage = nullptr; c->stage_bytes = 0; CK(cudaMalloc(&c->d_stage, bytes)); c->stage_bytes = bytes; }
        CK(cudaMemcpy2DAsync(c->d_stage, 49 * 8, elems, GXB_ELEM_STRIDE * 8, 40 * 8, ne, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpy2DAsync(c->d_stage + 40, 49 * 8, elems + 76, GXB_ELEM_STRIDE * 8, 9 * 8, ne, cudaMemcpyHostToDevice, c->stream));
        CK(cudaEventRecord(c->evC, c->stream));
        k_pack_elements_compact<<<grid1(c->nbatch * c->ne, 128), 128, 0, c->stream>>>(c->d_stage, c->d_elemS, (int)c->ne, c->nbatch);
        CK(cudaGetLastError());
        return wait_copy(c);
    }
    int rc = stage_upload(c, elems, (size_t)c->nbatch * c->ne * GXB_ELEM_STRIDE * 8); if (rc) return rc;
    k_pack_elements<<<grid1(c->nbatch * c->ne, 128), 128, 0, c->stream>>>(c->d_stage, c->d_elemS, c->d_exS, (int)c->ne, c->nbatch);
    CK(cudaGetLastError());
    return wait_copy(c);                     // the caller may reuse its buffer
}
int gxb_set_points(gxb_ctx* c, const double* xyz) {
    NEED_BATCH("gxb_set_points");
    // point coordinates only enter the dynamic residuals (Δx in point.jl:643, element.jl:186-187); unused by the static path
    if (c->kind != 1) return GXB_OK;
    if (!xyz) return fail(GXB_ERR_INVALID, "gxb_set_points: null");
    int rc = stage_upload(c, xyz, (size_t)c->nbatch * c->np * 3 * 8); if (rc) return rc;
    k_pack_transpose<<<grid1(c->nbatch * c->np * 3, 256), 256, 0, c->stream>>>(c->d_stage, c->d_xyz, (int)c->np, 3, c->nbatch);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    return GXB_OK;
}
int gxb_set_bc(gxb_ctx* c, const double* bc) {
    NEED_BATCH("gxb_set_bc");
    if (!bc) return fail(GXB_ERR_INVALID, "gxb_set_bc: null");
    int rc = stage_upload(c, bc, (size_t)c->nbatch * c->np * 18 * 8); if (rc) return rc;
    k_pack_transpose<<<grid1(c->nbatch * c->np * 18, 256), 256, 0, c->stream>>>(c->d_stage, c->d_bcS, (int)c->np, 18, c->nbatch);
    CK(cudaGetLastError());
    return wait_copy(c);
}
int gxb_set_dloads(gxb_ctx* c, const double* dl) {
    NEED_BATCH("gxb_set_dloads");
    if (!dl) { cudaFree(c->d_dlS); c->d_dlS = nullptr; return GXB_OK; }
    if (!c->d_dlS) { int rc = dalloc(&c->d_dlS, (size_t)c->nbatch * 24 * c->ne); if (rc) return rc; }
    int rc = stage_upload(c, dl, (size_t)c->nbatch * c->ne * 24 * 8); if (rc) return rc;
    k_pack_transpose<<<grid1(c->nbatch * c->ne * 24, 256), 256, 0, c->stream>>>(c->d_stage, c->d_dlS, (int)c->ne, 24, c->nbatch);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    return GXB_OK;
}
int gxb_set_motion(gxb_ctx* c, const double* motion) {
    NEED_BATCH("gxb_set_motion");
    if (c->kind != 1) return fail(GXB_ERR_INVALID, "gxb_set_motion: body-frame motion only applies to kind=1 (DynamicSystem)");
    if (motion) CK(cudaMemcpyAsync(c->d_motion, motion, (size_t)c->nbatch * 96, cudaMemcpyHostToDevice, c->stream));
    else CK(cudaMemsetAsync(c->d_motion, 0, (size_t)c->nbatch * 96, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return GXB_OK;
}
int gxb_set_pmass(gxb_ctx* c, const double* pm) {
    NEED_BATCH("gxb_set_pmass");
    if (!pm) { cudaFree(c->d_pmS); c->d_pmS = nullptr; return GXB_OK; }
    if (!c->d_pmS) { int rc = dalloc(&c->d_pmS, (size_t)c->nbatch * 36 * c->np); if (rc) return rc; }
    int rc = stage_upload(c, pm, (size_t)c->nbatch * c->np * 36 * 8); if (rc) return rc;
    k_pack_pmass<<<grid1(c->nbatch * c->np * 36, 256), 256, 0, c->stream>>>(c->d_stage, c->d_pmS, (int)c->np, c->nbatch);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    return GXB_OK;
}
// The caller promises that gravity is zero in this static context, so the element mass (36), damping (6) and midpoint fields are
// never read (46 of the 91 element doubles suffice for static analyses without gravity): gxb_set_elements then
// fetches 49 of the 91 doubles of every element record with two strided copies -- 54 % of the PCIe bytes.
int gxb_hint_no_gravity(gxb_ctx* c, int on) {
    if (!c) return fail(GXB_ERR_INVALID, "gxb_hint_no_gravity: null");
    if (on && c->kind != 0) return fail(GXB_ERR_INVALID, "gxb_hint_no_gravity: only static contexts (kind = 0) can skip the element mass");
    if (on && c->has_grav) return fail(GXB_ERR_INVALID, "gxb_hint_no_gravity: a non-zero gravity vector is set");
    c->skip_mass = on != 0;
    return GXB_OK;
}
int gxb_set_body(gxb_ctx* c, const double* gravity, const double* force_scaling) {
    NEED_BATCH("gxb_set_body");
    c->has_grav = 0;
    if (gravity) {
        for (int64_t i = 0; i < c->nbatch * 3; i++) if (gravity[i] != 0.0) { c->has_grav = 1; break; }
        if (c->has_grav && c->skip_mass) { c->has_grav = 0; return fail(GXB_ERR_INVALID, "gxb_set_body: non-zero gravity after gxb_hint_no_gravity (the element mass was not uploaded)"); }
        CK(cudaMemcpyAsync(c->d_grav, gravity, (size_t)c->nbatch * 24, cudaMemcpyHostToDevice, c->stream));
    } else CK(cudaMemsetAsync(c->d_grav, 0, (size_t)c->nbatch * 24, c->stream));
    if (force_scaling) {
        for (int64_t i = 0; i < c->nbatch; i++) if (!(force_scaling[i] > 0.0)) return fail(GXB_ERR_INVALID, "gxb_set_body: force_scaling must be positive");
        CK(cudaMemcpyAsync(c->d_fs, force_scaling, (size_t)c->nbatch * 8, cudaMemcpyHostToDevice, c->stream));
    }
    CK(cudaEventRecord(c->evC, c->stream));
    return wait_copy(c);
}

int gxb_upload_x0(gxb_ctx* c, const double* x0) {
    NEED_BATCH("gxb_upload_x0");
    if (x0) { CK(cudaMemcpyAsync(c->d_x, x0, (size_t)c->nbatch * c->n * 8, cudaMemcpyHostToDevice, c->stream)); c->have_x0 = true; }
    else c->have_x0 = false;
    return GXB_OK;
}

}  // extern "C"

// ---- launch helpers ----
struct Prof {
    gxb_ctx* c; bool on;
    void begin() { if (on) cudaEventRecord(c->evA, c->stream); }
    void end(double& acc) { if (on) { cudaEventRecord(c->evB, c->stream); cudaEventSynchronize(c->evB); float ms = 0; cudaEventElapsedTime(&ms, c->evA, c->evB); acc += ms; } }
};

static StaticArgs make_static_args(gxb_ctx* c, const gxb_opts* o, const double* x, double* r, bool want_j, int want_status, bool use_status) {
    StaticArgs A{};
    A.N = (int)c->ne; A.n = (int)c->n; A.two_d = o->two_dimensional; A.has_grav = c->has_grav;
    A.elemS = c->d_elemS; A.elem_nf = GXB_ELEM_NF; A.bcS = c->d_bcS; A.dlS = c->d_dlS; A.pmS = c->d_pmS; A.grav = c->d_grav; A.fs = c->d_fs;
    A.flags = c->d_flags; A.x = x; A.r = r; A.Jrow = want_j ? c->d_J : nullptr;
    A.status = use_status ? c->ns.status : nullptr; A.want_status = want_status;
    return A;
}
static int launch_assemble(gxb_ctx* c, const StaticArgs& A, bool want_j) {
    int bs = (int)((c->np + 31) / 32) * 32;
    if (c->kind == 0) {
        size_t sm = (size_t)bs * (GXB_XCH + (want_j ? GXB_TILE : 0)) * 8;
        if (want_j) {
            static thread_local size_t configured = 0;
            if (sm > configured) { CK(cudaFuncSetAttribute(k_assemble<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); configured = sm; }
            k_assemble<true><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(A);
        } else k_assemble<false><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(A);
    } else {
        DynArgs D{};
        D.s = A; D.exS = c->d_exS; D.pxS = c->d_xyz; D.motion = c->d_motion; D.init = c->d_paug; D.c2dt = c->c2dt; D.damping = c->damping;
        for (int i = 0; i < 6; i++) D.ibody[i] = c->ibody[i];
        // the shared-memory tile holds `chunk` stations; longer chains are flushed chunk by chunk
        const size_t budget = 200 * 1024;
        int chunk = (int)std::min<size_t>((size_t)bs, (budget - (size_t)bs * GXB_XCH * 8) / (GXB_DTILE * 8));
        D.chunk = chunk;
        size_t sm = ((size_t)bs * GXB_XCH + (want_j ? (size_t)chunk * GXB_DTILE : 0)) * 8;
        if (want_j) {
            static thread_local size_t configured = 0;
            if (sm > configured) {
                CK(cudaFuncSetAttribute(k_assemble_dynamic<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                CK(cudaFuncSetAttribute(k_assemble_dynamic<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                CK(cudaFuncSetAttribute(k_assemble_initial<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                configured = sm;
            }
            if (c->mode == 0) k_assemble_dynamic<true, 0><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(D);
            else if (c->mode == 1) k_assemble_dynamic<true, 1><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(D);
            else { InitArgs I{D, c->d_icS, c->d_rvm}; k_assemble_initial<true><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(I); }
        } else {
            if (c->mode == 0) k_assemble_dynamic<false, 0><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(D);
            else if (c->mode == 1) k_assemble_dynamic<false, 1><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(D);
            else { InitArgs I{D, c->d_icS, c->d_rvm}; k_assemble_initial<false><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(I); }
        }
    }
    CK(cudaGetLastError());
    c->stats.launches++; c->stats.n_assemble++;
    if (want_j && c->kind == 1 && c->nbody > 0 && c->mode != 1) {
        // dense Jacobian columns of the body-acceleration states + the unit entries the band matrix keeps in their place
        const size_t nb = (size_t)c->nbatch, n = (size_t)c->n;
        if (!c->d_D) {
            CK(cudaMalloc(&c->d_D, 6 * nb * n * 8)); CK(cudaMalloc(&c->d_Z, 6 * nb * n * 8));
            CK(cudaMemsetAsync(c->d_D, 0, 6 * nb * n * 8, c->stream)); CK(cudaMemsetAsync(c->d_Z, 0, 6 * nb * n * 8, c->stream));
        }
        BodyArgs B{};
        B.d.s = A; B.d.exS = c->d_exS; B.d.pxS = c->d_xyz; B.d.motion = c->d_motion;
        for (int i = 0; i < 6; i++) B.d.ibody[i] = c->ibody[i];
        B.icS = c->d_icS; B.rvm = c->d_rvm; B.mode = c->mode; B.Dn = c->d_D;
        k_body_columns<<<(unsigned)c->nbatch, bs, 0, c->stream>>>(B);
        CK(cudaGetLastError());
        c->stats.launches++;
    }
    return GXB_OK;
}
template <int KLB, int KUB> static int launch_factor_t(gxb_ctx* c, const SolveArgs& A) {
    using S = LuShape<KLB, KUB>;
    static const int wpc_env = []() { const char* e = getenv("GXB_LU_WPC"); int v = e ? atoi(e) : 4; return (v == 1 || v == 2 || v == 4) ? v : 4; }();
    const int wpc = wpc_env;      // warps (= instances) per CTA; the kernel is compiled for up to 128 threads per CTA
    size_t sm = (size_t)wpc * S::SM_DOUBLES * 8;
    static thread_local size_t configured = 0;
    if (sm > 48 * 1024 && sm > configured) { CK(cudaFuncSetAttribute(k_factor_solve<KLB, KUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); configured = sm; }
    k_factor_solve<KLB, KUB><<<grid1(c->nbatch, wpc), 32 * wpc, sm, c->stream>>>(A);
    CK(cudaGetLastError());
    return GXB_OK;
}
template <int KLB, int KUB> static int launch_factor_t(gxb_ctx* c, const SolveArgs& A);
static int launch_factor_body(gxb_ctx* c, const gxb_opts* o);
static int launch_factor(gxb_ctx* c, const gxb_opts* o, int direction_only, double sign = -1.0) {
    if (c->nbody > 0 && !direction_only) return launch_factor_body(c, o);
    SolveArgs A{};
    A.sign = sign;
    A.nb = (int)(c->n / 6); A.n = (int)c->n; A.nbatch = c->nbatch;
    A.J = c->d_J; A.F = direction_only ? c->d_Ft : c->d_F; A.x = c->d_x; A.U = c->d_U; A.p = c->d_p; A.xt = c->d_xt; A.ns = c->ns;
    A.maxstep = o->maxstep; A.linear = o->linear; A.direction_only = direction_only;
    int rc = c->kind == 0 ? launch_factor_t<2, 2>(c, A) : launch_factor_t<3, 4>(c, A);
    if (rc) return rc;
    c->stats.launches++; c->stats.n_factor++;
    return GXB_OK;
}
// Newton step with body-acceleration columns: 1 + nbody band solves (L is not stored, so each one re-factorises B) and the
// Woodbury correction.  A niche path (free-flying structures): clarity over speed.
static int launch_factor_body(gxb_ctx* c, const gxb_opts* o) {
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n;
    SolveArgs A{};
    A.sign = 1.0;
    A.nb = (int)(c->n / 6); A.n = (int)c->n; A.nbatch = c->nbatch;
    A.J = c->d_J; A.F = c->d_F; A.x = c->d_x; A.U = c->d_U; A.p = c->d_p; A.xt = c->d_xt; A.ns = c->ns;
    A.maxstep = o->maxstep; A.linear = o->linear; A.direction_only = 2;
    int rc;
    if ((rc = launch_factor_t<3, 4>(c, A))) return rc;
    for (int i = 0; i < 6; i++) {
        if (c->ibody[i] < 0) continue;
        SolveArgs Ai = A; Ai.F = c->d_D + (size_t)i * nb * n; Ai.p = c->d_Z + (size_t)i * nb * n;
        if ((rc = launch_factor_t<3, 4>(c, Ai))) return rc;
        c->stats.launches++; c->stats.n_factor++;
    }
    WoodArgs W{};
    W.S = A; W.S.direction_only = 0; W.Z = c->d_Z;
    for (int i = 0; i < 6; i++) W.ibody[i] = c->ibody[i];
    k_woodbury<<<grid1(c->nbatch, 4), 128, 0, c->stream>>>(W);
    CK(cudaGetLastError());
    c->stats.launches += 2; c->stats.n_factor++;
    return GXB_OK;
}
static int launch_update(gxb_ctx* c, const gxb_opts* o, int mode) {
    UpdateArgs A{};
    A.n = (int)c->n; A.mode = mode; A.x = c->d_x; A.xt = c->d_xt; A.F = c->d_F; A.Ft = c->d_Ft; A.p = c->d_p; A.ns = c->ns;
    A.ftol = o->ftol; A.c1 = o->c1; A.rho_hi = o->rho_hi; A.rho_lo = o->rho_lo; A.iterations = o->iterations; A.ls_iterations = 1000;
    A.counter = c->d_counter;
    k_update<<<(unsigned)c->nbatch, 128, 0, c->stream>>>(A);
    CK(cudaGetLastError());
    c->stats.launches++; c->stats.n_update++;
    return GXB_OK;
}

extern "C" {

static int solve_resident(gxb_ctx* c, const gxb_opts* opts, int mode = 0);
struct MarchBuffers;
static bool use_persistent(const gxb_ctx* c, const gxb_opts& o, bool marching);
static int64_t straggler_threshold(const gxb_ctx* c, const gxb_opts& o);
static int launch_persistent(gxb_ctx* c, const gxb_opts& o, int mode, const MarchBuffers* mb, int resume = 0);

int gxb_static_solve_resident(gxb_ctx* c, const gxb_opts* opts) {
    NEED_BATCH("gxb_static_solve_resident");
    if (c->kind != 0) return fail(GXB_ERR_INVALID, "gxb_static_solve_resident: context topology is not kind=0 (StaticSystem)");
    return solve_resident(c, opts);
}
int gxb_steady_solve_resident(gxb_ctx* c, const gxb_opts* opts) {
    NEED_BATCH("gxb_steady_solve_resident");
    if (c->kind != 1) return fail(GXB_ERR_INVALID, "gxb_steady_solve_resident: context topology is not kind=1 (DynamicSystem)");
    c->damping = (opts && opts->structural_damping) ? 1 : 0;
    int rc = solve_resident(c, opts);
    c->damping = 0;
    return rc;
}
int gxb_steady_solve(gxb_ctx* c, const gxb_opts* opts, const double* x0, double* x, int32_t* converged, int32_t* iters, double* resid_inf) {
    int rc;
    if ((rc = gxb_upload_x0(c, x0))) return rc;
    if ((rc = gxb_steady_solve_resident(c, opts))) return rc;
    return gxb_fetch(c, x, converged, iters, resid_inf);
}
// Newton rounds on whatever is in d_xt / status (instances in state TRIAL are evaluated first).  `burst` rounds are launched
// between two reads of the unfinished-instance counter (kernels early-exit on finished instances, so surplus rounds are cheap).
static int newton_rounds(gxb_ctx* c, const gxb_opts& o, Prof& prof, int burst) {
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n;
    int rc;
    CK(cudaMemsetAsync(c->d_counter, 0, 4, c->stream));
    // initial value_jacobian!!(df, x0)
    StaticArgs A = make_static_args(c, &o, c->d_xt, c->d_Ft, true, ST_TRIAL, true);
    prof.begin(); if ((rc = launch_assemble(c, A, true))) return rc; prof.end(c->stats.ms_assemble);
    prof.begin(); if ((rc = launch_update(c, &o, 0))) return rc; prof.end(c->stats.ms_update);
    if (o.linear) {
        // lsolve! (analyses.jl:3506-3524): x = x0 - J\r, converged = true.  State after mode 0 may be DONE (r == 0): force one step.
        k_fill_int<<<grid1(nb, 256), 256, 0, c->stream>>>(c->ns.status, ST_SOLVE, nb);
        prof.begin(); if ((rc = launch_factor(c, &o, 0))) return rc; prof.end(c->stats.ms_factor);
        // final residual r(x) is not re-evaluated by the reference; keep F = r(x0) like lsolve! leaves `resid`
        CK(cudaMemcpyAsync(c->d_Ft, c->d_F, nb * n * 8, cudaMemcpyDeviceToDevice, c->stream));
        prof.begin(); if ((rc = launch_update(c, &o, 2))) return rc; prof.end(c->stats.ms_update);
        c->stats.rounds += 1;
        return GXB_OK;
    }
    // Trial points are evaluated residual-first: the Jacobian (5-8x the work of the residual) is assembled only for the instances
    // whose step was accepted AND that have to iterate again -- the Jacobian at the converged point is never needed (NLsolve
    // evaluates value_jacobian!! at the top of the next iteration only), nor is it at rejected line-search points.
    // Measured (B200): config 4 1.72 -> 1.46 s, config 3 steady 294 -> 130 ms, config 1 8.1 -> 7.5 ms, config 2 9.00 -> 8.91 ms.
    static const bool residual_first = !(getenv("GXB_RESIDUAL_FIRST") && atoi(getenv("GXB_RESIDUAL_FIRST")) == 0);   // A/B switch
    StaticArgs A_next = make_static_args(c, &o, c->d_xt, c->d_Ft, true, ST_SOLVE, true);
    // the first burst goes out without looking at the counter: instances that were already converged at x0 early-exit
    for (;;) {
        for (int q = 0; q < burst; q++) {
            CK(cudaMemsetAsync(c->d_counter, 0, 4, c->stream));
            prof.begin(); if ((rc = launch_factor(c, &o, 0))) return rc; prof.end(c->stats.ms_factor);
            prof.begin(); if ((rc = launch_assemble(c, A, !residual_first))) return rc; prof.end(c->stats.ms_assemble);
            prof.begin(); if ((rc = launch_update(c, &o, 1))) return rc; prof.end(c->stats.ms_update);
            if (residual_first) { prof.begin(); if ((rc = launch_assemble(c, A_next, true))) return rc; prof.end(c->stats.ms_assemble); }
            c->stats.rounds += 1;
        }
        CK(cudaMemcpyAsync(c->h_counter, c->d_counter, 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (*c->h_counter == 0) break;
        // straggler hand-off: the few instances that are still iterating finish in the persistent kernel (one CTA each, no more
        // batch-wide rounds paced by the slowest instance); finished instances exit at once
        if (c->mode != 1 && *c->h_counter <= straggler_threshold(c, o)) {
            if ((rc = launch_persistent(c, o, c->mode, nullptr, 1))) return rc;
            break;
        }
    }
    return GXB_OK;
}

static int solve_resident(gxb_ctx* c, const gxb_opts* opts, int mode) {
    gxb_opts o; if (opts) o = *opts; else gxb_default_opts(&o);
    c->stats = gxb_stats{};
    Prof prof{c, o.profile != 0};
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n;
    int rc;
    c->mode = mode;
    CK(cudaEventRecord(c->ev0, c->stream));
    if (!c->have_x0) CK(cudaMemsetAsync(c->d_x, 0, nb * n * 8, c->stream));   // reset_state (system.jl:383-386)
    CK(cudaMemcpyAsync(c->d_xt, c->d_x, nb * n * 8, cudaMemcpyDeviceToDevice, c->stream));
    k_fill_int<<<grid1(nb, 256), 256, 0, c->stream>>>(c->ns.status, ST_TRIAL, nb);
    CK(cudaMemsetAsync(c->ns.failcode, 0, nb * 4, c->stream)); CK(cudaMemsetAsync(c->ns.iters, 0, nb * 4, c->stream));
    CK(cudaMemsetAsync(c->ns.fcalls, 0, nb * 4, c->stream));
    if (use_persistent(c, o, false)) rc = launch_persistent(c, o, mode, nullptr);
    else rc = newton_rounds(c, o, prof, 1);
    if (rc) { c->mode = 0; return rc; }
    c->mode = 0;
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->stats.ms_total = ms;
    // totals
    std::vector<int> it(nb), fc(nb);
    CK(cudaMemcpy(it.data(), c->ns.iters, nb * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(fc.data(), c->ns.fcalls, nb * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < nb; i++) { c->stats.newton_iters += it[i]; c->stats.resid_evals += fc[i]; }
    c->have_x0 = false;
    return GXB_OK;
}

// ---- Newmark time marching (time_domain_analysis! from a given initial state, src/analyses.jl:2600-2790) ----
// per-instance bookkeeping of the time loop: an instance whose step does not converge stops marching (analyses.jl:2781-2790)
// one launch per time step: prescribed values of the new time level, `paug` update (analyses.jl:2693-2738), Newton state reset.
// One CTA per instance (the series entries are written before the points read bcS: same CTA, one barrier).
__global__ void k_step_prepare(double* __restrict__ bcS, const double* __restrict__ series, const int* __restrict__ ser_point,
                               const int* __restrict__ ser_field, int K, int64_t ser_nt, int64_t step,
                               const double* __restrict__ x, const unsigned short* __restrict__ flags, double* __restrict__ paug,
                               const double* __restrict__ rates0, int NP, int n, int first, double c_prev, double c_new,
                               int* status, int* iters, int* fcalls, const int* alive) {
    const int64_t b = blockIdx.x;
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        bcS[((size_t)b * 18 + ser_field[k]) * NP + ser_point[k]] = series[((size_t)b * ser_nt + step) * K + k];
    __syncthreads();
    for (int p = threadIdx.x; p < NP; p += blockDim.x) {
        const double* xp = x + b * n + 18 * p;
        const double* bc = bcS + ((size_t)b * 18) * NP + p;
        const unsigned fl = flags[p];
        double* pa = paug + ((size_t)b * 12) * NP + p;
        for (int k = 0; k < 12; k++) {
            const double sv = (k < 6 && (fl & (1u << k))) ? bc[k * NP] : xp[k];    // u, θ honour prescribed displacements at the new time
            const double rate = first ? rates0[((size_t)b * NP + p) * 12 + k] : c_prev * sv - pa[(size_t)k * NP];
            pa[(size_t)k * NP] = c_new * sv + rate;
        }
    }
    if (threadIdx.x == 0) { status[b] = alive[b] ? ST_TRIAL : ST_FAILED; iters[b] = 0; fcalls[b] = 0; }
}
__global__ void k_step_end(const int* status, const int* iters, int* alive, int* steps_done, int* iters_total, int64_t nbatch) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= nbatch) return;
    if (!alive[b]) return;
    steps_done[b] += 1;
    iters_total[b] += iters[b];
    if (status[b] != ST_DONE) alive[b] = 0;
}

}  // extern "C"   (persistent kernels are templates)

// ------------------------------------------------------------------------------------------------------------------
// Persistent per-instance solver: ONE CTA owns one instance and runs its whole Newton iteration -- and, for time marching, its
// whole time loop -- without returning to the host: assembly by the CTA (thread per station), pivoted band LU by warp 0, Armijo /
// convergence logic by the CTA, all through the same device functions and the same per-instance state as the multi-kernel path.
// This is the latency regime's answer (few instances per SM: BASELINE config 4 has 512, config 3's stragglers run alone): no
// launch or host-synchronisation cost per round, and every instance does exactly the rounds IT needs instead of the batch maximum.
// MODE 0: steady solve, 1: Newmark march (nt time levels), 2: initial-condition solve.
struct PersistArgs {
    DynArgs D;                 // D.s.x = xt, D.s.r = Ft, D.s.Jrow = J (as in the multi-kernel rounds)
    const double* icS; const unsigned* rvm;     // MODE 2
    SolveArgs S;
    UpdateArgs Up;             // Up.counter = nullptr
    int resume;                // MODE 0/2: continue instances left in ST_SOLVE / ST_TRIAL by the multi-kernel rounds (stragglers)
    // time marching (MODE 1)
    int64_t nt; const double* tvec;             // device copy of the time vector
    double* bcS; const double* series; const int* ser_point; const int* ser_field; int K; int64_t ser_nt;
    double* paug; const double* rates0;
    int* alive; int* steps; int* itot;
    double* hist; int64_t nsave, save_every;
};

// separate (non-inlined) functions: each phase gets its own register allocation instead of one allocation for the fused kernel
__device__ __noinline__ void persist_factor(const SolveArgs& S, int b, double* smem) { factor_instance<3, 4>(S, b, smem); }
__device__ __noinline__ int persist_update(const UpdateArgs& U, int b) { return update_instance(U, b); }
template <int MODE, bool WANT_J> __device__ __noinline__ void persist_assemble(const PersistArgs& P, const DynArgs& D, int b, double* smem) {
    if (MODE == 2) { InitArgs I{D, P.icS, P.rvm}; initial_station<WANT_J>(I, b, threadIdx.x, smem); }
    else dynamic_station<WANT_J, (MODE == 2 ? 0 : MODE)>(D, b, threadIdx.x, smem);
}

template <int MODE> __global__ void __launch_bounds__(256) k_persistent_dynamic(PersistArgs P) {
    extern __shared__ __align__(16) double psm[];
    const int b = blockIdx.x;
    const int n = P.Up.n, NP = P.D.s.N + 1;
    double* x = P.Up.x + (size_t)b * n;
    double* xt = P.Up.xt + (size_t)b * n;
    DynArgs D = P.D;
    const int64_t nsteps = MODE == 1 ? P.nt - 1 : 1;
#ifdef GXB_PERSIST_TIMING
    long long ptime[4] = {0, 0, 0, 0};
    const long long t_begin = clock64();
#endif
    for (int64_t it = 1; it <= nsteps; it++) {
        if (MODE == 1) {
            if (!P.alive[b]) break;                                   // uniform: written by thread 0 before a barrier
            const double dt = P.tvec[it] - P.tvec[it - 1];
            const double c_prev = it > 1 ? 2.0 / (P.tvec[it - 1] - P.tvec[it - 2]) : 0.0;
            const double c_new = 2.0 / dt;
            D.c2dt = c_new;
            // prescribed values of the new time level, then the `paug` update (analyses.jl:2693-2738)
            for (int k = threadIdx.x; k < P.K; k += blockDim.x)
                P.bcS[((size_t)b * 18 + P.ser_field[k]) * NP + P.ser_point[k]] = P.series[((size_t)b * P.ser_nt + it) * P.K + k];
            __syncthreads();
            for (int p = threadIdx.x; p < NP; p += blockDim.x) {
                const double* xp = x + 18 * p;
                const double* bc = P.bcS + ((size_t)b * 18) * NP + p;
                const unsigned fl = P.D.s.flags[p];
                double* pa = P.paug + ((size_t)b * 12) * NP + p;
                for (int k = 0; k < 12; k++) {
                    const double sv = (k < 6 && (fl & (1u << k))) ? bc[k * NP] : xp[k];
                    const double rate = it == 1 ? P.rates0[((size_t)b * NP + p) * 12 + k] : c_prev * sv - pa[(size_t)k * NP];
                    pa[(size_t)k * NP] = c_new * sv + rate;
                }
            }
            for (int c = threadIdx.x; c < n; c += blockDim.x) xt[c] = x[c];     // initial guess = previous state (analyses.jl:3592)
            if (threadIdx.x == 0) { P.Up.ns.iters[b] = 0; P.Up.ns.fcalls[b] = 0; }
            __syncthreads();
        }
        // value_jacobian!! at the initial guess, then Newton rounds until this instance is done.  One call site per device function:
        // the station code is tens of thousands of instructions, two inlined copies would thrash the instruction cache.
        int st = ST_TRIAL;
        UpdateArgs Uc = P.Up; Uc.mode = 0;
        if (MODE != 1 && P.resume) {
            st = P.Up.ns.status[b];
            if (st != ST_SOLVE && st != ST_TRIAL) return;
            Uc.mode = 1;
        }
#ifdef GXB_PERSIST_TIMING
        long long tq = clock64(), tn;
#define GXB_PTICK(i) do { tn = clock64(); ptime[i] += tn - tq; tq = tn; } while (0)
#else
#define GXB_PTICK(i) do { } while (0)
#endif
        for (;;) {
            if (st == ST_SOLVE) {
                // the assembly wrote its tiles into psm with ordinary stores; the factorisation fills the same bytes through the async
                // proxy (TMA bulk copies): order the two proxies before warp 0 starts
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                __syncthreads();
                if (threadIdx.x < 32) persist_factor(P.S, b, psm);
                __syncthreads();
                st = P.Up.ns.status[b];
                GXB_PTICK(0);
                if (st != ST_TRIAL) break;                            // singular Jacobian
            }
            const bool first_eval = Uc.mode == 0;
            if (first_eval) persist_assemble<MODE, true>(P, D, b, psm);       // value_jacobian!! at the initial guess
            else persist_assemble<MODE, false>(P, D, b, psm);                 // trial points: residual only
            __syncthreads();
            GXB_PTICK(1);
            st = persist_update(Uc, b);
            Uc.mode = 1;
            __syncthreads();
            GXB_PTICK(2);
            if (st != ST_SOLVE && st != ST_TRIAL) break;
            if (st == ST_SOLVE && !first_eval) {                              // accepted, iterate again: now the Jacobian at x
                persist_assemble<MODE, true>(P, D, b, psm);
                __syncthreads();
                GXB_PTICK(1);
            }
        }
        if (MODE == 1) {
            if (threadIdx.x == 0) {
                P.steps[b] += 1; P.itot[b] += P.Up.ns.iters[b];
                if (st != ST_DONE) P.alive[b] = 0;
            }
            if (P.hist && it % P.save_every == 0) {
                double* h = P.hist + ((size_t)b * P.nsave + it / P.save_every) * n;
                for (int c = threadIdx.x; c < n; c += blockDim.x) h[c] = x[c];
            }
            __syncthreads();
        }
    }
#ifdef GXB_PERSIST_TIMING
    if (threadIdx.x == 0 && (b == 0 || b == gridDim.x - 1))
        printf("[persistent MODE=%d b=%d] cycles: factor %lld, assemble %lld, update %lld, total %lld (steps %lld)\n", MODE, b, ptime[0], ptime[1], ptime[2],
               clock64() - t_begin, (long long)nsteps);
#endif
}

extern "C" {

// persistent per-instance path: policy + launch.  opts->persistent: 0 auto, 1 on, 2 off.  Auto picks it in the latency regime
// (all CTAs co-resident with few warps per SM), where the multi-kernel rounds are bound by launch/synchronisation and by the
// slowest instance of the batch.
static bool use_persistent(const gxb_ctx* c, const gxb_opts& o, bool marching) {
    if (c->kind != 1 || o.linear || o.profile || c->nbody > 0) return false;
    if (o.persistent == 1) return true;
    if (o.persistent == 2) return false;
    // measured on B200: with every instance needing about the same number of rounds the multi-kernel path wins or
    // ties (its factorisation kernel keeps one warp per SM sub-partition); the persistent kernel pays off when a few instances need
    // many more rounds than the rest -- which the multi-kernel driver detects itself and hands over (straggler hand-off below)
    (void)marching;
    return false;
}
// straggler hand-off: once at most this many instances are unfinished, the remaining rounds run in the persistent kernel
static int64_t straggler_threshold(const gxb_ctx* c, const gxb_opts& o) {
    if (c->kind != 1 || o.linear || o.profile || o.persistent == 2 || c->nbody > 0) return -1;
    return std::min<int64_t>(c->sm_count, std::max<int64_t>(1, c->nbatch / 8));
}
struct MarchBuffers { int64_t nt = 0; const double* d_tvec = nullptr; int* alive = nullptr; int* steps = nullptr; int* itot = nullptr; double* hist = nullptr; int64_t nsave = 0, save_every = 1; };
static int launch_persistent(gxb_ctx* c, const gxb_opts& o, int mode, const MarchBuffers* mb, int resume) {
    PersistArgs P{};
    P.resume = resume;
    StaticArgs A = make_static_args(c, &o, c->d_xt, c->d_Ft, true, 0, false);
    P.D.s = A; P.D.exS = c->d_exS; P.D.pxS = c->d_xyz; P.D.motion = c->d_motion; P.D.init = c->d_paug; P.D.c2dt = c->c2dt; P.D.damping = c->damping;
    for (int i = 0; i < 6; i++) P.D.ibody[i] = c->ibody[i];
    P.icS = c->d_icS; P.rvm = c->d_rvm;
    const int bs = (int)((c->np + 31) / 32) * 32;
    // a smaller tile chunk than the multi-kernel assembly: shared memory per CTA decides how many instances are resident per SM
    const size_t budget = 64 * 1024;
    int chunk = (int)std::min<size_t>((size_t)bs, (budget - (size_t)bs * GXB_XCH * 8) / (GXB_DTILE * 8));
    P.D.chunk = chunk;
    size_t sm = std::max(((size_t)bs * GXB_XCH + (size_t)chunk * GXB_DTILE) * 8, (size_t)LuShape<3, 4>::SM_DOUBLES * 8);
    P.S.sign = -1.0; P.S.nb = (int)(c->n / 6); P.S.n = (int)c->n; P.S.nbatch = c->nbatch;
    P.S.J = c->d_J; P.S.F = c->d_F; P.S.x = c->d_x; P.S.U = c->d_U; P.S.p = c->d_p; P.S.xt = c->d_xt; P.S.ns = c->ns;
    P.S.maxstep = o.maxstep; P.S.linear = 0; P.S.direction_only = 0;
    P.Up.n = (int)c->n; P.Up.x = c->d_x; P.Up.xt = c->d_xt; P.Up.F = c->d_F; P.Up.Ft = c->d_Ft; P.Up.p = c->d_p; P.Up.ns = c->ns;
    P.Up.ftol = o.ftol; P.Up.c1 = o.c1; P.Up.rho_hi = o.rho_hi; P.Up.rho_lo = o.rho_lo; P.Up.iterations = o.iterations; P.Up.ls_iterations = 1000;
    P.Up.counter = nullptr;
    if (mb) {
        P.nt = mb->nt; P.tvec = mb->d_tvec; P.bcS = c->d_bcS; P.series = c->d_series; P.ser_point = c->d_ser_point; P.ser_field = c->d_ser_field;
        P.K = c->ser_K; P.ser_nt = c->ser_nt; P.paug = c->d_paug; P.rates0 = c->d_rates0;
        P.alive = mb->alive; P.steps = mb->steps; P.itot = mb->itot; P.hist = mb->hist; P.nsave = mb->nsave; P.save_every = mb->save_every;
    }
    static thread_local size_t configured = 0;
    if (sm > configured) {
        CK(cudaFuncSetAttribute(k_persistent_dynamic<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CK(cudaFuncSetAttribute(k_persistent_dynamic<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CK(cudaFuncSetAttribute(k_persistent_dynamic<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        configured = sm;
    }
    if (mode == 0) k_persistent_dynamic<0><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(P);
    else if (mode == 1) k_persistent_dynamic<1><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(P);
    else k_persistent_dynamic<2><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(P);
    CK(cudaGetLastError());
    c->stats.launches += 1; c->stats.rounds += 1;
    return GXB_OK;
}

int gxb_set_bc_series(gxb_ctx* c, int64_t nt, int32_t K, const int32_t* point, const int32_t* field, const double* values) {
    NEED_BATCH("gxb_set_bc_series");
    cudaFree(c->d_series); cudaFree(c->d_ser_point); cudaFree(c->d_ser_field);
    c->d_series = nullptr; c->d_ser_point = c->d_ser_field = nullptr; c->ser_nt = 0; c->ser_K = 0;
    if (K == 0 || !values) return GXB_OK;
    if (nt < 1 || K < 0 || !point || !field) return fail(GXB_ERR_INVALID, "gxb_set_bc_series: bad arguments");
    std::vector<int> pt(K), fd(K);
    for (int k = 0; k < K; k++) {
        if (point[k] < 1 || point[k] > c->np || field[k] < 0 || field[k] >= 18) return fail(GXB_ERR_INVALID, "gxb_set_bc_series: point is 1-based, field in 0..17");
        pt[k] = point[k] - 1; fd[k] = field[k];
    }
    CK(cudaMalloc(&c->d_series, (size_t)c->nbatch * nt * K * 8));
    CK(cudaMalloc(&c->d_ser_point, K * 4)); CK(cudaMalloc(&c->d_ser_field, K * 4));
    CK(cudaMemcpy(c->d_series, values, (size_t)c->nbatch * nt * K * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_ser_point, pt.data(), K * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_ser_field, fd.data(), K * 4, cudaMemcpyHostToDevice));
    c->ser_nt = nt; c->ser_K = K;
    return GXB_OK;
}

int gxb_newmark_run(gxb_ctx* c, const gxb_opts* opts, int64_t nt, const double* tvec, const double* x0, const double* rates0,
                    int64_t save_every, double* x_hist, int32_t* steps_done, int32_t* iters_total, int32_t* converged) {
    NEED_BATCH("gxb_newmark_run");
    if (c->kind != 1) return fail(GXB_ERR_INVALID, "gxb_newmark_run: context topology is not kind=1 (DynamicSystem)");
    if (nt < 2 || !tvec) return fail(GXB_ERR_INVALID, "gxb_newmark_run: need nt >= 2 and tvec");
    if ((x0 == nullptr) != (rates0 == nullptr)) return fail(GXB_ERR_INVALID, "gxb_newmark_run: give both x0 and the initial rates, or neither");
    if (!x0 && !c->have_initial) return fail(GXB_ERR_INVALID, "gxb_newmark_run: no initial state (pass x0 and rates0, or call gxb_initial_condition first)");
    gxb_opts o; if (opts) o = *opts; else gxb_default_opts(&o);
    c->damping = o.structural_damping ? 1 : 0;
    // linear=true: one lsolve! per step, about the current state (update_linearization) or about the initial state (analyses.jl:2741-2755)
    double* d_xlin = nullptr;
    if (c->ser_K && c->ser_nt < nt) return fail(GXB_ERR_INVALID, "gxb_newmark_run: the bc series has fewer time levels than tvec");
    if (save_every < 1) save_every = 1;
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n, P = (size_t)c->np;
    const int64_t nsave = (nt - 1) / save_every + 1;      // time levels 0, save_every, 2 save_every, ...
    c->stats = gxb_stats{};
    Prof prof{c, o.profile != 0};
    int rc;
    int *d_alive = nullptr, *d_steps = nullptr, *d_itot = nullptr; double* d_hist = nullptr;
    CK(cudaMalloc(&d_alive, nb * 4)); CK(cudaMalloc(&d_steps, nb * 4)); CK(cudaMalloc(&d_itot, nb * 4));
    if (x_hist) CK(cudaMalloc(&d_hist, nb * nsave * n * 8));
    k_fill_int<<<grid1(nb, 256), 256, 0, c->stream>>>(d_alive, 1, nb);
    k_fill_int<<<grid1(nb, 256), 256, 0, c->stream>>>(d_steps, 1, nb);
    CK(cudaMemsetAsync(d_itot, 0, nb * 4, c->stream));
    CK(cudaMemsetAsync(c->ns.failcode, 0, nb * 4, c->stream));
    if (x0) {
        CK(cudaMemcpyAsync(c->d_x, x0, nb * n * 8, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_rates0, rates0, nb * P * 12 * 8, cudaMemcpyHostToDevice, c->stream));
    }
    c->have_initial = false;
    if (o.linear && !o.update_linearization) {
        CK(cudaMalloc(&d_xlin, nb * n * 8));
        CK(cudaMemcpyAsync(d_xlin, c->d_x, nb * n * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    CK(cudaMemsetAsync(c->d_paug, 0, nb * 12 * P * 8, c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    if (d_hist) k_save_state<<<grid1(nb * n, 256), 256, 0, c->stream>>>(c->d_x, d_hist, (int)n, nsave, 0, c->nbatch);
    c->mode = 1;
    for (int64_t it = 1; it < nt; it++)
        if (!(tvec[it] - tvec[it - 1] > 0.0)) { c->mode = 0; return fail(GXB_ERR_INVALID, "gxb_newmark_run: tvec must be strictly increasing"); }
    double* d_tvec = nullptr;
    const bool persistent = use_persistent(c, o, true);
    if (persistent) {
        CK(cudaMalloc(&d_tvec, nt * 8));
        CK(cudaMemcpyAsync(d_tvec, tvec, nt * 8, cudaMemcpyHostToDevice, c->stream));
        MarchBuffers mb; mb.nt = nt; mb.d_tvec = d_tvec; mb.alive = d_alive; mb.steps = d_steps; mb.itot = d_itot; mb.hist = d_hist; mb.nsave = nsave; mb.save_every = save_every;
        if ((rc = launch_persistent(c, o, 1, &mb))) { c->mode = 0; cudaFree(d_tvec); return rc; }
    }
    for (int64_t it = 1; it < nt && !persistent; it++) {
        const double dt = tvec[it] - tvec[it - 1];
        const double c_prev = it > 1 ? 2.0 / (tvec[it - 1] - tvec[it - 2]) : 0.0;
        c->c2dt = 2.0 / dt;
        k_step_prepare<<<(unsigned)nb, 32, 0, c->stream>>>(c->d_bcS, c->d_series, c->d_ser_point, c->d_ser_field, c->ser_K, c->ser_nt, it, c->d_x,
                                                           c->d_flags, c->d_paug, c->d_rates0, (int)P, (int)n, it == 1, c_prev, c->c2dt,
                                                           c->ns.status, c->ns.iters, c->ns.fcalls, d_alive);
        CK(cudaMemcpyAsync(c->d_xt, d_xlin ? d_xlin : c->d_x, nb * n * 8, cudaMemcpyDeviceToDevice, c->stream));   // initial guess = previous state (analyses.jl:3592)
        c->stats.launches += 1;
        if ((rc = newton_rounds(c, o, prof, 2))) { c->mode = 0; return rc; }
        k_step_end<<<grid1(nb, 256), 256, 0, c->stream>>>(c->ns.status, c->ns.iters, d_alive, d_steps, d_itot, c->nbatch);
        c->stats.launches += 1;
        if (d_hist && it % save_every == 0) { k_save_state<<<grid1(nb * n, 256), 256, 0, c->stream>>>(c->d_x, d_hist, (int)n, nsave, it / save_every, c->nbatch); c->stats.launches += 1; }
    }
    c->mode = 0;
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->stats.ms_total = ms;
    std::vector<int> al(nb), itot(nb);
    CK(cudaMemcpy(al.data(), d_alive, nb * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(itot.data(), d_itot, nb * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < nb; i++) c->stats.newton_iters += itot[i];
    if (converged) for (size_t i = 0; i < nb; i++) converged[i] = al[i];
    if (iters_total) memcpy(iters_total, itot.data(), nb * 4);
    if (steps_done) CK(cudaMemcpy(steps_done, d_steps, nb * 4, cudaMemcpyDeviceToHost));
    if (x_hist) CK(cudaMemcpy(x_hist, d_hist, nb * nsave * n * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_alive); cudaFree(d_steps); cudaFree(d_itot); cudaFree(d_hist); cudaFree(d_tvec); cudaFree(d_xlin);
    CK(cudaGetLastError());
    return GXB_OK;
}

int gxb_fetch(gxb_ctx* c, double* x, int32_t* converged, int32_t* iters, double* resid_inf) {
    NEED_BATCH("gxb_fetch");
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n;
    if (x) CK(cudaMemcpyAsync(x, c->d_x, nb * n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (iters) CK(cudaMemcpyAsync(iters, c->ns.iters, nb * 4, cudaMemcpyDeviceToHost, c->stream));
    if (resid_inf) CK(cudaMemcpyAsync(resid_inf, c->ns.finf, nb * 8, cudaMemcpyDeviceToHost, c->stream));
    std::vector<int> st;
    if (converged) { st.resize(nb); CK(cudaMemcpyAsync(st.data(), c->ns.status, nb * 4, cudaMemcpyDeviceToHost, c->stream)); }
    CK(cudaStreamSynchronize(c->stream));
    if (converged) for (size_t i = 0; i < nb; i++) converged[i] = st[i] == ST_DONE;
    return GXB_OK;
}

int gxb_get_stats(gxb_ctx* c, gxb_stats* st) {
    if (!c || !st) return fail(GXB_ERR_INVALID, "gxb_get_stats: null");
    *st = c->stats;
    return GXB_OK;
}

int gxb_static_solve(gxb_ctx* c, const gxb_opts* opts, const double* x0, double* x, int32_t* converged, int32_t* iters, double* resid_inf) {
    int rc;
    if ((rc = gxb_upload_x0(c, x0))) return rc;
    if ((rc = gxb_static_solve_resident(c, opts))) return rc;
    return gxb_fetch(c, x, converged, iters, resid_inf);
}

int gxb_residual_jacobian(gxb_ctx* c, const gxb_opts* opts, const double* x, double* r, double* Jblk) {
    NEED_BATCH("gxb_residual_jacobian");
    c->damping = (c->kind == 1 && opts && opts->structural_damping) ? 1 : 0;
    if (!x) return fail(GXB_ERR_INVALID, "gxb_residual_jacobian: x is null");
    gxb_opts o; if (opts) o = *opts; else gxb_default_opts(&o);
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n;
    CK(cudaMemcpyAsync(c->d_xt, x, nb * n * 8, cudaMemcpyHostToDevice, c->stream));
    StaticArgs A = make_static_args(c, &o, c->d_xt, c->d_Ft, Jblk != nullptr, 0, false);
    int rc = launch_assemble(c, A, Jblk != nullptr); if (rc) return rc;
    if (r) CK(cudaMemcpyAsync(r, c->d_Ft, nb * n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (Jblk) {
        const size_t nblk = c->brow.size();
        size_t bytes = nb * nblk * 9 * 8;
        if (bytes > c->stage_bytes) { cudaFree(c->d_stage); c->d_stage = nullptr; c->stage_bytes = 0; CK(cudaMalloc(&c->d_stage, bytes)); c->stage_bytes = bytes; }
        k_extract_blocks<<<grid1(nb * nblk * 9, 256), 256, 0, c->stream>>>(c->d_J, c->d_brow, c->d_bcol, c->d_stage, (int64_t)nblk, (int)n, c->W, c->KLB, c->nbatch);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(Jblk, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return GXB_OK;
}

int gxb_newmark_residual_jacobian(gxb_ctx* c, const gxb_opts* opts, const double* init, double dt, const double* x, double* r, double* Jblk) {
    NEED_BATCH("gxb_newmark_residual_jacobian");
    if (c->kind != 1) return fail(GXB_ERR_INVALID, "gxb_newmark_residual_jacobian: context topology is not kind=1 (DynamicSystem)");
    if (!init || !(dt > 0.0)) return fail(GXB_ERR_INVALID, "gxb_newmark_residual_jacobian: need the initialisation terms and dt > 0");
    int rc = stage_upload(c, init, (size_t)c->nbatch * c->np * 12 * 8); if (rc) return rc;
    k_pack_transpose<<<grid1(c->nbatch * c->np * 12, 256), 256, 0, c->stream>>>(c->d_stage, c->d_paug, (int)c->np, 12, c->nbatch);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    c->mode = 1; c->c2dt = 2.0 / dt;
    rc = gxb_residual_jacobian(c, opts, x, r, Jblk);
    c->mode = 0;
    return rc;
}

// ---- initial condition analysis (initial_condition_analysis!, src/analyses.jl:1695-1945) ----
// uploads the initial values (ic: nbatch x np x 18 = u0, θ0, V0, Ω0, Vdot0, Ωdot0 per point, NULL = zeros) and builds the rate_vars
// masks: computed on the device (analyses.jl:1835-1873) unless both rate_vars1 / rate_vars2 (nstates each, shared by the batch) are given
static int initial_setup(gxb_ctx* c, const gxb_opts& o, const double* ic, const uint8_t* rv1, const uint8_t* rv2, const char* who) {
    if (c->kind != 1) return fail(GXB_ERR_INVALID, std::string(who) + ": context topology is not kind=1 (DynamicSystem)");
    if ((rv1 == nullptr) != (rv2 == nullptr)) return fail(GXB_ERR_INVALID, std::string(who) + ": give both rate_vars1 and rate_vars2 or neither");
    const size_t nb = (size_t)c->nbatch, P = (size_t)c->np;
    if (!c->d_icS) { CK(cudaMalloc(&c->d_icS, nb * 18 * P * 8)); CK(cudaMalloc(&c->d_rvm, nb * P * 4)); }
    if (ic) {
        int rc = stage_upload(c, ic, nb * P * 18 * 8); if (rc) return rc;
        k_pack_transpose<<<grid1(nb * P * 18, 256), 256, 0, c->stream>>>(c->d_stage, c->d_icS, (int)P, 18, c->nbatch);
        CK(cudaGetLastError());
    } else CK(cudaMemsetAsync(c->d_icS, 0, nb * 18 * P * 8, c->stream));
    if (rv1) {
        std::vector<unsigned> m(nb * P, 0u);
        for (size_t p = 0; p < P; p++) {
            unsigned v = 0;
            for (int k = 0; k < 6; k++) {
                const bool r1 = rv1[18 * p + 6 + k] != 0, r2 = rv2[18 * p + 6 + k] != 0, pd = c->pd[6 * p + k] != 0;
                if (r2) v |= 1u << k;
                if (!pd && !r1 && r2) v |= 1u << (8 + k);
                if (r1) v |= 1u << (16 + k);
            }
            for (size_t b = 0; b < nb; b++) m[b * P + p] = v;
        }
        CK(cudaMemcpy(c->d_rvm, m.data(), nb * P * 4, cudaMemcpyHostToDevice));
    } else {
        k_initial_rate_vars<<<grid1(nb * P, 128), 128, 0, c->stream>>>(c->d_elemS, GXB_ELEM_NF, c->d_bcS, c->d_pmS, c->d_flags, (int)c->ne,
                                                                       o.two_dimensional, c->d_rvm, c->nbatch);
        CK(cudaGetLastError());
    }
    c->stats.launches += 1;
    return GXB_OK;
}

int gxb_initial_residual_jacobian(gxb_ctx* c, const gxb_opts* opts, const double* ic, const uint8_t* rate_vars1, const uint8_t* rate_vars2,
                                  const double* x, double* r, double* Jblk) {
    NEED_BATCH("gxb_initial_residual_jacobian");
    gxb_opts o; if (opts) o = *opts; else gxb_default_opts(&o);
    int rc = initial_setup(c, o, ic, rate_vars1, rate_vars2, "gxb_initial_residual_jacobian"); if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    c->mode = 2;
    rc = gxb_residual_jacobian(c, opts, x, r, Jblk);
    c->mode = 0;
    return rc;
}

int gxb_initial_condition(gxb_ctx* c, const gxb_opts* opts, const double* ic, const double* x0, double* x, double* rates, int32_t* converged,
                          int32_t* iters, double* resid_inf, uint8_t* rate_vars) {
    NEED_BATCH("gxb_initial_condition");
    gxb_opts o; if (opts) o = *opts; else gxb_default_opts(&o);
    int rc = initial_setup(c, o, ic, nullptr, nullptr, "gxb_initial_condition"); if (rc) return rc;
    if ((rc = gxb_upload_x0(c, x0))) return rc;
    c->damping = o.structural_damping ? 1 : 0;
    c->have_initial = false;
    rc = solve_resident(c, &o, 2);
    c->damping = 0;
    if (rc) return rc;
    const size_t nb = (size_t)c->nbatch, P = (size_t)c->np;
    DynArgs body{};
    for (int i = 0; i < 6; i++) body.ibody[i] = c->ibody[i];
    k_initial_output<<<grid1(nb * P, 128), 128, 0, c->stream>>>(c->d_x, c->d_bcS, c->d_icS, c->d_rvm, c->d_flags, c->d_xyz, c->d_motion, c->d_rates0,
                                                                (int)P, (int)c->n, c->nbatch, body);
    CK(cudaGetLastError());
    c->stats.launches += 1;
    c->have_initial = true;
    if (rates) CK(cudaMemcpyAsync(rates, c->d_rates0, nb * P * 12 * 8, cudaMemcpyDeviceToHost, c->stream));
    if (rate_vars) {
        std::vector<unsigned> m(nb * P);
        CK(cudaMemcpyAsync(m.data(), c->d_rvm, nb * P * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (size_t t = 0; t < nb * P; t++)
            for (int k = 0; k < 6; k++) { rate_vars[12 * t + k] = (m[t] >> (16 + k)) & 1u; rate_vars[12 * t + 6 + k] = (m[t] >> k) & 1u; }
    }
    return gxb_fetch(c, x, converged, iters, resid_inf);
}

// dense Jacobian columns of the body-acceleration states as left by the last Jacobian assembly (parity tests)
int gxb_body_columns(gxb_ctx* c, int64_t* icol_body, double* D) {
    NEED_BATCH("gxb_body_columns");
    if (icol_body) for (int i = 0; i < 6; i++) icol_body[i] = c->ibody[i] + 1;      // 1-based like SystemIndices, 0 = none
    if (D) {
        if (!c->d_D) return fail(GXB_ERR_INVALID, "gxb_body_columns: no Jacobian with body-acceleration columns has been assembled");
        CK(cudaMemcpyAsync(D, c->d_D, (size_t)6 * c->nbatch * c->n * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return GXB_OK;
}

// ---- eigenvalue analysis: mass matrix and the Krylov part of solve_eigensystem (src/analyses.jl:943-965) ----
static int assemble_mass_matrix(gxb_ctx* c, const gxb_opts& o, const double* d_x) {
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n;
    if (!c->d_M) { CK(cudaMalloc(&c->d_M, nb * n * c->W * 8)); CK(cudaMemsetAsync(c->d_M, 0, nb * n * c->W * 8, c->stream)); }
    MassArgs A{};
    A.s = make_static_args(c, &o, d_x, nullptr, false, 0, false);
    A.Mrow = c->d_M;
    int bs = (int)((c->np + 31) / 32) * 32;
    k_mass_matrix<<<(unsigned)nb, bs, 0, c->stream>>>(A);
    CK(cudaGetLastError());
    c->stats.launches++;
    return GXB_OK;
}

int gxb_mass_matrix(gxb_ctx* c, const gxb_opts* opts, const double* x, double* Mblk) {
    NEED_BATCH("gxb_mass_matrix");
    if (c->kind != 1) return fail(GXB_ERR_INVALID, "gxb_mass_matrix: context topology is not kind=1 (DynamicSystem)");
    if (!x || !Mblk) return fail(GXB_ERR_INVALID, "gxb_mass_matrix: null");
    gxb_opts o; if (opts) o = *opts; else gxb_default_opts(&o);
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n;
    CK(cudaMemcpyAsync(c->d_xt, x, nb * n * 8, cudaMemcpyHostToDevice, c->stream));
    int rc = assemble_mass_matrix(c, o, c->d_xt); if (rc) return rc;
    const size_t nblk = c->brow.size();
    size_t bytes = nb * nblk * 9 * 8;
    if (bytes > c->stage_bytes) { cudaFree(c->d_stage); c->d_stage = nullptr; c->stage_bytes = 0; CK(cudaMalloc(&c->d_stage, bytes)); c->stage_bytes = bytes; }
    k_extract_blocks<<<grid1(nb * nblk * 9, 256), 256, 0, c->stream>>>(c->d_M, c->d_brow, c->d_bcol, c->d_stage, (int64_t)nblk, (int)n, c->W, c->KLB, c->nbatch);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(Mblk, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return GXB_OK;
}

int gxb_eigen_arnoldi(gxb_ctx* c, const gxb_opts* opts, const double* x, int32_t m, double* H, double* V) {
    NEED_BATCH("gxb_eigen_arnoldi");
    if (c->kind != 1) return fail(GXB_ERR_INVALID, "gxb_eigen_arnoldi: context topology is not kind=1 (DynamicSystem)");
    if (!x || !H || m < 1 || m > c->n) return fail(GXB_ERR_INVALID, "gxb_eigen_arnoldi: need x, H and 1 <= m <= nstates");
    if (c->nbody > 0) return fail(GXB_ERR_UNSUPPORTED, "gxb_eigen_arnoldi: body-frame acceleration states (dense Jacobian columns) are not supported by the eigen path");
    gxb_opts o; if (opts) o = *opts; else gxb_default_opts(&o);
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n;
    c->stats = gxb_stats{};
    c->damping = o.structural_damping ? 1 : 0;
    c->mode = 0;
    int rc;
    double *d_V = nullptr, *d_H = nullptr;
    cudaFree(c->d_Vk); c->d_Vk = nullptr; c->vk_m = 0;
    CK(cudaMalloc(&d_V, nb * (size_t)(m + 1) * n * 8));
    CK(cudaMalloc(&d_H, nb * (size_t)(m + 1) * m * 8));
    CK(cudaMemsetAsync(d_H, 0, nb * (size_t)(m + 1) * m * 8, c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    // K = steady Jacobian at x (steady_jacobian!, analyses.jl:1358), M = mass matrix at x (mass_matrix!, analyses.jl:1359)
    CK(cudaMemcpyAsync(c->d_x, x, nb * n * 8, cudaMemcpyHostToDevice, c->stream));
    StaticArgs A = make_static_args(c, &o, c->d_x, c->d_F, true, 0, false);
    if ((rc = launch_assemble(c, A, true))) { cudaFree(d_V); cudaFree(d_H); return rc; }
    if ((rc = assemble_mass_matrix(c, o, c->d_x))) { cudaFree(d_V); cudaFree(d_H); return rc; }
    k_arnoldi_start<<<(unsigned)nb, 256, 0, c->stream>>>(d_V, (int)n, m);
    for (int j = 0; j < m; j++) {
        // w = K^{-1} (M v_j)   (f! of analyses.jl:949)
        k_rowblock_spmv<<<grid1(nb * n, 256), 256, 0, c->stream>>>(c->d_M, d_V + (size_t)j * n, c->d_Ft, (int)n, c->W, c->KLB, c->nbatch, (int64_t)(m + 1) * n);
        if ((rc = launch_factor(c, &o, 1, 1.0))) { cudaFree(d_V); cudaFree(d_H); return rc; }
        k_arnoldi_orth<<<(unsigned)nb, 256, 0, c->stream>>>(d_V, d_H, c->d_p, (int)n, m, j);
        c->stats.launches += 2;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaMemcpyAsync(H, d_H, nb * (size_t)(m + 1) * m * 8, cudaMemcpyDeviceToHost, c->stream));
    if (V) CK(cudaMemcpyAsync(V, d_V, nb * (size_t)(m + 1) * n * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->stats.ms_total = ms;
    c->damping = 0;
    cudaFree(d_H);
    c->d_Vk = d_V; c->vk_m = m;          // stays resident: gxb_eigen_ritz_vectors forms V y on the device
    return GXB_OK;
}

// Ritz vectors of the last gxb_eigen_arnoldi: vec[b][:, j] = V_m[b] y[b][:, j]  (solve_eigensystem's `V = Vk * F.vectors`, analyses.jl:959).
// Y: nbatch x m x nev complex (interleaved re, im; the eigenvectors of the Hessenberg matrices chosen by the caller),
// vec: nbatch x nstates x nev complex.  Only nev vectors per instance cross PCIe instead of the m+1 basis vectors.
int gxb_eigen_ritz_vectors(gxb_ctx* c, int32_t nev, const double* Y, double* vec) {
    NEED_BATCH("gxb_eigen_ritz_vectors");
    if (!c->d_Vk) return fail(GXB_ERR_INVALID, "gxb_eigen_ritz_vectors: call gxb_eigen_arnoldi first");
    const int m = c->vk_m;
    if (!Y || !vec || nev < 1 || nev > m) return fail(GXB_ERR_INVALID, "gxb_eigen_ritz_vectors: need Y, vec and 1 <= nev <= m");
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n;
    double *d_Y = nullptr, *d_out = nullptr;
    CK(cudaMalloc(&d_Y, nb * m * nev * 16)); CK(cudaMalloc(&d_out, nb * n * nev * 16));
    CK(cudaMemcpyAsync(d_Y, Y, nb * m * nev * 16, cudaMemcpyHostToDevice, c->stream));
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)nb);
    k_ritz_vectors<<<grid, 128, 0, c->stream>>>(c->d_Vk, d_Y, d_out, (int)n, m, nev);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(vec, d_out, nb * n * nev * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_Y); cudaFree(d_out);
    c->stats.launches += 1;
    return GXB_OK;
}

int gxb_newton_direction(gxb_ctx* c, const gxb_opts* opts, const double* x, double* p) {
    NEED_BATCH("gxb_newton_direction");
    if (!x || !p) return fail(GXB_ERR_INVALID, "gxb_newton_direction: null");
    gxb_opts o; if (opts) o = *opts; else gxb_default_opts(&o);
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n;
    CK(cudaMemcpyAsync(c->d_xt, x, nb * n * 8, cudaMemcpyHostToDevice, c->stream));
    StaticArgs A = make_static_args(c, &o, c->d_xt, c->d_Ft, true, 0, false);
    int rc = launch_assemble(c, A, true); if (rc) return rc;
    if ((rc = launch_factor(c, &o, 1))) return rc;
    CK(cudaMemcpyAsync(p, c->d_p, nb * n * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return GXB_OK;
}

}  // extern "C"
