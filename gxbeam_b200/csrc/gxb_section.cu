// Batched cross-section analysis (BASELINE config 5): compliance_matrix and mass_matrix of src/section.jl for many layups
// that share one quad-mesh topology (materials, ply angles and node coordinates vary per instance).
//
//   reference (per layup, serial)                                      here (one thread per element, then one CTA per instance)
//   element_submatrix, 2 x 2 Gauss   src/section.jl:347-501            k_section_assemble: Q, the six element matrices, scatter
//   scatter into A, R, E, C, L, M    :661-677                          (FP64 atomics) into BAND storage (nodes renumbered by reverse
//   KKT system [E R D; R' A 0; D' 0 0], `factorize` (UMFPACK) :679-703   Cuthill-McKee on the host, once per topology)
//   two solves with 6 right-hand sides :699-732                        k_section_solve: partial L D L' of the bordered band matrix,
//   S = [X dX Y]' [E C R; C' M L; R' L' A] [X dX Y] :736-757            dense pivoted solve of the trailing 18 x 18 block, products, centres
//   shear / tension centre, transformation, reorder :759-781
//   mass_matrix :840-890                                               k_section_solve (element loop by the CTA)
//
// Why no pivoting is needed in the band part: E is symmetric positive SEMI-definite, its null space being the four rigid motions of the
// warping field (two in-plane translations, the in-plane rotation, the out-of-plane translation).  No such motion vanishes at two distinct
// nodes, so the leading block E11 that leaves out the LAST TWO nodes of the ordering is positive definite: L D L' without pivoting is stable
// on it.  The 12 border rows [R D]' are carried through that elimination as dense rows; what remains -- the 6 dof of the last two nodes
// and the 12 multipliers -- is an 18 x 18 symmetric indefinite system that one thread solves with partial pivoting.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <algorithm>
#include <queue>
#include <string>
#include <vector>

#include "../../include/gxbeam_b200.h"

int gxb_fail_(int code, const std::string& msg);      // gxb_lib.cu: sets the thread-local error string
#define SCK(call)                                                                                                                  \
    do {                                                                                                                           \
        cudaError_t e_ = (call);                                                                                                   \
        if (e_ != cudaSuccess) {                                                                                                   \
            cudaGetLastError();                                                                                                    \
            return gxb_fail_(GXB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (gxb_section.cu:" + std::to_string(__LINE__) + ")"); \
        }                                                                                                                          \
    } while (0)

#define SEC_ELEM 612          // doubles per element: Ae 36, Re 72, Ee 144, Ce 144, Le 72, Me 144 (row-major)
#define SEC_MTAIL 6           // dof left out of the unpivoted band elimination: the last two nodes of the ordering
#define SEC_NT (SEC_MTAIL + 12)

struct SecDims {
    int nn, ne, ndof, hb;     // hb: half bandwidth in dof of the renumbered mesh
    size_t per_instance;      // doubles of scratch per instance
    size_t oEb0, oEb, oCb, oMb, oR, oL, oBd, oA, oC22, oX1, oX2, oT1, oT2, oMass;
};

// ---- element matrices: section.jl:216-501 -------------------------------------------------------------------------------------
__device__ void sec_stiffness(const double* m, double (&Q)[6][6]) {
    const double E1 = m[0], E2 = m[1], E3 = m[2], G12 = m[3], G13 = m[4], G23 = m[5], nu12 = m[6], nu13 = m[7], nu23 = m[8];
    const double nu21 = nu12 * E2 / E1, nu31 = nu13 * E3 / E1, nu32 = nu23 * E3 / E2;
    const double delta = 1.0 / (1 - nu12 * nu21 - nu23 * nu32 - nu13 * nu31 - 2 * nu21 * nu32 * nu13);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Q[i][j] = 0.0;
    Q[5][5] = E1 * (1 - nu23 * nu32) * delta;
    Q[0][0] = E2 * (1 - nu13 * nu31) * delta;
    Q[1][1] = E3 * (1 - nu12 * nu21) * delta;
    Q[0][5] = Q[5][0] = E1 * (nu21 + nu31 * nu23) * delta;
    Q[0][1] = Q[1][0] = E2 * (nu32 + nu12 * nu31) * delta;
    Q[1][5] = Q[5][1] = E1 * (nu31 + nu21 * nu32) * delta;
    Q[3][3] = G12; Q[4][4] = G13; Q[2][2] = G23;
}
__device__ void sec_congruence(const double (&T)[6][6], double (&Q)[6][6]) {      // Q <- T Q T'
    double t[6][6];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double s = 0; for (int k = 0; k < 6; k++) s += T[i][k] * Q[k][j]; t[i][j] = s; }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double s = 0; for (int k = 0; k < 6; k++) s += t[i][k] * T[j][k]; Q[i][j] = s; }
}
__device__ void sec_Ttheta(double theta, double (&T)[6][6]) {      // get_Ttheta (:260-278)
    double s, c; sincos(theta, &s, &c);
    const double c2 = c * c, s2 = s * s, sc = s * c;
    const double t[6][6] = {{c2, 0, 0, 2 * sc, 0, s2}, {0, 1, 0, 0, 0, 0}, {0, 0, c, 0, s, 0}, {-sc, 0, 0, c2 - s2, 0, sc}, {0, 0, -s, 0, c, 0}, {s2, 0, 0, -2 * sc, 0, c2}};
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) T[i][j] = t[i][j];
}
__device__ void sec_Tbeta(const double (&xy)[4][2], double (&T)[6][6]) {      // element_orientation (:331-345), get_Tbeta (:293-309)
    const double dx = 0.5 * (xy[1][0] + xy[2][0]) - 0.5 * (xy[0][0] + xy[3][0]), dy = 0.5 * (xy[1][1] + xy[2][1]) - 0.5 * (xy[0][1] + xy[3][1]);
    const double ds = sqrt(dx * dx + dy * dy), c = dx / ds, s = dy / ds;
    const double c2 = c * c, s2 = s * s, sc = s * c;
    const double t[6][6] = {{c2, s2, -2 * sc, 0, 0, 0}, {s2, c2, 2 * sc, 0, 0, 0}, {sc, -sc, c2 - s2, 0, 0, 0}, {0, 0, 0, c, -s, 0}, {0, 0, 0, s, c, 0}, {0, 0, 0, 0, 0, 1}};
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) T[i][j] = t[i][j];
}
// element_submatrix: out[612] = Ae | Re | Ee | Ce | Le | Me, row-major.  xy: the four corner nodes (counterclockwise from bottom left)
__device__ void sec_element(const double* mat, double theta, const double (&xy)[4][2], double* out) {
    double Q[6][6], T[6][6];
    sec_stiffness(mat, Q);
    sec_Ttheta(theta, T); sec_congruence(T, Q);        // rotate_ply_to_element (:253-258)
    sec_Tbeta(xy, T); sec_congruence(T, Q);            // rotate_element_to_beam (:286-291)
    for (int i = 0; i < SEC_ELEM; i++) out[i] = 0.0;
    double* Ae = out; double* Re = out + 36; double* Ee = out + 108; double* Ce = out + 252; double* Le = out + 396; double* Me = out + 468;
    const double gp = 0.5773502691896258;      // 1 / sqrt(3)
    for (int q = 0; q < 4; q++) {
        const double ksi = (q < 2) ? gp : -gp, eta = (q % 2 == 0) ? gp : -gp;
        const double mk = 1 - ksi, pk = 1 + ksi, me = 1 - eta, pe = 1 + eta;
        const double N[4] = {mk * me / 4, pk * me / 4, pk * pe / 4, mk * pe / 4};
        const double Nk[4] = {-me / 4, me / 4, pe / 4, -pe / 4}, Ne[4] = {-mk / 4, -pk / 4, pk / 4, mk / 4};
        double x = 0, y = 0, xk = 0, xe = 0, yk = 0, ye = 0;
        for (int i = 0; i < 4; i++) { x += N[i] * xy[i][0]; y += N[i] * xy[i][1]; xk += Nk[i] * xy[i][0]; xe += Ne[i] * xy[i][0]; yk += Nk[i] * xy[i][1]; ye += Ne[i] * xy[i][1]; }
        const double det = xk * ye - yk * xe;
        const double j00 = ye / det, j01 = -yk / det, j10 = -xe / det, j11 = xk / det;
        double SZ[6][6], SN[6][12], BN[6][12];
        for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) SZ[i][j] = 0.0; for (int j = 0; j < 12; j++) { SN[i][j] = 0.0; BN[i][j] = 0.0; } }
        SZ[3][0] = 1; SZ[3][5] = -y; SZ[4][1] = 1; SZ[4][5] = x; SZ[5][2] = 1; SZ[5][3] = y; SZ[5][4] = -x;
        for (int i = 0; i < 4; i++) {
            SN[3][3 * i] = N[i]; SN[4][3 * i + 1] = N[i]; SN[5][3 * i + 2] = N[i];
            const double gx = j00 * Nk[i] + j01 * Ne[i], gy = j10 * Nk[i] + j11 * Ne[i];
            BN[0][3 * i] = gx; BN[1][3 * i + 1] = gy; BN[2][3 * i] = gy; BN[2][3 * i + 1] = gx; BN[3][3 * i + 2] = gx; BN[4][3 * i + 2] = gy;
        }
        double QSZ[6][6], QBN[6][12], QSN[6][12];
        for (int i = 0; i < 6; i++) {
            for (int j = 0; j < 6; j++) { double s = 0; for (int k = 0; k < 6; k++) s += Q[i][k] * SZ[k][j]; QSZ[i][j] = s * det; }
            for (int j = 0; j < 12; j++) { double s1 = 0, s2 = 0; for (int k = 0; k < 6; k++) { s1 += Q[i][k] * BN[k][j]; s2 += Q[i][k] * SN[k][j]; } QBN[i][j] = s1 * det; QSN[i][j] = s2 * det; }
        }
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double s = 0; for (int k = 0; k < 6; k++) s += SZ[k][i] * QSZ[k][j]; Ae[6 * i + j] += s; }
        for (int i = 0; i < 12; i++) {
            for (int j = 0; j < 6; j++) { double s1 = 0, s2 = 0; for (int k = 0; k < 6; k++) { s1 += BN[k][i] * QSZ[k][j]; s2 += SN[k][i] * QSZ[k][j]; } Re[6 * i + j] += s1; Le[6 * i + j] += s2; }
            for (int j = 0; j < 12; j++) {
                double s1 = 0, s2 = 0, s3 = 0;
                for (int k = 0; k < 6; k++) { s1 += BN[k][i] * QBN[k][j]; s2 += BN[k][i] * QSN[k][j]; s3 += SN[k][i] * QSN[k][j]; }
                Ee[12 * i + j] += s1; Ce[12 * i + j] += s2; Me[12 * i + j] += s3;
            }
        }
    }
}

struct SecArgs {
    SecDims d; int64_t nbatch;
    const int* conn;        // [ne][4] 0-based original node numbers
    const int* perm;        // [nn]: new number of original node i
    const double* nodes;    // [b][nn][2]
    const double* mats;     // [b][ne][10]
    const double* theta;    // [b][ne]
    double* scratch;        // [b][per_instance]
    double* elem_out;       // debug: [b][ne][612] or null
    int gxbeam_order, shear_center, pair_table, smem_window;
    double* S; double* sc; double* tc; double* M; double* mc; int* status;
};

__global__ void __launch_bounds__(64) k_section_assemble(SecArgs A) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= A.nbatch * A.d.ne) return;
    const int64_t b = t / A.d.ne; const int e = (int)(t % A.d.ne);
    const int* cn = A.conn + 4 * e;
    double xy[4][2];
    for (int i = 0; i < 4; i++) { xy[i][0] = A.nodes[((size_t)b * A.d.nn + cn[i]) * 2]; xy[i][1] = A.nodes[((size_t)b * A.d.nn + cn[i]) * 2 + 1]; }
    double em[SEC_ELEM];
    sec_element(A.mats + ((size_t)b * A.d.ne + e) * 10, A.theta[(size_t)b * A.d.ne + e], xy, em);
    if (A.elem_out) { double* o = A.elem_out + ((size_t)b * A.d.ne + e) * SEC_ELEM; for (int i = 0; i < SEC_ELEM; i++) o[i] = em[i]; return; }
    double* s = A.scratch + (size_t)b * A.d.per_instance;
    const int hb = A.d.hb;
    int g[12];
    for (int i = 0; i < 4; i++) for (int d = 0; d < 3; d++) g[3 * i + d] = 3 * A.perm[cn[i]] + d;
    const double* Ae = em; const double* Re = em + 36; const double* Ee = em + 108; const double* Ce = em + 252; const double* Le = em + 396; const double* Me = em + 468;
    for (int i = 0; i < 36; i++) atomicAdd(s + A.d.oA + i, Ae[i]);
    for (int i = 0; i < 12; i++) {
        for (int j = 0; j < 6; j++) { atomicAdd(s + A.d.oR + (size_t)g[i] * 6 + j, Re[6 * i + j]); atomicAdd(s + A.d.oL + (size_t)g[i] * 6 + j, Le[6 * i + j]); }
        for (int j = 0; j < 12; j++) {
            const int r = g[i], c = g[j];
            atomicAdd(s + A.d.oCb + (size_t)r * (2 * hb + 1) + (c - r + hb), Ce[12 * i + j]);
            if (r >= c) {           // symmetric matrices: lower band, column-oriented: X[c][r - c]
                atomicAdd(s + A.d.oEb0 + (size_t)c * (hb + 1) + (r - c), Ee[12 * i + j]);
                atomicAdd(s + A.d.oMb + (size_t)c * (hb + 1) + (r - c), Me[12 * i + j]);
            }
        }
    }
}

// y[ndof][6] (+)= op(B) x[ndof][6] for a band matrix; sym: lower band, column-oriented (hb + 1 per column); else full band, row-oriented
// (2 hb + 1 per row), transposed optionally.  One thread per (row, column-of-x) pair; sign: y = sgn * B x (+ y if accumulate).
// strain_recovery (section.jl:952-1040) + tsai_wu (:1108-1131): one thread per (layup, element), after k_section_solve has left X, Y (oX1)
// and dX (oX2) of the layup in its scratch (the reference's SectionCache).  Outputs [b][ne][6] = the memory layout of the reference's 6 x ne
// Julia matrices; any may be null.
struct SecRecoverArgs {
    const double* F; const double* Mo;      // [b][3] each
    const double* strengths;                // [b][ne][9] (S1t, S1c, S2t, S2c, S3t, S3c, S12, S13, S23: Material fields 11-19, :29-37) or null
    int gxbeam_order;
    double *strain_b, *stress_b, *strain_p, *stress_p, *failure;
};
__global__ void __launch_bounds__(64) k_section_recover(SecArgs A, SecRecoverArgs Rv) {
    const SecDims& D = A.d;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.nbatch * D.ne) return;
    const int64_t b = t / D.ne; const int e = (int)(t % D.ne);
    const int* cn = A.conn + 4 * e;
    double xy[4][2];
    for (int i = 0; i < 4; i++) { xy[i][0] = A.nodes[((size_t)b * D.nn + cn[i]) * 2]; xy[i][1] = A.nodes[((size_t)b * D.nn + cn[i]) * 2 + 1]; }
    const double* mat = A.mats + ((size_t)b * D.ne + e) * 10;
    const double theta = A.theta[(size_t)b * D.ne + e];
    // loads in the section's axes (z along the beam): GXBeam's x, y, z = the section's z, x, y
    double ld[6];
    {
        const double* F = Rv.F + 3 * b; const double* Mo = Rv.Mo + 3 * b;
        if (Rv.gxbeam_order) { ld[0] = F[1]; ld[1] = F[2]; ld[2] = F[0]; ld[3] = Mo[1]; ld[4] = Mo[2]; ld[5] = Mo[0]; }
        else { for (int i = 0; i < 3; i++) { ld[i] = F[i]; ld[3 + i] = Mo[i]; } }
    }
    const double* s = A.scratch + (size_t)b * D.per_instance;
    const double* X = s + D.oX1; const double* Y = X + (size_t)D.ndof * 6; const double* dX = s + D.oX2;
    double yl[6], xe[12], dxe[12];
    for (int i = 0; i < 6; i++) { double a = 0; for (int k = 0; k < 6; k++) a = fma(Y[6 * i + k], ld[k], a); yl[i] = a; }
    for (int i = 0; i < 4; i++) for (int d = 0; d < 3; d++) {
        const size_t row = (size_t)(3 * A.perm[cn[i]] + d) * 6;
        double a = 0, c = 0;
        for (int k = 0; k < 6; k++) { a = fma(X[row + k], ld[k], a); c = fma(dX[row + k], ld[k], c); }
        xe[3 * i + d] = a; dxe[3 * i + d] = c;
    }
    // element_matrices(0, 0, ...) (:347-462): shape functions 1/4 each at the element centre
    const double Nk[4] = {-0.25, 0.25, 0.25, -0.25}, Ne[4] = {-0.25, -0.25, 0.25, 0.25};
    double x = 0, y = 0, xk = 0, xet = 0, yk = 0, yet = 0;
    for (int i = 0; i < 4; i++) { x += 0.25 * xy[i][0]; y += 0.25 * xy[i][1]; xk += Nk[i] * xy[i][0]; xet += Ne[i] * xy[i][0]; yk += Nk[i] * xy[i][1]; yet += Ne[i] * xy[i][1]; }
    const double det = xk * yet - yk * xet;
    const double j00 = yet / det, j01 = -yk / det, j10 = -xet / det, j11 = xk / det;
    double eps[6] = {0, 0, 0, yl[0] - y * yl[5], yl[1] + x * yl[5], yl[2] + y * yl[3] - x * yl[4]};      // SZ Y theta
    for (int i = 0; i < 4; i++) {
        const double gx = j00 * Nk[i] + j01 * Ne[i], gy = j10 * Nk[i] + j11 * Ne[i];
        eps[0] = fma(gx, xe[3 * i], eps[0]); eps[1] = fma(gy, xe[3 * i + 1], eps[1]);
        eps[2] = fma(gy, xe[3 * i], fma(gx, xe[3 * i + 1], eps[2]));
        eps[3] = fma(gx, xe[3 * i + 2], eps[3]); eps[4] = fma(gy, xe[3 * i + 2], eps[4]);                  // BN Xe theta
        eps[3] = fma(0.25, dxe[3 * i], eps[3]); eps[4] = fma(0.25, dxe[3 * i + 1], eps[4]); eps[5] = fma(0.25, dxe[3 * i + 2], eps[5]);   // SN dXe theta
    }
    double Q[6][6], Qm[6][6], Tth[6][6], Tbe[6][6];
    sec_stiffness(mat, Qm);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Q[i][j] = Qm[i][j];
    sec_Ttheta(theta, Tth); sec_congruence(Tth, Q);
    sec_Tbeta(xy, Tbe); sec_congruence(Tbe, Q);
    double sig[6], tmp[6], epl[6], spl[6];
    for (int i = 0; i < 6; i++) { double a = 0; for (int k = 0; k < 6; k++) a = fma(Q[i][k], eps[k], a); sig[i] = a; }
    for (int i = 0; i < 6; i++) { double a = 0; for (int k = 0; k < 6; k++) a = fma(Tbe[k][i], eps[k], a); tmp[i] = a; }
    for (int i = 0; i < 6; i++) { double a = 0; for (int k = 0; k < 6; k++) a = fma(Tth[k][i], tmp[k], a); epl[i] = a; }
    for (int i = 0; i < 6; i++) { double a = 0; for (int k = 0; k < 6; k++) a = fma(Qm[i][k], epl[k], a); spl[i] = a; }
    // conventional orders (:1001-1008, :1033-1036): beam xx yy zz xy xz yz (GXBeam axes if asked), ply 11 22 33 12 13 23
    const int ib_g[6] = {5, 0, 1, 4, 2, 3}, ib_s[6] = {0, 1, 5, 2, 3, 4}, ip[6] = {5, 0, 1, 3, 4, 2};
    const int* ib = Rv.gxbeam_order ? ib_g : ib_s;
    const size_t o = (size_t)t * 6;
    for (int i = 0; i < 6; i++) {
        if (Rv.strain_b) Rv.strain_b[o + i] = eps[ib[i]];
        if (Rv.stress_b) Rv.stress_b[o + i] = sig[ib[i]];
        if (Rv.strain_p) Rv.strain_p[o + i] = epl[ip[i]];
        if (Rv.stress_p) Rv.stress_p[o + i] = spl[ip[i]];
    }
    if (Rv.failure && Rv.strengths) {
        const double* m = Rv.strengths + (size_t)t * 9;
        const double S1t = m[0], S1c = m[1], S2t = m[2], S2c = m[3], S3t = m[4], S3c = m[5], S12 = m[6], S13 = m[7], S23 = m[8];
        const double s1 = spl[ip[0]], s2 = spl[ip[1]], s3 = spl[ip[2]], s4 = spl[ip[3]], s5 = spl[ip[4]], s6 = spl[ip[5]];
        Rv.failure[t] = s1 * s1 / (S1t * S1c) + s2 * s2 / (S2t * S2c) + s3 * s3 / (S3t * S3c) + s4 * s4 / (S12 * S12) + s5 * s5 / (S13 * S13)
                        + s6 * s6 / (S23 * S23) + s1 * (1 / S1t - 1 / S1c) + s2 * (1 / S2t - 1 / S2c) + s3 * (1 / S3t - 1 / S3c)
                        - s1 * s2 / sqrt(S1t * S1c * S2t * S2c) - s1 * s3 / sqrt(S1t * S1c * S3t * S3c) - s2 * s3 / sqrt(S2t * S2c * S3t * S3c);
    }
}

__device__ void sec_band_mv(const double* Bm, bool sym, bool trans, const double* x, double* y, int ndof, int hb, double sgn, bool accumulate) {
    // one thread per row, all six right-hand sides at once: a band entry is read once for six FMAs (x rows are 48 contiguous bytes)
    for (int r = threadIdx.x; r < ndof; r += blockDim.x) {
        double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const int c0 = max(0, r - hb), c1 = min(ndof - 1, r + hb);
        for (int c = c0; c <= c1; c++) {
            double v;
            if (sym) v = (r >= c) ? Bm[(size_t)c * (hb + 1) + (r - c)] : Bm[(size_t)r * (hb + 1) + (c - r)];
            else v = trans ? Bm[(size_t)c * (2 * hb + 1) + (r - c + hb)] : Bm[(size_t)r * (2 * hb + 1) + (c - r + hb)];
            const double2* xc = reinterpret_cast<const double2*>(x + (size_t)c * 6);
            const double2 x0 = xc[0], x1 = xc[1], x2 = xc[2];
            acc[0] = fma(v, x0.x, acc[0]); acc[1] = fma(v, x0.y, acc[1]); acc[2] = fma(v, x1.x, acc[2]);
            acc[3] = fma(v, x1.y, acc[3]); acc[4] = fma(v, x2.x, acc[4]); acc[5] = fma(v, x2.y, acc[5]);
        }
        double* yr = y + (size_t)r * 6;
#pragma unroll
        for (int k = 0; k < 6; k++) yr[k] = accumulate ? yr[k] + sgn * acc[k] : sgn * acc[k];
    }
}
// G[6][6] (+)= P' Q for P, Q of ndof x 6 (row-major): CTA-wide reduction, result in shared red36
__device__ void sec_gram(const double* Pm, const double* Qm, int n, double* G, double* red /* [blockDim/32][36] */) {
    double a[36];
    for (int i = 0; i < 36; i++) a[i] = 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x)
        for (int i = 0; i < 6; i++) { const double p = Pm[(size_t)r * 6 + i]; for (int j = 0; j < 6; j++) a[6 * i + j] = fma(p, Qm[(size_t)r * 6 + j], a[6 * i + j]); }
    for (int i = 0; i < 36; i++) for (int o = 16; o; o >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) for (int i = 0; i < 36; i++) red[(threadIdx.x >> 5) * 36 + i] = a[i];
    __syncthreads();
    if (threadIdx.x < 36) { double t = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w * 36 + threadIdx.x]; G[threadIdx.x] += t; }
    __syncthreads();
}

// Solve K z = rhs for the 6 right-hand sides in z[(ndof + 12)][6] (in place) with the partial L D L' left in Eb / Bd and the dense LU of
// the trailing block in Ts (with its pivots).  n1 = ndof - SEC_MTAIL.
__device__ void sec_kkt_solve(const double* Eb, const double* Bd, const double* Ts, const int* piv, double* z, int ndof, int hb) {
    const int n1 = ndof - SEC_MTAIL;
    // forward: unit lower triangular (band + 12 border rows), column by column
    for (int k = 0; k < n1; k++) {
        const int mk = min(hb, ndof - 1 - k);
        for (int t = threadIdx.x; t < (mk + 12) * 6; t += blockDim.x) {
            const int q = t / 6, c = t % 6;
            const double zk = z[(size_t)k * 6 + c];
            if (q < mk) z[(size_t)(k + 1 + q) * 6 + c] -= Eb[(size_t)k * (hb + 1) + 1 + q] * zk;
            else z[(size_t)(ndof + q - mk) * 6 + c] -= Bd[(size_t)k * 12 + (q - mk)] * zk;
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < n1 * 6; t += blockDim.x) z[t] /= Eb[(size_t)(t / 6) * (hb + 1)];
    __syncthreads();
    // trailing 18 x 18 block (dof of the last two nodes + 12 multipliers): dense LU with the recorded row interchanges, one thread per column
    if (threadIdx.x < 6) {
        const int c = threadIdx.x;
        double v[SEC_NT];
        for (int i = 0; i < SEC_NT; i++) v[i] = z[(size_t)(n1 + i) * 6 + c];
        // LAPACK getrs order: all row interchanges first (the factorisation swapped whole rows, multipliers included), then L, then U
        for (int k = 0; k < SEC_NT; k++) { const int p = piv[k]; const double tmp = v[k]; v[k] = v[p]; v[p] = tmp; }
        for (int k = 0; k < SEC_NT; k++)
            for (int i = k + 1; i < SEC_NT; i++) v[i] -= Ts[i * SEC_NT + k] * v[k];
        for (int k = SEC_NT - 1; k >= 0; k--) { double s = v[k]; for (int j = k + 1; j < SEC_NT; j++) s -= Ts[k * SEC_NT + j] * v[j]; v[k] = s / Ts[k * SEC_NT + k]; }
        for (int i = 0; i < SEC_NT; i++) z[(size_t)(n1 + i) * 6 + c] = v[i];
    }
    __syncthreads();
    // backward: z_k -= sum_d L[k+d][k] z_{k+d} + sum_j Lb[j][k] z_{ndof+j}
    for (int k = n1 - 1; k >= 0; k--) {
        const int mk = min(hb, ndof - 1 - k);
        // 6 columns x up to 32 partial sums each: threads (c, lane) with lane striding over the mk + 12 terms
        const int c = threadIdx.x / 32, lane = threadIdx.x % 32;
        double acc = 0.0;
        if (c < 6) {
            for (int q = lane; q < mk + 12; q += 32) {
                if (q < mk) acc = fma(Eb[(size_t)k * (hb + 1) + 1 + q], z[(size_t)(k + 1 + q) * 6 + c], acc);
                else acc = fma(Bd[(size_t)k * 12 + (q - mk)], z[(size_t)(ndof + q - mk) * 6 + c], acc);
            }
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (c < 6 && lane == 0) z[(size_t)k * 6 + c] -= acc;
        __syncthreads();
    }
}

// The same solve with the active part of the right-hand sides in shared memory: forward, rows k .. k + hb and the 12 border rows; backward,
// the solved rows k + 1 .. k + hb; the columns of L are prefetched one step ahead.  sm: hw * 6 + 72 + 2 (hw + 12) doubles.
__device__ void sec_kkt_solve_win(const double* Eb, const double* Bd, const double* Ts, const int* piv, double* z, int ndof, int hb, double* sm) {
    // columns of L (hw band entries + 12 border entries) come through a ring of SEC_SR buffers filled SEC_SP columns ahead by cp.async: the
    // factor does not fit the L2 next to the other resident instances, and a column's own work is shorter than a DRAM round trip
    constexpr int SEC_SR = 16, SEC_SP = 6;         // two columns per step: SEC_SP + 2 columns in flight, none of them in a slot still being read
    const int hw = hb + 1, n1 = ndof - SEC_MTAIL, tid = threadIdx.x, lw = hw + 12;
    double* zw = sm; double* zb = zw + hw * 6; double* Lc = zb + 72;            // Lc: [SEC_SR][lw]
    auto stage = [&](int col) {
        if (col >= 0 && col < n1 && tid < lw) {
            const double* src = tid < hw ? Eb + (size_t)col * hw + tid : Bd + (size_t)col * 12 + (tid - hw);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(Lc + (col % SEC_SR) * lw + tid);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src));
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    for (int t = tid; t < hw * 6; t += blockDim.x) { const int r = t / 6; zw[t] = r < ndof ? z[(size_t)r * 6 + t % 6] : 0.0; }
    for (int t = tid; t < 72; t += blockDim.x) zb[t] = z[(size_t)ndof * 6 + t];
#pragma unroll
    for (int q = 0; q < SEC_SP; q++) stage(q);
    // forward substitution, TWO columns per barrier pair: z_k is final, z_{k+1} needs L[k+1,k] z_k only (every thread forms both locally),
    // then each pending row takes both columns' contributions in one read-modify-write
    const int sa = tid / 6, ca = tid - 6 * sa;          // tid < 12: (which of the two columns, right-hand side) in the hand-over
    double zin = (tid < 12 && sa + hw < ndof) ? z[(size_t)(sa + hw) * 6 + ca] : 0.0;      // rows hw, hw + 1 take over the slots of columns 0, 1
    int ks = 0;
    int k = 0;
    for (; k + 1 < n1; k += 2, ks = (ks + 2 >= hw ? ks + 2 - hw : ks + 2)) {
        stage(k + SEC_SP); stage(k + 1 + SEC_SP);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(SEC_SP) : "memory");      // this thread's pieces of columns k, k + 1 have landed
        __syncthreads();
        const double* c0 = Lc + (k % SEC_SR) * lw;
        const double* c1 = Lc + ((k + 1) % SEC_SR) * lw;
        const int ks1 = ks + 1 == hw ? 0 : ks + 1;
        const double zin_next = (tid < 12 && k + 2 + sa + hw < ndof) ? z[(size_t)(k + 2 + sa + hw) * 6 + ca] : 0.0;
        const double rdiag = tid < 12 ? 1.0 / (sa == 0 ? c0[0] : c1[0]) : 0.0;     // issued early: the division's latency hides under the update loop
        const int mk0 = min(hb, ndof - 1 - k), mk1 = min(hb, ndof - 2 - k);
        const int nwin = min(mk1, hb - 1);               // pending rows k + 2 .. k + 1 + nwin are in the window; row k + hw (reached by column
        const double l10 = c0[1];                        // k + 1 only) enters it at the hand-over below
        for (int t = tid; t < (nwin + 12) * 6; t += blockDim.x) {
            const int q = t / 6, c = t - 6 * q;
            const double z0 = zw[ks * 6 + c];
            const double z1 = zw[ks1 * 6 + c] - l10 * z0;
            if (q < nwin) {
                int sl = ks + 2 + q; if (sl >= hw) sl -= hw;
                const double a0 = q + 2 <= mk0 ? c0[q + 2] : 0.0;             // L[k + 2 + q, k]
                zw[sl * 6 + c] -= a0 * z0 + c1[q + 1] * z1;                    // L[k + 2 + q, k + 1]
            } else {
                const int j = q - nwin;
                zb[j * 6 + c] -= c0[hw + j] * z0 + c1[hw + j] * z1;
            }
        }
        __syncthreads();
        if (tid < 32) {
            double newv = 0.0;
            if (tid < 12) {
                const double z0 = zw[ks * 6 + ca];
                const double z1 = zw[ks1 * 6 + ca] - l10 * z0;
                z[(size_t)(k + sa) * 6 + ca] = (sa == 0 ? z0 : z1) * rdiag;
                // slot of row k -> row k + hw, which column k + 1 reaches with L[k + hw, k + 1] = c1[hb]; slot of row k + 1 -> row k + 1 + hw
                newv = (sa == 0 && mk1 == hb) ? zin - c1[hb] * z1 : zin;
            }
            __syncwarp();                                 // both final values are read before either slot is handed over
            if (tid < 12) zw[(sa == 0 ? ks : ks1) * 6 + ca] = newv;
        }
        zin = zin_next;
    }
    if (k < n1) {                                          // odd number of columns: the last one alone
        stage(k + SEC_SP);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(SEC_SP - 1) : "memory");
        __syncthreads();
        const double* cur = Lc + (k % SEC_SR) * lw;
        const int mk = min(hb, ndof - 1 - k);
        for (int t = tid; t < (mk + 12) * 6; t += blockDim.x) {
            const int q = t / 6, c = t - 6 * q;
            const double zk = zw[ks * 6 + c];
            if (q < mk) { int sl = ks + 1 + q; if (sl >= hw) sl -= hw; zw[sl * 6 + c] -= cur[1 + q] * zk; }
            else zb[(q - mk) * 6 + c] -= cur[hw + (q - mk)] * zk;
        }
        __syncthreads();
        if (tid < 6) {
            z[(size_t)k * 6 + tid] = zw[ks * 6 + tid] / cur[0];
            zw[ks * 6 + tid] = (k + hw < ndof) ? z[(size_t)(k + hw) * 6 + tid] : 0.0;
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    for (int t = tid; t < SEC_MTAIL * 6; t += blockDim.x) { const int r = n1 + t / 6; z[(size_t)r * 6 + t % 6] = zw[(r % hw) * 6 + t % 6]; }
    for (int t = tid; t < 72; t += blockDim.x) z[(size_t)ndof * 6 + t] = zb[t];
    __syncthreads();
    if (tid < 6) {
        const int c = tid;
        double v[SEC_NT];
        for (int i = 0; i < SEC_NT; i++) v[i] = z[(size_t)(n1 + i) * 6 + c];
        for (int k = 0; k < SEC_NT; k++) { const int p = piv[k]; const double tmp = v[k]; v[k] = v[p]; v[p] = tmp; }
        for (int k = 0; k < SEC_NT; k++)
            for (int i = k + 1; i < SEC_NT; i++) v[i] -= Ts[i * SEC_NT + k] * v[k];
        for (int k = SEC_NT - 1; k >= 0; k--) { double a = v[k]; for (int j = k + 1; j < SEC_NT; j++) a -= Ts[k * SEC_NT + j] * v[j]; v[k] = a / Ts[k * SEC_NT + k]; }
        for (int i = 0; i < SEC_NT; i++) z[(size_t)(n1 + i) * 6 + c] = v[i];
    }
    __syncthreads();
    // backward: window = solved rows k + 1 .. k + hb (rows >= ndof do not exist: zero), border = the 12 solved multipliers
    for (int t = tid; t < hw * 6; t += blockDim.x) { const int r = n1 + t / 6; zw[(r % hw) * 6 + t % 6] = r < ndof ? z[(size_t)r * 6 + t % 6] : 0.0; }
    for (int t = tid; t < 72; t += blockDim.x) zb[t] = z[(size_t)ndof * 6 + t];
#pragma unroll
    for (int q = 0; q < SEC_SP; q++) stage(n1 - 1 - q);
    // backward substitution, TWO rows per barrier pair: the dot products of rows k and k - 1 with the solved rows k + 1 .. run side by side,
    // row k - 1 then takes L[k, k-1] x_k
    const int c = tid / 32, lane = tid % 32;
    const bool owner = c < 6 && lane == 0;
    k = n1 - 1;
    double ya = (owner && k >= 0) ? z[(size_t)k * 6 + c] : 0.0;                   // forward-solved values of rows k, k - 1
    double yb = (owner && k >= 1) ? z[(size_t)(k - 1) * 6 + c] : 0.0;
    for (; k >= 1; k -= 2) {
        stage(k - SEC_SP); stage(k - 1 - SEC_SP);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(SEC_SP) : "memory");
        __syncthreads();                     // columns k, k - 1 staged; the rows written by the previous step are visible
        const double* c0 = Lc + (k % SEC_SR) * lw;
        const double* c1 = Lc + ((k - 1) % SEC_SR) * lw;
        const double l10 = c1[1];                                                  // L[k, k - 1]
        const double y0 = ya, y1 = yb;
        ya = (owner && k >= 2) ? z[(size_t)(k - 2) * 6 + c] : 0.0;
        yb = (owner && k >= 3) ? z[(size_t)(k - 3) * 6 + c] : 0.0;
        const int mk0 = min(hb, ndof - 1 - k), mk1 = min(hb, ndof - k);
        double acc0 = 0.0, acc1 = 0.0;
        if (c < 6) {
            const int k1 = (k + 1) % hw;
            for (int q = lane; q < mk0 + 12; q += 32) {
                if (q < mk0) {
                    int sl = k1 + q; if (sl >= hw) sl -= hw;
                    const double xv = zw[sl * 6 + c];                              // x of row k + 1 + q
                    acc0 = fma(c0[1 + q], xv, acc0);
                    if (q + 2 <= mk1) acc1 = fma(c1[2 + q], xv, acc1);             // L[k + 1 + q, k - 1]
                } else {
                    const double bv = zb[(q - mk0) * 6 + c];
                    acc0 = fma(c0[hw + (q - mk0)], bv, acc0);
                    acc1 = fma(c1[hw + (q - mk0)], bv, acc1);
                }
            }
        }
        for (int o = 16; o; o >>= 1) { acc0 += __shfl_xor_sync(0xffffffffu, acc0, o); acc1 += __shfl_xor_sync(0xffffffffu, acc1, o); }
        __syncthreads();                     // every read of the slots that rows k, k - 1 take over (rows k + hw, k - 1 + hw) is done
        if (owner) {
            const double x0 = y0 - acc0;
            const double x1 = y1 - acc1 - l10 * x0;
            z[(size_t)k * 6 + c] = x0; z[(size_t)(k - 1) * 6 + c] = x1;
            zw[(k % hw) * 6 + c] = x0; zw[((k - 1) % hw) * 6 + c] = x1;
        }
    }
    if (k == 0) {                            // odd number of rows: row 0 alone
        stage(-1);
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        const double* cur = Lc;
        const int mk = min(hb, ndof - 1);
        double acc = 0.0;
        if (c < 6) {
            const int k1 = 1 % hw;
            for (int q = lane; q < mk + 12; q += 32) {
                if (q < mk) { int sl = k1 + q; if (sl >= hw) sl -= hw; acc = fma(cur[1 + q], zw[sl * 6 + c], acc); }
                else acc = fma(cur[hw + (q - mk)], zb[(q - mk) * 6 + c], acc);
            }
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (owner) z[(size_t)0 * 6 + c] = ya - acc;
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
}

__global__ void __launch_bounds__(256) k_section_solve(SecArgs A) {
    const int64_t b = blockIdx.x;
    const SecDims& D = A.d;
    const int ndof = D.ndof, hb = D.hb, nn = D.nn, n1 = ndof - SEC_MTAIL;
    double* s = A.scratch + (size_t)b * D.per_instance;
    double *Eb0 = s + D.oEb0, *Eb = s + D.oEb, *Cb = s + D.oCb, *Mb = s + D.oMb, *R = s + D.oR, *L = s + D.oL, *Bd = s + D.oBd, *Am = s + D.oA, *C22 = s + D.oC22;
    double *X1 = s + D.oX1, *X2 = s + D.oX2, *T1 = s + D.oT1, *T2 = s + D.oT2;
    const double* xy = A.nodes + (size_t)b * nn * 2;
    __shared__ double Ts[SEC_NT * SEC_NT]; __shared__ int piv[SEC_NT]; __shared__ double red[8 * 36]; __shared__ double G[36]; __shared__ int bad;
    __shared__ unsigned char c22_i[78], c22_j[78];            // (i >= j) of the u-th entry of the 12 x 12 corner's lower triangle
    if (threadIdx.x < 78) {
        const int u = threadIdx.x;
        int i = 0; while ((i + 1) * (i + 2) / 2 <= u) i++;
        c22_i[u] = (unsigned char)i; c22_j[u] = (unsigned char)(u - i * (i + 1) / 2);
    }
    // (row offset, column offset) of the t-th pair 1 <= dj <= dI <= hb of the rank-1 update, enumerated row by row (so that the first
    // mk (mk + 1) / 2 entries are the pairs of a shorter column): a table in dynamic shared memory when it fits (A.pair_table), else decoded
    extern __shared__ __align__(16) double sec_dyn[];
    const int hw = hb + 1;
    double* win = sec_dyn;                                     // [hw][hw]: slot c % hw holds column c of the band (window mode)
    double* bwin = win + (A.smem_window ? hw * hw : 0);        // [hw][12] border entries of the same columns
    double* c22s = bwin + (A.smem_window ? hw * 12 : 0);       // [144] corner
    double* stg = c22s + (A.smem_window ? 144 : 0);            // [8][hw + 12] staged incoming columns
    unsigned short* pair_tab = reinterpret_cast<unsigned short*>(stg + (A.smem_window ? 8 * (hw + 12) : 0));
    if (A.pair_table)
        for (int t = threadIdx.x; t < hb * (hb + 1) / 2; t += blockDim.x) {
            int di = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
            while ((di + 1) * (di + 2) / 2 <= t) di++;
            while (di * (di + 1) / 2 > t) di--;
            pair_tab[2 * t] = (unsigned short)(di + 1); pair_tab[2 * t + 1] = (unsigned short)(t - di * (di + 1) / 2 + 1);
        }
#ifdef SEC_TIMING
    long long tq = clock64(), tn, tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define SEC_TICK(i) do { __syncthreads(); tn = clock64(); tph[i] += tn - tq; tq = tn; } while (0)
#else
#define SEC_TICK(i) do { } while (0)
#endif
    // ---- border [R | D] (constraint columns of section.jl:683-692, in the renumbered dof), corner [A 0; 0 0], working copy of E ----
    if (!A.smem_window) for (int t = threadIdx.x; t < ndof * (hb + 1); t += blockDim.x) Eb[t] = Eb0[t];
    for (int i = threadIdx.x; i < nn; i += blockDim.x) {
        const int p = A.perm[i];
        const double x = xy[2 * i], y = xy[2 * i + 1];
        for (int d = 0; d < 3; d++) {
            double* row = Bd + (size_t)(3 * p + d) * 12;
            for (int j = 0; j < 6; j++) row[j] = R[(size_t)(3 * p + d) * 6 + j];
            row[6] = d == 0; row[7] = d == 1; row[8] = d == 2;
            row[9] = d == 2 ? y : 0.0; row[10] = d == 2 ? -x : 0.0; row[11] = d == 0 ? -y : (d == 1 ? x : 0.0);
        }
    }
    for (int t = threadIdx.x; t < 144; t += blockDim.x) { const int i = t / 12, j = t % 12; C22[t] = (i < 6 && j < 6) ? Am[6 * i + j] : 0.0; }
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    SEC_TICK(0);
    // ---- partial L D L' of the bordered band matrix: the first n1 columns (E11 is positive definite) ----
    // Window mode: the hb + 1 columns the rank-1 update of column k can touch live in shared memory (ring of column slots); a column is
    // read from global memory once, when the window reaches it, and written back once, when it has become a column of L.  (Updating the
    // band in global memory costs a DRAM round trip per read-modify-write: 24 us per column instead of ~0.3 us.)
    if (A.smem_window) {
        for (int t = threadIdx.x; t < hw * hw; t += blockDim.x) { const int c = t / hw, d = t % hw; win[t] = c < ndof ? Eb0[(size_t)c * hw + d] : 0.0; }
        for (int t = threadIdx.x; t < hw * 12; t += blockDim.x) { const int c = t / 12; bwin[t] = c < ndof ? Bd[(size_t)c * 12 + t % 12] : 0.0; }
        for (int t = threadIdx.x; t < 144; t += blockDim.x) c22s[t] = C22[t];
        __syncthreads();
        // TWO columns per step.  Column k first updates column k + 1 (phase A), which is then final too; every other target in the window
        // takes both columns' contributions in one read-modify-write (phase B: the pair enumeration relative to column k + 1, plus the
        // term of column k where it reaches); phase C scales and writes both columns of L and hands their slots to the incoming columns
        // k + hw, k + hw + 1.  Column k + 1 reaches one row further than the window: its contribution to column k + hw (diagonal entry
        // and border) is applied to the incoming column at the hand-over.  Incoming columns are staged six columns ahead by cp.async
        // into a ring of eight buffers (a step is shorter than a DRAM round trip with two CTAs per SM).
        const int lw = hw + 12;
        auto stage_col = [&](int cn) {
            if (cn < ndof && threadIdx.x < lw) {
                const int t = threadIdx.x;
                const double* src = t < hw ? Eb0 + (size_t)cn * hw + t : Bd + (size_t)cn * 12 + (t - hw);
                const unsigned dst = (unsigned)__cvta_generic_to_shared(stg + (cn & 7) * lw + t);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src));
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        };
#pragma unroll
        for (int q = 0; q < 6; q++) stage_col(hw + q);
        const int tid = threadIdx.x;
        int ks = 0, k = 0;                            // ks = k % hw, kept incrementally
        for (; k + 1 < n1; k += 2, ks = (ks + 2 >= hw ? ks + 2 - hw : ks + 2)) {
            const int ks1 = ks + 1 == hw ? 0 : ks + 1;
            double* ck = win + ks * hw; double* c1 = win + ks1 * hw;
            double* bk0 = bwin + ks * 12; double* bk1 = bwin + ks1 * 12;
            const double d0 = ck[0];
            const double rd0 = 1.0 / d0;
            const int mk0 = min(hb, ndof - 1 - k), mk1 = min(hb, ndof - 2 - k);
            stage_col(k + hw + 6); stage_col(k + hw + 7);
            // phase A: column k -> column k + 1 (rows k + 1 .. k + mk0) and its border entries.  Every thread forms the second pivot
            // d1 = c1[0] - ck[1]^2 / d0 itself (the same arithmetic as the thread that updates c1[0]), so that 1 / d1 is under way before the barrier
            const double l1 = ck[1] * rd0;
            const double d1 = c1[0] - ck[1] * l1;
            const double rd1 = 1.0 / d1;
            if (tid >= 1 && tid < mk0) c1[tid] -= ck[1 + tid] * l1;       // c1[0] stays as it came: nobody reads it after this point (d1 is in registers)
            else if (tid >= 244) bk1[tid - 244] -= bk0[tid - 244] * l1;   // (hb <= 243: the window needs (hb + 1)^2 doubles of shared memory anyway)
            __syncthreads();
            if ((!(d0 > 0.0) || !(d1 > 0.0)) && tid == 0) bad = 1;
            const bool reach_out = mk1 == hb;         // column k + 1 reaches row k + hw: the pair (hb, hb) and the border entry b = hb are not in the window
            const int npair = mk1 * (mk1 + 1) / 2 - (reach_out ? 1 : 0);
            const int nbb = 12 * (mk1 - (reach_out ? 1 : 0));
            // correction of the incoming column k + hw, formed now from the final column k + 1 (phase C overwrites that slot)
            double corr = 0.0;
            if (reach_out && tid < lw) {
                const double lh = c1[hb] * rd1;
                corr = tid == 0 ? c1[hb] * lh : (tid >= hw ? bk1[tid - hw] * lh : 0.0);
            }
            // phase B: four targets at a time per thread, all operands read before the first store
            for (int t0 = tid; t0 < npair; t0 += 4 * blockDim.x) {
                int idx[4]; double nv[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int t = t0 + u * blockDim.x;
                    idx[u] = -1;
                    if (t < npair) {
                        const int a = pair_tab[2 * t], b = pair_tab[2 * t + 1];      // rows / columns relative to column k + 1
                        int sl = ks1 + b; if (sl >= hw) sl -= hw;
                        idx[u] = sl * hw + (a - b);
                        double upd = c1[a] * c1[b] * rd1;
                        if (a + 1 <= mk0) upd = fma(ck[a + 1] * ck[b + 1], rd0, upd);
                        nv[u] = win[idx[u]] - upd;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) if (idx[u] >= 0) win[idx[u]] = nv[u];
            }
            for (int t = npair + tid; t < npair + nbb + 78; t += blockDim.x) {
                if (t < npair + nbb) {
                    const int u = t - npair, b = u / 12 + 1, j = u - 12 * (b - 1);
                    int sl = ks1 + b; if (sl >= hw) sl -= hw;
                    double upd = bk1[j] * c1[b] * rd1;
                    if (b + 1 <= mk0) upd = fma(bk0[j] * ck[b + 1], rd0, upd);
                    bwin[sl * 12 + j] -= upd;
                } else {
                    const int u = t - npair - nbb;
                    const int i = c22_i[u], j = c22_j[u];
                    const double v = bk0[i] * bk0[j] * rd0 + bk1[i] * bk1[j] * rd1;
                    c22s[12 * i + j] -= v; if (i != j) c22s[12 * j + i] -= v;
                }
            }
            __syncthreads();
            // phase C: columns k, k + 1 are final: scale to L, write them out, refill their slots with columns k + hw, k + hw + 1 (Bd
            // rows of incoming columns are still untouched: border entries are first modified when the column is in the window)
            asm volatile("cp.async.wait_group 6;\n" ::: "memory");      // this thread's pieces of columns k + hw, k + hw + 1 have landed
            for (int tt = tid; tt < 2 * lw; tt += blockDim.x) {             // lw <= 256: an item of column k (tt < lw) is its thread's first
                const int which = tt >= lw ? 1 : 0, t = tt - which * lw;
                const int cn = k + hw + which;
                double in = cn < ndof ? stg[(cn & 7) * lw + t] : 0.0;
                if (which == 0) in -= corr;
                double* col = which ? c1 : ck; double* bcol = which ? bk1 : bk0;
                const double rd = which ? rd1 : rd0, dd = which ? d1 : d0;
                const int mk = which ? mk1 : mk0;
                if (t < hw) {
                    Eb[(size_t)(k + which) * hw + t] = t == 0 ? dd : (t <= mk ? col[t] * rd : 0.0);
                    col[t] = in;
                } else {
                    const int j = t - hw;
                    Bd[(size_t)(k + which) * 12 + j] = bcol[j] * rd;
                    bcol[j] = in;
                }
            }
            __syncthreads();
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        if (k < n1) {                                 // odd number of columns: the last one alone (nothing comes in any more: k + hw >= ndof)
            double* ck = win + ks * hw;
            double* bk = bwin + ks * 12;
            const double dk = ck[0];
            if (!(dk > 0.0) && tid == 0) bad = 1;
            const double rd = 1.0 / dk;
            const int mk = min(hb, ndof - 1 - k);
            const int npair = mk * (mk + 1) / 2, nbb = 12 * mk;
            for (int t = tid; t < npair + nbb + 78; t += blockDim.x) {
                if (t < npair) {
                    const int dI = pair_tab[2 * t], dj = pair_tab[2 * t + 1];
                    int sl = ks + dj; if (sl >= hw) sl -= hw;
                    win[sl * hw + (dI - dj)] -= ck[dI] * ck[dj] * rd;
                } else if (t < npair + nbb) {
                    const int u = t - npair, dj = u / 12 + 1, j = u - 12 * (dj - 1);
                    int sl = ks + dj; if (sl >= hw) sl -= hw;
                    bwin[sl * 12 + j] -= bk[j] * ck[dj] * rd;
                } else {
                    const int u = t - npair - nbb;
                    const int i = c22_i[u], j = c22_j[u];
                    const double v = bk[i] * bk[j] * rd;
                    c22s[12 * i + j] -= v; if (i != j) c22s[12 * j + i] -= v;
                }
            }
            __syncthreads();
            if (tid < lw) {
                const int t = tid;
                if (t < hw) { Eb[(size_t)k * hw + t] = t == 0 ? dk : (t <= mk ? ck[t] * rd : 0.0); ck[t] = 0.0; }
                else { Bd[(size_t)k * 12 + (t - hw)] = bk[t - hw] * rd; bk[t - hw] = 0.0; }
            }
            __syncthreads();
        }
        // the tail columns (dof of the last two nodes) and the corner go back to global memory for the dense trailing block
        for (int t = threadIdx.x; t < SEC_MTAIL * hw; t += blockDim.x) { const int c = n1 + t / hw; Eb[(size_t)c * hw + t % hw] = win[(c % hw) * hw + t % hw]; }
        for (int t = threadIdx.x; t < SEC_MTAIL * 12; t += blockDim.x) { const int c = n1 + t / 12; Bd[(size_t)c * 12 + t % 12] = bwin[(c % hw) * 12 + t % 12]; }
        for (int t = threadIdx.x; t < 144; t += blockDim.x) C22[t] = c22s[t];
        __syncthreads();
    } else
    for (int k = 0; k < n1; k++) {
        double* ck = Eb + (size_t)k * (hb + 1);
        double* bk = Bd + (size_t)k * 12;
        const double dk = ck[0];
        if (!(dk > 0.0) && threadIdx.x == 0) bad = 1;
        const double rd = 1.0 / dk;
        const int mk = min(hb, ndof - 1 - k);
        const int npair = mk * (mk + 1) / 2, nbb = 12 * mk;
        for (int t = threadIdx.x; t < npair + nbb + 78; t += blockDim.x) {
            if (t < npair) {
                int dI, dj;
                if (A.pair_table) { dI = pair_tab[2 * t]; dj = pair_tab[2 * t + 1]; }
                else {
                    int di = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
                    while ((di + 1) * (di + 2) / 2 <= t) di++;
                    while (di * (di + 1) / 2 > t) di--;
                    dj = t - di * (di + 1) / 2 + 1; dI = di + 1;
                }
                Eb[(size_t)(k + dj) * (hb + 1) + (dI - dj)] -= ck[dI] * ck[dj] * rd;
            } else if (t < npair + nbb) {
                const int u = t - npair, dj = u / 12 + 1, j = u % 12;
                Bd[(size_t)(k + dj) * 12 + j] -= bk[j] * ck[dj] * rd;
            } else {
                int u = t - npair - nbb;          // 78 pairs i >= j of the 12 x 12 corner, kept symmetric
                int i = (int)((sqrt(8.0 * u + 1.0) - 1.0) * 0.5);
                while ((i + 1) * (i + 2) / 2 <= u) i++;
                while (i * (i + 1) / 2 > u) i--;
                const int j = u - i * (i + 1) / 2;
                const double v = bk[i] * bk[j] * rd;
                C22[12 * i + j] -= v; if (i != j) C22[12 * j + i] -= v;
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < mk + 12; t += blockDim.x) { if (t < mk) ck[1 + t] *= rd; else bk[t - mk] *= rd; }
        __syncthreads();
    }
    SEC_TICK(1);
    // ---- trailing block: [Schur complement of the last two nodes | their border rows; border' | corner], dense LU with partial pivoting ----
    for (int t = threadIdx.x; t < SEC_NT * SEC_NT; t += blockDim.x) {
        const int i = t / SEC_NT, j = t % SEC_NT;
        double v;
        if (i < SEC_MTAIL && j < SEC_MTAIL) { const int r = max(i, j), c = min(i, j); v = Eb[(size_t)(n1 + c) * (hb + 1) + (r - c)]; }
        else if (i < SEC_MTAIL) v = Bd[(size_t)(n1 + i) * 12 + (j - SEC_MTAIL)];
        else if (j < SEC_MTAIL) v = Bd[(size_t)(n1 + j) * 12 + (i - SEC_MTAIL)];
        else v = C22[12 * (i - SEC_MTAIL) + (j - SEC_MTAIL)];
        Ts[t] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < SEC_NT; k++) {
            int p = k; double best = fabs(Ts[k * SEC_NT + k]);
            for (int i = k + 1; i < SEC_NT; i++) if (fabs(Ts[i * SEC_NT + k]) > best) { best = fabs(Ts[i * SEC_NT + k]); p = i; }
            piv[k] = p;
            if (!(best > 0.0)) bad = 1;
            if (p != k) for (int j = 0; j < SEC_NT; j++) { const double tmp = Ts[k * SEC_NT + j]; Ts[k * SEC_NT + j] = Ts[p * SEC_NT + j]; Ts[p * SEC_NT + j] = tmp; }
            for (int i = k + 1; i < SEC_NT; i++) {
                const double f = Ts[i * SEC_NT + k] / Ts[k * SEC_NT + k];
                Ts[i * SEC_NT + k] = f;
                for (int j = k + 1; j < SEC_NT; j++) Ts[i * SEC_NT + j] -= f * Ts[k * SEC_NT + j];
            }
        }
    }
    __syncthreads();
    SEC_TICK(2);
    // ---- first solve: B2 = [0; Tr'; 0] (section.jl:694-697) -> dX, dY ----
    for (int t = threadIdx.x; t < (ndof + 12) * 6; t += blockDim.x) X2[t] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) { X2[(size_t)(ndof + 4) * 6 + 0] = -1.0; X2[(size_t)(ndof + 3) * 6 + 1] = 1.0; }
    __syncthreads();
    if (A.smem_window) sec_kkt_solve_win(Eb, Bd, Ts, piv, X2, ndof, hb, win); else sec_kkt_solve(Eb, Bd, Ts, piv, X2, ndof, hb);
    SEC_TICK(3);
    // ---- second right-hand side: B1 = [C' dX - C dX + L dY; I - L' dX; 0] (section.jl:712-724) ----
    const double* dX = X2; const double* dY = X2 + (size_t)ndof * 6;
    sec_band_mv(Cb, false, true, dX, X1, ndof, hb, 1.0, false);
    __syncthreads();
    sec_band_mv(Cb, false, false, dX, X1, ndof, hb, -1.0, true);
    __syncthreads();
    for (int t = threadIdx.x; t < ndof * 6; t += blockDim.x) {
        const int r = t / 6, k = t % 6; double acc = X1[t];
        for (int j = 0; j < 6; j++) acc = fma(L[(size_t)r * 6 + j], dY[j * 6 + k], acc);
        X1[t] = acc;
    }
    if (threadIdx.x < 36) G[threadIdx.x] = 0.0;
    __syncthreads();
    sec_gram(L, dX, ndof, G, red);                       // L' dX
    for (int t = threadIdx.x; t < 72; t += blockDim.x) { const int i = t / 6, j = t % 6; X1[(size_t)(ndof + i) * 6 + j] = i < 6 ? ((i == j) - G[6 * i + j]) : 0.0; }
    __syncthreads();
    SEC_TICK(4);
    if (A.smem_window) sec_kkt_solve_win(Eb, Bd, Ts, piv, X1, ndof, hb, win); else sec_kkt_solve(Eb, Bd, Ts, piv, X1, ndof, hb);
    SEC_TICK(5);
    // ---- S = X'(E X + C dX + R Y) + dX'(C' X + M dX + L Y) + Y'(R' X + L' dX + A Y)   (section.jl:736-757) ----
    const double* X = X1; const double* Y = X1 + (size_t)ndof * 6;
    sec_band_mv(Eb0, true, false, X, T1, ndof, hb, 1.0, false);
    __syncthreads();
    sec_band_mv(Cb, false, false, dX, T1, ndof, hb, 1.0, true);
    sec_band_mv(Cb, false, true, X, T2, ndof, hb, 1.0, false);
    __syncthreads();
    sec_band_mv(Mb, true, false, dX, T2, ndof, hb, 1.0, true);
    __syncthreads();
    for (int t = threadIdx.x; t < ndof * 6; t += blockDim.x) {
        const int r = t / 6, k = t % 6; double a1 = T1[t], a2 = T2[t];
        for (int j = 0; j < 6; j++) { a1 = fma(R[(size_t)r * 6 + j], Y[j * 6 + k], a1); a2 = fma(L[(size_t)r * 6 + j], Y[j * 6 + k], a2); }
        T1[t] = a1; T2[t] = a2;
    }
    if (threadIdx.x < 36) G[threadIdx.x] = 0.0;
    __syncthreads();
    sec_gram(X, T1, ndof, G, red);
    sec_gram(dX, T2, ndof, G, red);
    // Y'(R'X + L'dX + A Y): the two 6 x 6 Gram matrices R'X and L'dX, then a 6 x 6 product by one thread
    __shared__ double W[36];
    if (threadIdx.x < 36) W[threadIdx.x] = 0.0;
    __syncthreads();
    sec_gram(R, X, ndof, W, red);
    sec_gram(L, dX, ndof, W, red);
    if (threadIdx.x < 36) { const int i = threadIdx.x / 6, j = threadIdx.x % 6; double acc = W[threadIdx.x]; for (int k = 0; k < 6; k++) acc = fma(Am[6 * i + k], Y[k * 6 + j], acc); W[threadIdx.x] = acc; }
    __syncthreads();
    if (threadIdx.x < 36) { const int i = threadIdx.x / 6, j = threadIdx.x % 6; double acc = G[threadIdx.x]; for (int k = 0; k < 6; k++) acc = fma(Y[k * 6 + i], W[6 * k + j], acc); G[threadIdx.x] = acc; }
    __syncthreads();
    SEC_TICK(6);
    // ---- centres, transformation to the shear centre, GXBeam ordering (section.jl:759-781); mass matrix (:840-890) ----
    if (threadIdx.x == 0) {
        double Sm[6][6];
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Sm[i][j] = G[6 * i + j];
        const double xs = -Sm[5][1] / Sm[5][5], ys = Sm[5][0] / Sm[5][5];
        const double den = Sm[3][3] * Sm[4][4] - Sm[3][4] * Sm[3][4];
        const double xt = (Sm[3][3] * Sm[4][2] - Sm[3][4] * Sm[3][2]) / den, yt = (-Sm[3][2] * Sm[4][4] + Sm[3][4] * Sm[4][2]) / den;
        A.sc[2 * b] = xs; A.sc[2 * b + 1] = ys; A.tc[2 * b] = xt; A.tc[2 * b + 1] = yt;
        if (A.shear_center) {
            const double P[3][3] = {{0, 0, ys}, {0, 0, -xs}, {-ys, xs, 0}};
            double H[6][6], Ht[6][6], T[6][6];
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { H[i][j] = i == j; Ht[i][j] = i == j; }
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { H[i][3 + j] = P[j][i]; Ht[3 + i][j] = P[i][j]; }
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double a = 0; for (int k = 0; k < 6; k++) a += H[i][k] * Sm[k][j]; T[i][j] = a; }
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double a = 0; for (int k = 0; k < 6; k++) a += T[i][k] * Ht[k][j]; Sm[i][j] = a; }
        }
        const int idx[6] = {2, 0, 1, 5, 3, 4};
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) A.S[(size_t)b * 36 + 6 * i + j] = A.gxbeam_order ? Sm[idx[i]][idx[j]] : Sm[i][j];
        A.status[b] = bad;
    }
    // mass matrix: two passes over the elements (total mass and centre, then the moments of inertia about it)
    {
        __shared__ double mred[8 * 36]; __shared__ double macc[36];
        auto reduce3 = [&](double a0, double a1, double a2, int slot) {
            for (int o = 16; o; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
            __syncthreads();
            if ((threadIdx.x & 31) == 0) { mred[(threadIdx.x >> 5) * 3] = a0; mred[(threadIdx.x >> 5) * 3 + 1] = a1; mred[(threadIdx.x >> 5) * 3 + 2] = a2; }
            __syncthreads();
            if (threadIdx.x < 3) { double t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += mred[w * 3 + threadIdx.x]; macc[slot + threadIdx.x] = t; }
            __syncthreads();
        };
        auto geo = [&](int e, double& dm, double& xc, double& yc) {
            const int* cn = A.conn + 4 * e;
            double px[4], py[4];
            for (int i = 0; i < 4; i++) { px[i] = xy[2 * cn[i]]; py[i] = xy[2 * cn[i] + 1]; }
            double ar = 0; for (int i = 0; i < 4; i++) ar += px[i] * py[(i + 1) & 3] - px[(i + 1) & 3] * py[i];
            dm = A.mats[((size_t)b * D.ne + e) * 10 + 9] * 0.5 * ar;
            xc = 0.25 * (px[0] + px[1] + px[2] + px[3]); yc = 0.25 * (py[0] + py[1] + py[2] + py[3]);
        };
        double a0 = 0, a1 = 0, a2 = 0;
        for (int e = threadIdx.x; e < D.ne; e += blockDim.x) { double dm, xc, yc; geo(e, dm, xc, yc); a0 += dm; a1 += xc * dm; a2 += yc * dm; }
        reduce3(a0, a1, a2, 0);
        const double m = macc[0], xm = macc[1] / m, ym = macc[2] / m;
        a0 = a1 = a2 = 0;
        for (int e = threadIdx.x; e < D.ne; e += blockDim.x) { double dm, xc, yc; geo(e, dm, xc, yc); a0 += (yc - ym) * (yc - ym) * dm; a1 += (xc - xm) * (xc - xm) * dm; a2 += (xc - xm) * (yc - ym) * dm; }
        reduce3(a0, a1, a2, 3);
        if (threadIdx.x == 0) {
            const double Ixx = macc[3], Iyy = macc[4], Ixy = macc[5];
            const double Mm[36] = {m, 0, 0, 0, m * ym, -m * xm, 0, m, 0, -m * ym, 0, 0, 0, 0, m, m * xm, 0, 0, 0, -m * ym, m * xm, Ixx + Iyy, 0, 0,
                                   m * ym, 0, 0, 0, Ixx, -Ixy, -m * xm, 0, 0, 0, -Ixy, Iyy};
            for (int i = 0; i < 36; i++) A.M[(size_t)b * 36 + i] = Mm[i];
            A.mc[2 * b] = xm; A.mc[2 * b + 1] = ym;
        }
    }
#ifdef SEC_TIMING
    SEC_TICK(7);
    if (b == 0 && threadIdx.x == 0)
        printf("[section timing] cycles: init %lld, factor %lld, tail LU %lld, solve1 %lld, rhs2 %lld, solve2 %lld, products %lld, centres+mass %lld\n",
               tph[0], tph[1], tph[2], tph[3], tph[4], tph[5], tph[6], tph[7]);
#endif
}

// ------------------------------------------------------------------------------------------------------------------ host side
struct gxb_section {
    int device = 0; cudaStream_t stream = nullptr;
    SecDims d{};
    std::vector<int> conn, perm;
    int *d_conn = nullptr, *d_perm = nullptr;
};

// reverse Cuthill-McKee on the node graph (nodes of one element are mutually adjacent); start from a minimum-degree node
static std::vector<int> rcm_order(int nn, int ne, const std::vector<int>& conn) {
    std::vector<std::vector<int>> adj(nn);
    for (int e = 0; e < ne; e++) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) if (i != j) adj[conn[4 * e + i]].push_back(conn[4 * e + j]);
    for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    std::vector<int> order; order.reserve(nn);
    std::vector<char> seen(nn, 0);
    for (;;) {
        int start = -1;
        for (int i = 0; i < nn; i++) if (!seen[i] && (start < 0 || adj[i].size() < adj[start].size())) start = i;
        if (start < 0) break;
        std::queue<int> q; q.push(start); seen[start] = 1;
        while (!q.empty()) {
            const int u = q.front(); q.pop(); order.push_back(u);
            std::vector<int> nb;
            for (int v : adj[u]) if (!seen[v]) { seen[v] = 1; nb.push_back(v); }
            std::sort(nb.begin(), nb.end(), [&](int a, int b) { return adj[a].size() < adj[b].size() || (adj[a].size() == adj[b].size() && a < b); });
            for (int v : nb) q.push(v);
        }
    }
    std::reverse(order.begin(), order.end());
    std::vector<int> perm(nn);
    for (int k = 0; k < nn; k++) perm[order[k]] = k;
    return perm;
}

extern "C" {

int gxb_section_create(gxb_section** out, int device, int64_t nn, int64_t ne, const int64_t* conn) {
    if (!out || nn < 4 || ne < 1 || !conn) return gxb_fail_(GXB_ERR_INVALID, "gxb_section_create: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return gxb_fail_(GXB_ERR_CUDA, "gxb_section_create: no usable CUDA device; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return gxb_fail_(GXB_ERR_INVALID, "gxb_section_create: bad device index");
    SCK(cudaSetDevice(device));
    gxb_section* s = new gxb_section();
    s->device = device;
    s->conn.resize(4 * ne);
    for (int64_t i = 0; i < 4 * ne; i++) {
        if (conn[i] < 1 || conn[i] > nn) { delete s; return gxb_fail_(GXB_ERR_INVALID, "gxb_section_create: node numbers are 1-based (MeshElement.nodenum)"); }
        s->conn[i] = (int)conn[i] - 1;
    }
    s->perm = rcm_order((int)nn, (int)ne, s->conn);
    int span = 0;
    for (int64_t e = 0; e < ne; e++) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) span = std::max(span, std::abs(s->perm[s->conn[4 * e + i]] - s->perm[s->conn[4 * e + j]]));
    SecDims& d = s->d;
    d.nn = (int)nn; d.ne = (int)ne; d.ndof = 3 * (int)nn; d.hb = 3 * span + 2;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 1) & ~(size_t)1; return r; };
    d.oEb0 = take((size_t)d.ndof * (d.hb + 1)); d.oEb = take((size_t)d.ndof * (d.hb + 1)); d.oCb = take((size_t)d.ndof * (2 * d.hb + 1)); d.oMb = take((size_t)d.ndof * (d.hb + 1));
    d.oR = take((size_t)d.ndof * 6); d.oL = take((size_t)d.ndof * 6); d.oBd = take((size_t)d.ndof * 12); d.oA = take(36); d.oC22 = take(144);
    d.oX1 = take((size_t)(d.ndof + 12) * 6); d.oX2 = take((size_t)(d.ndof + 12) * 6); d.oT1 = take((size_t)d.ndof * 6); d.oT2 = take((size_t)d.ndof * 6); d.oMass = take(8);
    d.per_instance = o;
    SCK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    SCK(cudaMalloc(&s->d_conn, 4 * ne * sizeof(int))); SCK(cudaMalloc(&s->d_perm, nn * sizeof(int)));
    SCK(cudaMemcpy(s->d_conn, s->conn.data(), 4 * ne * sizeof(int), cudaMemcpyHostToDevice));
    SCK(cudaMemcpy(s->d_perm, s->perm.data(), nn * sizeof(int), cudaMemcpyHostToDevice));
    *out = s;
    return GXB_OK;
}

void gxb_section_destroy(gxb_section* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaFree(s->d_conn); cudaFree(s->d_perm);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int gxb_section_info(gxb_section* s, int64_t* half_bandwidth, int64_t* scratch_doubles_per_instance) {
    if (!s) return gxb_fail_(GXB_ERR_INVALID, "gxb_section_info: null");
    if (half_bandwidth) *half_bandwidth = s->d.hb;
    if (scratch_doubles_per_instance) *scratch_doubles_per_instance = (int64_t)s->d.per_instance;
    return GXB_OK;
}

static int section_run(gxb_section* s, int64_t nbatch, const double* nodes, const double* materials, const double* theta, int gxbeam_order,
                       int shear_center, double* S, double* sc, double* tc, double* M, double* mc, int32_t* ok, double* elem_out, float* ms_out,
                       const SecRecoverArgs* rec = nullptr /* HOST pointers */) {
    if (!s || nbatch < 1 || !nodes || !materials || !theta) return gxb_fail_(GXB_ERR_INVALID, "gxb_section: bad arguments");
    SCK(cudaSetDevice(s->device));
    const SecDims& d = s->d;
    // instances are processed in chunks so that the band scratch (a few MB per instance) stays within a fixed budget
    size_t free_b = 0, total_b = 0;
    SCK(cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = std::min<size_t>(free_b / 2, (size_t)48 << 30);
    int64_t chunk = elem_out ? nbatch : std::max<int64_t>(1, std::min<int64_t>(nbatch, (int64_t)(budget / (d.per_instance * 8))));
    double *d_nodes = nullptr, *d_mats = nullptr, *d_theta = nullptr, *d_scratch = nullptr, *d_out = nullptr, *d_elem = nullptr; int* d_status = nullptr;
    struct Guard { std::vector<void*> p; ~Guard() { for (void* q : p) cudaFree(q); } } g;
    auto alloc = [&](double** p, size_t n) { cudaError_t e = cudaMalloc((void**)p, n * 8); if (e == cudaSuccess) g.p.push_back(*p); return e; };
    SCK(alloc(&d_nodes, (size_t)chunk * d.nn * 2)); SCK(alloc(&d_mats, (size_t)chunk * d.ne * 10)); SCK(alloc(&d_theta, (size_t)chunk * d.ne));
    SCK(alloc(&d_out, (size_t)chunk * 80));
    if (elem_out) SCK(alloc(&d_elem, (size_t)chunk * d.ne * SEC_ELEM)); else SCK(alloc(&d_scratch, (size_t)chunk * d.per_instance));
    double *d_F = nullptr, *d_Mo = nullptr, *d_str = nullptr, *d_rec = nullptr;
    if (rec) {
        SCK(alloc(&d_F, (size_t)chunk * 3)); SCK(alloc(&d_Mo, (size_t)chunk * 3)); SCK(alloc(&d_rec, (size_t)chunk * d.ne * 25));
        if (rec->strengths) SCK(alloc(&d_str, (size_t)chunk * d.ne * 9));
    }
    { cudaError_t e = cudaMalloc((void**)&d_status, chunk * sizeof(int)); if (e != cudaSuccess) return gxb_fail_(GXB_ERR_CUDA, cudaGetErrorString(e)); g.p.push_back(d_status); }
    cudaEvent_t e0, e1; SCK(cudaEventCreate(&e0)); SCK(cudaEventCreate(&e1));
    float ms_total = 0.f;
    for (int64_t b0 = 0; b0 < nbatch; b0 += chunk) {
        const int64_t nb = std::min(chunk, nbatch - b0);
        SCK(cudaMemcpyAsync(d_nodes, nodes + (size_t)b0 * d.nn * 2, (size_t)nb * d.nn * 16, cudaMemcpyHostToDevice, s->stream));
        SCK(cudaMemcpyAsync(d_mats, materials + (size_t)b0 * d.ne * 10, (size_t)nb * d.ne * 80, cudaMemcpyHostToDevice, s->stream));
        SCK(cudaMemcpyAsync(d_theta, theta + (size_t)b0 * d.ne, (size_t)nb * d.ne * 8, cudaMemcpyHostToDevice, s->stream));
        SecArgs A{};
        A.d = d; A.nbatch = nb; A.conn = s->d_conn; A.perm = s->d_perm; A.nodes = d_nodes; A.mats = d_mats; A.theta = d_theta; A.scratch = d_scratch;
        A.elem_out = d_elem; A.gxbeam_order = gxbeam_order; A.shear_center = shear_center;
        A.S = d_out; A.sc = d_out + (size_t)nb * 36; A.tc = A.sc + (size_t)nb * 2; A.M = A.tc + (size_t)nb * 2; A.mc = A.M + (size_t)nb * 36; A.status = d_status;
        SCK(cudaEventRecord(e0, s->stream));
        if (!elem_out) SCK(cudaMemsetAsync(d_scratch, 0, (size_t)nb * d.per_instance * 8, s->stream));
        k_section_assemble<<<(unsigned)((nb * d.ne + 63) / 64), 64, 0, s->stream>>>(A);
        SCK(cudaGetLastError());
        if (elem_out) {
            SCK(cudaMemcpyAsync(elem_out + (size_t)b0 * d.ne * SEC_ELEM, d_elem, (size_t)nb * d.ne * SEC_ELEM * 8, cudaMemcpyDeviceToHost, s->stream));
        } else {
            // shared memory: the band window ((hb+1)^2 + 12 (hb+1) + 144 doubles) and the pair table when they fit the opt-in maximum
            int optin = 0;
            SCK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
            const size_t tab = (size_t)d.hb * (d.hb + 1) / 2 * 4;
            const size_t winb = ((size_t)(d.hb + 1) * (d.hb + 1) + (size_t)(d.hb + 1) * 12 + 144 + 8 * (size_t)(d.hb + 13)) * 8;
            const size_t statics = 8 * 1024;
            A.smem_window = winb + tab + statics <= (size_t)optin && d.hb + 13 <= 256 && d.hb >= 11;      // one thread per entry of an incoming column; the solves' 16-column ring needs hb >= 11 to fit the window's bytes
            A.pair_table = A.smem_window || (tab + statics <= (size_t)optin && d.hb < 65535);
            const size_t dyn = (A.smem_window ? winb : 0) + (A.pair_table ? tab : 0);
            if (dyn > 40 * 1024) SCK(cudaFuncSetAttribute(k_section_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            k_section_solve<<<(unsigned)nb, 256, dyn, s->stream>>>(A);
            SCK(cudaGetLastError());
            const size_t per = (size_t)nb * d.ne;
            if (rec) {
                SCK(cudaMemcpyAsync(d_F, rec->F + (size_t)b0 * 3, (size_t)nb * 24, cudaMemcpyHostToDevice, s->stream));
                SCK(cudaMemcpyAsync(d_Mo, rec->Mo + (size_t)b0 * 3, (size_t)nb * 24, cudaMemcpyHostToDevice, s->stream));
                if (d_str) SCK(cudaMemcpyAsync(d_str, rec->strengths + (size_t)b0 * d.ne * 9, per * 72, cudaMemcpyHostToDevice, s->stream));
                SecRecoverArgs Rv{};
                Rv.F = d_F; Rv.Mo = d_Mo; Rv.strengths = d_str; Rv.gxbeam_order = rec->gxbeam_order;
                Rv.strain_b = d_rec; Rv.stress_b = d_rec + per * 6; Rv.strain_p = d_rec + per * 12; Rv.stress_p = d_rec + per * 18;
                Rv.failure = d_str ? d_rec + per * 24 : nullptr;
                k_section_recover<<<(unsigned)((per + 63) / 64), 64, 0, s->stream>>>(A, Rv);
                SCK(cudaGetLastError());
            }
            SCK(cudaEventRecord(e1, s->stream));
            if (rec) {
                double* const dst[4] = {rec->strain_b, rec->stress_b, rec->strain_p, rec->stress_p};
                for (int k = 0; k < 4; k++)
                    if (dst[k]) SCK(cudaMemcpyAsync(dst[k] + (size_t)b0 * d.ne * 6, d_rec + per * 6 * k, per * 48, cudaMemcpyDeviceToHost, s->stream));
                if (rec->failure && d_str) SCK(cudaMemcpyAsync(rec->failure + (size_t)b0 * d.ne, d_rec + per * 24, per * 8, cudaMemcpyDeviceToHost, s->stream));
            }
            if (S) SCK(cudaMemcpyAsync(S + (size_t)b0 * 36, A.S, (size_t)nb * 36 * 8, cudaMemcpyDeviceToHost, s->stream));
            if (sc) SCK(cudaMemcpyAsync(sc + (size_t)b0 * 2, A.sc, (size_t)nb * 16, cudaMemcpyDeviceToHost, s->stream));
            if (tc) SCK(cudaMemcpyAsync(tc + (size_t)b0 * 2, A.tc, (size_t)nb * 16, cudaMemcpyDeviceToHost, s->stream));
            if (M) SCK(cudaMemcpyAsync(M + (size_t)b0 * 36, A.M, (size_t)nb * 36 * 8, cudaMemcpyDeviceToHost, s->stream));
            if (mc) SCK(cudaMemcpyAsync(mc + (size_t)b0 * 2, A.mc, (size_t)nb * 16, cudaMemcpyDeviceToHost, s->stream));
            std::vector<int> st(nb);
            SCK(cudaMemcpyAsync(st.data(), d_status, nb * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
            SCK(cudaStreamSynchronize(s->stream));
            float ms = 0; SCK(cudaEventElapsedTime(&ms, e0, e1)); ms_total += ms;
            if (ok) for (int64_t i = 0; i < nb; i++) ok[b0 + i] = st[i] == 0;
        }
        SCK(cudaStreamSynchronize(s->stream));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms_out) *ms_out = ms_total;
    return GXB_OK;
}

int gxb_section_properties(gxb_section* s, int64_t nbatch, const double* nodes, const double* materials, const double* theta, int gxbeam_order,
                           int shear_center, double* S, double* sc, double* tc, double* M, double* mc, int32_t* ok, double* device_ms) {
    float ms = 0.f;
    int rc = section_run(s, nbatch, nodes, materials, theta, gxbeam_order, shear_center, S, sc, tc, M, mc, ok, nullptr, &ms);
    if (device_ms) *device_ms = ms;
    return rc;
}

int gxb_section_strain_recovery(gxb_section* s, int64_t nbatch, const double* nodes, const double* materials, const double* theta,
                                const double* strengths, const double* F, const double* M, int gxbeam_order, double* strain_b, double* stress_b,
                                double* strain_p, double* stress_p, double* failure, double* S, int32_t* ok, double* device_ms) {
    if (!F || !M) return gxb_fail_(GXB_ERR_INVALID, "gxb_section_strain_recovery: F and M are required");
    if (failure && !strengths) return gxb_fail_(GXB_ERR_INVALID, "gxb_section_strain_recovery: the Tsai-Wu index needs the material strengths");
    SecRecoverArgs rec{};
    rec.F = F; rec.Mo = M; rec.strengths = strengths; rec.gxbeam_order = gxbeam_order;
    rec.strain_b = strain_b; rec.stress_b = stress_b; rec.strain_p = strain_p; rec.stress_p = stress_p; rec.failure = failure;
    float ms = 0.f;
    int rc = section_run(s, nbatch, nodes, materials, theta, gxbeam_order, 1, S, nullptr, nullptr, nullptr, nullptr, ok, nullptr, &ms, &rec);
    if (device_ms) *device_ms = ms;
    return rc;
}

int gxb_section_element_matrices(gxb_section* s, int64_t nbatch, const double* nodes, const double* materials, const double* theta, double* out) {
    if (!out) return gxb_fail_(GXB_ERR_INVALID, "gxb_section_element_matrices: null output");
    return section_run(s, nbatch, nodes, materials, theta, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, out, nullptr);
}

}  // extern "C"
