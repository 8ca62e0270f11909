// ORACLE — TEST INFRASTRUCTURE ONLY (see gxb_math.hpp).
//
// Matrix storages, a LAPACK-dgbtf2-style pivoted band LU, and the Newton / BackTracking driver that the reference
// reaches through NLsolve.jl + LineSearches.jl (third-party, not under /root/reference; compat "NLsolve 4",
// Project.toml:27-49).  Call sites restated: src/analyses.jl:3506-3524 (lsolve!), 3556-3599 (nlsolve!),
// defaults src/analyses.jl:94-97 (method=:newton, BackTracking(maxstep=1e6), ftol=1e-9, iterations=1000).
// The algorithm follows the published NLsolve `newton_` / LineSearches `BackTracking` (order 3) sources as
// summarised in SURVEY.md App. C.  Iteration counts are "parity unpinned" w.r.t. Julia (no Julia available).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace gxo {

template <class T> struct DenseMat {  // column-major n x n
    int64_t n;
    T* a;
    void zero() { std::fill(a, a + n * n, T(0)); }
    void set(int64_t i, int64_t j, T v) { a[i + n * j] = v; }
    void add(int64_t i, int64_t j, T v) { a[i + n * j] = a[i + n * j] + v; }
    void zero_row(int64_t i) { for (int64_t j = 0; j < n; j++) a[i + n * j] = T(0); }
};

struct BandMat {  // LAPACK general band storage with room for pivoting fill: ld = 2kl+ku+1
    int64_t n = 0, kl = 0, ku = 0, ld = 0;
    std::vector<double> ab;
    std::vector<int64_t> ipiv;
    void init(int64_t n_, int64_t kl_, int64_t ku_) { n = n_; kl = kl_; ku = ku_; ld = 2 * kl + ku + 1; ab.assign(ld * n, 0.0); ipiv.assign(n, 0); }
    double& at(int64_t i, int64_t j) { return ab[(kl + ku + i - j) + ld * j]; }
    void zero() { std::fill(ab.begin(), ab.end(), 0.0); }
    void set(int64_t i, int64_t j, double v) { at(i, j) = v; }
    void add(int64_t i, int64_t j, double v) { at(i, j) += v; }
    void zero_row(int64_t i) { for (int64_t j = std::max<int64_t>(0, i - kl); j <= std::min(n - 1, i + ku); j++) at(i, j) = 0.0; }
    // y = A x   (original, un-factored entries)
    void matvec(const double* x, double* y) {
        for (int64_t i = 0; i < n; i++) y[i] = 0.0;
        for (int64_t j = 0; j < n; j++)
            for (int64_t i = std::max<int64_t>(0, j - ku); i <= std::min(n - 1, j + kl); i++) y[i] += at(i, j) * x[j];
    }
    // y = A' x
    void matvec_t(const double* x, double* y) {
        for (int64_t j = 0; j < n; j++) {
            double s = 0.0;
            for (int64_t i = std::max<int64_t>(0, j - ku); i <= std::min(n - 1, j + kl); i++) s += at(i, j) * x[i];
            y[j] = s;
        }
    }
    // dgbtf2: returns 0 or (index+1) of the first exactly-zero pivot
    int64_t factor() {
        int64_t kv = kl + ku, info = 0, ju = 0;
        for (int64_t j = 0; j < n; j++) {
            int64_t km = std::min(kl, n - 1 - j);
            int64_t jp = 0; double amax = std::fabs(ab[kv + ld * j]);
            for (int64_t i = 1; i <= km; i++) { double v = std::fabs(ab[kv + i + ld * j]); if (v > amax) { amax = v; jp = i; } }
            ipiv[j] = j + jp;
            if (ab[kv + jp + ld * j] != 0.0) {
                ju = std::max(ju, std::min(j + ku + jp, n - 1));
                if (jp != 0)
                    for (int64_t c = j; c <= ju; c++) std::swap(ab[kv + jp + j - c + ld * c], ab[kv + j - c + ld * c]);
                if (km > 0) {
                    double rinv = 1.0 / ab[kv + ld * j];
                    for (int64_t i = 1; i <= km; i++) ab[kv + i + ld * j] *= rinv;
                    for (int64_t c = j + 1; c <= ju; c++) {
                        double ujc = ab[kv + j - c + ld * c];
                        if (ujc != 0.0) {
                            double* colc = &ab[kv + j - c + ld * c];
                            const double* l = &ab[kv + ld * j];
                            for (int64_t i = 1; i <= km; i++) colc[i] -= l[i] * ujc;
                        }
                    }
                }
            } else if (info == 0) info = j + 1;
        }
        return info;
    }
    // dgbtrs, no transpose, one right-hand side
    void solve(double* b) {
        int64_t kv = kl + ku;
        for (int64_t j = 0; j < n - 1; j++) {
            int64_t lm = std::min(kl, n - 1 - j);
            int64_t l = ipiv[j];
            if (l != j) std::swap(b[l], b[j]);
            double bj = b[j];
            const double* lc = &ab[kv + ld * j];
            for (int64_t i = 1; i <= lm; i++) b[j + i] -= lc[i] * bj;
        }
        for (int64_t j = n - 1; j >= 0; j--) {
            b[j] /= ab[kv + ld * j];
            double bj = b[j];
            int64_t i0 = std::max<int64_t>(0, j - kv);
            for (int64_t i = i0; i < j; i++) b[i] -= ab[kv + i - j + ld * j] * bj;
        }
    }
};

// dense LU with partial pivoting, col-major, in place; returns 0 or index+1 of first zero pivot
inline int64_t dense_lu_solve(int64_t n, double* a, double* b) {
    std::vector<int64_t> piv(n);
    for (int64_t j = 0; j < n; j++) {
        int64_t p = j; double amax = std::fabs(a[j + n * j]);
        for (int64_t i = j + 1; i < n; i++) if (std::fabs(a[i + n * j]) > amax) { amax = std::fabs(a[i + n * j]); p = i; }
        if (amax == 0.0) return j + 1;
        if (p != j) { for (int64_t c = 0; c < n; c++) std::swap(a[p + n * c], a[j + n * c]); std::swap(b[p], b[j]); }
        double rinv = 1.0 / a[j + n * j];
        for (int64_t i = j + 1; i < n; i++) a[i + n * j] *= rinv;
        for (int64_t c = j + 1; c < n; c++) { double u = a[j + n * c]; if (u != 0.0) for (int64_t i = j + 1; i < n; i++) a[i + n * c] -= a[i + n * j] * u; }
        for (int64_t i = j + 1; i < n; i++) b[i] -= a[i + n * j] * b[j];
    }
    for (int64_t j = n - 1; j >= 0; j--) { b[j] /= a[j + n * j]; for (int64_t i = 0; i < j; i++) b[i] -= a[i + n * j] * b[j]; }
    return 0;
}

struct NewtonOpts {
    int linear = 0;
    double ftol = 1e-9;
    int64_t iterations = 1000;
    double maxstep = 1e6, c1 = 1e-4, rho_hi = 0.5, rho_lo = 0.1;
    int64_t ls_iterations = 1000;
};
struct NewtonResult {
    int converged = 0;
    int64_t iters = 0;       // Newton iterations (= linear solves)
    int64_t f_calls = 0;     // residual evaluations
    int64_t ls_backtracks = 0;
    double resid_inf = 0.0;
    int error = 0;           // 1: non-finite residual at start, 2: line search failed
};

inline double norm_inf(int64_t n, const double* v) { double m = 0.0; for (int64_t i = 0; i < n; i++) { double a = std::fabs(v[i]); if (a > m || std::isnan(a)) m = a; } return m; }
inline bool any_nan(int64_t n, const double* v) { for (int64_t i = 0; i < n; i++) if (std::isnan(v[i])) return true; return false; }

// p <- -(J \ F) with the NLsolve singular-Jacobian fallback.  J (band) is destroyed.
inline void newton_direction(BandMat& J, const double* F, double* p) {
    int64_t n = J.n;
    std::vector<double> keep;           // copy for the (rare) fallback
    keep = J.ab;
    for (int64_t i = 0; i < n; i++) p[i] = F[i];
    int64_t info = J.factor();
    if (info == 0) {
        J.solve(p);
        for (int64_t i = 0; i < n; i++) p[i] = -p[i];
        return;
    }
    // fjac2 = J'J ; lambda = 1e6*sqrt(n*eps)*norm(fjac2,1) ; p = -(fjac2 + lambda I) \ (J'F)
    J.ab = keep;
    std::vector<double> A(n * n, 0.0), D(n * n, 0.0), rhs(n, 0.0);
    for (int64_t j = 0; j < n; j++)
        for (int64_t i = std::max<int64_t>(0, j - J.ku); i <= std::min(n - 1, j + J.kl); i++) D[i + n * j] = J.at(i, j);
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < n; j++) { double s = 0.0; for (int64_t k = 0; k < n; k++) s += D[k + n * i] * D[k + n * j]; A[i + n * j] = s; }
    double n1 = 0.0;
    for (int64_t j = 0; j < n; j++) { double s = 0.0; for (int64_t i = 0; i < n; i++) s += std::fabs(A[i + n * j]); n1 = std::max(n1, s); }
    double lambda = 1e6 * std::sqrt(double(n) * std::numeric_limits<double>::epsilon()) * n1;
    for (int64_t j = 0; j < n; j++) { double s = 0.0; for (int64_t k = 0; k < n; k++) s += D[k + n * j] * F[k]; rhs[j] = s; }
    for (int64_t i = 0; i < n * n; i++) A[i] = -A[i];
    for (int64_t i = 0; i < n; i++) A[i + n * i] -= lambda;
    dense_lu_solve(n, A.data(), rhs.data());
    for (int64_t i = 0; i < n; i++) p[i] = rhs[i];
}

// Generic damped Newton.  `resid(x, F)` fills F; `jacob(x, J)` fills the band matrix (after J.zero()).
template <class RF, class JF>
NewtonResult newton_solve(int64_t n, double* x, double* F, BandMat& J, RF&& resid, JF&& jacob, const NewtonOpts& o) {
    NewtonResult res;
    std::vector<double> p(n), xold(n), xls(n), g(n), Jp(n);
    if (o.linear) {  // analyses.jl:3506-3524
        resid(x, F); res.f_calls++;
        jacob(x, J);
        newton_direction(J, F, p.data());
        for (int64_t i = 0; i < n; i++) x[i] = x[i] + p[i];
        res.converged = 1; res.iters = 1; res.resid_inf = norm_inf(n, F);
        return res;
    }
    resid(x, F); res.f_calls++;
    jacob(x, J);
    for (int64_t i = 0; i < n; i++) if (!std::isfinite(F[i])) { res.error = 1; res.resid_inf = norm_inf(n, F); return res; }
    bool converged = norm_inf(n, F) <= o.ftol;
    bool stopped = any_nan(n, x) || any_nan(n, F);
    int64_t it = 0;
    while (!stopped && !converged && it < o.iterations) {
        it++;
        if (it > 1) jacob(x, J);
        // g = J'F and dphi0 = g'p need the un-factored J: do them before newton_direction destroys it
        J.matvec_t(F, g.data());
        newton_direction(J, F, p.data());
        std::copy(x, x + n, xold.begin());
        double phi0 = 0.0; for (int64_t i = 0; i < n; i++) phi0 += F[i] * F[i]; phi0 /= 2;
        double dphi0 = 0.0; for (int64_t i = 0; i < n; i++) dphi0 += g[i] * p[i];
        double pinf = norm_inf(n, p.data());
        double a0 = std::min(1.0, o.maxstep / pinf);
        auto phi = [&](double a) {
            for (int64_t i = 0; i < n; i++) xls[i] = xold[i] + a * p[i];
            resid(xls.data(), F); res.f_calls++;
            double s = 0.0; for (int64_t i = 0; i < n; i++) s += F[i] * F[i];
            return s / 2;
        };
        double phx0 = phi0, a1 = a0, a2 = a0;
        double phx1 = phi(a1);
        int iterfinite = 0; const int iterfinitemax = 52;  // -log2(eps)
        while (!std::isfinite(phx1) && iterfinite < iterfinitemax) { iterfinite++; a1 = a2; a2 = a1 / 2; phx1 = phi(a2); }
        int64_t k = 0;
        while (phx1 > phi0 + o.c1 * a2 * dphi0) {
            k++;
            if (k > o.ls_iterations) { res.error = 2; break; }
            double atmp;
            if (k == 1) {
                atmp = -(dphi0 * a2 * a2) / (2 * (phx1 - phi0 - dphi0 * a2));
            } else {
                double div = 1.0 / (a1 * a1 * a2 * a2 * (a2 - a1));
                double a = (a1 * a1 * (phx1 - phi0 - dphi0 * a2) - a2 * a2 * (phx0 - phi0 - dphi0 * a1)) * div;
                double b = (-a1 * a1 * a1 * (phx1 - phi0 - dphi0 * a2) + a2 * a2 * a2 * (phx0 - phi0 - dphi0 * a1)) * div;
                if (std::fabs(a) <= std::numeric_limits<double>::epsilon()) atmp = dphi0 / (2 * b);
                else { double d = std::max(b * b - 3 * a * dphi0, 0.0); atmp = (-b + std::sqrt(d)) / (3 * a); }
            }
            a1 = a2;
            // NaNMath.min / max: ignore NaN operands
            double hi = a2 * o.rho_hi, lo = a2 * o.rho_lo;
            atmp = std::isnan(atmp) ? hi : std::min(atmp, hi);
            a2 = std::max(atmp, lo);
            phx0 = phx1; phx1 = phi(a2);
            res.ls_backtracks++;
        }
        if (res.error) break;
        std::copy(xls.begin(), xls.end(), x);
        double dx = 0.0; for (int64_t i = 0; i < n; i++) dx = std::max(dx, std::fabs(x[i] - xold[i]));
        bool x_conv = dx <= 0.0;
        bool f_conv = norm_inf(n, F) <= o.ftol;
        stopped = any_nan(n, x) || any_nan(n, F);
        converged = x_conv || f_conv;
        res.converged = f_conv ? 1 : 0;  // analyses.jl:3595 : converged[] = result.f_converged
    }
    if (it == 0) res.converged = converged ? 1 : 0;
    res.iters = it;
    res.resid_inf = norm_inf(n, F);
    return res;
}

}  // namespace gxo
