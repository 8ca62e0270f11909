// ORACLE — TEST INFRASTRUCTURE ONLY.
// Freezes SURVEY §8(d)'s per-element / per-point flop counts f_e, f_p: the static (and steady / Newmark) residual + Jacobian restatements
// are instantiated with a scalar type that counts every FP64 add, multiply, divide and square root it performs (a multiply-add counts 2,
// like an FMA).  The reference evaluates the element/point properties twice per Newton iteration (once in *_residual!, once in
// *_jacobian!, src/system.jl:400-418 and :663-683); "fused" below counts them once, which is what the device kernels do.
//   g++ -O1 -std=c++17 -o _build/opcount opcount.cpp && _build/opcount
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>

struct Ops { long long add = 0, mul = 0, div = 0, sqr = 0; long long flops() const { return add + mul + div + sqr; } };
static Ops g_ops;
struct Cnt {
    double v;
    Cnt() : v(0) {}
    Cnt(double x) : v(x) {}
    Cnt& operator+=(const Cnt& o) { g_ops.add++; v += o.v; return *this; }
    Cnt& operator-=(const Cnt& o) { g_ops.add++; v -= o.v; return *this; }
    Cnt& operator*=(const Cnt& o) { g_ops.mul++; v *= o.v; return *this; }
    Cnt& operator/=(const Cnt& o) { g_ops.div++; v /= o.v; return *this; }
};
inline Cnt operator+(Cnt a, const Cnt& b) { g_ops.add++; return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, const Cnt& b) { g_ops.add++; return Cnt(a.v - b.v); }
inline Cnt operator*(Cnt a, const Cnt& b) { g_ops.mul++; return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, const Cnt& b) { g_ops.div++; return Cnt(a.v / b.v); }
inline Cnt operator-(const Cnt& a) { return Cnt(-a.v); }            // sign flips are free (operand modifiers on the device)
inline Cnt sqrt(const Cnt& a) { g_ops.sqr++; return Cnt(std::sqrt(a.v)); }
inline double realpart(const Cnt& a) { return a.v; }
inline bool operator==(const Cnt& a, const Cnt& b) { return a.v == b.v; }
inline bool operator!=(const Cnt& a, const Cnt& b) { return a.v != b.v; }
inline bool operator>(const Cnt& a, const Cnt& b) { return a.v > b.v; }
inline bool operator<(const Cnt& a, const Cnt& b) { return a.v < b.v; }

#include "gxb_math.hpp"
#include "gxb_static.hpp"

struct NullJac {           // the sparse setindex! of the reference is not arithmetic
    void set(int64_t, int64_t, const Cnt&) {}
    void add(int64_t, int64_t, const Cnt&) { g_ops.add++; }
    void zero() {}
    void zero_row(int64_t) {}
};

int main() {
    using namespace gxo;
    // three-element cantilever, anisotropic (fully populated) compliance, free interior points, follower-free tip load: the generic case
    const int ne = 3, np = 4;
    int64_t start[ne] = {1, 2, 3}, stop[ne] = {2, 3, 4};
    std::vector<double> elems(ne * 91, 0.0), points(np * 3, 0.0), bc(np * 18, 0.0);
    std::vector<uint8_t> pd(np * 6, 0), pl(np * 6, 1);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return 0.1 + (s >> 8) / 16777216.0; };
    for (int e = 0; e < ne; e++) {
        double* E = elems.data() + 91 * e;
        E[0] = 1.0; E[1] = e + 0.5;
        for (int i = 0; i < 6; i++) for (int j = 0; j <= i; j++) { double v = rnd() * 1e-6; E[4 + 6 * j + i] = v; E[4 + 6 * i + j] = v; }
        for (int i = 0; i < 6; i++) for (int j = 0; j <= i; j++) { double v = rnd(); E[40 + 6 * j + i] = v; E[40 + 6 * i + j] = v; }
        E[76] = E[80] = E[84] = 1.0;
    }
    for (int p = 0; p < np; p++) points[3 * p] = p;
    for (int k = 0; k < 6; k++) { pd[k] = 1; pl[k] = 0; }     // clamped root
    bc[18 * 3 + 6 + 5] = 10.0;                                // tip moment
    Problem P; P.np = np; P.ne = ne; P.start = start; P.stop = stop; P.elems = elems.data(); P.points = points.data();
    P.pd = pd.data(); P.pl = pl.data(); P.bc = bc.data(); P.force_scaling = 1024.0;
    Indices idx = system_indices(ne, start, stop, true);
    std::vector<Cnt> x(idx.nstates), r(idx.nstates);
    for (auto& v : x) v = Cnt(rnd());
    NullJac J;
    auto count = [&](auto&& f) { g_ops = Ops(); f(); return g_ops; };
    const int ie = 1, ip = 1;   // interior element / free interior point
    Ops props = count([&] { auto p = static_element_properties(P, idx, x.data(), ie); (void)p; });
    Ops eres = count([&] { static_element_residual(P, idx, x.data(), ie, r.data()); });
    Ops ejac = count([&] { static_element_jacobian(P, idx, x.data(), ie, J); });
    Ops pres = count([&] { static_point_residual(P, idx, x.data(), ip, r.data()); });
    Ops pjac = count([&] { static_point_jacobian(P, idx, x.data(), ip, J); });
    auto pr = [](const char* name, const Ops& o) { printf("%-44s add %5lld  mul %5lld  div %4lld  sqrt %2lld  flops %6lld\n", name, o.add, o.mul, o.div, o.sqr, o.flops()); };
    pr("static_element_properties", props);
    pr("static_element_residual!", eres);
    pr("static_element_jacobian!", ejac);
    pr("static_point_residual!", pres);
    pr("static_point_jacobian!", pjac);
    const long long fe_ref = eres.flops() + ejac.flops(), fe_fused = fe_ref - props.flops();
    const long long fp = pres.flops() + pjac.flops();
    printf("f_e (reference: residual + Jacobian, properties evaluated twice) = %lld\n", fe_ref);
    printf("f_e (fused: properties once)                                     = %lld\n", fe_fused);
    printf("f_p (residual + Jacobian, free interior point)                   = %lld\n", fp);
    for (int N : {100}) {
        const long long n = 12LL * N + 6, kl = 14, ku = 17;
        const long long lu = 2 * n * kl * (kl + ku) + 2 * n * (2 * kl + ku);
        printf("C2 (N = %d, n = %lld): assembly %lld (fused) + band LU/solve %lld = %lld flops per Newton iteration (SURVEY 8d formula)\n",
               N, n, N * fe_fused + (N + 1) * fp, lu, N * fe_fused + (N + 1) * fp + lu);
    }
    return 0;
}
