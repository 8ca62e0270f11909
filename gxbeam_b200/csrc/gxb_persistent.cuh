// Time-marching bookkeeping kernels and the persistent per-instance solver (one CTA runs an instance's whole Newton loop / time loop).
#pragma once
#include "gxb_kernels.cuh"

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
// newmark_output (src/analyses.jl:3461-3503): point rates of the current time level from the states and the initialisation terms,
// udot = 2/dt u - udot_init etc.; udot, θdot are zeroed where the displacement is prescribed (udot_udot mask).  rates: [NP][12] of one instance.
__device__ __forceinline__ void newmark_rates_point(const double* __restrict__ x, const double* __restrict__ bcS, const unsigned short* __restrict__ flags,
                                                    const double* __restrict__ paug, double c2dt, int NP, int n, int64_t b, int p, double* __restrict__ out) {
    const double* xp = x + b * n + 18 * p;
    const double* bc = bcS + ((size_t)b * 18) * NP + p;
    const double* pa = paug + ((size_t)b * 12) * NP + p;
    const unsigned fl = flags[p];
    for (int k = 0; k < 12; k++) {
        const bool pd = k < 6 && (fl & (1u << k));
        const double sv = pd ? bc[k * NP] : xp[k];
        out[12 * p + k] = pd ? 0.0 : c2dt * sv - pa[(size_t)k * NP];
    }
}
// rates_hist[b][slot][:] for every instance (slot 0 = the given initial rates)
__global__ void k_save_rates(const double* __restrict__ x, const double* __restrict__ bcS, const unsigned short* __restrict__ flags,
                             const double* __restrict__ paug, const double* __restrict__ rates0, double c2dt, int NP, int n, int64_t nsave, int64_t slot,
                             int64_t nbatch, double* __restrict__ rates_hist) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * NP) return;
    const int64_t b = t / NP; const int p = (int)(t % NP);
    double* out = rates_hist + ((size_t)b * nsave + slot) * NP * 12;
    if (slot == 0) { for (int k = 0; k < 12; k++) out[12 * p + k] = rates0[((size_t)b * NP + p) * 12 + k]; }
    else newmark_rates_point(x, bcS, flags, paug, c2dt, NP, n, b, p, out);
}
__global__ void k_step_end(const int* status, const int* iters, int* alive, int* steps_done, int* iters_total, int64_t nbatch) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= nbatch) return;
    if (!alive[b]) return;
    steps_done[b] += 1;
    iters_total[b] += iters[b];
    if (status[b] != ST_DONE) alive[b] = 0;
}


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
    double* hist; double* rates_hist; int64_t nsave, save_every;
};

// separate (non-inlined) functions: each phase gets its own register allocation instead of one allocation for the fused kernel
#ifndef GXB_PERSIST_MMA
#define GXB_PERSIST_MMA 1     // 1: tensor-core factorisation (gxb_lu_mma.cuh) inside the persistent kernel, 0: row-per-lane kernel
#endif
__device__ __noinline__ void persist_factor(const SolveArgs& S, int b, double* smem) { factor_instance<3, 4, GXB_PERSIST_MMA>(S, b, smem); }
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
            if (P.rates_hist && it % P.save_every == 0) {
                double* out = P.rates_hist + ((size_t)b * P.nsave + it / P.save_every) * NP * 12;
                for (int p = threadIdx.x; p < NP; p += blockDim.x) newmark_rates_point(P.Up.x, P.bcS, P.D.s.flags, P.paug, D.c2dt, NP, n, b, p, out);
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

