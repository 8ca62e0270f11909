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
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <condition_variable>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/gxbeam_b200.h"
#ifndef GXB_LU_MINBLOCKS
#define GXB_LU_MINBLOCKS 4
#endif
#include "gxb_kernels.cuh"
#include "gxb_persistent.cuh"

// ------------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
int gxb_fail_(int code, const std::string& msg) { return fail(code, msg); }      // for the other translation units of the library
#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            cudaGetLastError(); /* a non-sticky error (e.g. a failed cudaMalloc) must not be reported again by the next launch check */ \
            return fail(GXB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
        }                                                                                                    \
    } while (0)

struct gxb_ctx {
    int device = 0;
    size_t smem_optin = 0;            // cudaDevAttrMaxSharedMemoryPerBlockOptin of the device
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
    bool elems_compact = false;   // the last gxb_set_elements was a compact upload: d_elemS holds zeros in the mass / damping fields
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
    void* d_tot = nullptr; long long* h_tot = nullptr;          // device-side totals of iters / fcalls and their pinned mirror
    bool have_x0 = false;
    std::vector<int32_t> eig_singular;   // per instance: K's factorisation hit a zero pivot in the last gxb_eigen_arnoldi
    bool fallback_armed = false;      // a singular Jacobian was seen in this solve: k_singular_fallback follows every factorisation
    gxb_stats stats{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr, evC = nullptr;   // evC: end of the last host->device copy
};

// Per-call device temporaries: freed on every exit path of the entry point that owns them (CK returns early on CUDA errors)
struct DevTemps {
    std::vector<void*> ptrs;
    template <class T> cudaError_t alloc(T** p, size_t bytes) { cudaError_t e = cudaMalloc((void**)p, bytes); if (e == cudaSuccess) ptrs.push_back(*p); return e; }
    void release(void* p) { for (auto& q : ptrs) if (q == p) q = nullptr; }      // ownership moves elsewhere
    ~DevTemps() { for (void* q : ptrs) if (q) cudaFree(q); }
};
// context state that an entry point sets for the duration of one call
struct ModeGuard {
    gxb_ctx* c;
    ~ModeGuard() { c->mode = 0; c->damping = 0; }
};

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

// cudaFuncAttributeMaxDynamicSharedMemorySize is per function AND per device, and setting it replaces the previous value: every kernel
// that uses dynamic shared memory is raised to the device's opt-in maximum once per device, so that contexts with different chain lengths
// (or on different GPUs) cannot undercut each other.
static int configure_kernels(int device, size_t* optin_out) {
    static std::mutex mu;
    static std::map<int, size_t> done;
    std::lock_guard<std::mutex> lk(mu);
    auto it = done.find(device);
    if (it != done.end()) { *optin_out = it->second; return GXB_OK; }
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    // the limit covers static + dynamic shared memory of the kernel
#define GXB_SMEM_MAX(k) do { cudaFuncAttributes fa_; CK(cudaFuncGetAttributes(&fa_, k)); \
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa_.sharedSizeBytes)); } while (0)
    GXB_SMEM_MAX(k_assemble<true>); GXB_SMEM_MAX(k_assemble<false>); GXB_SMEM_MAX(k_assemble_residual_128); GXB_SMEM_MAX(k_residual_update_128);
    GXB_SMEM_MAX((k_assemble_dynamic<true, 0>)); GXB_SMEM_MAX((k_assemble_dynamic<true, 1>));
    GXB_SMEM_MAX(k_assemble_dynamic_split<0>); GXB_SMEM_MAX(k_assemble_dynamic_split<1>);
    GXB_SMEM_MAX((k_assemble_dynamic<false, 0>)); GXB_SMEM_MAX((k_assemble_dynamic<false, 1>));
    GXB_SMEM_MAX(k_assemble_initial<true>); GXB_SMEM_MAX(k_assemble_initial<false>);
    GXB_SMEM_MAX((k_factor_solve<2, 2>)); GXB_SMEM_MAX((k_factor_solve<3, 4>)); GXB_SMEM_MAX((k_factor_solve<4, 3>));
    GXB_SMEM_MAX((k_factor_solve_mma<2, 2>)); GXB_SMEM_MAX((k_factor_solve_mma<3, 4>)); GXB_SMEM_MAX((k_resolve_mma<3, 4>));
    GXB_SMEM_MAX((k_factor_solve_two<3, 4, 3, 2>)); GXB_SMEM_MAX((k_factor_solve_two<2, 2, 2, 1>));
    GXB_SMEM_MAX(k_factor_solve_part<2>); GXB_SMEM_MAX(k_factor_solve_part<3>); GXB_SMEM_MAX(k_factor_solve_part<4>);
    GXB_SMEM_MAX(k_factor_solve_part<6>); GXB_SMEM_MAX(k_factor_solve_part<8>);
    GXB_SMEM_MAX(k_arnoldi_orth);
    GXB_SMEM_MAX(k_persistent_dynamic<0>); GXB_SMEM_MAX(k_persistent_dynamic<1>); GXB_SMEM_MAX(k_persistent_dynamic<2>);
#undef GXB_SMEM_MAX
    done[device] = (size_t)optin;
    *optin_out = (size_t)optin;
    return GXB_OK;
}

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
    { int rc = configure_kernels(device, &c->smem_optin); if (rc) { delete c; return rc; } }
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&c->d_counter, 2 * sizeof(int)));          // [0] unfinished instances of the round, [1] sticky: a singular Jacobian was seen
    CK(cudaMallocHost(&c->h_counter, 2 * sizeof(int)));
    CK(cudaMalloc(&c->d_tot, 16)); CK(cudaMallocHost(&c->h_tot, 16));
    CK(cudaEventCreate(&c->ev0)); CK(cudaEventCreate(&c->ev1)); CK(cudaEventCreate(&c->evA)); CK(cudaEventCreate(&c->evB)); CK(cudaEventCreateWithFlags(&c->evC, cudaEventDisableTiming));
    *out = c;
    return GXB_OK;
}

void gxb_destroy(gxb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_batch(c);
    cudaFree(c->d_flags); cudaFree(c->d_brow); cudaFree(c->d_bcol); cudaFree(c->d_counter); cudaFreeHost(c->h_counter); cudaFree(c->d_tot); cudaFreeHost(c->h_tot);
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

static int set_batch_impl(gxb_ctx* c, int64_t nbatch);
int gxb_set_batch(gxb_ctx* c, int64_t nbatch) {
    int rc = require_chain(c, "gxb_set_batch"); if (rc) return rc;
    if (nbatch < 1) return fail(GXB_ERR_INVALID, "gxb_set_batch: nbatch < 1");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    free_batch(c);
    rc = set_batch_impl(c, nbatch);
    if (rc) {
        // a batch that does not fit (or any other failure half-way): give back what was allocated, leave the context without a batch but
        // usable, and clear CUDA's last-error slot (a failed cudaMalloc stays there and would be reported by the next launch check)
        const std::string msg = g_err;
        free_batch(c);
        c->nbatch = 0;
        cudaGetLastError();
        g_err = msg;
    }
    return rc;
}
static int set_batch_impl(gxb_ctx* c, int64_t nbatch) {
    int rc;
    c->nbatch = nbatch;
    const size_t nb = (size_t)nbatch, n = (size_t)c->n, N = (size_t)c->ne, P = (size_t)c->np;
    if ((rc = dalloc(&c->d_elemS, nb * GXB_ELEM_NF * N))) return rc;
    if ((rc = dalloc(&c->d_bcS, nb * 18 * P))) return rc;
    if ((rc = dalloc(&c->d_fs, nb))) return rc;
    if ((rc = dalloc(&c->d_grav, nb * 3))) return rc;
    if ((rc = dalloc(&c->d_x, nb * n)) || (rc = dalloc(&c->d_xt, nb * n)) || (rc = dalloc(&c->d_F, nb * n)) || (rc = dalloc(&c->d_Ft, nb * n)) ||
        (rc = dalloc(&c->d_p, nb * n))) return rc;
    // U scratch: the row-per-lane factorisation stores n rows of UW doubles, the tensor-core one n rounded up to 8 rows of
    // MmaShape::UW doubles
    size_t u_per_instance = n * c->UW;
    u_per_instance = std::max(u_per_instance, ((n + 7) & ~(size_t)7) * (size_t)(c->kind == 0 ? MmaShape<2, 2>::UW : MmaShape<3, 4>::UW));
    if (c->kind == 0) u_per_instance = std::max(u_per_instance, n * (size_t)PartShape::UW);
    if ((rc = dalloc(&c->d_J, nb * n * c->W)) || (rc = dalloc(&c->d_U, nb * u_per_instance))) return rc;
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
        c->elems_compact = true;
        return wait_copy(c);
    }
    c->elems_compact = false;
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
        if (c->has_grav && (c->skip_mass || c->elems_compact)) {
            c->has_grav = 0;
            return fail(GXB_ERR_INVALID, "gxb_set_body: non-zero gravity, but the elements were uploaded under gxb_hint_no_gravity (the element mass was "
                                         "not uploaded): switch the hint off and call gxb_set_elements again first");
        }
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
        // exchange records for every thread, Jacobian tiles for the N + 1 stations only: a smaller footprint leaves more of the SM's
        // 256 KB as L1 for the register spills of this 255-register kernel
        size_t sm = ((size_t)bs * GXB_XCH + (want_j ? (size_t)c->np * GXB_TILE : 0)) * 8;
        if (want_j) {
            k_assemble<true><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(A);
        } else if (bs <= 128) k_assemble_residual_128<<<(unsigned)c->nbatch, bs, sm, c->stream>>>(A);
        else k_assemble<false><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(A);
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
            // small batches: three CTAs per instance, one per group of row passes (GXB_DYN_SPLIT=0 / 1 forces it off / on)
            static const int split_env = []() { const char* e = getenv("GXB_DYN_SPLIT"); return e ? atoi(e) : -1; }();
            // measured on config 4's topology (B200): 64 instances 0.239 -> 0.215 s per 400 steps with three CTAs per instance; 256 instances
            // 0.294 -> 0.304 s, 512 instances 1.21 -> 1.59 s (three-way) / 1.37 s (two-way): the repeated core evaluation outweighs the
            // shorter chain as soon as the extra CTAs compete for SMs -- tiny batches only
            const int ways = split_env > 0 ? (split_env >= 3 ? 3 : 2) : (c->nbatch <= 128 ? 3 : 1);
            const bool split = c->mode != 2 && split_env != 0 && ways > 1;
            const dim3 grid3((unsigned)c->nbatch, ways);
            if (split && c->mode == 0) k_assemble_dynamic_split<0><<<grid3, bs, sm, c->stream>>>(D);
            else if (split && c->mode == 1) k_assemble_dynamic_split<1><<<grid3, bs, sm, c->stream>>>(D);
            else if (c->mode == 0) k_assemble_dynamic<true, 0><<<(unsigned)c->nbatch, bs, sm, c->stream>>>(D);
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
    {
        // tensor-core factorisation (gxb_lu_mma.cuh).  GXB_LU_MMA is a mask: bit 0 static systems, bit 1 dynamic systems; 0 keeps the
        // row-per-lane kernel everywhere (A/B runs)
        static const int mma_mask = []() { const char* e = getenv("GXB_LU_MMA"); return e ? atoi(e) : GXB_LU_MMA; }();
        if (mma_mask & (KLB == 2 ? 1 : 2)) {
            using M = MmaShape<KLB, KUB>;
            size_t smm = (size_t)wpc * M::SM_DOUBLES * 8;
            k_factor_solve_mma<KLB, KUB><<<grid1(c->nbatch, wpc), 32 * wpc, smm, c->stream>>>(A);
            CK(cudaGetLastError());
            return GXB_OK;
        }
    }
    k_factor_solve<KLB, KUB><<<grid1(c->nbatch, wpc), 32 * wpc, sm, c->stream>>>(A);
    CK(cudaGetLastError());
    return GXB_OK;
}
template <int KLB, int KUB> static int launch_factor_t(gxb_ctx* c, const SolveArgs& A);
static int launch_factor_body(gxb_ctx* c, const gxb_opts* o);
// Partitioned factorisation (gxb_lu_part.cuh): how many warps per instance.  One warp per instance (the tensor-core kernel) fills the GPU
// from ~2400 instances on; for small batches the chain is cut into P partitions.  The partition kernel runs at 170 registers = 12 warps
// per SM, and the whole batch should be resident in ONE wave: P = 12 * #SMs / nbatch, rounded down to one of 8, 6, 4, 3, 2.
// opts.lu_parts / GXB_LU_PARTS: 0 auto, 1 off, 2 / 3 / 4 / 6 / 8 forced.
static int choose_parts(const gxb_ctx* c, const gxb_opts* o) {
    if (c->kind != 0) return 1;
    static const int forced_env = []() { const char* e = getenv("GXB_LU_PARTS"); return e ? atoi(e) : 0; }();
    const int forced = (o && o->lu_parts > 0) ? o->lu_parts : forced_env;
    const int stations = (int)c->np;
    static const int allowed[5] = {8, 6, 4, 3, 2};
    int P = 1;
    const int64_t want = forced > 0 ? forced : ((int64_t)c->sm_count * 12) / c->nbatch;
    for (int cand : allowed) if (cand <= want) { P = cand; break; }
    while (P > 1 && stations < 4 * P) P = P == 8 ? 6 : (P == 6 ? 4 : (P == 4 ? 3 : (P == 3 ? 2 : 1)));
    return P;
}
// Cuts in front of a point (block 2 m).  The first partition carries no spike, so its block step is cheaper (measured at 512 instances,
// P = 3: 4850 vs 5750 cycles): it gets proportionally more stations.
static PartPlan make_plan(const gxb_ctx* c, int P) {
    PartPlan pl{};
    pl.P = P;
    const double w0 = 1.19;
    const double unit = (double)c->np / (w0 + (P - 1));
    double acc = w0 * unit;
    pl.R0[0] = 0;
    for (int q = 1; q < P; q++) {
        int m = (int)(acc + 0.5);
        const int lo = pl.R0[q - 1] / 2 + 4, hi = (int)c->np - 4 * (P - q);
        m = std::max(lo, std::min(hi, m));
        pl.R0[q] = 2 * m;
        acc += unit;
    }
    pl.R0[P] = (int)(c->n / 6);
    return pl;
}
// What follows a factorisation launch inside a Newton round: the singular-Jacobian fallback for the instances the factorisation parked
// (launched only once the host has seen the sticky flag: the instance waits at most one counter read-back), and the exact dφ0.
static int after_factor(gxb_ctx* c, const gxb_opts* o, const SolveArgs& A, int direction_only) {
    if (direction_only) return GXB_OK;
    if (c->fallback_armed) {
        FallbackArgs Fa{};
        Fa.S = A; Fa.W = c->W; Fa.KLB = c->KLB;
        const int KUB = c->W / 6 - 1 - c->KLB;
        Fa.u_stride = (size_t)c->n * (size_t)(6 * (c->KLB + KUB) + 6);
        k_singular_fallback<<<(unsigned)c->nbatch, 256, 0, c->stream>>>(Fa);
        CK(cudaGetLastError());
        c->stats.launches++;
    }
    if (o->exact_dphi0) {
        DphiArgs D{};
        D.S = A; D.W = c->W; D.KLB = c->KLB;
        k_exact_dphi0<<<(unsigned)c->nbatch, 128, 0, c->stream>>>(D);
        CK(cudaGetLastError());
        c->stats.launches++;
    }
    return GXB_OK;
}
static bool mma_enabled(const gxb_ctx* c) {
    static const int mma_mask = []() { const char* e = getenv("GXB_LU_MMA"); return e ? atoi(e) : GXB_LU_MMA; }();
    return (mma_mask & (c->kind == 0 ? 1 : 2)) != 0;
}
static int launch_factor(gxb_ctx* c, const gxb_opts* o, int direction_only, double sign = -1.0, double* Lst = nullptr, size_t l_stride = 0) {
    if (c->nbody > 0 && !direction_only) return launch_factor_body(c, o);
    SolveArgs A{};
    A.sign = sign; A.Lst = Lst; A.l_stride = l_stride;
    A.nb = (int)(c->n / 6); A.n = (int)c->n; A.nbatch = c->nbatch;
    A.J = c->d_J; A.F = direction_only ? c->d_Ft : c->d_F; A.x = c->d_x; A.U = c->d_U; A.p = c->d_p; A.xt = c->d_xt; A.ns = c->ns;
    A.maxstep = o->maxstep; A.linear = o->linear; A.direction_only = direction_only;
    A.fallback = direction_only ? 0 : 1; A.counter = c->d_counter; A.singular_seen = c->d_counter + 1;
    // dynamic systems, small batches: two warps per instance from both ends (gxb_lu_two.cuh).  GXB_LU_TWO = 0 / 1 forces it off / on;
    // opts.lu_parts = 1 switches it off like the partitioned kernel, = 2 forces it (static systems: the spike-free P = 2 variant).
    {
        static const int two_env = []() { const char* e = getenv("GXB_LU_TWO"); return e ? atoi(e) : -1; }();
        const int forced = o ? o->lu_parts : 0;
        // automatic choice: dynamic systems up to 740 instances (194 registers: 10 warps per SM, one wave); static systems between 3 and
        // 6 instances per SM, where it beats both the three-partition kernel (512 instances: 1.56 vs 1.60 ms per solve, no spike work) and
        // one warp per instance (800 instances: 2.05 vs 2.55 ms; at 1024 the tensor-core kernel is back in front, 2.70 vs 2.76 ms);
        // smaller static batches take more partitions (gxb_lu_part.cuh)
        const bool auto_two = c->kind == 1 ? 2 * c->nbatch <= (int64_t)10 * c->sm_count
                                           : (c->nbatch > (int64_t)3 * c->sm_count && c->nbatch <= (int64_t)6 * c->sm_count);
        bool two;
        if (forced == 2) two = true;                       // explicit
        else if (forced != 0) two = false;                 // another explicit choice (1 = one warp per instance, 3 .. 8 = partitions)
        else if (two_env == 0) two = false;
        else if (two_env > 0) two = c->kind == 1 || auto_two;
        else two = auto_two;
        two = two && c->np >= 6 && !Lst;                   // the eigen operator records the factorisation: tensor-core kernel only
        if (two) {
            const int m = (int)((c->np + 1) / 2);                 // cut in front of point m: halves of (almost) equal numbers of block steps
            if (c->kind == 1) {
                using T2 = TwoShape<3, 4, 3, 2>;
                k_factor_solve_two<3, 4, 3, 2><<<(unsigned)c->nbatch, 64, (size_t)T2::CTA_DOUBLES * 8, c->stream>>>(A, 3 * m);
            } else {
                using T2 = TwoShape<2, 2, 2, 1>;
                k_factor_solve_two<2, 2, 2, 1><<<(unsigned)c->nbatch, 64, (size_t)T2::CTA_DOUBLES * 8, c->stream>>>(A, 2 * m);
            }
            CK(cudaGetLastError());
            c->stats.launches++; c->stats.n_factor++;
            return after_factor(c, o, A, direction_only);
        }
    }
    const int P = choose_parts(c, o);
    if (P > 1) {
        const PartPlan plan = make_plan(c, P);
        const size_t sm = (size_t)PartShape::cta_doubles(P) * 8;
        if (P == 2) k_factor_solve_part<2><<<(unsigned)c->nbatch, 64, sm, c->stream>>>(A, plan);
        else if (P == 3) k_factor_solve_part<3><<<(unsigned)c->nbatch, 96, sm, c->stream>>>(A, plan);
        else if (P == 4) k_factor_solve_part<4><<<(unsigned)c->nbatch, 128, sm, c->stream>>>(A, plan);
        else if (P == 6) k_factor_solve_part<6><<<(unsigned)c->nbatch, 192, sm, c->stream>>>(A, plan);
        else k_factor_solve_part<8><<<(unsigned)c->nbatch, 256, sm, c->stream>>>(A, plan);
        CK(cudaGetLastError());
        c->stats.launches++; c->stats.n_factor++;
        return after_factor(c, o, A, direction_only);
    }
    int rc = c->kind == 0 ? launch_factor_t<2, 2>(c, A) : launch_factor_t<3, 4>(c, A);
    if (rc) return rc;
    c->stats.launches++; c->stats.n_factor++;
    return after_factor(c, o, A, direction_only);
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
static UpdateArgs make_update_args(gxb_ctx* c, const gxb_opts* o, int mode);
// fused residual + update launch (static chains that fit a 128-thread CTA)
static bool can_fuse_update(const gxb_ctx* c) {
    static const bool on = !(getenv("GXB_FUSE_UPDATE") && atoi(getenv("GXB_FUSE_UPDATE")) == 0);
    // measured (B200): 512 instances 1.607 -> 1.595 ms per solve; 4096 instances 6.96 -> 7.13 ms (the separate update kernel overlaps the
    // other lane's kernels better): small batches only
    return on && c->kind == 0 && c->np <= 128 && c->nbatch * 2 <= (int64_t)12 * c->sm_count;
}
static int launch_residual_update(gxb_ctx* c, const gxb_opts* o, const StaticArgs& A) {
    const int bs = (int)((c->np + 31) / 32) * 32;
    const size_t sm = (size_t)bs * GXB_XCH * 8;
    k_residual_update_128<<<(unsigned)c->nbatch, bs, sm, c->stream>>>(A, make_update_args(c, o, 1));
    CK(cudaGetLastError());
    c->stats.launches++; c->stats.n_assemble++; c->stats.n_update++;
    return GXB_OK;
}
static UpdateArgs make_update_args(gxb_ctx* c, const gxb_opts* o, int mode) {
    UpdateArgs A{};
    A.n = (int)c->n; A.mode = mode; A.x = c->d_x; A.xt = c->d_xt; A.F = c->d_F; A.Ft = c->d_Ft; A.p = c->d_p; A.ns = c->ns;
    A.ftol = o->ftol; A.c1 = o->c1; A.rho_hi = o->rho_hi; A.rho_lo = o->rho_lo; A.iterations = o->iterations; A.ls_iterations = 1000;
    A.counter = c->d_counter;
    return A;
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
            if (residual_first && can_fuse_update(c)) {
                prof.begin(); if ((rc = launch_residual_update(c, &o, A))) return rc; prof.end(c->stats.ms_assemble);
            } else {
                prof.begin(); if ((rc = launch_assemble(c, A, !residual_first))) return rc; prof.end(c->stats.ms_assemble);
                prof.begin(); if ((rc = launch_update(c, &o, 1))) return rc; prof.end(c->stats.ms_update);
            }
            if (residual_first) { prof.begin(); if ((rc = launch_assemble(c, A_next, true))) return rc; prof.end(c->stats.ms_assemble); }
            c->stats.rounds += 1;
        }
        CK(cudaMemcpyAsync(c->h_counter, c->d_counter, 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (c->h_counter[1] && !c->fallback_armed) {
            // first singular Jacobian of this solve: from now on k_singular_fallback follows every factorisation; the instances that are
            // parked in ST_SINGULAR right now get their (J'J + λI) step before the next round
            c->fallback_armed = true;
            SolveArgs A{};
            A.nb = (int)(c->n / 6); A.n = (int)c->n; A.nbatch = c->nbatch; A.sign = -1.0;
            A.J = c->d_J; A.F = c->d_F; A.x = c->d_x; A.U = c->d_U; A.p = c->d_p; A.xt = c->d_xt; A.ns = c->ns;
            A.maxstep = o.maxstep; A.linear = o.linear; A.fallback = 1;
            gxb_opts o2 = o; o2.exact_dphi0 = 0;
            if ((rc = after_factor(c, &o2, A, 0))) return rc;
            continue;
        }
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
    c->fallback_armed = false;
    CK(cudaMemsetAsync(c->d_counter, 0, 8, c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    if (!c->have_x0) CK(cudaMemsetAsync(c->d_x, 0, nb * n * 8, c->stream));   // reset_state (system.jl:383-386)
    CK(cudaMemcpyAsync(c->d_xt, c->d_x, nb * n * 8, cudaMemcpyDeviceToDevice, c->stream));
    k_fill_int<<<grid1(nb, 256), 256, 0, c->stream>>>(c->ns.status, ST_TRIAL, nb);
    CK(cudaMemsetAsync(c->ns.failcode, 0, nb * 4, c->stream)); CK(cudaMemsetAsync(c->ns.iters, 0, nb * 4, c->stream));
    CK(cudaMemsetAsync(c->ns.fcalls, 0, nb * 4, c->stream));
    if (use_persistent(c, o, false)) rc = launch_persistent(c, o, mode, nullptr);
    else {
        // static systems: two rounds between reads of the unfinished-instance counter (a surplus round costs four early-exit launches, a
        // host synchronisation ~20 us: 512 instances 1.557 -> 1.530 ms, 4096 instances 6.97 -> 6.89 ms; GXB_BURST overrides).  Steady
        // solves keep one round per read: their straggler hand-off wants the count after every round.
        static const int burst_env = []() { const char* e = getenv("GXB_BURST"); return e ? atoi(e) : 0; }();
        rc = newton_rounds(c, o, prof, burst_env > 0 ? burst_env : (c->kind == 0 ? 2 : 1));
    }
    if (rc) { c->mode = 0; return rc; }
    c->mode = 0;
    // totals of Newton iterations / residual evaluations: summed on the device, one 16-byte read-back with the end event
    {
        long long* d_tot = reinterpret_cast<long long*>(c->d_tot);
        CK(cudaMemsetAsync(d_tot, 0, 16, c->stream));
        k_sum_counts<<<grid1(nb, 256), 256, 0, c->stream>>>(c->ns.iters, c->ns.fcalls, d_tot, c->nbatch);
        CK(cudaGetLastError());
        c->stats.launches++;
        CK(cudaMemcpyAsync(c->h_tot, d_tot, 16, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->stats.ms_total = ms;
    c->stats.newton_iters += c->h_tot[0]; c->stats.resid_evals += c->h_tot[1];
    c->have_x0 = false;
    return GXB_OK;
}

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
struct MarchBuffers { int64_t nt = 0; const double* d_tvec = nullptr; int* alive = nullptr; int* steps = nullptr; int* itot = nullptr; double* hist = nullptr; double* rates_hist = nullptr; int64_t nsave = 0, save_every = 1; };
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
    size_t sm = std::max(((size_t)bs * GXB_XCH + (size_t)chunk * GXB_DTILE) * 8,
                         (size_t)std::max(LuShape<3, 4>::SM_DOUBLES, MmaShape<3, 4>::SM_DOUBLES) * 8);
    P.S.sign = -1.0; P.S.nb = (int)(c->n / 6); P.S.n = (int)c->n; P.S.nbatch = c->nbatch;
    P.S.J = c->d_J; P.S.F = c->d_F; P.S.x = c->d_x; P.S.U = c->d_U; P.S.p = c->d_p; P.S.xt = c->d_xt; P.S.ns = c->ns;
    P.S.maxstep = o.maxstep; P.S.linear = 0; P.S.direction_only = 0;
    P.Up.n = (int)c->n; P.Up.x = c->d_x; P.Up.xt = c->d_xt; P.Up.F = c->d_F; P.Up.Ft = c->d_Ft; P.Up.p = c->d_p; P.Up.ns = c->ns;
    P.Up.ftol = o.ftol; P.Up.c1 = o.c1; P.Up.rho_hi = o.rho_hi; P.Up.rho_lo = o.rho_lo; P.Up.iterations = o.iterations; P.Up.ls_iterations = 1000;
    P.Up.counter = nullptr;
    if (mb) {
        P.nt = mb->nt; P.tvec = mb->d_tvec; P.bcS = c->d_bcS; P.series = c->d_series; P.ser_point = c->d_ser_point; P.ser_field = c->d_ser_field;
        P.K = c->ser_K; P.ser_nt = c->ser_nt; P.paug = c->d_paug; P.rates0 = c->d_rates0;
        P.alive = mb->alive; P.steps = mb->steps; P.itot = mb->itot; P.hist = mb->hist; P.rates_hist = mb->rates_hist; P.nsave = mb->nsave; P.save_every = mb->save_every;
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
                    int64_t save_every, double* x_hist, int32_t* steps_done, int32_t* iters_total, int32_t* converged, double* rates_hist) {
    NEED_BATCH("gxb_newmark_run");
    if (c->kind != 1) return fail(GXB_ERR_INVALID, "gxb_newmark_run: context topology is not kind=1 (DynamicSystem)");
    if (nt < 2 || !tvec) return fail(GXB_ERR_INVALID, "gxb_newmark_run: need nt >= 2 and tvec");
    if ((x0 == nullptr) != (rates0 == nullptr)) return fail(GXB_ERR_INVALID, "gxb_newmark_run: give both x0 and the initial rates, or neither");
    if (!x0 && !c->have_initial) return fail(GXB_ERR_INVALID, "gxb_newmark_run: no initial state (pass x0 and rates0, or call gxb_initial_condition first)");
    gxb_opts o; if (opts) o = *opts; else gxb_default_opts(&o);
    if (c->ser_K && c->ser_nt < nt) return fail(GXB_ERR_INVALID, "gxb_newmark_run: the bc series has fewer time levels than tvec");
    // every argument is validated before any allocation or state change
    for (int64_t it = 1; it < nt; it++)
        if (!(tvec[it] - tvec[it - 1] > 0.0)) return fail(GXB_ERR_INVALID, "gxb_newmark_run: tvec must be strictly increasing");
    DevTemps tmp;
    ModeGuard guard{c};                     // c->mode / c->damping are reset on every exit path
    c->damping = o.structural_damping ? 1 : 0;
    // linear=true: one lsolve! per step, about the current state (update_linearization) or about the initial state (analyses.jl:2741-2755)
    double* d_xlin = nullptr;
    if (save_every < 1) save_every = 1;
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n, P = (size_t)c->np;
    const int64_t nsave = (nt - 1) / save_every + 1;      // time levels 0, save_every, 2 save_every, ...
    c->stats = gxb_stats{};
    Prof prof{c, o.profile != 0};
    int rc;
    int *d_alive = nullptr, *d_steps = nullptr, *d_itot = nullptr; double* d_hist = nullptr;
    CK(tmp.alloc(&d_alive, nb * 4)); CK(tmp.alloc(&d_steps, nb * 4)); CK(tmp.alloc(&d_itot, nb * 4));
    if (x_hist) CK(tmp.alloc(&d_hist, nb * nsave * n * 8));
    double* d_rhist = nullptr;
    if (rates_hist) CK(tmp.alloc(&d_rhist, nb * nsave * P * 12 * 8));
    k_fill_int<<<grid1(nb, 256), 256, 0, c->stream>>>(d_alive, 1, nb);
    k_fill_int<<<grid1(nb, 256), 256, 0, c->stream>>>(d_steps, 1, nb);
    CK(cudaMemsetAsync(d_itot, 0, nb * 4, c->stream));
    CK(cudaMemsetAsync(c->ns.failcode, 0, nb * 4, c->stream));
    if (x0) {
        CK(cudaMemcpyAsync(c->d_x, x0, nb * n * 8, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_rates0, rates0, nb * P * 12 * 8, cudaMemcpyHostToDevice, c->stream));
    }
    c->have_initial = false;
    c->fallback_armed = false;
    CK(cudaMemsetAsync(c->d_counter, 0, 8, c->stream));
    if (o.linear && !o.update_linearization) {
        CK(tmp.alloc(&d_xlin, nb * n * 8));
        CK(cudaMemcpyAsync(d_xlin, c->d_x, nb * n * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    CK(cudaMemsetAsync(c->d_paug, 0, nb * 12 * P * 8, c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    if (d_hist) k_save_state<<<grid1(nb * n, 256), 256, 0, c->stream>>>(c->d_x, d_hist, (int)n, nsave, 0, c->nbatch);
    if (d_rhist) k_save_rates<<<grid1(nb * P, 128), 128, 0, c->stream>>>(c->d_x, c->d_bcS, c->d_flags, c->d_paug, c->d_rates0, 0.0, (int)P, (int)n, nsave, 0, c->nbatch, d_rhist);
    c->mode = 1;
    double* d_tvec = nullptr;
    const bool persistent = use_persistent(c, o, true);
    if (persistent) {
        CK(tmp.alloc(&d_tvec, nt * 8));
        CK(cudaMemcpyAsync(d_tvec, tvec, nt * 8, cudaMemcpyHostToDevice, c->stream));
        MarchBuffers mb; mb.nt = nt; mb.d_tvec = d_tvec; mb.alive = d_alive; mb.steps = d_steps; mb.itot = d_itot; mb.hist = d_hist; mb.rates_hist = d_rhist; mb.nsave = nsave; mb.save_every = save_every;
        if ((rc = launch_persistent(c, o, 1, &mb))) return rc;
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
        if ((rc = newton_rounds(c, o, prof, 2))) return rc;
        k_step_end<<<grid1(nb, 256), 256, 0, c->stream>>>(c->ns.status, c->ns.iters, d_alive, d_steps, d_itot, c->nbatch);
        c->stats.launches += 1;
        if (d_hist && it % save_every == 0) { k_save_state<<<grid1(nb * n, 256), 256, 0, c->stream>>>(c->d_x, d_hist, (int)n, nsave, it / save_every, c->nbatch); c->stats.launches += 1; }
        if (d_rhist && it % save_every == 0) {
            k_save_rates<<<grid1(nb * P, 128), 128, 0, c->stream>>>(c->d_x, c->d_bcS, c->d_flags, c->d_paug, c->d_rates0, c->c2dt, (int)P, (int)n, nsave, it / save_every, c->nbatch, d_rhist);
            c->stats.launches += 1;
        }
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
    if (rates_hist) CK(cudaMemcpy(rates_hist, d_rhist, nb * nsave * P * 12 * 8, cudaMemcpyDeviceToHost));
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

int gxb_assembly_state(gxb_ctx* c, const double* rates, double* points, double* elements) {
    NEED_BATCH("gxb_assembly_state");
    if (!points || !elements) return fail(GXB_ERR_INVALID, "gxb_assembly_state: null output");
    const size_t nb = (size_t)c->nbatch, P = (size_t)c->np, N = (size_t)c->ne;
    DevTemps tmp;
    double *d_rates = nullptr, *d_pts = nullptr, *d_el = nullptr;
    CK(tmp.alloc(&d_pts, nb * P * 30 * 8)); CK(tmp.alloc(&d_el, nb * N * 30 * 8));
    if (rates) { CK(tmp.alloc(&d_rates, nb * P * 12 * 8)); CK(cudaMemcpyAsync(d_rates, rates, nb * P * 12 * 8, cudaMemcpyHostToDevice, c->stream)); }
    StateArgs A{};
    A.kind = c->kind; A.N = (int)N; A.n = (int)c->n; A.nbatch = c->nbatch;
    A.x = c->d_x; A.bcS = c->d_bcS; A.flags = c->d_flags; A.fs = c->d_fs; A.rates = d_rates; A.points = d_pts; A.elems = d_el;
    k_assembly_state<<<grid1(nb * P, 128), 128, 0, c->stream>>>(A);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(points, d_pts, nb * P * 30 * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(elements, d_el, nb * N * 30 * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->stats.launches += 1;
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

static int eigen_arnoldi_impl(gxb_ctx* c, const gxb_opts* opts, const double* x, int32_t m, double* H, double* V, bool transposed);
int gxb_eigen_arnoldi(gxb_ctx* c, const gxb_opts* opts, const double* x, int32_t m, double* H, double* V) {
    return eigen_arnoldi_impl(c, opts, x, m, H, V, false);
}
int gxb_eigen_arnoldi_left(gxb_ctx* c, const gxb_opts* opts, const double* x, int32_t m, double* H, double* V) {
    return eigen_arnoldi_impl(c, opts, x, m, H, V, true);
}
static int eigen_arnoldi_impl(gxb_ctx* c, const gxb_opts* opts, const double* x, int32_t m, double* H, double* V, bool transposed) {
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
    DevTemps tmp;
    ModeGuard guard{c};
    cudaFree(c->d_Vk); c->d_Vk = nullptr; c->vk_m = 0;
    CK(tmp.alloc(&d_V, nb * (size_t)(m + 1) * n * 8));
    CK(tmp.alloc(&d_H, nb * (size_t)(m + 1) * m * 8));
    CK(cudaMemsetAsync(d_H, 0, nb * (size_t)(m + 1) * m * 8, c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    // K = steady Jacobian at x (steady_jacobian!, analyses.jl:1358), M = mass matrix at x (mass_matrix!, analyses.jl:1359)
    CK(cudaMemcpyAsync(c->d_x, x, nb * n * 8, cudaMemcpyHostToDevice, c->stream));
    StaticArgs A = make_static_args(c, &o, c->d_x, c->d_F, true, 0, false);
    if ((rc = launch_assemble(c, A, true))) return rc;
    if ((rc = assemble_mass_matrix(c, o, c->d_x))) return rc;
    k_arnoldi_start<<<(unsigned)nb, 256, 0, c->stream>>>(d_V, (int)n, m);
    // `Kfact = lu(K)` once (analyses.jl:948): the tensor-core factorisation records its multipliers in d_L during the first application;
    // every further application of K^-1 replays them on the new right-hand side and back-substitutes (k_resolve_mma).  With the
    // row-per-lane kernel (GXB_LU_MMA without bit 1) L is not recorded and every application factorises again.
    double* d_L = nullptr;
    const size_t l_stride = (size_t)(((n + 7) & ~(size_t)7) / 8) * GXB_MMA_LSTEP;
    // Left eigenvectors are the right eigenvectors of the transposed pencil (K', M'): the same Arnoldi process with K' and M' in
    // row-block storage of swapped block bandwidths (4 / 3); K' goes through the row-per-lane factorisation (its window needs 30 row
    // slots; the tensor-core kernel's fragment layout has no shape for it), which factorises on every application.
    double *d_KT = nullptr, *d_MT = nullptr;
    if (transposed) {
        CK(tmp.alloc(&d_KT, nb * n * c->W * 8)); CK(tmp.alloc(&d_MT, nb * n * c->W * 8));
        const int KUB = c->W / 6 - 1 - c->KLB;
        k_rowblock_transpose<<<grid1(nb * n * c->W, 256), 256, 0, c->stream>>>(c->d_J, d_KT, (int)n, c->W, c->KLB, KUB, c->nbatch);
        k_rowblock_transpose<<<grid1(nb * n * c->W, 256), 256, 0, c->stream>>>(c->d_M, d_MT, (int)n, c->W, c->KLB, KUB, c->nbatch);
        CK(cudaGetLastError());
        c->stats.launches += 2;
    }
    if (mma_enabled(c) && !transposed) CK(tmp.alloc(&d_L, nb * l_stride * 8));
    CK(cudaMemsetAsync(c->ns.failcode, 0, nb * 4, c->stream));
    for (int j = 0; j < m; j++) {
        // w = K^{-1} (M v_j)   (f! of analyses.jl:949)
        if (transposed) {
            const int KUB = c->W / 6 - 1 - c->KLB;
            k_rowblock_spmv<<<grid1(nb * n * 8, 256), 256, 0, c->stream>>>(d_MT, d_V + (size_t)j * n, c->d_Ft, (int)n, c->W, KUB, c->nbatch, (int64_t)(m + 1) * n, 0, 0);
            SolveArgs T{};
            T.sign = 1.0; T.nb = (int)(c->n / 6); T.n = (int)c->n; T.nbatch = c->nbatch; T.direction_only = 1;
            T.J = d_KT; T.F = c->d_Ft; T.x = c->d_x; T.U = c->d_U; T.p = c->d_p; T.xt = c->d_xt; T.ns = c->ns;
            const int wpc = 4;
            k_factor_solve<4, 3><<<grid1(c->nbatch, wpc), 32 * wpc, (size_t)wpc * LuShape<4, 3>::SM_DOUBLES * 8, c->stream>>>(T);
            CK(cudaGetLastError());
            c->stats.launches++; c->stats.n_factor++;
            k_arnoldi_orth<<<(unsigned)nb, 256, (n + 2 * (size_t)(m + 1)) * 8, c->stream>>>(d_V, d_H, c->d_p, (int)n, m, j);
            c->stats.launches += 2;
            continue;
        }
        k_rowblock_spmv<<<grid1(nb * n * 8, 256), 256, 0, c->stream>>>(c->d_M, d_V + (size_t)j * n, c->d_Ft, (int)n, c->W, c->KLB, c->nbatch, (int64_t)(m + 1) * n, 18, 12);
        if (d_L && j > 0) {
            SolveArgs R{};
            R.sign = 1.0; R.nb = (int)(c->n / 6); R.n = (int)c->n; R.nbatch = c->nbatch;
            R.F = c->d_Ft; R.U = c->d_U; R.p = c->d_p; R.Lst = d_L; R.l_stride = l_stride;
            const int wpc = 4;
            k_resolve_mma<3, 4><<<grid1(c->nbatch, wpc), 32 * wpc, (size_t)wpc * MmaShape<3, 4>::SM_DOUBLES * 8, c->stream>>>(R);
            CK(cudaGetLastError());
            c->stats.launches++;
        } else if ((rc = launch_factor(c, &o, 1, 1.0, d_L, l_stride))) return rc;
        k_arnoldi_orth<<<(unsigned)nb, 256, (n + 2 * (size_t)(m + 1)) * 8, c->stream>>>(d_V, d_H, c->d_p, (int)n, m, j);
        c->stats.launches += 2;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaMemcpyAsync(H, d_H, nb * (size_t)(m + 1) * m * 8, cudaMemcpyDeviceToHost, c->stream));
    if (V) CK(cudaMemcpyAsync(V, d_V, nb * (size_t)(m + 1) * n * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->stats.ms_total = ms;
    {
        std::vector<int> fc(nb);
        CK(cudaMemcpy(fc.data(), c->ns.failcode, nb * 4, cudaMemcpyDeviceToHost));
        c->eig_singular.assign(nb, 0);
        for (size_t i = 0; i < nb; i++) c->eig_singular[i] = fc[i] == FAIL_SINGULAR;
    }
    tmp.release(d_V);
    c->d_Vk = d_V; c->vk_m = m;          // stays resident: gxb_eigen_ritz_vectors forms V y on the device
    return GXB_OK;
}

int gxb_eigen_status(gxb_ctx* c, int32_t* singular) {
    NEED_BATCH("gxb_eigen_status");
    if (!singular || c->eig_singular.size() != (size_t)c->nbatch) return fail(GXB_ERR_INVALID, "gxb_eigen_status: call gxb_eigen_arnoldi first");
    memcpy(singular, c->eig_singular.data(), (size_t)c->nbatch * 4);
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
    DevTemps tmp;
    CK(tmp.alloc(&d_Y, nb * m * nev * 16)); CK(tmp.alloc(&d_out, nb * n * nev * 16));
    CK(cudaMemcpyAsync(d_Y, Y, nb * m * nev * 16, cudaMemcpyHostToDevice, c->stream));
    dim3 grid((unsigned)nb, (unsigned)((n + 127) / 128));
    k_ritz_vectors<<<grid, 128, 0, c->stream>>>(c->d_Vk, d_Y, d_out, (int)n, m, nev);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(vec, d_out, nb * n * nev * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->stats.launches += 1;
    return GXB_OK;
}

int gxb_newton_direction(gxb_ctx* c, const gxb_opts* opts, const double* x, double* p) {
    NEED_BATCH("gxb_newton_direction");
    if (!x || !p) return fail(GXB_ERR_INVALID, "gxb_newton_direction: null");
    if (c->nbody > 0) return fail(GXB_ERR_UNSUPPORTED, "gxb_newton_direction: not available with body-frame acceleration states (use the solves)");
    gxb_opts o; if (opts) o = *opts; else gxb_default_opts(&o);
    const size_t nb = (size_t)c->nbatch, n = (size_t)c->n;
    ModeGuard guard{c};
    c->damping = (c->kind == 1 && o.structural_damping) ? 1 : 0;
    CK(cudaMemcpyAsync(c->d_xt, x, nb * n * 8, cudaMemcpyHostToDevice, c->stream));
    StaticArgs A = make_static_args(c, &o, c->d_xt, c->d_Ft, true, 0, false);
    int rc = launch_assemble(c, A, true); if (rc) return rc;
    if ((rc = launch_factor(c, &o, 1))) return rc;
    CK(cudaMemcpyAsync(p, c->d_p, nb * n * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return GXB_OK;
}


// expand the shared base record and overwrite the K per-instance fields: bcS[b][f][p] layout
__global__ void k_bc_compact(const double* __restrict__ base, const double* __restrict__ vals, const int* __restrict__ point, const int* __restrict__ field,
                             int K, int P, int64_t nbatch, double* __restrict__ bcS) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * 18 * P) return;
    const int p = (int)(t % P); const int64_t q = t / P; const int f = (int)(q % 18); const int64_t b = q / 18;
    double v = base[p * 18 + f];
    for (int k = 0; k < K; k++) if (point[k] == p && field[k] == f) v = vals[b * K + k];
    bcS[t] = v;
}

int gxb_set_bc_compact(gxb_ctx* c, const double* base, int32_t K, const int32_t* point, const int32_t* field, const double* values) {
    NEED_BATCH("gxb_set_bc_compact");
    if (!base || K < 0 || (K > 0 && (!point || !field || !values))) return fail(GXB_ERR_INVALID, "gxb_set_bc_compact: bad arguments");
    std::vector<int> pf(2 * (size_t)std::max(K, 1));
    for (int k = 0; k < K; k++) {
        if (point[k] < 1 || point[k] > c->np || field[k] < 0 || field[k] >= 18) return fail(GXB_ERR_INVALID, "gxb_set_bc_compact: point is 1-based, field in 0..17");
        pf[k] = point[k] - 1; pf[K + k] = field[k];
    }
    const size_t nb = (size_t)c->nbatch, P = (size_t)c->np;
    const size_t bytes = P * 18 * 8 + nb * K * 8 + 2 * (size_t)K * 4 + 64;
    if (bytes > c->stage_bytes) { cudaFree(c->d_stage); c->d_stage = nullptr; c->stage_bytes = 0; CK(cudaMalloc(&c->d_stage, bytes)); c->stage_bytes = bytes; }
    double* d_base = c->d_stage; double* d_vals = d_base + P * 18; int* d_pf = reinterpret_cast<int*>(d_vals + nb * K);
    CK(cudaMemcpyAsync(d_base, base, P * 18 * 8, cudaMemcpyHostToDevice, c->stream));
    if (K) {
        CK(cudaMemcpyAsync(d_vals, values, nb * K * 8, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(d_pf, pf.data(), 2 * (size_t)K * 4, cudaMemcpyHostToDevice, c->stream));     // pageable: staged before the call returns
    }
    CK(cudaEventRecord(c->evC, c->stream));
    k_bc_compact<<<grid1(nb * 18 * P, 256), 256, 0, c->stream>>>(d_base, d_vals, d_pf, d_pf + K, K, (int)P, c->nbatch, c->d_bcS);
    CK(cudaGetLastError());
    return wait_copy(c);
}

}  // extern "C"

// ---- several GPUs of one box: one context + one worker thread per GPU, contiguous shards, no collective ----
struct gxb_multi {
    std::vector<gxb_ctx*> ctx;
    std::vector<int> device;           // device of every shard (several shards may share one GPU: "lanes")
    std::vector<int64_t> off;          // shard b-range [off[g], off[g+1])
    int kind = -1; int64_t np = 0, ne = 0, n = 0, nbatch = 0;
    // shards on the same GPU share one PCIe link: their uploads are taken one after the other, in shard order, so that the first shard's
    // Newton solve starts under the second shard's transfer instead of all transfers finishing together
    std::mutex up_mu; std::condition_variable up_cv; std::vector<int> up_next;      // per device: next shard allowed to upload
};
// run f(g) on one host thread per GPU; the first failure's code and message are reported
template <class F> static int multi_fanout(gxb_multi* m, F&& f) {
    const size_t G = m->ctx.size();
    std::vector<int> rc(G, 0); std::vector<std::string> msg(G);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++)
        th.emplace_back([&, g]() { rc[g] = f((int)g); if (rc[g]) msg[g] = g_err; });        // g_err is thread-local: copy it out
    for (auto& t : th) t.join();
    for (size_t g = 0; g < G; g++) if (rc[g]) return fail(rc[g], "GPU shard " + std::to_string(g) + ": " + msg[g]);
    return GXB_OK;
}
extern "C" {

int gxb_multi_create(gxb_multi** out, int32_t ngpu, const int32_t* device_ids) {
    if (!out || ngpu < 1) return fail(GXB_ERR_INVALID, "gxb_multi_create: need ngpu >= 1");
    gxb_multi* m = new gxb_multi();
    for (int g = 0; g < ngpu; g++) {
        gxb_ctx* c = nullptr;
        int rc = gxb_create(&c, device_ids ? device_ids[g] : g);
        if (rc) { for (auto* q : m->ctx) gxb_destroy(q); delete m; return rc; }
        m->ctx.push_back(c); m->device.push_back(device_ids ? device_ids[g] : g);
    }
    *out = m;
    return GXB_OK;
}
void gxb_multi_destroy(gxb_multi* m) {
    if (!m) return;
    for (auto* c : m->ctx) gxb_destroy(c);
    delete m;
}
int gxb_multi_topology(gxb_multi* m, int kind, int64_t np, int64_t ne, const int64_t* start, const int64_t* stop, const uint8_t* pd, const uint8_t* pl,
                       int64_t* nstates, int64_t* irow_point, int64_t* irow_elem, int64_t* icol_point, int64_t* icol_elem) {
    if (!m) return fail(GXB_ERR_INVALID, "gxb_multi_topology: null");
    for (size_t g = 0; g < m->ctx.size(); g++) {
        const bool first = g == 0;
        int rc = gxb_topology(m->ctx[g], kind, np, ne, start, stop, pd, pl, first ? nstates : nullptr, first ? irow_point : nullptr,
                              first ? irow_elem : nullptr, first ? icol_point : nullptr, first ? icol_elem : nullptr);
        if (rc) return rc;
    }
    m->kind = kind; m->np = np; m->ne = ne; m->n = m->ctx[0]->n; m->nbatch = 0;
    return GXB_OK;
}
int gxb_multi_set_batch(gxb_multi* m, int64_t nbatch) {
    if (!m || m->kind < 0) return fail(GXB_ERR_INVALID, "gxb_multi_set_batch: call gxb_multi_topology first");
    const int64_t G = (int64_t)m->ctx.size();
    if (nbatch < G) return fail(GXB_ERR_INVALID, "gxb_multi_set_batch: fewer instances than GPUs");
    m->off.assign(G + 1, 0);
    for (int64_t g = 0; g <= G; g++) m->off[g] = nbatch * g / G;
    int rc = multi_fanout(m, [&](int g) { return gxb_set_batch(m->ctx[g], m->off[g + 1] - m->off[g]); });
    if (rc) return rc;
    m->nbatch = nbatch;
    return GXB_OK;
}
int gxb_multi_solve(gxb_multi* m, const gxb_opts* opts, const double* elems, const double* points, const double* bc, const double* dloads,
                    const double* pmass, const double* gravity, const double* force_scaling, const double* motion, const double* x0,
                    double* x, int32_t* converged, int32_t* iters, double* resid_inf, double* shard_ms) {
    if (!m || m->nbatch < 1) return fail(GXB_ERR_INVALID, "gxb_multi_solve: call gxb_multi_set_batch first");
    if (!elems || !bc) return fail(GXB_ERR_INVALID, "gxb_multi_solve: elems and bc are required");
    const int64_t P = m->np, N = m->ne, n = m->n;
    return multi_fanout(m, [&](int g) {
        gxb_ctx* c = m->ctx[g];
        const int64_t b0 = m->off[g];
        int rc;
        if ((rc = gxb_set_elements(c, elems + b0 * N * GXB_ELEM_STRIDE))) return rc;
        if (m->kind == 1 && points && (rc = gxb_set_points(c, points + b0 * P * 3))) return rc;
        if ((rc = gxb_set_bc(c, bc + b0 * P * 18))) return rc;
        if ((rc = gxb_set_dloads(c, dloads ? dloads + b0 * N * 24 : nullptr))) return rc;
        if ((rc = gxb_set_pmass(c, pmass ? pmass + b0 * P * 36 : nullptr))) return rc;
        if ((rc = gxb_set_body(c, gravity ? gravity + b0 * 3 : nullptr, force_scaling ? force_scaling + b0 : nullptr))) return rc;
        if (m->kind == 1 && (rc = gxb_set_motion(c, motion ? motion + b0 * 12 : nullptr))) return rc;
        auto solve = m->kind == 0 ? gxb_static_solve : gxb_steady_solve;
        rc = solve(c, opts, x0 ? x0 + b0 * n : nullptr, x ? x + b0 * n : nullptr, converged ? converged + b0 : nullptr, iters ? iters + b0 : nullptr,
                   resid_inf ? resid_inf + b0 : nullptr);
        if (!rc && shard_ms) shard_ms[g] = c->stats.ms_total;
        return rc;
    });
}
int gxb_multi_hint_no_gravity(gxb_multi* m, int on) {
    if (!m) return fail(GXB_ERR_INVALID, "gxb_multi_hint_no_gravity: null");
    for (auto* c : m->ctx) { int rc = gxb_hint_no_gravity(c, on); if (rc) return rc; }
    return GXB_OK;
}
int gxb_multi_solve_compact(gxb_multi* m, const gxb_opts* opts, const double* elems, const double* bc_base, int32_t K, const int32_t* bc_point,
                            const int32_t* bc_field, const double* bc_values, const double* gravity, const double* force_scaling, const double* x0,
                            double* x, int32_t* converged, int32_t* iters, double* resid_inf, double* shard_ms) {
    if (!m || m->nbatch < 1) return fail(GXB_ERR_INVALID, "gxb_multi_solve_compact: call gxb_multi_set_batch first");
    if (m->kind != 0) return fail(GXB_ERR_INVALID, "gxb_multi_solve_compact: static contexts (kind = 0) only");
    if (!elems || !bc_base || K < 0 || (K > 0 && (!bc_point || !bc_field || !bc_values)))
        return fail(GXB_ERR_INVALID, "gxb_multi_solve_compact: elems, bc_base and the K varying entries are required");
    const int64_t N = m->ne, n = m->n;
    const size_t G = m->ctx.size();
    // shard order per device
    std::vector<int> order(G, 0);
    {
        int maxdev = 0; for (int d : m->device) maxdev = std::max(maxdev, d);
        std::vector<int> cnt(maxdev + 1, 0);
        for (size_t g = 0; g < G; g++) order[g] = cnt[m->device[g]]++;
        std::lock_guard<std::mutex> lk(m->up_mu);
        m->up_next.assign(maxdev + 1, 0);
    }
    return multi_fanout(m, [&](int g) {
        gxb_ctx* c = m->ctx[g];
        const int64_t b0 = m->off[g];
        const int dev = m->device[g];
        int rc = 0;
        {
            std::unique_lock<std::mutex> lk(m->up_mu);
            m->up_cv.wait(lk, [&] { return m->up_next[dev] == order[g]; });
        }
        // gxb_set_* return once the host buffers have been consumed (the copies are complete), so the next shard may start its transfer
        if (!rc) rc = gxb_set_elements(c, elems + b0 * N * GXB_ELEM_STRIDE);
        if (!rc) rc = gxb_set_bc_compact(c, bc_base, K, bc_point, bc_field, bc_values + b0 * K);
        if (!rc) rc = gxb_set_body(c, gravity ? gravity + b0 * 3 : nullptr, force_scaling ? force_scaling + b0 : nullptr);
        {
            std::lock_guard<std::mutex> lk(m->up_mu);
            m->up_next[dev] = order[g] + 1;
        }
        m->up_cv.notify_all();
        if (rc) return rc;
        rc = gxb_static_solve(c, opts, x0 ? x0 + b0 * n : nullptr, x ? x + b0 * n : nullptr, converged ? converged + b0 : nullptr,
                              iters ? iters + b0 : nullptr, resid_inf ? resid_inf + b0 : nullptr);
        if (!rc && shard_ms) shard_ms[g] = c->stats.ms_total;
        return rc;
    });
}

}  // extern "C"
