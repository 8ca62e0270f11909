/* gxbeam_b200 — C-ABI of the B200-native batched Newton solver for GXBeam's residual/Jacobian/solve hot path.
 *
 * The reference (byuflowlab/GXBeam.jl, pure Julia) has no FFI layer.  The seam this library sits behind is the pair
 *     nlsolve!(p, constants, residual!, jacobian!)         src/analyses.jl:3556-3599
 *     lsolve!(x0, p, constants, residual!, jacobian!)      src/analyses.jl:3506-3524
 * as reached from static_analysis! (src/analyses.jl:78-179).  Every entry point below names the reference code it
 * replaces.  Julia binds these with `ccall` from a new `static_analysis_batch`; in this repository
 * the same symbols are bound with ctypes (gxbeam_b200/api.py).
 *
 * Conventions: every function returns 0 on success and a GXB_ERR_* code otherwise (message via gxb_last_error());
 * nothing throws across the boundary; the caller owns every host buffer (plain contiguous Float64 / Int64 / UInt8
 * arrays, Julia column-major == "last index slowest" as written below); the library owns all device memory.
 * A context is bound to one CUDA device and must be used by one host thread at a time.  There is no CPU fallback:
 * if no CUDA device is usable gxb_create fails.
 */
#ifndef GXBEAM_B200_H
#define GXBEAM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GXB_OK 0
#define GXB_ERR_INVALID 1      /* bad argument / call order */
#define GXB_ERR_CUDA 2         /* CUDA runtime error */
#define GXB_ERR_UNSUPPORTED 3  /* topology or option the device path does not implement (fails loudly) */

#define GXB_ELEM_STRIDE 91  /* Element{Float64}: L, x[3], compliance[36], mass[36], Cab[9], mu[6]  (src/assembly.jl:14-21) */
#define GXB_BC_STRIDE 18    /* PrescribedConditions: u, theta, F, M, Ff, Mf                      (src/loads.jl:16-25)   */
#define GXB_DLOAD_STRIDE 24 /* DistributedLoads: f1,f2,m1,m2,f1_follower,...,m2_follower          (src/loads.jl:262-271) */

typedef struct gxb_ctx gxb_ctx;

/* Keyword arguments of static_analysis! that reach the solver (src/analyses.jl:78-102) plus the BackTracking constants
 * of LineSearches (c_1, rho_hi, rho_lo). */
typedef struct gxb_opts {
    int32_t linear;          /* linear=true: one step x = x0 - J\r, converged = true   (analyses.jl:3506-3524) */
    int32_t two_dimensional; /* system.jl:1072-1103 */
    double ftol;             /* 1e-9 */
    int64_t iterations;      /* 1000 */
    double maxstep;          /* 1e6 */
    double c1, rho_hi, rho_lo; /* 1e-4, 0.5, 0.1 */
    int32_t profile;         /* !=0: time every kernel class with CUDA events (gxb_get_stats) */
    int32_t structural_damping; /* kind=1 only: structural_damping keyword (analyses.jl:397 default false, :1708 default true) */
    int32_t persistent;      /* kind=1 only: 0 auto, 1 force, 2 forbid the persistent per-instance kernel (one CTA runs an instance's
                                whole Newton loop / time loop; chosen automatically for small batches).  No effect on results beyond
                                round-off in the line search's sums. */
    int32_t update_linearization; /* gxb_newmark_run with linear != 0: 0 = linearise every step about the initial state (the reference's
                                default, analyses.jl:2536), 1 = about the current state */
    int32_t lu_parts;        /* warps per instance of the factorisation: 0 auto (by batch size), 1 one warp per instance (tensor-core kernel),
                                2 the two-sided kernel (top-down + mirrored bottom-up sweep meeting in a dense interface; static and dynamic
                                systems), 3 / 4 / 6 / 8 (kind = 0 only) the partitioned kernel: the chain is cut into lu_parts sub-chains that
                                are eliminated concurrently and coupled through a small interface system.  Any setting solves the same
                                linear system; results agree to round-off. */
    int32_t exact_dphi0;     /* line search slope dφ0 = g'p with g = J'F as NLsolve's newton evaluates it (one extra pass over J per Newton
                                iteration) instead of -F'F, which is the same number in exact arithmetic because J p = -F.  Default 0.
                                The singular-Jacobian fallback step p = -(J'J + λI) \ J'F always uses g'p. */
} gxb_opts;

/* Filled by gxb_get_stats after a solve. */
typedef struct gxb_stats {
    int64_t rounds;            /* host-visible Newton/line-search rounds */
    int64_t launches;          /* kernels launched by the last solve */
    int64_t newton_iters;      /* sum over instances of Newton iterations (= linear solves) */
    int64_t resid_evals;       /* sum over instances of residual evaluations */
    int64_t n_assemble, n_factor, n_update; /* launches per kernel class */
    double ms_assemble, ms_factor, ms_update; /* summed device time per class (only with opts.profile) */
    double ms_total;           /* device time of the whole resident solve */
} gxb_stats;

void gxb_default_opts(gxb_opts* o);

/* Create / destroy a context on CUDA device `device`. */
int gxb_create(gxb_ctx** ctx, int device);
void gxb_destroy(gxb_ctx* ctx);
const char* gxb_last_error(void);

/* SystemIndices(start, stop; static = kind==0)                                        src/system.jl:29-122
 * kind: 0 = StaticSystem (6 states / point), 1 = DynamicSystem (12 states / point).
 * start/stop: ne, 1-based.  pd/pl: np x 6 flags shared by the whole batch (they fix the sparsity pattern).
 * Outputs (may be NULL): nstates, irow_point[np], irow_elem[ne], icol_point[np], icol_elem[ne], 1-based, bit-exact. */
int gxb_topology(gxb_ctx* ctx, int kind, int64_t np, int64_t ne, const int64_t* start, const int64_t* stop, const uint8_t* pd,
                 const uint8_t* pl, int64_t* nstates, int64_t* irow_point, int64_t* irow_elem, int64_t* icol_point,
                 int64_t* icol_elem);

/* Canonical structural pattern of the Jacobian at 3x3-block granularity (what insert_static_point_jacobians!
 * src/point.jl:1785-1800 and insert_static_element_jacobians! src/element.jl:2487-2552 can touch).
 * For kind=1: insert_dynamic/steady_point_jacobians! src/point.jl:1907-1935, 1809-1826 and insert_dynamic_element_jacobians!
 * src/element.jl:2685-2759 (the structural-damping blocks included; body-acceleration columns excluded).
 * Returns the number of blocks in *nblocks; brow/bcol (may be NULL) receive the 0-based scalar row/col of each block. */
int gxb_pattern(gxb_ctx* ctx, int64_t* nblocks, int64_t* brow, int64_t* bcol);

/* Per-instance inputs.  All arrays are [nbatch][...] contiguous.  gxb_set_batch first. */
int gxb_set_batch(gxb_ctx* ctx, int64_t nbatch);
int gxb_set_elements(gxb_ctx* ctx, const double* elems);   /* nbatch x ne x 91   assembly.elements            */
/* Static contexts only: promise that gravity is zero, so that gxb_set_elements uploads only L, x, compliance and Cab (49 of the 91
 * doubles of an Element; static_element_properties reads nothing else without gravity, src/element.jl:62-107, 1860-1899).
 * A later gxb_set_body with non-zero gravity fails loudly. */
int gxb_hint_no_gravity(gxb_ctx* ctx, int on);
int gxb_set_points(gxb_ctx* ctx, const double* xyz);       /* nbatch x np x 3    assembly.points (dynamic only) */
int gxb_set_bc(gxb_ctx* ctx, const double* bc);            /* nbatch x np x 18   prescribed_conditions values  */
int gxb_set_dloads(gxb_ctx* ctx, const double* dl);        /* nbatch x ne x 24 or NULL   distributed_loads     */
int gxb_set_pmass(gxb_ctx* ctx, const double* pm);         /* nbatch x np x 36 (col-major 6x6) or NULL  point_masses */
int gxb_set_body(gxb_ctx* ctx, const double* gravity /* nbatch x 3 or NULL */, const double* force_scaling /* nbatch */);
/* kind=1 only: linear_velocity, angular_velocity, linear_acceleration, angular_acceleration of the body frame
 * (steady_state_analysis keyword arguments, src/analyses.jl:388-391): nbatch x 12 = vb, ωb, ab, αb; NULL = zeros. */
int gxb_set_motion(gxb_ctx* ctx, const double* motion);

/* static_analysis!: nlsolve!/lsolve! over the batch                       src/analyses.jl:137-171, 3506-3599
 * x0: nbatch x nstates initial guess or NULL (zeros, = reset_state).  Outputs: x (nbatch x nstates), converged, iters
 * (Newton iterations), resid_inf (||r||_inf of the final residual); each may be NULL. */
int gxb_static_solve(gxb_ctx* ctx, const gxb_opts* opts, const double* x0, double* x, int32_t* converged, int32_t* iters,
                     double* resid_inf);

/* steady_state_analysis!: the same Newton driver on a DynamicSystem (kind=1)          src/analyses.jl:383-527, 3556-3599
 * with steady_system_residual!/jacobian! (src/system.jl:427-455, 693-720). */
int gxb_steady_solve(gxb_ctx* ctx, const gxb_opts* opts, const double* x0, double* x, int32_t* converged, int32_t* iters,
                     double* resid_inf);
int gxb_steady_solve_resident(gxb_ctx* ctx, const gxb_opts* opts);

/* time_domain_analysis! from a given initial state: Newmark time marching                src/analyses.jl:2600-2790
 *   per step: prescribed conditions of the new time (gxb_set_bc_series), the `paug` initialisation terms (analyses.jl:2693-2738),
 *   nlsolve! with newmark_system_residual!/jacobian! (src/system.jl:530-554, 829-852) starting from the previous state.
 * The initial state is what the reference obtains from initial_condition_analysis! or takes from `initial_state=`
 * (analyses.jl:2549-2596): x0 (nbatch x nstates) and the point rates udot, θdot, Vdot, Ωdot at t0 (nbatch x np x 12).
 * tvec: nt times.  Saved time levels: 0, save_every, 2 save_every, ...  -> x_hist (nbatch x nsave x nstates, may be NULL).
 * steps_done[b]: time levels reached (an instance stops at its first unconverged step, analyses.jl:2781-2790);
 * iters_total[b]: Newton iterations over all steps; converged[b].  The reference's time-domain default is
 * structural_damping = true: set opts->structural_damping accordingly. */
int gxb_set_bc_series(gxb_ctx* ctx, int64_t nt, int32_t K, const int32_t* point /* 1-based */, const int32_t* field /* 0..17 */,
                      const double* values /* nbatch x nt x K */);
int gxb_newmark_run(gxb_ctx* ctx, const gxb_opts* opts, int64_t nt, const double* tvec, const double* x0, const double* rates0,
                    int64_t save_every, double* x_hist, int32_t* steps_done, int32_t* iters_total, int32_t* converged,
                    double* rates_hist /* optional: nbatch x nsave x np x 12, the dx of newmark_output (analyses.jl:3461-3503) */);

/* The same solve on inputs already resident in HBM (no host<->device traffic): x0 from gxb_upload_x0 or zeros. */
int gxb_upload_x0(gxb_ctx* ctx, const double* x0 /* or NULL */);
int gxb_static_solve_resident(gxb_ctx* ctx, const gxb_opts* opts);
int gxb_fetch(gxb_ctx* ctx, double* x, int32_t* converged, int32_t* iters, double* resid_inf);
int gxb_get_stats(gxb_ctx* ctx, gxb_stats* st);

/* {static,steady}_system_residual! + _jacobian! at given states (by the context's kind)
 *                                                          src/system.jl:400-418, 663-683 | 427-455, 693-720
 * x: nbatch x nstates.  r: nbatch x nstates.  Jblk: nbatch x nblocks x 9, each 3x3 row-major on gxb_pattern's blocks. */
int gxb_residual_jacobian(gxb_ctx* ctx, const gxb_opts* opts, const double* x, double* r, double* Jblk);
/* newmark_system_residual! + newmark_system_jacobian! (src/system.jl:530-554, 829-852) at given states, with the rate
 * initialisation terms init (nbatch x np x 12: udot_init, θdot_init, Vdot_init, Ωdot_init) and the step size dt. */
int gxb_newmark_residual_jacobian(gxb_ctx* ctx, const gxb_opts* opts, const double* init, double dt, const double* x, double* r,
                                  double* Jblk);

/* initial_condition_analysis! (consistent initial conditions of a time-domain run)        src/analyses.jl:1695-1945
 *   rate_vars1/2 from two mass-matrix assemblies (analyses.jl:1835-1873), one nlsolve! with initial_system_residual!/jacobian!
 *   (src/system.jl:466-521, 731-817), then initial_output (analyses.jl:2332-2390).
 * ic: nbatch x np x 18 = u0, theta0, V0, Omega0, Vdot0, Omegadot0 of every point (NULL = zeros, the reference's default).
 * x0: nbatch x nstates guess in the initial-condition layout or NULL (reset_state).  opts->structural_damping as in the reference
 * (its time-domain default is true).  Body-frame velocities AND accelerations come from gxb_set_motion.
 * Outputs: x (nbatch x nstates, DynamicSystem states), rates (nbatch x np x 12: udot, theta_dot, Vdot, Omegadot) -- exactly the
 * x0 / rates0 of gxb_newmark_run, which may instead be called with x0 = rates0 = NULL right after this function to march from
 * the state left in HBM.  rate_vars (optional): nbatch x np x 12 = rate_vars1[icol+6..11], rate_vars2[icol+6..11] per point. */
int gxb_initial_condition(gxb_ctx* ctx, const gxb_opts* opts, const double* ic, const double* x0, double* x, double* rates,
                          int32_t* converged, int32_t* iters, double* resid_inf, uint8_t* rate_vars);
/* initial_system_residual! + initial_system_jacobian! (src/system.jl:466-521, 731-817) at given vectors x (nbatch x nstates, in the
 * initial-condition layout).  rate_vars1 / rate_vars2: nstates flags each (shared by the batch), or both NULL to compute them
 * as initial_condition_analysis! does. */
int gxb_initial_residual_jacobian(gxb_ctx* ctx, const gxb_opts* opts, const double* ic, const uint8_t* rate_vars1,
                                  const uint8_t* rate_vars2, const double* x, double* r, double* Jblk);

/* Ritz vectors of the last gxb_eigen_arnoldi on the device: vec[b] = V_m[b] * Y[b]   (`V = Vk * F.vectors`, src/analyses.jl:959).
 * Y: nbatch x m x nev complex (interleaved re, im) -- the caller's chosen eigenvectors of the Hessenberg matrices;
 * vec: nbatch x nstates x nev complex.  The Krylov basis stays in HBM; only nev vectors per instance are copied back. */
int gxb_eigen_ritz_vectors(gxb_ctx* ctx, int32_t nev, const double* Y, double* vec);

/* Body-frame accelerations as state variables (a DOF with displacement AND load prescribed):
 * update_body_acceleration_indices! / body_accelerations (src/system.jl:138-182), their dense Jacobian columns
 * (point.jl:1829-1841, 1879-1895; element.jl:2567-2583).  Handled inside the steady / initial-condition solves; this call returns
 * icol_body (6 entries, 1-based, 0 = none) and, after a gxb_*residual_jacobian call, the dense columns D (6 x nbatch x nstates):
 * J[:, icol_body[i]] = D[i] + e_{icol_body[i]}; the pattern blocks of Jblk hold the unit entry in that column. */
int gxb_body_columns(gxb_ctx* ctx, int64_t* icol_body, double* D);

/* system_mass_matrix! (src/system.jl:961-995): M = d r / d xdot at the states x, on gxb_pattern's blocks (kind=1).
 * Mblk: nbatch x nblocks x 9. */
int gxb_mass_matrix(gxb_ctx* ctx, const gxb_opts* opts, const double* x, double* Mblk);

/* The Krylov part of solve_eigensystem (src/analyses.jl:943-965): K = steady Jacobian and M = mass matrix at the states x
 * (what eigenvalue_analysis! builds at src/analyses.jl:1358-1359), then m Arnoldi steps with the operator v -> K^{-1} (M v)
 * (the LinearMap `f!` of analyses.jl:949).  K is factorised ONCE, like `Kfact = lu(K)` (analyses.jl:948): the first application records the
 * multipliers of the batched pivoted block-banded LU, every further application replays them on the new right-hand side.
 * Outputs: H, nbatch x (m+1) x m row-major upper Hessenberg with H[j+1][j] the residual norms; V (may be NULL),
 * nbatch x (m+1) x nstates, the orthonormal Krylov basis.  The caller finishes like the reference: mu = eig(H[:m,:m]),
 * sort by (abs, imag) descending, lambda = -1/mu, eigenvectors V[:m]' y   (Julia: `eigen`, Python: numpy.linalg.eig). */
int gxb_eigen_arnoldi(gxb_ctx* ctx, const gxb_opts* opts, const double* x, int32_t m, double* H, double* V);
/* The same Arnoldi process for the TRANSPOSED pencil, v -> K'^{-1} (M' v): its Ritz vectors are the left eigenvectors of the system,
 * the rows of the matrix `U` of left_eigenvectors (src/analyses.jl:966-1079; the reference gets them by three passes of inverse
 * iteration with a complex shifted factorisation per mode -- the same vectors up to the normalisation U M V = I, which the caller
 * applies).  H, V and gxb_eigen_ritz_vectors as for gxb_eigen_arnoldi (the retained Krylov basis is replaced). */
int gxb_eigen_arnoldi_left(gxb_ctx* ctx, const gxb_opts* opts, const double* x, int32_t m, double* H, double* V);
/* Per instance (nbatch flags): 1 if the factorisation of K in the last gxb_eigen_arnoldi hit an exactly-zero / non-finite pivot -- where
 * the reference's `lu(K)` (src/analyses.jl:948) throws SingularException; that instance's H is meaningless. */
int gxb_eigen_status(gxb_ctx* ctx, int32_t* singular);

/* One Newton direction p = -(J(x) \ r(x)) per instance (the `linsolve` of analyses.jl:3585), for parity tests of the
 * batched pivoted block-banded LU in isolation.  p: nbatch x nstates. */
int gxb_newton_direction(gxb_ctx* ctx, const gxb_opts* opts, const double* x, double* p);

/* AssemblyState(x, system, assembly; prescribed_conditions) on the device (src/output.jl:106, extract_point_state :146-203,
 * extract_element_state :317-403) for the states of the last solve, still resident in HBM: points = nbatch x np x 30 and
 * elements = nbatch x ne x 30 doubles, each record in the field order of the reference's PointState / ElementState structs
 * (u, udot, theta, thetadot, V, Vdot, Omega, Omegadot, F, M; theta as Wiener-Milenkovic parameters), so Julia can
 * `reinterpret(PointState{Float64}, points)`.  rates (optional): nbatch x np x 12 = udot, thetadot, Vdot, Omegadot per point (the dx of
 * newmark_output / gxb_newmark_run's rates_hist); NULL = zeros (StaticSystem) or NaN (DynamicSystem), as the reference fills them. */
int gxb_assembly_state(gxb_ctx* ctx, const double* rates, double* points, double* elements);

/* Compact upload of the prescribed conditions (the e2e path of design sweeps): a PrescribedConditions record is 18 doubles of which
 * all but a handful are NaN / zero and identical across the batch (src/loads.jl:16-25, :94).  `base` (np x 18, shared by the batch, NaN
 * where unused exactly like gxb_set_bc) is replicated on the device; `values` (nbatch x K) then overwrites field[k] (0..17: u, theta, F,
 * M, Ff, Mf) of point[k] (1-based) per instance.  Equivalent to gxb_set_bc with the expanded array; K * 8 bytes per instance cross PCIe
 * instead of np * 144. */
int gxb_set_bc_compact(gxb_ctx* ctx, const double* base, int32_t K, const int32_t* point, const int32_t* field, const double* values);

/* ---- one box, several GPUs: SURVEY 8(e) ----------------------------------------------------------------------------------------
 * A gxb_multi owns one gxb_ctx and one host worker thread per GPU.  The batch is cut into contiguous shards (instances are
 * independent: no collective, no peer traffic); every call fans out to the workers, each of which uploads ITS slice of the caller's
 * host arrays and writes ITS slice of the caller's output arrays -- the "final host gather" is the caller's single array.
 * Results are bit-identical to a single-GPU solve of the same batch (an instance never sees its neighbours) as long as opts.lu_parts is
 * fixed: its automatic choice depends on the shard size, and different partitionings agree to round-off only. */
typedef struct gxb_multi gxb_multi;
int gxb_multi_create(gxb_multi** m, int32_t ngpu, const int32_t* device_ids);
void gxb_multi_destroy(gxb_multi* m);
/* gxb_topology on every GPU (outputs as there, may be NULL), then gxb_set_batch with the shard sizes */
int gxb_multi_topology(gxb_multi* m, int kind, int64_t np, int64_t ne, const int64_t* start, const int64_t* stop, const uint8_t* pd,
                       const uint8_t* pl, int64_t* nstates, int64_t* irow_point, int64_t* irow_elem, int64_t* icol_point,
                       int64_t* icol_elem);
int gxb_multi_set_batch(gxb_multi* m, int64_t nbatch);
/* static_analysis (kind = 0) / steady_state_analysis (kind = 1) of the whole batch.  Arrays are the full-batch versions of the
 * gxb_set_* / gxb_static_solve arguments; dloads, pmass, gravity, force_scaling, x0 may be NULL; points / motion are read for kind = 1
 * only.  shard_ms (optional, ngpu doubles): device time of each shard's solve. */
int gxb_multi_solve(gxb_multi* m, const gxb_opts* opts, const double* elems, const double* points, const double* bc, const double* dloads,
                    const double* pmass, const double* gravity, const double* force_scaling, const double* motion, const double* x0,
                    double* x, int32_t* converged, int32_t* iters, double* resid_inf, double* shard_ms);

/* Design sweeps through host buffers (what bench.py times as `e2e`): gxb_hint_no_gravity on every shard, and gxb_multi_solve with the
 * prescribed conditions in the compact form of gxb_set_bc_compact (bc_base np x 18 shared, bc_values nbatch x K).  device_ids may name
 * the same GPU several times: the shards then act as pipeline lanes of that GPU -- their uploads are taken one after the other in
 * shard order (one PCIe link), so that shard k solves while shard k + 1 is being transferred and shard k - 1 is copied back. */
int gxb_multi_hint_no_gravity(gxb_multi* m, int on);
int gxb_multi_solve_compact(gxb_multi* m, const gxb_opts* opts, const double* elems, const double* bc_base, int32_t K, const int32_t* bc_point,
                            const int32_t* bc_field, const double* bc_values, const double* gravity, const double* force_scaling, const double* x0,
                            double* x, int32_t* converged, int32_t* iters, double* resid_inf, double* shard_ms);

/* ---- cross-section analysis (src/section.jl; BASELINE config 5) ---------------------------------------------------------------------
 * compliance_matrix(nodes, elements; gxbeam_order, shear_center) (src/section.jl:641-783) and mass_matrix(nodes, elements) (:840-890)
 * for nbatch layups that share one quad-mesh topology.  A gxb_section holds the topology: conn = ne x 4 node numbers, 1-based and
 * counterclockwise from the bottom-left node exactly like MeshElement.nodenum (:101-105); the nodes are renumbered once (reverse
 * Cuthill-McKee) so that the global matrices E, C, M of :661-677 are banded.
 * Per instance: nodes nbatch x nn x 2 (Node.x, Node.y), materials nbatch x ne x 10 (E1, E2, E3, G12, G13, G23, nu12, nu13, nu23, rho: the
 * first ten fields of Material, :18-28), theta nbatch x ne (MeshElement.theta).
 * Outputs (any may be NULL): S nbatch x 36 (row-major 6 x 6), sc / tc nbatch x 2 (shear / tension centre), M nbatch x 36, mc nbatch x 2
 * (mass centre), ok nbatch (0 where the elimination met a non-positive pivot: degenerate mesh), device_ms: kernel time. */
typedef struct gxb_section gxb_section;
int gxb_section_create(gxb_section** s, int device, int64_t nn, int64_t ne, const int64_t* conn);
void gxb_section_destroy(gxb_section* s);
int gxb_section_info(gxb_section* s, int64_t* half_bandwidth, int64_t* scratch_doubles_per_instance);
int gxb_section_properties(gxb_section* s, int64_t nbatch, const double* nodes, const double* materials, const double* theta, int gxbeam_order,
                           int shear_center, double* S, double* sc, double* tc, double* M, double* mc, int32_t* ok, double* device_ms);
/* strain_recovery(F, M, nodes, elements, cache; gxbeam_order) (src/section.jl:952-1040) + tsai_wu(stress_p, elements) (:1108-1131) for every
 * layup: the compliance solve of gxb_section_properties (its X, dX, Y are the reference's SectionCache and never leave the GPU) followed by the
 * element-centre strains and stresses under the section loads F, M (nbatch x 3 each; GXBeam axes, x along the beam, if gxbeam_order).
 * strengths: nbatch x ne x 9 (S1t, S1c, S2t, S2c, S3t, S3c, S12, S13, S23: Material fields 11-19, :29-37), NULL if failure is NULL.
 * Outputs (any may be NULL): strain_b, stress_b (beam axes: xx yy zz xy xz yz), strain_p, stress_p (ply axes: 11 22 33 12 13 23), each
 * nbatch x ne x 6 = the memory of the reference's 6 x ne matrices; failure nbatch x ne (Tsai-Wu index, fails if >= 1); S nbatch x 36
 * (compliance matrix about the shear centre, as gxb_section_properties with shear_center = 1); ok, device_ms as there. */
int gxb_section_strain_recovery(gxb_section* s, int64_t nbatch, const double* nodes, const double* materials, const double* theta,
                                const double* strengths, const double* F, const double* M, int gxbeam_order, double* strain_b, double* stress_b,
                                double* strain_p, double* stress_p, double* failure, double* S, int32_t* ok, double* device_ms);
/* parity entry: element_submatrix (src/section.jl:486-501) of every element: out = nbatch x ne x 612 doubles = Ae (6x6), Re (12x6),
 * Ee (12x12), Ce (12x12), Le (12x6), Me (12x12), row-major. */
int gxb_section_element_matrices(gxb_section* s, int64_t nbatch, const double* nodes, const double* materials, const double* theta, double* out);

#ifdef __cplusplus
}
#endif
#endif
