/*
 * fctpdeco.h -- C ABI of libfctpdeco.so: the B200 (sm_100a) implementation of the hot path of
 * KarolinaBenkova/FEM-FCT-PDECO (P1 FEM flux-corrected transport step, element assembly into a fixed
 * CSR pattern, sparse state/adjoint solves, norms / cost functional, PDECO time loops).
 *
 * The reference has no FFI: its boundary is the Python namespace of helpers.py.  Each
 * entry point below names the reference function (file:line in the reference repo) whose arithmetic it
 * replaces; the Python package `fem-fct-pdeco_b200/` binds these with ctypes and re-exports the
 * reference names (FCT_alg_ref, ChebSI, ...).  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C, no exceptions: every call returns 0 on success, nonzero on error; fct_last_error()
 *     returns a message for the calling thread's last failure.
 *   - one context per (process, device); a context is NOT thread-safe; calls enqueue work on the
 *     context's CUDA stream.  `_dev` pointers are device pointers (fp64 unless said otherwise) valid
 *     on the context's device; `_host` pointers are host pointers and the call synchronises before
 *     returning.  Functions without host outputs return as soon as the work is enqueued.
 *   - all matrices are *value arrays of length nnz on the context's fixed CSR pattern* (rowptr, colidx,
 *     columns ascending, diagonal present, structurally symmetric) -- what helpers.py:87-104
 *     (assemble_sparse: PETSc getValuesCSR) returns.  Vectors are in DoF order.
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef FCTPDECO_H
#define FCTPDECO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fct_ctx fct_ctx;

/* per-step diagnostics written by fct_step (host struct) */
typedef struct fct_step_info {
    int32_t solver_sweeps;     /* Jacobi sweeps executed for the low-order system            */
    int32_t converged;         /* 1 if the stopping test was met within max_sweeps           */
    double  last_delta;        /* ||x_k - x_{k-1}||_inf at exit                               */
    double  x_norm;            /* ||x_k||_inf at exit                                         */
    double  min_rowsum_low;    /* min_i sum_j (M_L + dt(A-D+S))_ij: <= 0 reproduces the       */
                               /* reference's "3: False" diagnostic (helpers.py:1796-1809)    */
} fct_step_info;

/* ---- library ---------------------------------------------------------------------------------- */
const char* fct_last_error(void);
int  fct_version(void);
int  fct_device_count(void);

/* ---- structured mesh bookkeeping (host side, closed form) --------------------------------------
 * Replaces dolfin RectangleMesh(Point(a1,a1),Point(a2,a2),n,n) + FunctionSpace(mesh,'CG',1) +
 * vertex_to_dof_map + find_node_neighbours (advection_solidbody_FCT.py:48-50,82-86,
 * helpers.py:271-307) and the CSR pattern of assemble_sparse (helpers.py:87-104).
 * Sizes: nodes=(n+1)^2, cells=2n^2, nnz = nodes + 2*(2(n+1)n + n^2). */
int fct_mesh_rect_sizes(int32_t n, int64_t* nodes, int64_t* cells, int64_t* nnz);
int fct_mesh_rect_build(int32_t n, double a1, double a2,
                        int32_t* vertex_to_dof,   /* [nodes]            or NULL */
                        int32_t* cell_dofs,       /* [cells*3] DoF idx  or NULL */
                        double*  dof_xy,          /* [nodes*2] DoF order or NULL */
                        int32_t* rowptr,          /* [nodes+1]          or NULL */
                        int32_t* colidx);         /* [nnz]              or NULL */

/* ---- context ---------------------------------------------------------------------------------- */
/* rowptr/colidx are HOST arrays of the (local) pattern; they are copied to the device.
 * row_begin/row_end: the rows this context owns and computes (0,n for a single GPU).  Rows outside
 * are halo rows: their vector entries are inputs refreshed by fct_halo_exchange. */
int fct_ctx_create(fct_ctx** out, int device, int32_t n, const int32_t* rowptr, const int32_t* colidx,
                   int32_t row_begin, int32_t row_end);
int fct_ctx_destroy(fct_ctx* ctx);
int fct_ctx_set_stream(fct_ctx* ctx, void* cuda_stream);      /* cudaStream_t; NULL = legacy default */
int fct_ctx_sync(fct_ctx* ctx);
int fct_ctx_sizes(fct_ctx* ctx, int32_t* n, int64_t* nnz, int32_t* row_begin, int32_t* row_end);
/* device pointers of the pattern (int32) */
int fct_ctx_pattern_dev(fct_ctx* ctx, const int32_t** rowptr_dev, const int32_t** colidx_dev,
                        const int32_t** tpos_dev);
/* P1 mesh for the assembly kernels: DoF-indexed cells and DoF-ordered coordinates (host arrays, copied) */
int fct_ctx_set_mesh(fct_ctx* ctx, int64_t ncells, const int32_t* cell_dofs, const double* dof_xy);
/* static matrices: consistent mass M (values on the pattern), lumped mass ML, diag(M).
 * fct_ctx_set_mass copies from device arrays; fct_assemble_static computes M, ML, K on the device
 * (replaces assemble_sparse_lil(u*v*dx), row_lump, assemble_sparse(dot(grad(u),grad(v))*dx):
 * helpers.py:553-555, 309-328). */
int fct_ctx_set_mass(fct_ctx* ctx, const double* M_dev);
/* override the lumped mass computed by fct_ctx_set_mass with the caller's M_lumped diagonal
 * (FCT_alg_ref takes M_lumped as an argument, helpers.py:1715) */
int fct_ctx_set_lumped(fct_ctx* ctx, const double* ML_dev);
int fct_assemble_static(fct_ctx* ctx);
int fct_ctx_static_dev(fct_ctx* ctx, const double** M_dev, const double** ML_dev, const double** Mdiag_dev,
                       const double** K_dev);
/* Structured numbering: declares that local row i is DoF g0 + i of dolfin's CG1 numbering on
 * RectangleMesh(n_cells x n_cells, diagonal "right") -- what fct_mesh_rect_build produces (advection_solidbody_FCT.py:48-50,82;
 * SURVEY.md App. B.2).  The pattern is verified on the device; when it holds (and the mass matrix has row templates) the
 * low-order Jacobi solve and ChebSI run K sweeps / iterations per launch on overlapped (diagonal, position) tiles, with
 * results bit-identical to the per-sweep kernels.  Other meshes simply keep the per-sweep kernels. */
int fct_ctx_set_rect(fct_ctx* ctx, int32_t n_cells, int64_t g0);
int fct_tiles_active(fct_ctx* ctx, int32_t* active_out);
/* test hook (host only, no device needed): interior origins (diagonal, position) of the tiles that cover rows
 * [row_begin,row_end) of a block starting at global DoF g0, for K fused passes; geom_out = {region diagonals, region positions} */
int fct_debug_tile_list(int32_t n_cells, int64_t g0, int32_t row_begin, int32_t row_end, int32_t K, int32_t* d0p0_out,
                        int32_t cap, int32_t* count_out, int32_t* geom_out);
/* solver options for the low-order system (defaults: rtol 1e-14, max_sweeps 200, check_every 2) */
int fct_ctx_set_solver(fct_ctx* ctx, double rtol, int32_t max_sweeps);

/* ---- device memory helpers (optional; any device allocation works, e.g. torch tensors) -------- */
int fct_malloc(fct_ctx* ctx, void** ptr_dev, int64_t bytes);
int fct_free(fct_ctx* ctx, void* ptr_dev);
int fct_h2d(fct_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes);   /* async on ctx stream */
int fct_d2h(fct_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes);   /* synchronises       */
/* page-locked host memory (so that the host-buffer entry points overlap copies with compute) */
int fct_host_alloc(fct_ctx* ctx, void** ptr_host, int64_t bytes);
int fct_host_free(fct_ctx* ctx, void* ptr_host);

/* ---- sparse kernels --------------------------------------------------------------------------- */
/* y = alpha * A x + beta * z   (z may be NULL when beta == 0; y may alias z).  scipy `A @ x`. */
int fct_spmv(fct_ctx* ctx, const double* A_dev, const double* x_dev, double alpha, double beta,
             const double* z_dev, double* y_dev);
/* ChebSI (helpers.py:143-185): exactly `iters` iterations of the Jacobi-preconditioned Chebyshev
 * recurrence for M y = b with eigenvalue bounds [lmin,lmax]; Md_dev is the `Md` argument (diag(M)). */
int fct_chebsi(fct_ctx* ctx, const double* M_dev, const double* Md_dev, const double* b_dev,
               double* y_dev, int32_t iters, double lmin, double lmax);
/* artificial_diffusion_mat (helpers.py:206-242): D_ij = max(0,-m_ij,-m_ji), D_ii = -sum_j D_ij. */
int fct_artificial_diffusion(fct_ctx* ctx, const double* mat_dev, double* D_dev);
/* row_lump (helpers.py:309-328): out_i = sum_j mat_ij. */
int fct_row_lump(fct_ctx* ctx, const double* mat_dev, double* out_dev);

/* ---- the FCT step ----------------------------------------------------------------------------- */
/* FCT_alg_ref (helpers.py:1715-1872), sign = +1:  [M + dt (A + S)] u+ = M u^n + dt rhs.
 * Legacy FCT_alg (old_helpers.py:115-204), sign = -1: the same with A -> -A.
 * S_dev (non_flux_mat / source_mat) and rhs_dev may be NULL.  Uses the context's M, ML.
 * info_host may be NULL (then nothing is copied back and the call does not synchronise). */
int fct_step(fct_ctx* ctx, const double* A_dev, double sign, const double* S_dev, const double* rhs_dev,
             const double* u_n_dev, double dt, double* u_out_dev, fct_step_info* info_host);
/* same call with HOST buffers (A_host[nnz], S_host[nnz]|NULL, rhs_host[n]|NULL, u_n_host[n] -> u_out_host[n]);
 * copies are inside the call.  This is what the FCT_alg_ref / FCT_alg Python shims use. */
int fct_step_host(fct_ctx* ctx, const double* A_host, double sign, const double* S_host, const double* rhs_host,
                  const double* u_n_host, double dt, double* u_out_host, fct_step_info* info_host);

/* ---- linear solvers for the second-species systems -------------------------------------------- */
/* spsolve replacements (helpers.py:596,686,1342,1538).  kind: 0 = Jacobi, 1 = Jacobi-PCG (SPD),
 * 2 = Jacobi-BiCGStab (nonsymmetric), 3 = CG preconditioned with a degree-8 Chebyshev polynomial of the Jacobi-scaled matrix
 * (SPD; FCT_CHEB_PCG_DEGREE overrides the degree; single GPU).  x_dev holds the initial guess on entry.  its_host/res_host
 * may be NULL. */
int fct_solve(fct_ctx* ctx, int32_t kind, const double* mat_dev, const double* b_dev, double* x_dev,
              double rtol, int32_t maxit, int32_t* its_host, double* res_host);
/* out = a*X + b*Y on the pattern values (Y may be NULL), e.g. M + dt*(Df*K + delta*M) */
int fct_vals_axpby(fct_ctx* ctx, double a, const double* X_dev, double b, const double* Y_dev, double* out_dev);

/* ---- norms / cost functional ------------------------------------------------------------------ */
/* x^T M y over owned rows (L2_norm_sq_Omega, helpers.py:362-381, when x == y) */
int fct_dot_M(fct_ctx* ctx, const double* M_dev, const double* x_dev, const double* y_dev, double* out_host);
/* L2_norm_sq_Q (helpers.py:330-360) of (phi - target) (target may be NULL): trapezoid in time.
 * phi/target are time-major trajectories [(num_steps+1) * n]. */
int fct_norm_sq_Q(fct_ctx* ctx, const double* M_dev, const double* phi_dev, const double* target_dev,
                  int32_t num_steps, double dt, double* out_host);
/* elementwise helpers on vectors / trajectories: out = clip(x + s*d, lo, hi); out = a*x + b*y */
int fct_clip_axpy(fct_ctx* ctx, int64_t len, const double* x_dev, double s, const double* d_dev,
                  double lo, double hi, double* out_dev);
int fct_axpby(fct_ctx* ctx, int64_t len, double a, const double* x_dev, double b, const double* y_dev,
              double* out_dev);

/* ---- element assembly into the fixed pattern (needs fct_ctx_set_mesh) ------------------------- */
/* Form kinds (i = test, j = trial):
 *   FCT_FORM_MASS          u v                         (helpers.py:553)
 *   FCT_FORM_STIFFNESS     grad u . grad v             (helpers.py:555)
 *   FCT_FORM_DRIFT         (b.grad c) u v + (b.grad v) c u   coef0 = c (P1), s0,s1 = b
 *                          (advection_solidbody_FCT_PDECO_alltime.py:222-223; old_helpers.py:62-63)
 *   FCT_FORM_WIND_P1       (w.grad v) u, w P1 nodal     coef0 = wx, coef1 = wy   (helpers.py:581)
 *   FCT_FORM_WIND_P1_T     (w.grad u) v                 coef0 = wx, coef1 = wy   (helpers.py:681)
 *   FCT_FORM_WMASS1/2/3    (f0 [f1 [f2]]) u v, P1 coefficient fields  (helpers.py:591,683,692,953,1032)
 *   FCT_FORM_CHTX          (grad f . grad v) u          coef0 = f   (old_helpers.py:102)
 *   FCT_FORM_CHTX_EXP      exp(-s0 m)(grad f . grad v) u, coef0 = f, coef1 = m, quadrature degree 4
 *                          (helpers.py:1350-1351)
 *   FCT_FORM_CHTX_ADJ      (1-s0 u)exp(-s0 u)(grad u_trial . grad vn) w, coef0 = vn, coef1 = u, degree 5
 *                          (helpers.py:1499-1500)
 *   FCT_FORM_WIND_POLY3    (w.grad v) u with an analytic wind that is a polynomial of degree <= 3 in (x,y):
 *   FCT_FORM_WIND_POLY3_T  (w.grad u) v   coef0 = 20 device doubles, wx then wy, monomial order
 *                          1,x,y,x^2,xy,y^2,x^3,x^2y,xy^2,y^3; integrated exactly (7-point degree-5 rule), which
 *                          is what dolfin does for Expression(..., degree=4) winds such as
 *                          helpers.py:506-508 (Schnakenberg) and :876-878 (nonlinear)
 *   FCT_FORM_DRIFT_MASS    (b.grad c) u v      the two parts of FCT_FORM_DRIFT on their own (the legacy scripts
 *   FCT_FORM_DRIFT_CONV    (b.grad v) c u      assemble them separately: advection_solidbody_FCT_PDECO_alltime.py:222-223)
 *   FCT_FORM_DIVW_MASS     div(w_h) u v, w_h P1 nodal   coef0 = wx, coef1 = wy: together with FCT_FORM_WIND_P1_T the legacy
 *                          div(w_h u) v of a projected wind (Schnak_FCT_PDECO.py:70,242,256)
 * out_vals = scale * form (+ out_vals if accumulate != 0). */
enum {
    FCT_FORM_MASS = 0, FCT_FORM_STIFFNESS = 1, FCT_FORM_DRIFT = 2, FCT_FORM_WIND_P1 = 3,
    FCT_FORM_WIND_P1_T = 4, FCT_FORM_WMASS1 = 5, FCT_FORM_WMASS2 = 6, FCT_FORM_WMASS3 = 7,
    FCT_FORM_CHTX = 8, FCT_FORM_CHTX_EXP = 9, FCT_FORM_CHTX_ADJ = 10, FCT_FORM_WIND_POLY3 = 11,
    FCT_FORM_WIND_POLY3_T = 12, FCT_FORM_DRIFT_MASS = 13, FCT_FORM_DRIFT_CONV = 14, FCT_FORM_DIVW_MASS = 15
};
int fct_assemble_matrix(fct_ctx* ctx, int32_t kind, const double* coef0_dev, const double* coef1_dev,
                        const double* coef2_dev, double s0, double s1, double scale, int32_t accumulate,
                        double* out_vals_dev);
/* Linear forms:
 *   FCT_LOAD_P1_1/2/3/4    (f0 [f1 [f2 [f3]]]) v       (helpers.py:584-585,684,693,956,1339-1340)
 *   FCT_LOAD_CONST         s0 * v                       (helpers.py:594)
 *   FCT_LOAD_DRIFT_GRAD    p (b.grad u) v, coef0 = p, coef1 = u, s0,s1 = b
 *                          (advection_solidbody_FCT_PDECO_alltime.py:273)
 *   FCT_LOAD_CHTX_ADJ      s1 * u exp(-s0 u) grad p . grad w, coef0 = p, coef1 = u, degree 4 (helpers.py:1531-1532)
 *   FCT_LOAD_POLY3         p(x,y) v with a polynomial of degree <= 3, coef0 = 10 device doubles (monomial order as for
 *                          FCT_FORM_WIND_POLY3): the right-hand sides of project(wind, W) (Schnak_FCT_PDECO.py:70,242)
 * out = scale * form (+ out if accumulate). */
enum {
    FCT_LOAD_P1_1 = 0, FCT_LOAD_P1_2 = 1, FCT_LOAD_P1_3 = 2, FCT_LOAD_P1_4 = 3, FCT_LOAD_CONST = 4,
    FCT_LOAD_DRIFT_GRAD = 5, FCT_LOAD_CHTX_ADJ = 6, FCT_LOAD_POLY3 = 7
};
int fct_assemble_vector(fct_ctx* ctx, int32_t kind, const double* coef0_dev, const double* coef1_dev,
                        const double* coef2_dev, const double* coef3_dev, double s0, double s1, double scale,
                        int32_t accumulate, double* out_dev);

/* ---- device-resident time loops of the three PDE systems of the refactored API --
 * solve_nonlinear_equation / solve_adjoint_nonlinear_equation (helpers.py:881-1038), solve_schnak_system /
 * solve_adjoint_schnak_system (:511-698), solve_chtxs_system / solve_adjoint_chtxs_system (:1250-1581) on device
 * trajectories [(num_steps+1) * n], time-major.  Level 0 of a state trajectory holds the initial condition on entry,
 * level num_steps of an adjoint trajectory the terminal condition.  control_dev: the ONE control vector the reference
 * builds from the first step's slice and reuses (App. D-1), or NULL for the constant control_const.  The model
 * parameters of get_*_params are arguments: params6 = {Du, Dv, c_b, gamma, omega1, omega2}, params5 = {delta, Dm, Df,
 * chi, eta}, wind20 = polynomial wind coefficients as for FCT_FORM_WIND_POLY3 (host arrays).  Single GPU.
 * total_sweeps_host (may be NULL): Jacobi sweeps of all FCT steps. */
int fct_forward_nonlinear(fct_ctx* ctx, const double* control_dev, double control_const, double* var1_traj_dev,
                          int32_t num_steps, double dt, double eps, const double* wind20_host, int32_t* total_sweeps_host);
int fct_adjoint_nonlinear(fct_ctx* ctx, const double* u_traj_dev, double* p_traj_dev, int32_t num_steps, double dt,
                          double eps, const double* wind20_host, int32_t* total_sweeps_host);
int fct_forward_schnak(fct_ctx* ctx, const double* control_dev, double control_const, double* var1_traj_dev,
                       double* var2_traj_dev, int32_t num_steps, double dt, const double* params6_host,
                       const double* wind20_host, double rescaling, int32_t* total_sweeps_host);
int fct_adjoint_schnak(fct_ctx* ctx, const double* u_traj_dev, const double* v_traj_dev, double* p_traj_dev,
                       double* q_traj_dev, int32_t num_steps, double dt, const double* params6_host,
                       const double* wind20_host, int32_t* total_sweeps_host);
int fct_forward_chtxs(fct_ctx* ctx, const double* control_dev, double control_const, double* var1_traj_dev,
                      double* var2_traj_dev, int32_t num_steps, double dt, const double* params5_host, double rescaling,
                      int32_t* total_sweeps_host);
/* uhat/vhat_traj_dev: both NULL ("finaltime") or both given ("alltime": nodal tracking terms, helpers.py:1509,1535) */
int fct_adjoint_chtxs(fct_ctx* ctx, const double* u_traj_dev, const double* v_traj_dev, const double* uhat_traj_dev,
                      const double* vhat_traj_dev, double* p_traj_dev, double* q_traj_dev, const double* control_traj_dev,
                      int32_t num_steps, double dt, const double* params5_host, double rescaling,
                      int32_t* total_sweeps_host);

/* ---- device-resident time loops of the drift-control advection PDECO ------------------------------
 * (advection_solidbody_FCT_PDECO_alltime.py:210-275, the shape of the 4096^2 benchmark).  Trajectories
 * are time-major device arrays [(num_steps+1) * n]; slice 0 of u must hold the initial condition. */
int fct_advdrift_state(fct_ctx* ctx, const double* c_traj_dev, double* u_traj_dev, int32_t num_steps, double dt,
                       double bx, double by, double eps, int32_t* total_sweeps_host);
int fct_advdrift_adjoint(fct_ctx* ctx, const double* c_traj_dev, const double* u_traj_dev,
                         const double* uhat_traj_dev, double* p_traj_dev, int32_t num_steps, double dt,
                         double bx, double by, double eps, int32_t* total_sweeps_host);
int fct_advdrift_gradient(fct_ctx* ctx, const double* c_traj_dev, const double* u_traj_dev,
                          const double* p_traj_dev, double* d_traj_dev, int32_t num_steps, double beta,
                          double bx, double by);
/* the same forward loop with HOST trajectories (pinned or pageable): c slices are streamed in and u slices
 * streamed out on a copy stream, overlapped with compute.  This is the end-to-end path bench.py times. */
int fct_advdrift_state_host(fct_ctx* ctx, const double* c_traj_host, double* u_traj_host, int32_t num_steps,
                            double dt, double bx, double by, double eps, int32_t* total_sweeps_host);

/* ---- multi-GPU (row-block partition, one-ring halo; one process per GPU) ----------------------- */
int fct_nccl_unique_id(void* id_out_128bytes);                       /* rank 0 calls, broadcasts the bytes */
int fct_ctx_init_comm(fct_ctx* ctx, const void* id_128bytes, int32_t rank, int32_t world,
                      int32_t send_lo_begin, int32_t send_lo_end,    /* local rows sent to rank-1          */
                      int32_t send_hi_begin, int32_t send_hi_end);   /* local rows sent to rank+1          */
int fct_halo_exchange(fct_ctx* ctx, double* vec_dev);                /* refresh halo entries of a vector    */
/* deep halos: ring j = local rows within j mesh rings of the owned rows (ring 0 = [row_begin,row_end), ring `depth`
 * = [0,n)).  With depth k the dependent passes of the Jacobi / Chebyshev loops exchange every k-th pass only. */
int fct_ctx_set_rings(fct_ctx* ctx, int32_t depth, const int32_t* ring_lo, const int32_t* ring_hi);
/* NVLink peer-memory mailboxes (CUDA IPC) replacing NCCL send/recv for the halo exchanges and the Jacobi stopping
 * test: fct_p2p_create exports this rank's region (64-byte cudaIpcMemHandle); after gathering all ranks' handles
 * (rank order) fct_p2p_connect maps them.  Once connected every halo exchange is one kernel (push over NVLink, wait,
 * unpack) and the multi-GPU Jacobi loop runs as a CUDA-graph WHILE node.  fct_p2p_error reports a timed-out wait. */
int fct_p2p_create(fct_ctx* ctx, int32_t rank, int32_t world, int32_t max_halo, void* handle_out_64bytes);
int fct_p2p_connect(fct_ctx* ctx, const void* all_handles_world_x_64bytes);
int fct_p2p_error(fct_ctx* ctx, int32_t* error_host);

/* ---- instrumentation ---------------------------------------------------------------------------- */
/* times `reps` back-to-back Jacobi sweeps of the low-order system of (A, u_n, dt) with CUDA events on the context's
 * stream (the dominant kernel of the FCT step once the mass matrix runs on row templates); bench.py's roofline */
int fct_bench_jacobi_sweeps(fct_ctx* ctx, const double* A_dev, const double* u_n_dev, double dt, int32_t reps,
                            float* ms_per_sweep_host);
/* the same system through the overlapped-tile kernel (fct_tile.cu): `reps` launches of `sweeps` (2..4) fused Jacobi sweeps
 * each; returns the CUDA-event time per sweep.  Fails if the context has no tile kernels (fct_ctx_set_rect not called or
 * not verified, no row templates, FCT_NO_TILES=1). */
int fct_bench_jacobi_fused(fct_ctx* ctx, const double* A_dev, const double* u_n_dev, double dt, int32_t sweeps,
                           int32_t reps, float* ms_per_sweep_host);
/* test hook: exactly `sweeps` Jacobi sweeps on the low-order system of (A, u_n, dt) from the initial guess u_n, one launch
 * per sweep (fused = 0) or as tile launches of `fused` (2..4) sweeps each; the iterate is copied to x_out_dev */
int fct_debug_jacobi_fixed(fct_ctx* ctx, const double* A_dev, const double* u_n_dev, double dt, int32_t sweeps,
                           int32_t fused, double* x_out_dev);
/* number of row templates the static mass matrix compressed to (0 = CSR kernels in use) */
int fct_template_count(fct_ctx* ctx, int32_t* count);
/* number of geometry templates the mesh compressed to (0 = generic assembly kernels in use) */
int fct_geom_template_count(fct_ctx* ctx, int32_t* count);
/* CUDA events on the context's stream (what bench.py times kernels with) */
int fct_event_create(fct_ctx* ctx, void** event_out);
int fct_event_record(fct_ctx* ctx, void* event);
int fct_event_elapsed_ms(fct_ctx* ctx, void* start, void* stop, float* ms_host);   /* synchronises on stop */
int fct_event_destroy(fct_ctx* ctx, void* event);
/* cudaProfilerStart (start != 0) / cudaProfilerStop of the CUDA runtime inside this library: delimits the region an
 * `ncu --profile-from-start off` capture records (tools/kprof_step.py) */
int fct_profiler_range(int32_t start);
/* number of kernels this library has launched on this context since creation */
int fct_launch_count(fct_ctx* ctx, int64_t* count);
/* number of halo exchanges executed on this context since creation (multi-GPU; exchanges inside CUDA-graph bodies included
 * on the peer-memory path) */
int fct_exchange_count(fct_ctx* ctx, int64_t* count);
/* Checked allocator (FCT_GUARD=1 in the environment: canary bands around, and a 0xFF fill of, every device buffer of the
 * library).  *corrupted = buffers, live now or freed since the last call, whose canaries were overwritten; *live = buffers
 * checked.  Both 0 when the mode is off.  The test suite asserts corrupted == 0 after every GPU test. */
int fct_guard_check(int64_t* corrupted, int64_t* live);

#ifdef __cplusplus
}
#endif
#endif /* FCTPDECO_H */
