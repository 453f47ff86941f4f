/* cmpc.h — C ABI of libcmpc_b200.so, the B200 (sm_100a) SCP hot path of ahmadgazar/centroidal-MPC.
 *
 * The reference has no FFI: its hot path is Python calling JAX (XLA) and OSQP (C).  Each entry
 * point below names the reference interface it replaces (paths relative to the reference
 * repository root).  Plain C types only; every array is a raw pointer + the sizes in cmpc_dims.
 * Unless a function says "host", pointers are DEVICE pointers owned by the caller (e.g.
 * torch.Tensor.data_ptr()); the library allocates only its workspace, inside cmpc_create.
 *
 * Array layouts (row-major, instance-major; B = batch, N = horizon, nc = contacts, nu = 3*nc):
 *   x_init, x_final   [B][9]        (src/centroidal_model.py:87-89)
 *   X_ref             [B][N+1][9]   model._init_trajectories['state'] transposed (:174,:185)
 *   U_init            [B][N][nu]    model._init_trajectories['control'] transposed (:176-186)
 *   contact_pos       [Bp][N][nc][3]   model._contact_data['contacts_position'] (:127-156)
 *   contact_R         [Bp][N][nc][3][3] model._contact_data['contacts_orient']; NULL = identity
 *   contact_active    [Bp][N][nc]   int32, model._contact_data['contacts_logic']
 *       Bp = 1 when dims.shared_plan != 0 (one contact plan for the whole batch), else B
 *   X_out             [B][N+1][9]   last accepted state trajectory  (all_solution['state'][-1].T)
 *   U_out             [B][N][nu]    last accepted control trajectory (all_solution['control'][-1].T)
 *
 * Return value: 0 on success, negative on a usage / CUDA error (message via cmpc_last_error()).
 * Per-instance solver outcomes are reported in status[] and never through the return value.
 */
#ifndef CMPC_H
#define CMPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cmpc_handle_s* cmpc_handle;

typedef struct {
  int32_t batch;        /* B  */
  int32_t N;            /* horizon (knots of control) */
  int32_t nc;           /* contacts: 4 (solo12) or 2 (bolt); point-contact model, n_u = 3*nc */
  int32_t shared_plan;  /* 1: contact arrays have leading dimension 1 */
  int32_t contact_model; /* CMPC_CONTACT_POINT (0): n_u = 3*nc, controls (fx,fy,fz) per contact (solo12, bolt);
                            CMPC_CONTACT_WRENCH (1): flat feet, nc <= 2, n_u = 6*nc, controls (cop_x,cop_y,fx,fy,fz,tau_z)
                            per foot in the reference's order (TALOS: src/centroidal_model.py:204-208,
                            src/optimizer.py:48-64); contact_R is then mandatory, the CoP box rows
                            (src/constraints.py:111-145) come from cmpc_model.foot_range, and the cost has no tracking
                            gradient (src/scp_solver.py:13-20) */
} cmpc_dims;
enum { CMPC_CONTACT_POINT = 0, CMPC_CONTACT_WRENCH = 1 };

/* conf attributes copied by Centroidal_model.__init__ (src/centroidal_model.py:27-32) */
typedef struct {
  double robot_mass, gravity_constant, dt, mu;
  double state_cost_weights[9];     /* diagonal of conf.state_cost_weights   */
  double control_cost_weights[12];  /* diagonal of conf.control_cost_weights */
  double foot_range[4];             /* wrench model only: conf.robot_foot_range, -[1] <= cop_x <= [0], -[3] <= cop_y <= [2]
                                       (src/constraints.py:127-137) */
} cmpc_model;

/* conf.scp_params, keys read at src/scp_solver.py:120-128 */
typedef struct {
  double trust_region_radius0, omega0, omega_max, rho0, rho1;
  double beta_succ, beta_fail, gamma_fail, convergence_threshold;
  int32_t max_iterations;
} cmpc_scp_params;

/* QP settings; the defaults reproduce the reference's OSQP call (src/scp_solver.py:61-63:
 * eps_abs = eps_rel = 1e-7, polish on) with OSQP's own defaults where the setting has the same
 * meaning.  The device QP solver is an ADMM whose x-update treats the dynamics exactly (Riccati),
 * interleaved with OSQP-style polishes that are accepted early when they certify a KKT point. */
typedef struct {
  double eps_abs, eps_rel;
  double sigma;                  /* accepted for OSQP compatibility, unused: W_x, W_u > 0 make the
                                    x-update strictly convex without a proximal term */
  double alpha /* relaxation; default 1.8 (OSQP: 1.6), 1.6 for the wrench model */, rho, delta, adaptive_rho_tolerance;
  int32_t max_iter, check_termination /* default 25 as in OSQP; every polish attempt checks too */, polish,
      polish_refine_iter, adaptive_rho;
  int32_t adaptive_rho_start;    /* first ADMM iteration at which rho may be adapted */
  int32_t polish_active_set_rounds; /* extra polish rounds with a corrected active set (0 = OSQP) */
  int32_t active_set_start;      /* ADMM iteration of the first early polish (0 = only after termination) */
  int32_t active_set_step;       /* ADMM iterations between early polishes */
  double active_set_tol;         /* certificate: primal residual / row violation <= tol * (1 + norm), the form of
                                    OSQP's test (default 1e-9: 100x tighter than the reference's eps = 1e-7) */
  double warm_start_tol;         /* warm_start: a friction / CoP row of the warm start U_init with G u - ub >= -tol (1 + |ub|)
                                    starts in the active set (default 1e-7) */
  int32_t warm_start;            /* 1: receding-horizon use -- U_init is (the shift of) a previous solution: the first QP
                                    of the solve starts with a certified polish on the active set read off U_init, before
                                    any ADMM iteration; when the certificate fails, ADMM runs as in a cold solve.  The
                                    answer is the same certified KKT point either way (default 0) */
} cmpc_qp_settings;

/* status[] values */
enum { CMPC_OK = 0, CMPC_QP_MAX_ITER = 1, CMPC_QP_NUMERIC = 2,
       CMPC_DEVICE_ERROR = 3 /* a staged copy of the instance's tile never completed; the other tiles are unaffected */ };

void cmpc_default_qp_settings(cmpc_qp_settings* s);

/* Allocates the per-batch solver workspace on the current CUDA device. */
int cmpc_create(const cmpc_dims* dims, cmpc_handle* out);
int cmpc_destroy(cmpc_handle h);
int64_t cmpc_workspace_bytes(cmpc_handle h);

/* Binds problem data (device pointers are kept, not copied).
 * Replaces Centroidal_model construction as seen by the solver: src/centroidal_model.py:15-47. */
int cmpc_set_problem(cmpc_handle h, const cmpc_model* model, const double* x_init, const double* x_final,
                     const double* X_ref, const double* U_init, const double* contact_pos,
                     const double* contact_R, const int32_t* contact_active);

/* Stochastic mode (Centroidal_model(conf, STOCHASTIC_OCP=True), src/constraints.py:157-163,187-214):
 * binds the upper bounds of the friction-pyramid rows, friction_ub [B][N][nc][4] (device pointer, kept,
 * not copied; the output of cmpc_friction_backoffs), used by every following cmpc_solve_scp /
 * cmpc_solve_scp_host on this handle.  NULL returns to the nominal rows G f <= 0.
 * Recommended cmpc_qp_settings with upper bounds: polish_refine_iter = 10, polish_active_set_rounds = 19
 * (the defaults are correct but leave some instances to several polish attempts). */
int cmpc_set_friction_ub(cmpc_handle h, const double* friction_ub);

/* solve_scp(model, scp_params) for the whole batch: src/scp_solver.py:118-179, including
 * compute_trajectory_data (src/centroidal_model.py:257-291), the QP assembly
 * (src/cost.py:9-39, src/constraints.py:12-50,104-109,153-185,260-293), the OSQP solve
 * (src/scp_solver.py:59-68) and the trust-region logic.  Asynchronous on `stream`
 * (a cudaStream_t, may be NULL); no host synchronisation inside.
 * scp_iters[b]: SCP iterations executed; status[b]: CMPC_OK, or the QP failure that makes the
 * reference return False (:146-148).  n_accepted (nullable): accepted solutions (0 => the
 * reference returns empty lists, :119,:179). */
int cmpc_solve_scp(cmpc_handle h, const cmpc_scp_params* scp, const cmpc_qp_settings* qp, double* X_out,
                   double* U_out, int32_t* scp_iters, int32_t* status, int32_t* n_accepted, void* stream);

/* Same with HOST buffers: copies inputs to the device, solves, copies the results back and
 * synchronises.  This is the end-to-end call a host-side caller of solve_scp would make. */
int cmpc_solve_scp_host(cmpc_handle h, const cmpc_model* model, const cmpc_scp_params* scp,
                        const cmpc_qp_settings* qp, const double* x_init, const double* x_final,
                        const double* X_ref, const double* U_init, const double* contact_pos,
                        const double* contact_R, const int32_t* contact_active, double* X_out,
                        double* U_out, int32_t* scp_iters, int32_t* status, int32_t* n_accepted);

/* Per-instance statistics of the last solve (device pointers, each nullable):
 * qp_iters[B] total ADMM iterations, n_factor[B] Riccati factorisations, info[B][12] =
 * {sigma_max(X-Xbar), accuracy ratio, primal res, dual res, rho, radius, weight, polished,
 *  multiplier-method sweeps, polish attempts, certified, 0}.  certified = 1: the last QP ended at a certified KKT
 * point (active set consistent, primal residual <= active_set_tol (1 + norm), stationarity exact); 0: it ended by OSQP's
 * termination test at eps_abs / eps_rel, i.e. with the accuracy of the reference's own solver setting.  Asynchronous copies on `stream` (pass the stream
 * of the solve, so that they are ordered after it). */
int cmpc_get_stats(cmpc_handle h, int32_t* qp_iters, int32_t* n_factor, double* info, void* stream);

/* compute_trajectory_data (src/centroidal_model.py:257-291): f, A=df/dx, B=df/du along (X,U).
 * X [B][N+1][9], U [B][N][nu] -> f [B][N][9], fx [B][N][9][9], fu [B][N][9][nu]. */
int cmpc_linearize(const cmpc_dims* dims, const cmpc_model* model, const double* X, const double* U,
                   const double* contact_pos, const int32_t* contact_active, double* f, double* fx,
                   double* fu, void* stream);

/* integrate_dynamics_trajectory (src/centroidal_model.py:243-255): f(x_k,u_k), k < N -> f [B][N][9]. */
int cmpc_rollout(const cmpc_dims* dims, const cmpc_model* model, const double* X, const double* U,
                 const double* contact_pos, const int32_t* contact_active, double* f, void* stream);

/* conf.Q, conf.R, conf.cov_w, conf.cov_white_noise (src/centroidal_model.py:34-35,41-42): dense,
 * row-major; R and cov_w are nu x nu with leading dimension nu (the first nu*nu entries are read). */
/* Wrench contact model (dims.contact_model = CMPC_CONTACT_WRENCH): f(x_k,u_k) and, when fx and fu are not
   null, the Jacobians A_k [B][N][9][9], B_k [B][N][9][6*nc] along (X, U) in the reference's control order
   -- integrate_model_one_step / jacfwd for robot == 'TALOS', /root/reference/src/centroidal_model.py:
   189-212,229-231,243-255.  contact_R [Bp][N][nc][3][3] row-major. */
int cmpc_linearize_wrench(const cmpc_dims* dims, const cmpc_model* model, const double* X, const double* U,
                          const double* contact_pos, const double* contact_R, const int32_t* contact_active, double* f,
                          double* fx, double* fu, void* stream);
typedef struct {
  double Q[81], R[144], cov_w[144], cov_eta[81];
} cmpc_lqr_weights;
#define CMPC_LQR_SCRATCH_BYTES 4096

/* LQR feedback gains and state covariances along (X,U): the `LQR_gains` and `Covs` entries of
 * compute_trajectory_data (src/centroidal_model.py:215-227,233-238,284-285) that solve_scp hands
 * back as all_solution['gains'] / ['covs'] (src/scp_solver.py:165-166).
 * gains [B][N][nu][9]; covs [B][N+1][9][9] with covs[b][0] = 0 (nullable: gains only).
 * `w` is a HOST pointer; `scratch` is CMPC_LQR_SCRATCH_BYTES of caller-owned DEVICE memory.
 * A knot whose R + B'PB is not positive definite gets NaN gains; cmpc_friction_backoffs turns those into NaN
 * bounds, so that the instance's solve ends with status != 0 instead of running without back-offs. */
int cmpc_lqr_covs(const cmpc_dims* dims, const cmpc_model* model, const cmpc_lqr_weights* w, const double* X,
                  const double* U, const double* contact_pos, const int32_t* contact_active, double* gains,
                  double* covs, void* scratch, void* stream);

/* Stochastic mode (Centroidal_model(conf, STOCHASTIC_OCP=True)): upper bounds of the friction-pyramid
 * rows with the chance-constraint back-offs of construct_friction_pyramid_constraints
 * (src/constraints.py:157-163,187-214): friction_ub [B][N][nc][4] =
 * - sum_u xi 2 G_ju sqrt((K_c Sigma_k K_c')_uu) over the entries with G_ju > 1e-6 and sqrt(.) > 1e-6, zero at
 * k = 0 and for inactive contacts.  xi = Phi^-1(1 - beta_u/5*3) is computed by the caller (:157);
 * gains / covs are the outputs of cmpc_lqr_covs; contact_R as in cmpc_set_problem (NULL = identity). */
/* The same for the wrench contact model (dims.contact_model = CMPC_CONTACT_WRENCH): R, cov_w with leading dimension
   n_u = 6*nc; the contact-position noise has three components per foot, so cov_w is conf.cov_w (3 nc x 3 nc) in the
   leading block of the 12 x 12 array.  contact_R [Bp][N][nc][3][3]. */
int cmpc_lqr_covs_wrench(const cmpc_dims* dims, const cmpc_model* model, const cmpc_lqr_weights* w, const double* X,
                         const double* U, const double* contact_pos, const double* contact_R, const int32_t* contact_active,
                         double* gains, double* covs, void* scratch, void* stream);
int cmpc_friction_backoffs(const cmpc_dims* dims, const cmpc_model* model, double xi, const double* gains,
                           const double* covs, const double* contact_R, const int32_t* contact_active,
                           double* friction_ub, void* stream);

/* Diagnostics of a profiling build of the library (-DCMPC_PROFILE, scripts/variant.sh prof): cycle counters per
 * operation kind summed over all tiles since the last call, out32[32] (layout: scripts/prof_cycles.py).
 * Returns -1 in a normal build.  Not part of the reference's interface. */
/* Multi-GPU (one process per GPU, instances sharded, no collective on the solve path): instead of gathering the
   solutions afterwards, every rank can hand cmpc_solve_scp result pointers INTO A BUFFER OF THE DESTINATION RANK'S
   GPU; the kernel's write-back then goes over NVLink as the tiles finish (posted stores, like the mapped host
   buffers of cmpc_solve_scp_host) and no gather kernel competes with the next solve for the SMs.
   cmpc_peer_alloc: the destination rank allocates the buffer and gets a 64-byte handle to send to the other
   processes (any channel; bench.py uses torch.distributed.broadcast_object_list); cmpc_peer_open: the other
   ranks map it (peer access is enabled by the driver) and get a device pointer valid in THEIR process; the caller
   offsets it to its shard.  The destination reads the buffer after all ranks have synchronised their streams.
   No counterpart in the reference (it solves one problem per call on the CPU). */
int cmpc_peer_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64);
int cmpc_peer_open(const unsigned char* handle64, void** dev_ptr);
int cmpc_peer_close(void* dev_ptr);
int cmpc_peer_free(void* dev_ptr);
int cmpc_debug_profile(double* out32);

/* DFMA micro-benchmark on the current device: achieved FP64 TFLOP/s and the SM clock (MHz) seen. */
int cmpc_fp64_peak(double* tflops, double* ms);

/* Counts kernel launches issued by this library since load (for bench.py's gpu_launches). */
int64_t cmpc_launch_count(void);

const char* cmpc_last_error(void);
const char* cmpc_version(void);
/* SHA-256 (first 16 hex digits) of the sources this library was built from, stamped by __graft_entry__.build();
 * "unstamped" for a hand build.  bench.py matches it against profiles/r2_traffic.json. */
const char* cmpc_build_id(void);

#ifdef __cplusplus
}
#endif
#endif /* CMPC_H */
