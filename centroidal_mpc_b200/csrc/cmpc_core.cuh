// cmpc_core.cuh — the per-instance SCP / Riccati-ADMM solver, one warp per MPC instance.
//
// What this replaces in the reference (paths relative to /root/reference):
//   src/centroidal_model.py:189-232,257-291   dynamics + Jacobians  -> linearize_knot()
//   src/cost.py:9-39, src/constraints.py:12-50,104-109,153-185,260-293, src/scp_solver.py:10-48
//                                             QP assembly           -> never materialised: the
//                                             stage records below ARE the block-banded KKT data
//   src/scp_solver.py:59-68 (OSQP)            QP solve              -> admm_solve() + polish()
//   src/scp_solver.py:71-87,151               accuracy ratio, spectral trust test -> evaluate()
//   src/scp_solver.py:118-179                 trust-region loop     -> solve_instance()
//
// Execution model.  All code here is warp-uniform "driver" code with lane-parallel PHASES:
//     CMPC_LANES(l) { ... work of lane l ... }  CMPC_SYNC();
// Within a phase no lane reads what another lane writes; everything that crosses a phase
// boundary lives in the per-warp shared-memory struct WarpMem or in global memory.  On the GPU
// CMPC_LANES runs once with l = lane id and CMPC_SYNC is __syncwarp(); compiled for the host
// (tests/emu, a test-only build; never part of libcmpc_b200.so) CMPC_LANES is a loop over
// 32 lanes, which executes the identical arithmetic in the identical order.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CMPC_HD __host__ __device__ __forceinline__
#else
#define CMPC_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define CMPC_LANES(l) for (int l = (int)(threadIdx.x & 31u), _once = 1; _once; _once = 0)
#define CMPC_SYNC() __syncwarp()
#else
#define CMPC_LANES(l) for (int l = 0; l < 32; ++l)
#define CMPC_SYNC() ((void)0)
#endif

namespace cmpc {

constexpr int NX = 9;
constexpr int MAXC = 4;    // contacts
constexpr int MAXU = 12;   // 3 * MAXC
constexpr int STG = 40;    // stage record   : q[9] c[9] S[3] kbar[3] d[12] actmask pad[3]
constexpr int STA = 40;    // iterate record : x[9] u[12] vf[16] vk[3]
constexpr int FAC = 264;   // factor record  : K[na*9] @0, Hinv[na*na] @108, Pc[9] @252
constexpr int DVC = 12;    // feed-forward d_k (compact)
constexpr int POL = 20;    // polish record  : yf[16] yk[4]
constexpr int O_Q = 0, O_C = 9, O_S = 18, O_KB = 21, O_D = 24, O_ACT = 36;
constexpr int O_X = 0, O_U = 9, O_VF = 21, O_VK = 37;
constexpr int O_K = 0, O_HI = 108, O_PC = 252;

enum Status { ST_OK = 0, ST_QP_MAXITER = 1, ST_QP_NUMERIC = 2 };

struct Params {
  int N, nc, nu, identity_R;
  double m, g, dt, mu;
  double Wx[NX], Wu[MAXU];
  // QP solver settings (OSQP's where they have the same meaning; scp_solver.py:61-63)
  double sigma, alpha, rho0, eps_abs, eps_rel, delta, adapt_tol, rho_e_rel, rho_k_rel, rho_e_pol_rel;
  int max_iter, check_every, polish, refine, adaptive_rho, adapt_start, polish_rounds;
  // SCP parameters (scp_solver.py:120-128)
  double radius0, omega0, omega_max, acc_rho0, acc_rho1, beta_succ, beta_fail, gamma_fail, conv_thresh;
  int max_scp;
};

// Per-batch global-memory views (all device pointers on the GPU; plain host pointers in the
// emulation).  `plan_stride` is 0 when the whole batch shares one contact plan.
struct Batch {
  int B;
  const double* x_init;   // [B][9]
  const double* x_final;  // [B][9]
  const double* X_ref;    // [B][N+1][9]
  const double* U_init;   // [B][N][nu]
  const double* cpos;     // [Bp][N][nc][3]
  const double* cR;       // [Bp][N][nc][9]
  const int* cact;        // [Bp][N][nc]
  long plan_stride;       // 0 or 1 (multiplies the per-instance plan size)
  // workspace
  double* stg;            // [B][N+1][STG]
  double* sta;            // [B][N+1][STA]
  double* sta2;           // [B][N+1][STA]  (ADMM iterate kept while polishing)
  double* fac;            // [B][N][FAC]
  double* dvec;           // [B][N][DVC]
  double* pol;            // [B][N+1][POL]
  int* pmask;             // [B][N+1]
  // outputs
  double* X_out;          // [B][N+1][9]
  double* U_out;          // [B][N][nu]
  int* scp_iters;         // [B]
  int* status;            // [B]
  int* n_accepted;        // [B]
  int* qp_iters;          // [B] total ADMM iterations
  int* n_factor;          // [B]
  double* info;           // [B][8]: snorm, acc ratio, pri, dua, rho, radius, weight, polished
};

// Per-warp shared memory, 7.6 KB: 28 warps (instances) per SM fit in 227 KB.
struct WarpMem {
  // ---- factor phase scratch; the solve phase reuses [fK|fHuu] as the loaded K and Hinv
  double P[81];
  double PB[2 * 9 * MAXU];   // [0,108): (P B)^T ; [108,216): H_ux.  Outside factor(): the
                             // four 32-lane reduction slots red[slot][lane] = PB[32*slot+lane]
  double Huu[MAXU * MAXU];   // H_uu, inverted in place  (solve phase: loaded Hinv)
  double K[MAXU * 9];        // (solve phase: loaded K)
  double T[81];
  double prow[MAXU], pcol[MAXU];
  // ---- per-knot records
  double stg[STG];
  double sta[STA];
  double G[MAXC][12];        // friction rows (4x3 per contact), contact frame -> world
  double rrow[16], lrow[16]; // per friction row: penalty, linear coefficient
  double kM[9], kl[3];       // kappa block: 3x3 penalty matrix and linear term
  double ef2[16];            // friction row equilibration factors e^2
  // ---- vectors
  double p[9], g[9], hu[MAXU], dd[MAXU], xk[9], xn[9], ut[MAXU], Pc[9], ye[9], lam[9];
  double sc[16];             // scalars shared between lanes
};

// scalars in WarpMem::sc
enum { SC_RHO = 0, SC_RADIUS, SC_WEIGHT, SC_PRI, SC_DUA, SC_NPRI, SC_NDUA, SC_NUM, SC_DEN, SC_SNORM,
       SC_RHOE, SC_RHOK, SC_RHOEP };

struct Ctx {
  const Params* prm;
  Batch bt;
  WarpMem* s;
  int b;   // instance
  const double* cpos; const double* cR; const int* cact;   // this instance's plan
  double* stg; double* sta; double* sta2; double* fac; double* dvec; double* pol; int* pmask;
};

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
CMPC_HD void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// warp copy global -> shared (n doubles)
#define CMPC_COPY(dst, src, n)                                             \
  do {                                                                     \
    CMPC_LANES(l_) {                                                       \
      for (int e_ = l_; e_ < (n); e_ += 32) (dst)[e_] = (src)[e_];         \
    }                                                                      \
  } while (0)

// max-reduce red[slot][0..31] into sc[dst] (called between phases)
#define CMPC_REDUCE_MAX(ctx, slot, dst)                                    \
  do {                                                                     \
    CMPC_SYNC();                                                           \
    CMPC_LANES(l_) {                                                       \
      if (l_ == 0) {                                                       \
        double m_ = 0.0;                                                   \
        for (int e_ = 0; e_ < 32; ++e_) m_ = fmax(m_, (ctx).s->PB[32 * (slot) + e_]); \
        (ctx).s->sc[dst] = m_;                                             \
      }                                                                    \
    }                                                                      \
    CMPC_SYNC();                                                           \
  } while (0)

#define CMPC_REDUCE_SUM(ctx, slot, dst)                                    \
  do {                                                                     \
    CMPC_SYNC();                                                           \
    CMPC_LANES(l_) {                                                       \
      if (l_ == 0) {                                                       \
        double m_ = 0.0;                                                   \
        for (int e_ = 0; e_ < 32; ++e_) m_ += (ctx).s->PB[32 * (slot) + e_];      \
        (ctx).s->sc[dst] = m_;                                             \
      }                                                                    \
    }                                                                      \
    CMPC_SYNC();                                                           \
  } while (0)

// ------------------------------------------------------------------------------------------
// K1: dynamics, closed-form Jacobian data and affine residual for one knot
//     (centroidal_model.py:189-232).  Point-contact model.
//     A_k = I + dt [[0, I/m, 0],[0,0,0],[[S]x,0,0]],  S = sum_i a_i fbar_i
//     B_k[:,3i:3i+3] = dt a_i [0; I; [d_i]x],          d_i = p_i - cbar
//     c_k = fbar - A xbar - B ubar = [0; dt m g e_z; -dt S x cbar]
// ------------------------------------------------------------------------------------------
CMPC_HD void linearize_knot(const Params& P, const double* xbar, const double* ubar, const double* cpos,
                            const int* cact, int k, double* rec) {
  // q = -Wx xbar  (cost.py:21-29; the tracking reference IS the linearisation point)
  for (int i = 0; i < NX; ++i) rec[O_Q + i] = -P.Wx[i] * xbar[i];
  for (int i = O_C; i < STG; ++i) rec[i] = 0.0;
  for (int i = 0; i < 3; ++i) rec[O_KB + i] = xbar[6 + i];
  if (k == P.N) return;   // terminal knot: no dynamics, no controls
  double S[3] = {0.0, 0.0, 0.0};
  int mask = 0;
  for (int c = 0; c < P.nc; ++c) {
    if (cact[c]) {
      mask |= 1 << c;
      for (int a = 0; a < 3; ++a) {
        S[a] += ubar[3 * c + a];
        rec[O_D + 3 * c + a] = cpos[3 * c + a] - xbar[a];
      }
    }
  }
  double Sxc[3];
  cross3(S, xbar, Sxc);
  rec[O_S + 0] = S[0]; rec[O_S + 1] = S[1]; rec[O_S + 2] = S[2];
  rec[O_C + 5] = P.dt * P.m * P.g;
  rec[O_C + 6] = -P.dt * Sxc[0];
  rec[O_C + 7] = -P.dt * Sxc[1];
  rec[O_C + 8] = -P.dt * Sxc[2];
  rec[O_ACT] = (double)mask;
}

// x+ = f(x,u) for the point-contact model (centroidal_model.py:189-212)
CMPC_HD void step_knot(const Params& P, const double* x, const double* u, const double* cpos, const int* cact,
                       double* xn) {
  double F[3] = {0, 0, 0}, Tq[3] = {0, 0, 0};
  for (int c = 0; c < P.nc; ++c) {
    if (cact[c]) {
      double d[3] = {cpos[3 * c] - x[0], cpos[3 * c + 1] - x[1], cpos[3 * c + 2] - x[2]};
      double t[3];
      cross3(d, u + 3 * c, t);
      for (int a = 0; a < 3; ++a) { F[a] += u[3 * c + a]; Tq[a] += t[a]; }
    }
  }
  for (int a = 0; a < 3; ++a) {
    xn[a] = x[a] + P.dt * (x[3 + a] / P.m);
    xn[3 + a] = x[3 + a] + P.dt * (F[a] + (a == 2 ? P.m * P.g : 0.0));
    xn[6 + a] = x[6 + a] + P.dt * Tq[a];
  }
}

// y = A_k v  and  y = A_k^T v  for the structured A (S from the stage record)
CMPC_HD double Av_elem(const Params& P, const double* S, const double* v, int i) {
  if (i < 3) return v[i] + (P.dt / P.m) * v[3 + i];
  if (i < 6) return v[i];
  int a = i - 6;   // (S x c)[a]
  int a1 = (a + 1) % 3, a2 = (a + 2) % 3;
  return v[i] + P.dt * (S[a1] * v[a2] - S[a2] * v[a1]);
}
CMPC_HD double ATv_elem(const Params& P, const double* S, const double* v, int i, int st = 1) {
  if (i < 3) {     // ([S]x)^T vk = -(S x vk) = vk x S
    int a = i, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
    return v[i * st] + P.dt * (v[(6 + a1) * st] * S[a2] - v[(6 + a2) * st] * S[a1]);
  }
  if (i < 6) return v[i * st] + (P.dt / P.m) * v[(i - 3) * st];
  return v[i * st];
}
// column j of B (compact control j -> contact c, axis a): B[:,j] = dt [0; e_a; d x e_a]
// (B^T v)_j = dt (v_l[a] + (v_k x d)[a])
CMPC_HD double BTv_elem(const Params& P, const double* d, const double* v, int a) {
  int a1 = (a + 1) % 3, a2 = (a + 2) % 3;
  return P.dt * (v[3 + a] + (v[6 + a1] * d[a2] - v[6 + a2] * d[a1]));
}
// element i of B_col(c,a) : rows 3..5 -> delta(i-3,a); rows 6..8 -> (d x e_a)[i-6]
CMPC_HD double Bcol_elem(const Params& P, const double* d, int a, int i) {
  if (i < 3) return 0.0;
  if (i < 6) return (i - 3 == a) ? P.dt : 0.0;
  int r = i - 6;
  if (r == a) return 0.0;
  // (d x e_a)[r]: e_a unit.  d x e_a = (d1*ea2 - d2*ea1, d2*ea0 - d0*ea2, d0*ea1 - d1*ea0)
  int r1 = (r + 1) % 3, r2 = (r + 2) % 3;
  return P.dt * ((r2 == a ? d[r1] : 0.0) - (r1 == a ? d[r2] : 0.0));
}

}  // namespace cmpc
