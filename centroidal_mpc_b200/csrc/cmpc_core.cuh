// cmpc_core.cuh — data layout, parameters and the closed-form centroidal model.
// One warp per MPC instance; cmpc_simt.cuh is the execution model, cmpc_solver.cuh the solver.
//
// What this replaces in the reference (paths relative to /root/reference):
//   src/centroidal_model.py:189-232,257-291   dynamics + Jacobians  -> linearize_knot()
//   src/cost.py:9-39, src/constraints.py:12-50,104-109,153-185,260-293, src/scp_solver.py:10-48
//                                             QP assembly           -> never materialised: the
//                                             stage records below ARE the block-banded KKT data
//   src/scp_solver.py:59-68 (OSQP)            QP solve              -> admm + active-set polish
//   src/scp_solver.py:71-87,151               accuracy ratio, spectral trust test -> evaluate()
//   src/scp_solver.py:118-179                 trust-region loop     -> solve_instance()
#pragma once
#include "cmpc_simt.cuh"

#if defined(__CUDACC__)
#define CMPC_HD __host__ __device__ __forceinline__
#else
#define CMPC_HD inline
#endif

namespace cmpc {

constexpr int NX = 9;
constexpr int MAXC = 4;    // contacts
constexpr int MAXU = 12;   // 3 * MAXC

// ---- per-instance, per-knot records in global memory (doubles).  "slot" = position of a contact
// among the knot's ACTIVE contacts; controls (3 per slot) and friction rows (4 per slot) are
// stored compactly by slot, na = 3 * slots.  Every record starts 32-byte aligned.
constexpr int SG = 28;       // stage record, constant during a solve (built by setup_instance)
constexpr int SG_XB = 0;     //   xbar[9]   linearisation point: q = -Wx xbar (cost.py:21-29), kbar = xbar[6:9]
constexpr int SG_S = 9;      //   S[3]      sum of active fbar:  A_k = I + dt[[0,I/m,0],[0,0,0],[[S]x,0,0]]
constexpr int SG_CK = 12;    //   ck[3]     affine term rows 6..8: -dt S x cbar  (row 5 is dt m g, rows 0..4 are 0)
constexpr int SG_D = 16;     //   d[slot][3] = p_contact - cbar:  B_k[:,3s:3s+3] = dt [0; I; [d]x]
constexpr int ST = 20;       // ADMM iterate record
constexpr int ST_VF = 0;     //   vf[slot*4+row] friction rows: w = min(v,0), y = rho e2 max(v,0)
constexpr int ST_VK = 16;    //   vk[3]          kappa copy:    w = prox(v),  y = rho_k (v - w)
constexpr int FAC = 372;     // factor record
constexpr int F_HI = 0;      //   Hinv[na x na] (symmetric), then K[na x 9] at even(na*na), Pc[9] after it
constexpr int F_KT = 264;    //   Kt[9 x na]   (K transposed: the forward sweep reads columns)
constexpr int DVC = 12;      // feed-forward d_k (compact)
constexpr int PM = 20;       // multiplier-method record: yf[16] (compact rows), yk[4] (3 pins + surface row)
constexpr int SOL = 24;      // solution record: x[9] at 0, u[12] (compact) at 12
constexpr int SOL_U = 12;
constexpr int INFO = 12;     // per-instance statistics (cmpc_get_stats)

CMPC_HD int even_up(int n) { return (n + 1) & ~1; }
CMPC_HD int fac_off_K(int na) { return even_up(na * na); }
CMPC_HD int fac_off_Pc(int na) { return even_up(na * na) + even_up(9 * na); }

enum Status { ST_OK = 0, ST_QP_MAXITER = 1, ST_QP_NUMERIC = 2 };

struct Params {
  int N, nc, nu, identity_R, fast;   // fast: identity R and the same W_u for every contact
  double m, g, dt, mu, kf, dt_m, dtmg;
  double Wx[NX], Wu[MAXU];
  double e2[4];                       // fast path: friction-row equilibration factors e^2 per pyramid row
  // QP solver settings (OSQP's where they have the same meaning; scp_solver.py:61-63)
  double alpha, rho0, eps_abs, eps_rel, delta, adapt_tol, rho_e_rel, rho_k_rel, rho_e_pol_rel, as_tol;
  int max_iter, check_every, polish, refine, adaptive_rho, adapt_start;
  int as_start, as_step, as_rounds;   // early active-set polish: first attempt, retry interval, rounds
  // SCP parameters (scp_solver.py:120-128)
  double radius0, omega0, omega_max, acc_rho0, acc_rho1, beta_succ, beta_fail, gamma_fail, conv_thresh;
  int max_scp;
};

// Per-batch global-memory views (device pointers on the GPU; host pointers in the emulation).
struct Batch {
  int B;
  const double* x_init;   // [B][9]
  const double* x_final;  // [B][9]
  const double* X_ref;    // [B][N+1][9]
  const double* U_init;   // [B][N][nu]
  const double* cpos;     // [Bp][N][nc][3]
  const double* cR;       // [Bp][N][nc][9] or null (identity)
  const int* cact;        // [Bp][N][nc]
  long plan_stride;       // 0 (shared plan) or 1
  // workspace
  double* stg;            // [B][N+1][SG]
  double* sta;            // [B][N+1][ST]
  double* fac;            // [B][N][FAC]
  double* dvec;           // [B][N][DVC]
  double* pm;             // [B][N+1][PM]
  double* sol;            // [B][N+1][SOL]
  double* gtab;           // [B][N][MAXC][16]  general friction rows G (12) + e2 (4) per slot; null when fast
  int* meta;              // [B][N+1]  bits 0..2: slots, bits 4..11: contact id per slot (2 bits each)
  int* pmask;             // [B][N+1]  active set of the multiplier method
  // outputs
  double* X_out;          // [B][N+1][9]
  double* U_out;          // [B][N][nu]
  int* scp_iters;         // [B]
  int* status;            // [B]
  int* n_accepted;        // [B]
  int* qp_iters;          // [B] total ADMM iterations
  int* n_factor;          // [B]
  double* info;           // [B][INFO]: snorm, acc ratio, pri, dua, rho, radius, weight, polished,
                          //            multiplier-method sweeps, polish attempts, 2 spare
};

// Per-warp shared-memory scratch (factorisation, evaluate); 4.6 KB -> 28 instances per SM.
constexpr int MS = 21;     // row stride of the Gauss-Jordan tableau [Huu | Hux]
struct WarpMem {
  double P[81];            // P_{k+1} (row-major 9x9)
  double PA[81];           // P A
  double W[9 * MAXU];      // P B  (row stride 12)
  double M[MAXU * MS];     // [Huu | Hux] -> [Hinv | Huu^-1 Hux]
  double HX[MAXU * 9];     // Hux kept for the P update
  double pcol[MAXU + 4];   // pivot column
  double T[81];            // P_k before symmetrisation / Gram matrix in evaluate()
};

struct Ctx {
  const Params* prm;
  Batch bt;
  WarpMem* s;
  int b;   // instance
  const double* cpos; const double* cR; const int* cact;   // this instance's plan
  const double* Xr; const double* Ui; const double* xi; const double* xf;
  double* stg; double* sta; double* fac; double* dvec; double* pm; double* sol; double* gtab;
  int* meta; int* pmask;
};

// ------------------------------------------------------------------------------------------
// closed-form model pieces (scalar; used by setup, evaluate and the linearise/rollout kernels)
// ------------------------------------------------------------------------------------------
CMPC_HD void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
CMPC_HD int nxt3(int a) { return a == 2 ? 0 : a + 1; }
CMPC_HD int prv3(int a) { return a == 0 ? 2 : a - 1; }

// K1: dynamics, closed-form Jacobian data and affine term for one knot
//     (centroidal_model.py:189-232), point-contact model:
//     A_k = I + dt [[0, I/m, 0],[0,0,0],[[S]x,0,0]],  S = sum_i a_i fbar_i
//     B_k[:,3s:3s+3] = dt [0; I; [d_s]x],              d_s = p_s - cbar
//     c_k = fbar - A xbar - B ubar = [0; dt m g e_z; -dt S x cbar]
// Writes the stage record and returns the meta word (slots | contact ids << 4).
CMPC_HD int linearize_knot(const Params& P, const double* xbar, const double* ubar, const double* cpos,
                           const int* cact, int terminal, double* rec) {
  for (int i = 0; i < SG; ++i) rec[i] = 0.0;
  for (int i = 0; i < NX; ++i) rec[SG_XB + i] = xbar[i];
  if (terminal) return 0;   // terminal knot: no dynamics, no controls
  double S[3] = {0.0, 0.0, 0.0};
  int slot = 0, code = 0;
  for (int c = 0; c < P.nc; ++c) {
    if (cact[c]) {
      for (int a = 0; a < 3; ++a) {
        S[a] += ubar[3 * c + a];
        rec[SG_D + 3 * slot + a] = cpos[3 * c + a] - xbar[a];
      }
      code |= c << (2 * slot);
      ++slot;
    }
  }
  double Sxc[3];
  cross3(S, xbar, Sxc);
  for (int a = 0; a < 3; ++a) { rec[SG_S + a] = S[a]; rec[SG_CK + a] = -P.dt * Sxc[a]; }
  return slot | (code << 4);
}

// x+ = f(x,u) (centroidal_model.py:189-212), full (per-contact) control layout
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

// dense Jacobians from the structured form (cmpc_linearize): A (9x9 row-major), one column of B
CMPC_HD void dense_A(const Params& P, const double* S, double* A) {
  for (int i = 0; i < 81; ++i) A[i] = 0.0;
  for (int i = 0; i < 9; ++i) A[i * 9 + i] = 1.0;
  for (int a = 0; a < 3; ++a) {
    A[a * 9 + 3 + a] = P.dt / P.m;
    const int a1 = nxt3(a), a2 = prv3(a);      // ([S]x c)[a] = S[a1] c[a2] - S[a2] c[a1]
    A[(6 + a) * 9 + a2] += P.dt * S[a1];
    A[(6 + a) * 9 + a1] -= P.dt * S[a2];
  }
}
CMPC_HD void dense_Bcol(const Params& P, const double* d, int a, double* col) {
  for (int i = 0; i < 9; ++i) col[i] = 0.0;
  col[3 + a] = P.dt;
  col[6 + nxt3(a)] = P.dt * d[prv3(a)];      // rows 6..8: dt (d x e_a)
  col[6 + prv3(a)] = -P.dt * d[nxt3(a)];
}

}  // namespace cmpc
