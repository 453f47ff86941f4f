// cmpc_core.cuh — data layout, parameters and the closed-form centroidal model.
// Execution model: a TEAM of NL = 8 lanes per MPC instance, TL = 4 instances ("a tile") per warp; every
// per-instance quantity of the knot records is stored instance-interleaved ([field][TL]), so that a range
// of fields is one contiguous block for a bulk copy.  cmpc_tile.cuh is the solver written against this layout.
//
// What this replaces in the reference (paths relative to /root/reference):
//   src/centroidal_model.py:189-232,257-291   dynamics + Jacobians  -> linearize_knot()
//   src/cost.py:9-39, src/constraints.py:12-50,104-109,153-185,260-293, src/scp_solver.py:10-48
//                                             QP assembly           -> never materialised: the
//                                             knot records below ARE the block-banded KKT data
//   src/scp_solver.py:59-68 (OSQP)            QP solve              -> admm + active-set polish
//   src/scp_solver.py:71-87,151               accuracy ratio, spectral trust test -> evaluate
//   src/scp_solver.py:118-179                 trust-region loop     -> the per-lane driver
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)   // the solver is device code in the CUDA build, host code in the test build (g++)
#define CMPC_HD __device__ __forceinline__
#define CMPC_FN __device__ __forceinline__
#if defined(CMPC_INLINE_OPS)
#define CMPC_OP __device__ __forceinline__
#else
#define CMPC_OP __device__ __noinline__   // whole-horizon operations: real calls, so that each gets its own register allocation
#endif
#define CMPC_CX __host__ __device__ constexpr inline
#else
#define CMPC_HD inline
#define CMPC_FN
#define CMPC_OP
#define CMPC_CX constexpr inline
#endif

namespace cmpc {

// Contact model of this compilation of the solver.  0: point contacts, three controls (fx, fy, fz) per
// contact (solo12, bolt).  1 (cmpc_wrench.cu, namespace cmpc_wr): flat feet with six controls
// (cop_x, cop_y, fx, fy, fz, tau_z) per foot (TALOS, centroidal_model.py:204-208).  A foot is carried as TWO
// slots of three controls: a force slot (fx, fy, fz) whose lever arm is p + R[:,0:2] cop_bar - c_bar, and a
// wrench slot (cop_x, cop_y, tau_z) that acts on the angular momentum only; the CoP box (constraints.py:
// 111-145) takes the place of the wrench slot's four "friction" rows.  "Contact" c of the solver is then
// the pseudo-contact 2 foot + kind.
#ifndef CMPC_WRENCH
#define CMPC_WRENCH 0
#endif
constexpr bool WR = CMPC_WRENCH != 0;
constexpr int NX = 9;
constexpr int MAXC = 4;    // contacts (wrench model: two feet = four pseudo-contacts)
constexpr int MAXU = 12;   // 3 * MAXC
// Execution model: NL lanes of a warp cooperate on one MPC instance (each lane owns rows of the
// knot's small dense systems; vectors that every lane needs are exchanged through shared memory
// and __syncwarp), TL = 32 / NL instances share a warp ("a tile").  Every output element is
// computed by exactly one lane with a fixed operation order and cross-lane reductions are only
// max / or / integer sums, so the results are bit-identical for every NL -- the host build runs
// NL = 1 (tests/emu) and a lock-step host build runs NL = 8 (tests/emu, coroutines).
#ifndef CMPC_NL
#if defined(__CUDACC__)
#define CMPC_NL 8
#else
#define CMPC_NL 1
#endif
#endif
constexpr int NL = CMPC_NL;
#ifndef CMPC_TL
#if defined(__CUDACC__)
#define CMPC_TL (32 / CMPC_NL)
#else
#define CMPC_TL 4
#endif
#endif
constexpr int TL = CMPC_TL;   // instances per tile

// ---- knot record.  A tile's workspace is [N+1 knots][rstride doubles]; field f of instance-lane t
// at knot k lives at k*rstride + f*TL + t, so the TL values of a field are one contiguous 8*TL-byte
// row and a range of fields is one contiguous block (one cp.async.bulk).  The field offsets depend on
// the knot's slot count ns ("slot" = position of a contact among the knot's ACTIVE contacts; tile-
// uniform: the maximum over the tile's instances, an instance with fewer active contacts pads with
// slots whose B columns are zero) and on whether the general friction table is present (gen):
//   A  Pc[9]                       factor -> backward sweep            P_{k+1} c_k
//   M  Hn[na][nap], Kt[9][nap]     factor -> sweeps     Hn = -Huu^-1 (full square), Kt[i][j] = -K[j][i]; nap = na | 1
//   C  meta, xbar[9], S[3], ck[3], d[na]     stage data, constant during a solve
//   G  per slot G[12] e2[4] ub[4]            general friction rows (rotated contacts / stochastic mode)
//   D  vk[3], vf[4 ns]             ADMM iterate
//   E  dv[na]                      backward -> forward sweep (feed-forward)
//   F  yk[4], yf[4 ns]             multiplier method
//   -  x[9], u[na]                 solution
// The order makes the fields of the ADMM sweeps contiguous: backward = [A .. D], forward = [Kt .. E].
struct Lay {
  int na, nap, pc, hn, kt, meta, xb, s, ck, d, g, vk, vf, dv, yk, yf, x, u, end;
};
CMPC_CX Lay lay_of(int ns, bool gen) {
  Lay L{};
  L.na = 3 * ns;
  L.nap = ns > 0 ? (L.na | 1) : 0;   // row stride of [Hn; Kt]: odd, so that the rows of the lanes of a team fall into different banks
  L.pc = 0;
  L.hn = 9;
  L.kt = L.hn + L.na * L.nap;
  L.meta = L.kt + 9 * L.nap;   // one field = 2*TL int32: [t] meta (bits 0..2 slots, 4..11 contact id per slot), [TL+t] active set
  L.xb = L.meta + 1;          // xbar[9]: linearisation point, q = -Wx xbar (cost.py:21-29), kbar = xbar[6:9]
  L.s = L.xb + 9;             // S[3]: sum of active fbar, A_k = I + dt[[0,I/m,0],[0,0,0],[[S]x,0,0]]
  L.ck = L.s + 3;             // ck[3]: affine term rows 6..8, -dt S x cbar (row 5 is dt m g)
  L.d = L.ck + 3;             // d[slot][3] = p_contact - cbar: B_k[:,3s:3s+3] = dt [0; I; [d]x]
  L.g = L.d + (WR ? 3 : 1) * L.na;   // wrench model: M[slot][9], kappa rows of B_k[:, slot] = dt M (column a at 3a)
  L.vk = L.g + (gen ? 20 * ns : 0);   // vk[3]: kappa copy, w = prox(v), y = rho_k (v - w)
  L.vf = L.vk + 3;            // vf[4*slot+row]: friction rows, w = min(v,0), y = rho e2 max(v,0)
  L.dv = L.vf + 4 * ns;
  L.yk = L.dv + L.na;         // multiplier method: yk[4] (3 pins + surface row)
  L.yf = L.yk + 4;            //                    yf[4 ns]
  L.x = L.yf + 4 * ns;
  L.u = L.x + 9;
  L.end = L.u + L.na;
  return L;
}
constexpr int GS = 20;       // general friction table, per slot: G (12, row-major 4x3), e2 (4), upper bound (4)
CMPC_CX int rec_fields(int nc, bool gen) { return (lay_of(nc, gen).end + 3) & ~3; }
constexpr int REC_MAX = rec_fields(MAXC, true);
constexpr int INFO = 12;     // per-instance statistics (cmpc_get_stats)

enum Status { ST_OK = 0, ST_QP_MAXITER = 1, ST_QP_NUMERIC = 2, ST_DEVICE = 3 };

struct Params {
  int N, nc, nu, identity_R, fast;   // fast: identity R and the same W_u for every contact
  double m, g, dt, mu, kf, dt_m, dtmg;
  double Wx[NX], Wu[MAXU];
  double e2[4];                       // fast path: friction-row equilibration factors e^2 per pyramid row
  int nf;                             // rows of the contact arrays per knot: nc, or the feet (nc / 2) of the wrench model
  double qs;                          // 1: tracking gradient q = -Wx xbar (cost.py:21-29); 0: none (TALOS, scp_solver.py:13-20)
  double foot_range[4];               // wrench model: cop_x <= [0], -cop_x <= [1], cop_y <= [2], -cop_y <= [3]
  // QP solver settings (OSQP's where they have the same meaning; scp_solver.py:61-63)
  double alpha, rho0, eps_abs, eps_rel, delta, inv_delta, adapt_tol, rho_e_rel, rho_k_rel, rho_e_pol_rel, as_tol;
  int max_iter, check_every, polish, refine, adaptive_rho, adapt_start;
  int as_start, as_step, as_rounds;   // early active-set polish: first attempt, retry interval, rounds
  int warm;                           // first QP starts with a polish on the active set read off the warm start
  double warm_tol;
  double as_tol_loose;                // certificate tolerance at the rounding floor of the multiplier iteration (= as_tol: off)
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
  const double* fub;      // [B][N][nc][4] friction-row upper bounds (stochastic mode) or null (all zero)
  // workspace
  double* ws;             // [tiles][N+1][rfields][TL]
  int rfields;            // fields per knot record (rec_fields(nc, gen))
  int* nst;               // [tiles][N+1]         slots per knot of the tile
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

// ------------------------------------------------------------------------------------------
// closed-form model pieces (scalar; used by setup, evaluate and the linearise/rollout kernels)
// ------------------------------------------------------------------------------------------
CMPC_HD void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
CMPC_CX int nxt3(int a) { return a == 2 ? 0 : a + 1; }
CMPC_CX int prv3(int a) { return a == 0 ? 2 : a - 1; }

// position of control a of (pseudo-)contact c in the caller's control vector
CMPC_CX int uix(int c, int a) {
  return !WR ? 3 * c + a : 6 * (c >> 1) + ((c & 1) ? (a < 2 ? a : 5) : 2 + a);
}

// K1: closed-form Jacobian data and affine term of one knot
//     (centroidal_model.py:189-232), point-contact model:
//     A_k = I + dt [[0, I/m, 0],[0,0,0],[[S]x,0,0]],  S = sum_i a_i fbar_i
//     B_k[:,3s:3s+3] = dt [0; I; [d_s]x],              d_s = p_s - cbar
//     c_k = fbar - A xbar - B ubar = [0; dt m g e_z; -dt S x cbar]
struct KnotLin {
  double S[3], ck[3], d[WR ? 3 * MAXU : MAXU];   // wrench model: M[slot][9] instead of d[slot][3]
  int meta;   // slots | contact id per slot << 4
};
//     Wrench model (WR; centroidal_model.py:204-208), foot i with frame R = [r1 r2 r3]:
//     kappa_dot += (p - c + r1 cop_x + r2 cop_y) x f + r3 tau_z.  Force slot: lever arm e = p + R[:,0:2] cop_bar
//     - c_bar; wrench slot: kappa rows dt [r1 x fbar, r2 x fbar, r3], no momentum rows.  Both are stored as
//     the 3x3 block M (column a at 3a) with B_k[6:9, slot] = dt M;
//     c_k = [0; dt m g e_z; -dt S x cbar - dt sum_i (R[:,0:2] cop_bar_i) x fbar_i].
CMPC_HD void linearize_knot(const Params& P, const double* xbar, const double* ubar, const double* cpos,
                            const int* cact, int terminal, KnotLin& o, const double* cR = nullptr) {
  for (int i = 0; i < 3; ++i) o.S[i] = o.ck[i] = 0.0;
  for (int i = 0; i < (WR ? 3 * MAXU : MAXU); ++i) o.d[i] = 0.0;
  o.meta = 0;
  if (terminal) return;   // terminal knot: no dynamics, no controls
  int slot = 0, code = 0;
  if (WR) {
    double bil[3] = {0.0, 0.0, 0.0};
    for (int ft = 0; ft < P.nf; ++ft) {
      if (!cact[ft]) continue;
      const double* R = cR + 9 * ft;
      const double* uf = ubar + 6 * ft;
      const double fb[3] = {uf[2], uf[3], uf[4]};
      double rc[3], e[3], t[3];
      for (int a = 0; a < 3; ++a) {
        rc[a] = R[3 * a] * uf[0] + R[3 * a + 1] * uf[1];     // R[:,0:2] cop_bar
        e[a] = cpos[3 * ft + a] - xbar[a] + rc[a];
        o.S[a] += fb[a];
      }
      cross3(rc, fb, t);
      for (int a = 0; a < 3; ++a) bil[a] += t[a];
      double* Mf = o.d + 9 * slot;        // force slot: [e]x, column a = e x e_a
      for (int a = 0; a < 3; ++a) {
        Mf[3 * a + a] = 0.0;
        Mf[3 * a + nxt3(a)] = e[prv3(a)];
        Mf[3 * a + prv3(a)] = -e[nxt3(a)];
      }
      code |= (2 * ft) << (2 * slot);
      ++slot;
      double* Mw = o.d + 9 * slot;        // wrench slot: [r1 x fbar, r2 x fbar, r3]
      for (int a = 0; a < 2; ++a) {
        const double ra[3] = {R[a], R[3 + a], R[6 + a]};
        cross3(ra, fb, t);
        for (int j = 0; j < 3; ++j) Mw[3 * a + j] = t[j];
      }
      for (int j = 0; j < 3; ++j) Mw[6 + j] = R[3 * j + 2];
      code |= (2 * ft + 1) << (2 * slot);
      ++slot;
    }
    double Sxc[3];
    cross3(o.S, xbar, Sxc);
    for (int a = 0; a < 3; ++a) o.ck[a] = -P.dt * Sxc[a] - P.dt * bil[a];
    o.meta = slot | (code << 4);
    return;
  }
  for (int c = 0; c < P.nc; ++c) {
    if (cact[c]) {
      for (int a = 0; a < 3; ++a) {
        o.S[a] += ubar[3 * c + a];
        o.d[3 * slot + a] = cpos[3 * c + a] - xbar[a];
      }
      code |= c << (2 * slot);
      ++slot;
    }
  }
  double Sxc[3];
  cross3(o.S, xbar, Sxc);
  for (int a = 0; a < 3; ++a) o.ck[a] = -P.dt * Sxc[a];
  o.meta = slot | (code << 4);
}

// x+ = f(x,u) (centroidal_model.py:189-212), full (per-contact) control layout
CMPC_HD void step_knot(const Params& P, const double* x, const double* u, const double* cpos, const int* cact,
                       double* xn, const double* cR = nullptr) {
  double F[3] = {0, 0, 0}, Tq[3] = {0, 0, 0};
  for (int c = 0; c < P.nf; ++c) {
    if (WR) {   // six controls per foot: (cop_x, cop_y, fx, fy, fz, tau_z)
      if (!cact[c]) continue;
      const double* R = cR + 9 * c;
      const double* uf = u + 6 * c;
      double arm[3], t[3];
      for (int a = 0; a < 3; ++a) arm[a] = cpos[3 * c + a] - x[a] + (R[3 * a] * uf[0] + R[3 * a + 1] * uf[1]);
      cross3(arm, uf + 2, t);
      for (int a = 0; a < 3; ++a) { F[a] += uf[2 + a]; Tq[a] += t[a] + R[3 * a + 2] * uf[5]; }
      continue;
    }
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
