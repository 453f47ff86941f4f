// cmpc_solver.cuh — Riccati factorisation, ADMM / multiplier-method sweeps, active-set polish,
// trust-region loop.  One warp per MPC instance (cmpc_simt.cuh is the execution model,
// cmpc_core.cuh the data layout); DESIGN.md "device algorithm" has the mathematics and
// oracle/device_model.py is the executable numpy specification of this file.
//
// Lane roles inside a knot step (fixed for the whole kernel):
//   lanes 0..11   u-lanes    control j = 3*slot + axis            (active: j < na)
//   lanes 0..15   row-lanes  friction row r = 4*slot + pyramid row (active: r < 4*slots)
//   lanes 16..24  x-lanes    state component i = lane - 16  (c: 0..2, l: 3..5, kappa: 6..8)
//   lanes 16..27  mirrors of the u-lanes, carry (d_slot x f_slot)[axis] in the forward sweep
// Vectors live distributed over lanes in registers; matrix-vector products read one matrix
// column per lane (coalesced) and broadcast the vector with shuffles.
#pragma once
#include "cmpc_core.cuh"
#if !defined(__CUDA_ARCH__) && defined(CMPC_TRACE)
#include <stdio.h>
#endif

// cold, scalar-style code: a loop over lanes on the host, the thread's own lane on the device
#if defined(__CUDA_ARCH__)
#define CMPC_LANES(l) for (int l = (int)(threadIdx.x & 31u), _once = 1; _once; _once = 0)
#define CMPC_LANE0 if ((threadIdx.x & 31u) == 0)
#else
#define CMPC_LANES(l) for (int l = 0; l < 32; ++l)
#define CMPC_LANE0
#endif

namespace cmpc {

constexpr int MODE_ADMM = 0, MODE_PMM = 1;
constexpr int XL = 16;   // first x-lane

struct Lane {
  vi lane, jj, s3, a, a1, a2, s4, row, rax, xi, xa, xa1, xa2;
  vb is_u, is_m, is_r, is_x, is_c, is_l, is_k;
  vd Wx, sgn;
};

CMPC_F Lane make_lane(const Params& P) {
  Lane L;
  L.lane = lane_id();
  L.jj = L.lane & 15;
  L.s3 = L.jj / 3;
  L.a = L.jj - L.s3 * 3;
  L.a1 = seli(L.a == 2, viconst(0), L.a + 1);
  L.a2 = seli(L.a == 0, viconst(2), L.a - 1);
  L.s4 = L.jj >> 2;
  L.row = L.jj & 3;
  L.rax = L.row >> 1;
  L.xi = L.lane - XL;
  vi xs = seli(L.xi < 0, viconst(0), L.xi);
  vi xq = xs / 3;
  L.xa = xs - xq * 3;
  L.xa1 = seli(L.xa == 2, viconst(0), L.xa + 1);
  L.xa2 = seli(L.xa == 0, viconst(2), L.xa - 1);
  L.is_u = L.lane < 12;
  L.is_m = L.jj < 12;
  L.is_r = L.lane < 16;
  L.is_x = (L.lane >= XL) && (L.lane < XL + 9);
  L.is_c = L.is_x && (L.xi < 3);
  L.is_l = L.is_x && (L.xi >= 3) && (L.xi < 6);
  L.is_k = L.is_x && (L.xi >= 6);
  L.Wx = ldif(L.is_x, P.Wx, xs);
  L.sgn = sel((L.row & 1) == 1, vconst(-1.0), vconst(1.0));
  return L;
}

// Solver scalars (uniform) and the lane-resident terminal multiplier.
struct Sv {
  double rho, rhok, rhoe, rhoep, radius, weight;
  double pri, dua, npri, ndua;     // last residuals
  double nq, dynrow;               // constant parts of the residual norms
  int kap;                         // multiplier method: some knot has trust-region rows
  int fail;                        // a pivot was not positive
  int n_pmm, n_polish;             // statistics: multiplier-method sweeps, polish attempts
  vd ye;                           // x-lanes: multiplier of x_N = x_final
};

// ---------------------------------------------------------------- friction rows
// G = pyr4 * R^T, pyr4 = [[1,0,-k],[-1,0,-k],[0,1,-k],[0,-1,-k]], k = mu/sqrt2
// (utils.py:9-16, constraints.py:178-184); e2 = row equilibration under D_u = 1/sqrt(W_u).
// Fast path (identity R, same W_u for all contacts): lane constants; else the per-knot table.
CMPC_HD int popc32(unsigned v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
CMPC_HD double pyr4(const Params& P, int row, int a) {
  if (a == 2) return -P.kf;
  if (a == 0) return row == 0 ? 1.0 : (row == 1 ? -1.0 : 0.0);
  return row == 2 ? 1.0 : (row == 3 ? -1.0 : 0.0);
}
// u-lane view: g[row] = G_slot[row][axis of the lane], e2[row]
CMPC_F void fric_u(const Ctx& c, const Lane& L, int k, vb act, vd* g, vd* e2) {
  const Params& P = *c.prm;
  if (P.fast) {
    const vd mk = vconst(-P.kf), z = vconst(0.0);
    g[0] = sel(L.a == 0, vconst(1.0), sel(L.a == 2, mk, z));
    g[1] = sel(L.a == 0, vconst(-1.0), sel(L.a == 2, mk, z));
    g[2] = sel(L.a == 1, vconst(1.0), sel(L.a == 2, mk, z));
    g[3] = sel(L.a == 1, vconst(-1.0), sel(L.a == 2, mk, z));
    for (int r = 0; r < 4; ++r) e2[r] = vconst(P.e2[r]);
  } else {
    const double* t = c.gtab + (long)k * MAXC * 16;
    for (int r = 0; r < 4; ++r) {
      g[r] = ldif(act, t, L.s3 * 16 + L.a + 3 * r);
      e2[r] = ldif(act, t, L.s3 * 16 + 12 + r);
    }
  }
}
// row-lane view: gr[a] = G_slot[row of the lane][a], e2 of the row
CMPC_F void fric_r(const Ctx& c, const Lane& L, int k, vb act, vd* gr, vd& e2r) {
  const Params& P = *c.prm;
  if (P.fast) {
    gr[0] = sel(L.rax == 0, L.sgn, vconst(0.0));
    gr[1] = sel(L.rax == 1, L.sgn, vconst(0.0));
    gr[2] = vconst(-P.kf);
    e2r = sel(L.rax == 0, vconst(P.e2[0]), vconst(P.e2[2]));
  } else {
    const double* t = c.gtab + (long)k * MAXC * 16;
    for (int a = 0; a < 3; ++a) gr[a] = ldif(act, t, L.s4 * 16 + L.row * 3 + a);
    e2r = ldif(act, t, L.s4 * 16 + 12 + L.row);
  }
}
// scalar access for the cold code: G of contact slot s at knot k
CMPC_HD double fric_G(const Ctx& c, int k, int s, int row, int a) {
  return c.prm->fast ? pyr4(*c.prm, row, a) : c.gtab[((long)k * MAXC + s) * 16 + row * 3 + a];
}
CMPC_HD double fric_e2(const Ctx& c, int k, int s, int row) {
  return c.prm->fast ? c.prm->e2[row] : c.gtab[((long)k * MAXC + s) * 16 + 12 + row];
}

// ---------------------------------------------------------------- trust-region prox
// argmin_v omega*max(0, |v - kbar|_1 - r) + rho/2 |v - a|^2  (oracle/device_model.py prox_trust)
// branch 0: inside the L1 ball; 1: outside after soft-thresholding; 2: on the surface.
CMPC_HD int prox_trust(const double* a, const double* kbar, double r, double omega, double rho, double* w) {
  double b[3], ab[3];
  double s1 = 0.0;
  for (int i = 0; i < 3; ++i) { b[i] = a[i] - kbar[i]; ab[i] = fabs(b[i]); s1 += ab[i]; }
  if (s1 <= r) { for (int i = 0; i < 3; ++i) w[i] = a[i]; return 0; }
  double tau = omega / rho, s2 = 0.0, d[3];
  for (int i = 0; i < 3; ++i) { d[i] = fmax(ab[i] - tau, 0.0); s2 += d[i]; }
  if (s2 >= r) {
    for (int i = 0; i < 3; ++i) w[i] = kbar[i] + (b[i] < 0.0 ? -d[i] : d[i]);
    return 1;
  }
  double s[3] = {ab[0], ab[1], ab[2]};   // projection onto the L1 ball: sort descending
  if (s[0] < s[1]) { double t = s[0]; s[0] = s[1]; s[1] = t; }
  if (s[1] < s[2]) { double t = s[1]; s[1] = s[2]; s[2] = t; }
  if (s[0] < s[1]) { double t = s[0]; s[0] = s[1]; s[1] = t; }
  double css = 0.0;
  tau = 0.0;
  for (int j = 0; j < 3; ++j) {
    css += s[j];
    double t = (css - r) / (double)(j + 1);
    if (s[j] - t > 0.0) tau = t;
  }
  for (int i = 0; i < 3; ++i) {
    double di = fmax(ab[i] - tau, 0.0);
    w[i] = kbar[i] + (b[i] < 0.0 ? -di : di);
  }
  return 2;
}

// kappa-lanes: w = prox(v) for the lane's component.  Fast path: all three components inside the
// ball (two shuffles); otherwise every lane runs the scalar prox on the gathered triple.
CMPC_F vd prox_lanes(const Lane& L, const Sv& S, vd v, vd kbar) {
  const vd ab = vabs(v - kbar);
  const vd s1 = ab + shfl(ab, XL + 6 + L.xa1) + shfl(ab, XL + 6 + L.xa2);
  if (!vany(L.is_k && (s1 > vconst(S.radius)))) return v;
  double a3[3], k3[3], w3[3];
  for (int i = 0; i < 3; ++i) { a3[i] = uni(v, XL + 6 + i); k3[i] = uni(kbar, XL + 6 + i); }
  prox_trust(a3, k3, S.radius, S.weight, S.rhok, w3);
  return sel(L.xa == 0, vconst(w3[0]), sel(L.xa == 1, vconst(w3[1]), vconst(w3[2])));
}

// ---------------------------------------------------------------- multiplier-method kappa rows
// Penalty block kM (3x3) and linear term kl (3) of knot k from the active-set word pm and the
// multipliers yk (uniform scalar code; only reached when S.kap).  Codes per component:
// 0 pinned to kbar, 1 sign +, 2 sign -; branch (bits 16..17): 1 linear penalty, 2 surface row.
CMPC_HD void pmm_kappa_terms(const Params& P, const Sv& S, int pm, const double* kbar, const double* yk, double* M,
                             double* kl) {
  for (int i = 0; i < 9; ++i) M[i] = 0.0;
  for (int i = 0; i < 3; ++i) kl[i] = 0.0;
  const int br = (pm >> 16) & 3;
  if (br == 0) return;
  const double inv = 1.0 / P.delta;
  double sg[3];
  for (int i = 0; i < 3; ++i) {
    const int code = (pm >> (18 + 2 * i)) & 3;
    sg[i] = code == 1 ? 1.0 : (code == 2 ? -1.0 : 0.0);
    if (code == 0) {   // pinned component: kappa_i = kbar_i
      M[4 * i] += inv;
      kl[i] -= inv * kbar[i] - yk[i];
    }
  }
  if (br == 1) {
    for (int i = 0; i < 3; ++i) kl[i] += S.weight * sg[i];
  } else {             // surface: sg'(kappa - kbar) = radius
    double bb = S.radius;
    for (int i = 0; i < 3; ++i) bb += sg[i] * kbar[i];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) M[3 * i + j] += inv * sg[i] * sg[j];
      kl[i] -= sg[i] * (inv * bb - yk[3]);
    }
  }
}

// ---------------------------------------------------------------- Riccati factorisation
// Backward over k: H_uu = R + B'PB, H_ux = B'PA, Gauss-Jordan on [H_uu | H_ux] (SPD, no pivoting)
// -> [Hinv | Huu^-1 Hux], K = -Huu^-1 Hux, Pc = P c, P <- Q + A'PA + H_ux' K.
// Writes fac[k] = {Hinv, K, Pc, Kt}.  ADMM mode: R = W_u + rho G'E2G, Q = W_x + rho_k I (kappa);
// multiplier mode: active friction rows and kappa rows carry the penalty 1/delta.
CMPC_F void factor(Ctx& c, const Lane& L, Sv& S, int mode) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const int N = P.N;
  const double rho_e = (mode == MODE_ADMM) ? S.rhoe : S.rhoep;
  const double inv = 1.0 / P.delta;
  const vi lane = L.lane;
  double kM[9], kl[3];
  // terminal knot: P = W_x + rho_e I (+ kappa block)
  {
    for (int t = 0; t < 3; ++t) {
      const vi e = lane + 32 * t;
      const vb in = e < 81;
      const vi i = e / 9, r = e - i * 9;
      vd v = sel(i == r, ldif(in, P.Wx, seli(in, i, viconst(0))) + vconst(rho_e), vconst(0.0));
      if (mode == MODE_ADMM) v = v + sel((i == r) && (i >= 6), vconst(S.rhok), vconst(0.0));
      stif(in, s.P, e, v);
    }
    wsync();
    if (mode == MODE_PMM && S.kap) {
      pmm_kappa_terms(P, S, c.pmask[N], c.stg + (long)N * SG + SG_XB + 6, c.pm + (long)N * PM + 16, kM, kl);
      const vb in = lane < 9;
      const vi i = lane / 3, r = lane - i * 3;
      vd add = vconst(0.0);
      for (int e = 0; e < 9; ++e) add = sel(lane == e, vconst(kM[e]), add);
      const vi idx = (i + 6) * 9 + 6 + r;
      stif(in, s.P, idx, ldif(in, s.P, idx) + add);
      wsync();
    }
  }
  const vb is9 = lane < 9;
  for (int k = N - 1; k >= 0; --k) {
    const double* sg = c.stg + (long)k * SG;
    double* fk = c.fac + (long)k * FAC;
    const int mt = c.meta[k];
    const int ns = mt & 7, na = 3 * ns;
    const int oK = fac_off_K(na), oPc = fac_off_Pc(na);
    const double S0 = sg[SG_S], S1 = sg[SG_S + 1], S2 = sg[SG_S + 2];
    const int pm = (mode == MODE_PMM) ? c.pmask[k] : 0;
    // ---- F1: lanes 0..8 own a row of P: PA = P A, W = P B, Pc = P c
    {
      vd pr[9];
      for (int q = 0; q < 9; ++q) pr[q] = ldif(is9, s.P, lane * 9 + q);
      const double Sv3[3] = {S0, S1, S2};
      for (int q = 0; q < 3; ++q) {     // (P [S]x)[i][q] = P[i][6+q1] S[q2] - P[i][6+q2] S[q1]
        const int q1 = nxt3(q), q2 = prv3(q);
        const vd t = vfma(pr[6 + q1], vconst(Sv3[q2]), -(pr[6 + q2] * vconst(Sv3[q1])));
        stif(is9, s.PA, lane * 9 + q, vfma(vconst(P.dt), t, pr[q]));
        stif(is9, s.PA, lane * 9 + 3 + q, vfma(vconst(P.dt_m), pr[q], pr[3 + q]));
        stif(is9, s.PA, lane * 9 + 6 + q, pr[6 + q]);
      }
      for (int sl = 0; sl < ns; ++sl) {
        const double d0 = sg[SG_D + 3 * sl], d1 = sg[SG_D + 3 * sl + 1], d2 = sg[SG_D + 3 * sl + 2];
        const double dd[3] = {d0, d1, d2};
        for (int a = 0; a < 3; ++a) {   // (P B)[i][(sl,a)] = dt (P[i][3+a] + P[i][6+a1] d[a2] - P[i][6+a2] d[a1])
          const int a1 = nxt3(a), a2 = prv3(a);
          const vd t = vfma(pr[6 + a1], vconst(dd[a2]), vfma(-pr[6 + a2], vconst(dd[a1]), pr[3 + a]));
          stif(is9, s.W, lane * 12 + 3 * sl + a, vconst(P.dt) * t);
        }
      }
      vd pc = pr[5] * vconst(P.dtmg);
      pc = vfma(pr[6], vconst(sg[SG_CK]), pc);
      pc = vfma(pr[7], vconst(sg[SG_CK + 1]), pc);
      pc = vfma(pr[8], vconst(sg[SG_CK + 2]), pc);
      stif(is9, fk, lane + oPc, pc);
    }
    wsync();
    // ---- F2: lane c < na+9 owns column c of the tableau [Huu | Hux] = B' [W | PA] (+ R)
    const vb cact = lane < na + 9;
    const vb cu = lane < na;
    {
      const vi off = seli(cu, lane, lane - na);
      const vi str = seli(cu, viconst(12), viconst(9));
      const double* base_w = s.W;
      const double* base_pa = s.PA;
      vd xr[9];
      for (int r = 3; r < 9; ++r)
        xr[r] = sel(cu, ldif(cu, base_w, r * 12 + off), ldif(cact && !cu, base_pa, r * 9 + off));
      (void)str;
      vd g[4], e2[4];
      if (na) fric_u(c, L, k, cu, g, e2);
      for (int sl = 0; sl < ns; ++sl) {
        const double dd[3] = {sg[SG_D + 3 * sl], sg[SG_D + 3 * sl + 1], sg[SG_D + 3 * sl + 2]};
        const int cid = (mt >> (4 + 2 * sl)) & 3;
        for (int a = 0; a < 3; ++a) {   // row (sl,a) of B' v = dt (v[3+a] + v[6+a1] d[a2] - v[6+a2] d[a1])
          const int a1 = nxt3(a), a2 = prv3(a);
          vd val = vconst(P.dt) * vfma(xr[6 + a1], vconst(dd[a2]), vfma(-xr[6 + a2], vconst(dd[a1]), xr[3 + a]));
          stif(cact && !cu, s.HX, (3 * sl + a) * 9 + off, val);
          // R block of the lane's own slot: W_u + G' diag(rr) G
          vd radd = sel(L.a == a, vconst(P.Wu[3 * cid + a]), vconst(0.0));
          for (int row = 0; row < 4; ++row) {
            vd rr;
            if (mode == MODE_ADMM) rr = vconst(S.rho) * e2[row];
            else rr = vconst(((pm >> (4 * sl + row)) & 1) ? inv : 0.0);
            const double gra_f = pyr4(P, row, a);
            const vd gra = P.fast ? vconst(gra_f) : ldif(cu, c.gtab + (long)k * MAXC * 16, L.s3 * 16 + row * 3 + a);
            radd = vfma(rr * gra, g[row], radd);
          }
          val = val + sel(cu && (L.s3 == sl), radd, vconst(0.0));
          stif(cact, s.M, (3 * sl + a) * MS + lane, val);
        }
      }
    }
    wsync();
    // ---- F3: in-place Gauss-Jordan, column per lane
    for (int pv = 0; pv < na; ++pv) {
      stif(cu, s.pcol, lane, ldif(cu, s.M, lane * MS + pv));
      wsync();
      const double piv = s.pcol[pv];
      if (!(piv > 0.0)) S.fail = 1;
      const double ip = 1.0 / piv;
      const vd prow = ldif(cact, s.M, pv * MS + lane) * vconst(ip);
      const vb ispv = lane == pv;
      for (int i = 0; i < na; ++i) {
        const double pci = s.pcol[i];
        vd nv;
        if (i == pv) {
          nv = sel(ispv, vconst(ip), prow);
        } else {
          const vd old = ldif(cact, s.M, i * MS + lane);
          nv = sel(ispv, vconst(-pci * ip), vfma(vconst(-pci), prow, old));
        }
        stif(cact, s.M, i * MS + lane, nv);
      }
      wsync();
    }
    // ---- F4: factor record.  Hinv (symmetrised) and K = -Huu^-1 Hux rows; Kt columns
    for (int j = 0; j < na; ++j) {
      const vd v = ldif(cact, s.M, j * MS + lane);
      const vd vt = ldif(cu, s.M, lane * MS + j);
      const vd out = sel(cu, vconst(0.5) * (v + vt), -v);
      const vi addr = seli(cu, lane + (F_HI + j * na), lane + (oK + j * 9 - na));
      stif(cact, fk, addr, out);
    }
    for (int i = 0; i < 9; ++i) stif(cu, fk, lane + (F_KT + i * na), -ldif(cu, s.M, lane * MS + na + i));
    // ---- F5: P_k = Q + A'(PA) - Hux' (Huu^-1 Hux), 27 lanes x 3 rows, then symmetrise
    {
      const vb in = lane < 27;
      const vi g3 = lane / 9, cc = lane - g3 * 9;
      const double Sv3[3] = {S0, S1, S2};
      const vb gc = in && (g3 == 0), gl = in && (g3 == 1);
      for (int a = 0; a < 3; ++a) {
        const int a1 = nxt3(a), a2 = prv3(a);
        const vi r = g3 * 3 + a;
        vd val = ldif(in, s.PA, r * 9 + cc);
        // (A'Y)[r]: r<3: + dt (Y[6+a1] S[a2] - Y[6+a2] S[a1]);  3<=r<6: + dt/m Y[r-3]
        const vd yA = ldif(gc, s.PA, cc + (6 + a1) * 9), yB = ldif(gc, s.PA, cc + (6 + a2) * 9);
        const vd yC = ldif(gl, s.PA, cc + a * 9);
        val = vfma(vconst(P.dt), vfma(yA, vconst(Sv3[a2]), -(yB * vconst(Sv3[a1]))), val);
        val = vfma(vconst(P.dt_m), yC, val);
        for (int j = 0; j < na; ++j)
          val = vfma(-ldif(in, s.HX, r + j * 9), ldif(in, s.M, cc + (j * MS + na)), val);
        stif(in, s.T, r * 9 + cc, val);
      }
      wsync();
      for (int a = 0; a < 3; ++a) {
        const vi r = g3 * 3 + a;
        vd val = vconst(0.5) * (ldif(in, s.T, r * 9 + cc) + ldif(in, s.T, cc * 9 + r));
        const vb dg = in && (r == cc);
        val = val + sel(dg, ldif(dg, P.Wx, seli(dg, r, viconst(0))), vconst(0.0));
        if (mode == MODE_ADMM && k >= 1) val = val + sel(dg && (r >= 6), vconst(S.rhok), vconst(0.0));
        stif(in, s.P, r * 9 + cc, val);
      }
      wsync();
      if (mode == MODE_PMM && S.kap && k >= 1) {
        pmm_kappa_terms(P, S, pm, sg + SG_XB + 6, c.pm + (long)k * PM + 16, kM, kl);
        const vb in9 = lane < 9;
        const vi i = lane / 3, r = lane - i * 3;
        vd add = vconst(0.0);
        for (int e = 0; e < 9; ++e) add = sel(lane == e, vconst(0.5 * (kM[e] + kM[(e % 3) * 3 + e / 3])), add);
        const vi idx = (i + 6) * 9 + 6 + r;
        stif(in9, s.P, idx, ldif(in9, s.P, idx) + add);
        wsync();
      }
    }
  }
}

// ---------------------------------------------------------------- backward sweep (linear term)
// p_N = qx_N;  g = p + Pc;  hu = ru + B'g;  d = -Hinv hu;  p = qx + A'g + K'hu.
// qx = -Wx xbar (+ kappa / terminal penalty terms), ru = friction penalty terms.
CMPC_F void backward_sweep(Ctx& c, const Lane& L, const Sv& S, int mode) {
  const Params& P = *c.prm;
  const int N = P.N;
  const double rho_e = (mode == MODE_ADMM) ? S.rhoe : S.rhoep;
  double kM[9], kl3[3];
  vd p;
  {
    const double* sg = c.stg + (long)N * SG;
    const vd xbar = ldif(L.is_x, sg + SG_XB, L.xi);
    const vd xf = ldif(L.is_x, c.xf, L.xi);
    p = -(L.Wx * xbar) - (vconst(rho_e) * xf - S.ye);
    if (mode == MODE_ADMM) {
      const vd vk = ldif(L.is_k, c.sta + (long)N * ST + ST_VK, L.xa);
      const vd w = prox_lanes(L, S, vk, xbar);
      p = p + sel(L.is_k, vconst(-S.rhok) * (w + w - vk), vconst(0.0));
    } else if (S.kap) {
      pmm_kappa_terms(P, S, c.pmask[N], sg + SG_XB + 6, c.pm + (long)N * PM + 16, kM, kl3);
      p = p + sel(L.is_k, sel(L.xa == 0, vconst(kl3[0]), sel(L.xa == 1, vconst(kl3[1]), vconst(kl3[2]))), vconst(0.0));
    }
  }
  // shuffle sources / coefficients of A'g on the x-lanes
  const vi srcA = seli(L.is_c, L.xa1 + (XL + 6), L.xi + (XL - 3));
  const vi srcB = L.xa2 + (XL + 6);
  for (int k = N - 1; k >= 0; --k) {
    const double* sg = c.stg + (long)k * SG;
    const double* fk = c.fac + (long)k * FAC;
    const int mt = c.meta[k];
    const int ns = mt & 7, na = 3 * ns;
    const int oK = fac_off_K(na), oPc = fac_off_Pc(na);
    const vb ua = L.is_u && (L.jj < na);
    // g = p + Pc
    const vd g = p + ldif(L.is_x, fk + oPc, L.xi);
    // u-lanes: hu = ru + B'g,  (B'g)_(s,a) = dt (g[3+a] + g[6+a1] d[a2] - g[6+a2] d[a1])
    const vd d1 = ldif(ua, sg + SG_D, L.s3 * 3 + L.a1), d2 = ldif(ua, sg + SG_D, L.s3 * 3 + L.a2);
    const vd g3 = shfl(g, L.a + (XL + 3)), gk1 = shfl(g, L.a1 + (XL + 6)), gk2 = shfl(g, L.a2 + (XL + 6));
    vd hu = vconst(P.dt) * vfma(gk1, d2, vfma(-gk2, d1, g3));
    if (na) {
      vd gg[4], e2[4];
      fric_u(c, L, k, ua, gg, e2);
      if (mode == MODE_ADMM) {
        const double* sk = c.sta + (long)k * ST + ST_VF;
        for (int r = 0; r < 4; ++r)   // -(rho e2 w - y) = rho e2 |v|
          hu = vfma(gg[r], vconst(S.rho) * e2[r] * vabs(ldif(ua, sk, L.s3 * 4 + r)), hu);
      } else {
        const double* pk = c.pm + (long)k * PM;
        const int pm = c.pmask[k];
        for (int r = 0; r < 4; ++r) {
          const vb on = ua && (((viconst(pm) >> (L.s3 * 4 + r)) & 1) == 1);
          hu = vfma(gg[r], ldif(on, pk, L.s3 * 4 + r), hu);
        }
      }
    }
    hu = sel(ua, hu, vconst(0.0));
    // d = -Hinv hu (u-lanes read a column of Hinv), K'hu (x-lanes read a column of K)
    const vb ma = ua || L.is_x;
    const vi off = seli(L.is_u, L.jj + F_HI, L.xi + oK);
    const vi str = seli(L.is_u, viconst(na), viconst(9));
    vd acc = vconst(0.0);
    vi idx = off;
    for (int j = 0; j < na; ++j) {
      acc = vfma(ldif(ma, fk, idx), vconst(uni(hu, j)), acc);
      idx = idx + str;
    }
    stif(ua, c.dvec + (long)k * DVC, L.jj, -acc);
    // p = qx + A'g + K'hu on the x-lanes
    const vd xbar = ldif(L.is_x, sg + SG_XB, L.xi);
    const vd t1 = shfl(g, srcA), t2 = shfl(g, srcB);
    const vd Sa1 = ldif(L.is_c, sg + SG_S, L.xa1), Sa2 = ldif(L.is_c, sg + SG_S, L.xa2);
    const vd cA = sel(L.is_c, vconst(P.dt) * Sa2, sel(L.is_l, vconst(P.dt_m), vconst(0.0)));
    const vd cB = vconst(-P.dt) * Sa1;   // zero off the c-lanes (Sa1 = 0 there)
    vd pn = vfma(cA, t1, vfma(cB, t2, g)) + acc - L.Wx * xbar;
    if (k >= 1) {
      if (mode == MODE_ADMM) {
        const vd vk = ldif(L.is_k, c.sta + (long)k * ST + ST_VK, L.xa);
        const vd w = prox_lanes(L, S, vk, xbar);
        pn = pn + sel(L.is_k, vconst(-S.rhok) * (w + w - vk), vconst(0.0));
      } else if (S.kap) {
        pmm_kappa_terms(P, S, c.pmask[k], sg + SG_XB + 6, c.pm + (long)k * PM + 16, kM, kl3);
        pn = pn + sel(L.is_k, sel(L.xa == 0, vconst(kl3[0]), sel(L.xa == 1, vconst(kl3[1]), vconst(kl3[2]))), vconst(0.0));
      }
    }
    p = sel(L.is_x, pn, vconst(0.0));
  }
}

// ---------------------------------------------------------------- forward sweep + local updates
// u~ = K x~ + d,  x~+ = A x~ + B u~ + c.
// ADMM mode: relaxation, friction / kappa / terminal updates of (w, y) stored as v, all
// knot-local; with `check` also OSQP's residuals (oracle/device_model.py iterate()).
// Multiplier mode: y += (1/delta) row on the active rows, solution record, primal residual,
// and (with `upd`) the corrected friction active set: violated rows join, rows with a negative
// multiplier leave; returns the number of changes through *changes.
struct Res { vd pri, dua, npri, ndua; };

CMPC_F void forward_sweep(Ctx& c, const Lane& L, Sv& S, int mode, int check, int upd, int* changes) {
  const Params& P = *c.prm;
  const int N = P.N;
  const double al = (mode == MODE_ADMM) ? P.alpha : 1.0;
  const double inv = 1.0 / P.delta;
  Res R;
  R.pri = R.dua = R.npri = R.ndua = vconst(0.0);
  int nchg = 0;
  vd x = ldif(L.is_x, c.xi, L.xi);
  const vd xf = ldif(L.is_x, c.xf, L.xi);
  // x+ = A x + B u + c:  c-lanes: + dt/m x[3+i];  kappa-lanes: + dt (S[a1] c[a2] - S[a2] c[a1])
  const vi srcA = seli(L.is_c, L.xi + (XL + 3), L.xa2 + XL);
  const vi srcB = L.xa1 + XL;
  const vi gsrc = seli(L.is_k, L.xa + XL, L.xa);
  for (int k = 0; k <= N; ++k) {
    const double* sg = c.stg + (long)k * SG;
    double* sk = c.sta + (long)k * ST;
    double* so = c.sol + (long)k * SOL;
    const vd xbar = ldif(L.is_x, sg + SG_XB, L.xi);
    vd rdx = vconst(0.0);   // stationarity residual of the x rows of this knot
    if (mode == MODE_PMM) stif(L.is_x, so, L.xi, x);
    // ---- kappa copy of knot k
    if (k >= 1) {
      if (mode == MODE_ADMM) {
        const vd vk = ldif(L.is_k, sk + ST_VK, L.xa);
        const vd w = prox_lanes(L, S, vk, xbar);
        const vd vn = vfma(vconst(al), x, vfma(vconst(1.0 - al), w, vk - w));
        stif(L.is_k, sk + ST_VK, L.xa, vn);
        if (check) {
          const vd wn = prox_lanes(L, S, vn, xbar);
          R.pri = vmax(R.pri, sel(L.is_k, vabs(x - wn), vconst(0.0)));
          R.npri = vmax(R.npri, sel(L.is_k, vmax(vabs(x), vabs(wn)), vconst(0.0)));
          rdx = sel(L.is_k, vconst(S.rhok) * ((vn - wn) - (vk - w) - (x - w)), vconst(0.0));
        }
      } else {
        if (S.kap) {
          const int pm = c.pmask[k];
          const int br = (pm >> 16) & 3;
          if (br != 0) {
            double x3[3];
            for (int i = 0; i < 3; ++i) x3[i] = uni(x, XL + 6 + i);
            CMPC_LANE0 {
              double* yk = c.pm + (long)k * PM + 16;
              const double* kb = sg + SG_XB + 6;
              double accv = -S.radius;
              for (int i = 0; i < 3; ++i) {
                const int code = (pm >> (18 + 2 * i)) & 3;
                const double sgn = code == 1 ? 1.0 : (code == 2 ? -1.0 : 0.0);
                if (code == 0) yk[i] += inv * (x3[i] - kb[i]);
                accv += sgn * (x3[i] - kb[i]);
              }
              if (br == 2) yk[3] += inv * accv;
            }
            wsync();
          }
        }
        R.npri = vmax(R.npri, sel(L.is_k, vabs(x), vconst(0.0)));
      }
    }
    // ---- terminal equality
    if (k == N) {
      const double re = (mode == MODE_ADMM) ? S.rhoe : S.rhoep;
      const vd dx = sel(L.is_x, x - xf, vconst(0.0));
      S.ye = S.ye + vconst(re * al) * dx;
      if (check || mode == MODE_PMM) {
        R.pri = vmax(R.pri, vabs(dx));
        R.npri = vmax(R.npri, sel(L.is_x, vmax(vabs(x), vabs(xf)), vconst(0.0)));
        // no stationarity term: the certificate uses y_e + rho_e (x_N - x_f) for these equality
        // rows, the multiplier of the x-update itself (any value is admissible on an equality)
      }
    }
    if (check && k >= 1) {
      const vd Px = L.Wx * x;
      const vd aty = rdx - Px + L.Wx * xbar;     // (A'y)_x = r_d - P x - q,  q = -Wx xbar
      R.dua = vmax(R.dua, vabs(rdx));
      R.ndua = vmax(R.ndua, sel(L.is_x, vmax(vabs(Px), vabs(aty)), vconst(0.0)));
    }
    if (k == N) break;
    // ---- controls u~ = K x + d
    const double* fk = c.fac + (long)k * FAC;
    const int mt = c.meta[k];
    const int ns = mt & 7, na = 3 * ns;
    const vb ua = L.is_u && (L.jj < na);
    vd ut = ldif(ua, c.dvec + (long)k * DVC, L.jj);
    {
      vi idx = L.jj + F_KT;
      for (int i = 0; i < 9; ++i) {
        ut = vfma(ldif(ua, fk, idx), vconst(uni(x, XL + i)), ut);
        idx = idx + na;
      }
    }
    if (mode == MODE_PMM) stif(ua, so + SOL_U, L.jj, ut);
    // ---- next state.  mirrors: t = (d x f)[a] = d[a1] f[a2] - d[a2] f[a1]
    {
      const vb mact = L.is_m && (L.jj < na);
      const vd d1 = ldif(mact, sg + SG_D, L.s3 * 3 + L.a1), d2 = ldif(mact, sg + SG_D, L.s3 * 3 + L.a2);
      const vd f1 = shfl(ut, L.s3 * 3 + L.a1), f2 = shfl(ut, L.s3 * 3 + L.a2);
      const vd tq = vfma(d1, f2, -(d2 * f1));
      const vd m = sel(L.is_r, ut, sel(mact, tq, vconst(0.0)));
      vd sum = vconst(0.0);
      for (int sl = 0; sl < ns; ++sl) sum = sum + shfl(m, gsrc + 3 * sl);
      const vd t1 = shfl(x, srcA), t2 = shfl(x, srcB);
      const vd Sa1 = ldif(L.is_k, sg + SG_S, L.xa1), Sa2 = ldif(L.is_k, sg + SG_S, L.xa2);
      const vd cA = sel(L.is_c, vconst(P.dt_m), vconst(P.dt) * Sa1);   // zero on the l-lanes (Sa1 = 0)
      const vd cB = vconst(-P.dt) * Sa2;
      const vd cst = sel(L.is_k, ldif(L.is_k, sg + SG_CK, L.xa), sel(L.is_x && (L.xi == 5), vconst(P.dtmg), vconst(0.0)));
      const vd fsum = sel(L.is_l || L.is_k, sum, vconst(0.0));
      const vd xn = vfma(cA, t1, vfma(cB, t2, x)) + vfma(vconst(P.dt), fsum, cst);
      // ---- friction rows of knot k on the row-lanes
      if (na) {
        const vb ra = L.is_r && (L.jj < 4 * ns);
        vd gr[3], e2r;
        fric_r(c, L, k, ra, gr, e2r);
        vd cf;
        if (P.fast) {
          cf = vfma(gr[2], shfl(ut, L.s4 * 3 + 2), L.sgn * shfl(ut, L.s4 * 3 + L.rax));
        } else {
          cf = gr[0] * shfl(ut, L.s4 * 3);
          cf = vfma(gr[1], shfl(ut, L.s4 * 3 + 1), cf);
          cf = vfma(gr[2], shfl(ut, L.s4 * 3 + 2), cf);
        }
        vd delta_r = vconst(0.0);
        if (mode == MODE_ADMM) {
          const vd v = ldif(ra, sk + ST_VF, L.jj);
          const vd w0 = vmin(v, vconst(0.0)), y0 = vmax(v, vconst(0.0));
          const vd vn = vfma(vconst(al), cf, vfma(vconst(1.0 - al), w0, y0));
          stif(ra, sk + ST_VF, L.jj, vn);
          if (check) {
            const vd wn = vmin(vn, vconst(0.0));
            R.pri = vmax(R.pri, sel(ra, vabs(cf - wn), vconst(0.0)));
            R.npri = vmax(R.npri, sel(ra, vmax(vabs(cf), vabs(wn)), vconst(0.0)));
            delta_r = sel(ra, vconst(S.rho) * e2r * (vmax(vn, vconst(0.0)) - y0 - cf + w0), vconst(0.0));
          }
        } else {
          double* pk = c.pm + (long)k * PM;
          const int pm = c.pmask[k];
          const vb on = ra && (((viconst(pm) >> L.jj) & 1) == 1);
          const vd yn = vfma(vconst(inv), cf, ldif(on, pk, L.jj));
          R.pri = vmax(R.pri, sel(on, vabs(cf), sel(ra, vmax(cf, vconst(0.0)), vconst(0.0))));
          R.npri = vmax(R.npri, sel(ra, vabs(cf), vconst(0.0)));
          if (upd) {
            const vb keep = on && !(yn < vconst(0.0));
            const vb join = ra && !on && (cf > vconst(P.as_tol));
            const unsigned nm = vballot(keep || join);
            nchg += popc32((nm ^ (unsigned)pm) & 0xffffu);
            stif(ra, pk, L.jj, sel(keep, yn, vconst(0.0)));
            CMPC_LANE0 { c.pmask[k] = (pm & ~0xffff) | (int)(nm & 0xffffu); }
          } else {
            stif(on, pk, L.jj, yn);
          }
        }
        if (check) {
          // u rows of the stationarity residual: G' delta; norms of P u and A'y
          vd gg[4], e2[4];
          fric_u(c, L, k, ua, gg, e2);
          vd rdu = vconst(0.0);
          for (int r = 0; r < 4; ++r) rdu = vfma(gg[r], shfl(delta_r, L.s3 * 4 + r), rdu);
          const int code = mt >> 4;
          const vi cid = (viconst(code) >> (L.s3 * 2)) & 3;
          const vd Pu = ldif(ua, P.Wu, seli(ua, cid * 3 + L.a, viconst(0))) * ut;
          R.dua = vmax(R.dua, sel(ua, vabs(rdu), vconst(0.0)));
          R.ndua = vmax(R.ndua, sel(ua, vmax(vabs(Pu), vabs(rdu - Pu)), vconst(0.0)));
        }
      }
      x = sel(L.is_x, xn, vconst(0.0));
    }
  }
  if (check || mode == MODE_PMM) {
    S.pri = wmax(R.pri);
    S.npri = fmax(wmax(R.npri), S.dynrow);
    if (check) {
      S.dua = wmax(R.dua);
      S.ndua = fmax(wmax(R.ndua), S.nq);
    }
  }
  if (changes) *changes = nchg;
}

// ---------------------------------------------------------------- rho change: keep (w, y), move v
CMPC_F void rescale_duals(Ctx& c, const Lane& L, const Sv& S, double rho_new, double rhok_new) {
  const Params& P = *c.prm;
  const double ratio = S.rho / rho_new;
  for (int k = 0; k <= P.N; ++k) {
    double* sk = c.sta + (long)k * ST;
    if (k < P.N) {
      const int ns = c.meta[k] & 7;
      const vb ra = L.is_r && (L.jj < 4 * ns);
      const vd v = ldif(ra, sk + ST_VF, L.jj);
      stif(ra, sk + ST_VF, L.jj, vfma(vconst(ratio), vmax(v, vconst(0.0)), vmin(v, vconst(0.0))));
    }
    if (k >= 1) {
      const vd kbar = ldif(L.is_k, c.stg + (long)k * SG + SG_XB, L.xi);
      const vd vk = ldif(L.is_k, sk + ST_VK, L.xa);
      const vd w = prox_lanes(L, S, vk, kbar);
      stif(L.is_k, sk + ST_VK, L.xa, vfma(vconst(S.rhok / rhok_new), vk - w, w));
    }
  }
  wsync();
}

CMPC_F void set_rho(const Params& P, Sv& S, double rho) {
  const double wk = fmin(P.Wx[6], fmin(P.Wx[7], P.Wx[8]));
  double wm = 0.0;
  for (int i = 0; i < 9; ++i) wm = fmax(wm, P.Wx[i]);
  S.rho = rho;
  S.rhok = rho * P.rho_k_rel * wk;
  S.rhoe = P.rho_e_rel * wm;
  S.rhoep = P.rho_e_pol_rel * wm;
}

// ---------------------------------------------------------------- active-set polish
// Guess the active set from the ADMM iterate (friction row active iff its multiplier is
// positive, OSQP's rule; trust-region rows by the branch the prox took), solve the
// equality-constrained QP by the method of multipliers with penalty 1/delta on the active rows
// (OSQP: regularised KKT + iterative refinement), then correct the friction active set and
// repeat, at most 1 + rounds times.  Returns 1 when the point is CERTIFIED: no row wants to
// change, the primal residual is below as_tol, stationarity is exact by construction — a KKT
// point of the QP regardless of how far the ADMM iterate still was.  The ADMM iterate (sta) is
// left untouched; the polished point is in sol.
CMPC_F int build_active_set(Ctx& c, const Lane& L, Sv& S) {
  const Params& P = *c.prm;
  const int N = P.N;
  int kap = 0;
  for (int k = 0; k <= N; ++k) {
    const double* sk = c.sta + (long)k * ST;
    double* pk = c.pm + (long)k * PM;
    int pm = 0;
    if (k < N) {
      const int ns = c.meta[k] & 7;
      const vb ra = L.is_r && (L.jj < 4 * ns);
      vd gr[3], e2r;
      fric_r(c, L, k, ra, gr, e2r);
      const vd v = ldif(ra, sk + ST_VF, L.jj);
      const vb on = ra && (v > vconst(0.0));     // -w < y  <=>  v > 0
      pm = (int)(vballot(on) & 0xffffu);
      stif(L.is_r, pk, L.jj, sel(on, vconst(S.rho) * e2r * v, vconst(0.0)));
    }
    if (k >= 1) {
      const double* kb = c.stg + (long)k * SG + SG_XB + 6;
      const double a3[3] = {sk[ST_VK], sk[ST_VK + 1], sk[ST_VK + 2]};
      double w[3];
      const int br = prox_trust(a3, kb, S.radius, S.weight, S.rhok, w);
      if (br != 0) {
        kap = 1;
        pm |= br << 16;
        double yk4[4] = {0.0, 0.0, 0.0, 0.0}, msum = 0.0;
        int nz = 0;
        for (int i = 0; i < 3; ++i) {
          const double d = w[i] - kb[i];
          const double yk = S.rhok * (a3[i] - w[i]);
          const int code = d > 0.0 ? 1 : (d < 0.0 ? 2 : 0);
          pm |= code << (18 + 2 * i);
          if (code == 0) yk4[i] = yk;
          else { msum += (code == 1 ? yk : -yk); ++nz; }
        }
        if (br == 2) yk4[3] = nz ? msum / nz : 0.0;
        CMPC_LANE0 { for (int i = 0; i < 4; ++i) pk[16 + i] = yk4[i]; }
      } else {
        CMPC_LANE0 { for (int i = 0; i < 4; ++i) pk[16 + i] = 0.0; }
      }
    }
    CMPC_LANE0 { c.pmask[k] = pm; }
  }
  wsync();
  return kap;
}

CMPC_F int active_set_polish(Ctx& c, const Lane& L, Sv& S, int rounds, int* nfact) {
  const Params& P = *c.prm;
  const vd ye_keep = S.ye;
  const double pri_keep = S.pri, npri_keep = S.npri;
  S.kap = build_active_set(c, L, S);
  ++S.n_polish;
  int certified = 0, prev_chg = 1 << 30;
  for (int round = 0; round <= rounds; ++round) {
    S.fail = 0;
    factor(c, L, S, MODE_PMM);
    ++*nfact;
    if (S.fail) break;
    int chg = 0;
    backward_sweep(c, L, S, MODE_PMM);
    forward_sweep(c, L, S, MODE_PMM, 0, 0, nullptr);
    ++S.n_pmm;
    for (int sw = 0; sw < 1 + P.refine; ++sw) {
      ++S.n_pmm;
      backward_sweep(c, L, S, MODE_PMM);
      forward_sweep(c, L, S, MODE_PMM, 0, 1, &chg);
#if !defined(__CUDA_ARCH__) && defined(CMPC_TRACE)
      fprintf(stderr, "     sweep %d: changes %d pri %.3e\n", sw, chg, S.pri);
#endif
      if (chg || S.pri <= P.as_tol) break;
    }
    if (!(S.pri == S.pri)) break;   // NaN
#if !defined(__CUDA_ARCH__) && defined(CMPC_TRACE)
    fprintf(stderr, "  polish round %d: changes %d pri %.3e fail %d\n", round, chg, S.pri, S.fail);
#endif
    if (chg == 0) {
      certified = S.pri <= P.as_tol;
      break;
    }
    if (chg > prev_chg) break;      // the active set is not settling: back to ADMM
    prev_chg = chg;
  }
  S.kap = 0;
  // the polished residuals stay in S.pri / S.npri for the caller; the ADMM multiplier comes back
  const double ppri = S.pri, pnpri = S.npri;
  S.ye = ye_keep;
  (void)pri_keep; (void)npri_keep;
  S.pri = ppri; S.npri = pnpri;
  return certified;
}

// ---------------------------------------------------------------- QP driver
// ADMM (the active-set predictor and the globally convergent fallback) interleaved with
// certified active-set polishes.  Returns 1 = "solved" (certified polish, or OSQP's termination
// test passed), 0 on max_iter / numeric failure.  *polished: the answer in sol is a polished one.
CMPC_F void copy_iterate_to_sol(Ctx& c, const Lane& L, Sv& S);

CMPC_F int qp_solve(Ctx& c, const Lane& L, Sv& S, int* iters_out, int* nfact_out, int* polished) {
  const Params& P = *c.prm;
  int nfact = 0, solved = 0, it = 0;
  *polished = 0;
  S.fail = 0;
  factor(c, L, S, MODE_ADMM);
  ++nfact;
  if (S.fail) { *iters_out = 0; *nfact_out = nfact; return 0; }
  int next_as = (P.polish && P.as_start > 0) ? P.as_start : -1, as_step = P.as_step;
  for (it = 1; it <= P.max_iter; ++it) {
    const int check = (it % P.check_every == 0) || (it == next_as);
    backward_sweep(c, L, S, MODE_ADMM);
    forward_sweep(c, L, S, MODE_ADMM, check, 0, nullptr);
    if (!check) continue;
#if !defined(__CUDA_ARCH__) && defined(CMPC_TRACE)
    if (it % 20 == 0) fprintf(stderr, "it %d pri %.3e/%.3e dua %.3e/%.3e rho %.3g\n", it, S.pri, P.eps_abs + P.eps_rel * S.npri, S.dua, P.eps_abs + P.eps_rel * S.ndua, S.rho);
#endif
    if (!(S.pri == S.pri) || !(S.dua == S.dua)) break;   // NaN
    const int term = S.pri <= P.eps_abs + P.eps_rel * S.npri && S.dua <= P.eps_abs + P.eps_rel * S.ndua;
    if (term || it == next_as) {
      const double pri0 = S.pri, dua0 = S.dua, npri0 = S.npri, ndua0 = S.ndua;
      if (P.polish) {
        const int cert = active_set_polish(c, L, S, P.as_rounds, &nfact);
        // OSQP's rule for an uncertified polish after normal termination: keep it if it improves
        const double m0 = fmax(pri0 / (P.eps_abs + P.eps_rel * npri0), dua0 / (P.eps_abs + P.eps_rel * ndua0));
        const double m1 = S.pri / (P.eps_abs + P.eps_rel * S.npri);
        if (cert || (term && !S.fail && m1 < m0)) {
          solved = 1; *polished = 1;
          S.dua = 0.0; S.ndua = ndua0;
          break;
        }
        S.pri = pri0; S.dua = dua0; S.npri = npri0; S.ndua = ndua0;
        S.fail = 0;
      }
      if (term) { solved = 1; break; }
      next_as = it + as_step;
      as_step *= 2;
      factor(c, L, S, MODE_ADMM);     // the polish overwrote the factor records
      ++nfact;
      if (S.fail) break;
    }
    if (P.adaptive_rho && it >= P.adapt_start && it % P.check_every == 0) {
      double est = S.rho * sqrt((S.pri / (S.npri + 1e-10)) / (S.dua / (S.ndua + 1e-10) + 1e-10));
      est = fmin(fmax(est, 1e-6), 1e6);
      if (est > S.rho * P.adapt_tol || est < S.rho / P.adapt_tol) {
        Sv T = S;
        set_rho(P, T, est);
        rescale_duals(c, L, S, T.rho, T.rhok);
        S.rho = T.rho; S.rhok = T.rhok;
        factor(c, L, S, MODE_ADMM);
        ++nfact;
        if (S.fail) break;
      }
    }
  }
  if (solved && !*polished) copy_iterate_to_sol(c, L, S);
  *iters_out = it > P.max_iter ? P.max_iter : it;
  *nfact_out = nfact;
  return solved;
}

// unpolished answer: one more x-update, written to sol through the multiplier-mode forward sweep
// code path would disturb the iterate; instead re-run the ADMM sweeps' LQR solve read-only.
CMPC_F void copy_iterate_to_sol(Ctx& c, const Lane& L, Sv& S) {
  const Params& P = *c.prm;
  const int N = P.N;
  backward_sweep(c, L, S, MODE_ADMM);
  vd x = ldif(L.is_x, c.xi, L.xi);
  const vi srcA = seli(L.is_c, L.xi + (XL + 3), L.xa2 + XL);
  const vi srcB = L.xa1 + XL;
  const vi gsrc = seli(L.is_k, L.xa + XL, L.xa);
  for (int k = 0; k <= N; ++k) {
    const double* sg = c.stg + (long)k * SG;
    double* so = c.sol + (long)k * SOL;
    stif(L.is_x, so, L.xi, x);
    if (k == N) break;
    const double* fk = c.fac + (long)k * FAC;
    const int ns = c.meta[k] & 7, na = 3 * ns;
    const vb ua = L.is_u && (L.jj < na);
    vd ut = ldif(ua, c.dvec + (long)k * DVC, L.jj);
    vi idx = L.jj + F_KT;
    for (int i = 0; i < 9; ++i) {
      ut = vfma(ldif(ua, fk, idx), vconst(uni(x, XL + i)), ut);
      idx = idx + na;
    }
    stif(ua, so + SOL_U, L.jj, ut);
    const vb mact = L.is_m && (L.jj < na);
    const vd d1 = ldif(mact, sg + SG_D, L.s3 * 3 + L.a1), d2 = ldif(mact, sg + SG_D, L.s3 * 3 + L.a2);
    const vd f1 = shfl(ut, L.s3 * 3 + L.a1), f2 = shfl(ut, L.s3 * 3 + L.a2);
    const vd tq = vfma(d1, f2, -(d2 * f1));
    const vd m = sel(L.is_r, ut, sel(mact, tq, vconst(0.0)));
    vd sum = vconst(0.0);
    for (int sl = 0; sl < ns; ++sl) sum = sum + shfl(m, gsrc + 3 * sl);
    const vd t1 = shfl(x, srcA), t2 = shfl(x, srcB);
    const vd Sa1 = ldif(L.is_k, sg + SG_S, L.xa1), Sa2 = ldif(L.is_k, sg + SG_S, L.xa2);
    const vd cA = sel(L.is_c, vconst(P.dt_m), vconst(P.dt) * Sa1);
    const vd cB = vconst(-P.dt) * Sa2;
    const vd cst = sel(L.is_k, ldif(L.is_k, sg + SG_CK, L.xa), sel(L.is_x && (L.xi == 5), vconst(P.dtmg), vconst(0.0)));
    const vd fsum = sel(L.is_l || L.is_k, sum, vconst(0.0));
    x = sel(L.is_x, vfma(cA, t1, vfma(cB, t2, x)) + vfma(vconst(P.dt), fsum, cst), vconst(0.0));
  }
  wsync();
}

// ---------------------------------------------------------------- trust test and accuracy ratio
// sigma_max(X - Xbar) via the 9x9 Gram matrix + cyclic Jacobi (scp_solver.py:151: np.linalg.norm(.,2));
// rho = sum ||(f(x,u) - lin)[6:9]||^2 / sum ||lin||^2 (scp_solver.py:71-87).
CMPC_F void evaluate(Ctx& c, const Lane& L, double* snorm, double* num_out, double* den_out) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const int N = P.N;
  const double* Xr = c.Xr;
  wsync();
  // Gram matrix G = D D^T, D = X - Xbar (upper triangle, 45 entries)
  CMPC_LANES(l) {
    for (int e = l; e < 45; e += 32) {
      int i = 0, rem = e;
      while (rem >= 9 - i) { rem -= 9 - i; ++i; }
      int r = i + rem;
      double acc = 0.0;
      for (int k = 0; k <= N; ++k) {
        double di = c.sol[(long)k * SOL + i] - Xr[k * 9 + i];
        double dr = c.sol[(long)k * SOL + r] - Xr[k * 9 + r];
        acc += di * dr;
      }
      s.T[i * 9 + r] = acc;
      s.T[r * 9 + i] = acc;
    }
  }
  wsync();
  // accuracy ratio, lanes over knots
  vd vnum = vconst(0.0), vden = vconst(0.0);
  CMPC_LANES(l) {
    double num = 0.0, den = 0.0;
    for (int k = l; k < N; k += 32) {
      const double* so = c.sol + (long)k * SOL;
      const double* gk = c.stg + (long)k * SG;
      const int mt = c.meta[k];
      const int ns = mt & 7;
      double u[MAXU];
      for (int i = 0; i < MAXU; ++i) u[i] = 0.0;
      double F[3] = {0, 0, 0}, Tq[3] = {0, 0, 0};
      for (int sl = 0; sl < ns; ++sl) {
        const int cid = (mt >> (4 + 2 * sl)) & 3;
        double t[3];
        cross3(gk + SG_D + 3 * sl, so + SOL_U + 3 * sl, t);
        for (int a = 0; a < 3; ++a) {
          u[3 * cid + a] = so[SOL_U + 3 * sl + a];
          F[a] += so[SOL_U + 3 * sl + a];
          Tq[a] += t[a];
        }
      }
      // lin = A x + B u + c with the structured A, B
      double lin[9], nl[9], Sxc[3];
      cross3(gk + SG_S, so, Sxc);
      for (int a = 0; a < 3; ++a) {
        lin[a] = so[a] + P.dt_m * so[3 + a];
        lin[3 + a] = so[3 + a] + P.dt * F[a] + (a == 2 ? P.dtmg : 0.0);
        lin[6 + a] = so[6 + a] + P.dt * Sxc[a] + P.dt * Tq[a] + gk[SG_CK + a];
      }
      step_knot(P, so, u, c.cpos + (long)k * P.nc * 3, c.cact + (long)k * P.nc, nl);
      for (int i = 6; i < 9; ++i) num += (nl[i] - lin[i]) * (nl[i] - lin[i]);
      for (int i = 0; i < 9; ++i) den += lin[i] * lin[i];
    }
#if defined(__CUDA_ARCH__)
    vnum = num; vden = den;
#else
    vnum.v[l] = num; vden.v[l] = den;
#endif
  }
  *num_out = wsum(vnum);
  *den_out = wsum(vden);
  // largest eigenvalue of the Gram matrix: cyclic Jacobi on one lane
  CMPC_LANE0 {
    double* A = s.T;
    for (int sweep = 0; sweep < 12; ++sweep) {
      double off = 0.0;
      for (int i = 0; i < 9; ++i)
        for (int j = i + 1; j < 9; ++j) off += A[i * 9 + j] * A[i * 9 + j];
      double dg = 0.0;
      for (int i = 0; i < 9; ++i) dg += A[i * 9 + i] * A[i * 9 + i];
      if (off <= 1e-30 * dg || off == 0.0) break;
      for (int p = 0; p < 8; ++p) {
        for (int q = p + 1; q < 9; ++q) {
          double apq = A[p * 9 + q];
          if (apq == 0.0) continue;
          double th = (A[q * 9 + q] - A[p * 9 + p]) / (2.0 * apq);
          double t = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
          double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
          for (int r = 0; r < 9; ++r) {
            double arp = A[r * 9 + p], arq = A[r * 9 + q];
            A[r * 9 + p] = cs * arp - sn * arq;
            A[r * 9 + q] = sn * arp + cs * arq;
          }
          for (int r = 0; r < 9; ++r) {
            double apr = A[p * 9 + r], aqr = A[q * 9 + r];
            A[p * 9 + r] = cs * apr - sn * aqr;
            A[q * 9 + r] = sn * apr + cs * aqr;
          }
        }
      }
    }
    double mx = 0.0;
    for (int i = 0; i < 9; ++i) mx = fmax(mx, A[i * 9 + i]);
    A[0] = sqrt(mx);
  }
  wsync();
  *snorm = s.T[0];
  wsync();
}

// ---------------------------------------------------------------- per-instance setup
// K1 for every knot (lanes over knots), friction table when not on the fast path, start of the
// iterate at the linearisation point, constant parts of the residual norms.
CMPC_F void setup_instance(Ctx& c, const Lane& L, Sv& S) {
  const Params& P = *c.prm;
  const int N = P.N;
  vd vq = vconst(0.0), vc = vconst(0.0);
  CMPC_LANES(l) {
    double mq = 0.0, mc = 0.0;
    for (int k = l; k <= N; k += 32) {
      double* gk = c.stg + (long)k * SG;
      double* sk = c.sta + (long)k * ST;
      const int kk = k < N ? k : N - 1;
      const int* act = c.cact + (long)kk * P.nc;
      const int mt = linearize_knot(P, c.Xr + k * 9, c.Ui + kk * P.nu, c.cpos + (long)kk * P.nc * 3, act, k == N, gk);
      c.meta[k] = mt;
      for (int i = 0; i < 9; ++i) mq = fmax(mq, fabs(P.Wx[i] * gk[SG_XB + i]));
      for (int r = 0; r < ST; ++r) sk[r] = 0.0;
      for (int i = 0; i < 3; ++i) sk[ST_VK + i] = c.Xr[k * 9 + 6 + i];
      if (k < N) {
        mc = fmax(mc, fabs(P.dtmg));
        for (int i = 0; i < 3; ++i) mc = fmax(mc, fabs(gk[SG_CK + i]));
        const int ns = mt & 7;
        for (int sl = 0; sl < ns; ++sl) {
          const int cid = (mt >> (4 + 2 * sl)) & 3;
          const double* ub = c.Ui + k * P.nu + 3 * cid;
          if (!P.fast) {
            double* gt = c.gtab + ((long)k * MAXC + sl) * 16;
            for (int row = 0; row < 4; ++row) {
              double mx = 0.0;
              for (int a = 0; a < 3; ++a) {
                double g = 0.0;
                if (c.cR) {
                  const double* R = c.cR + ((long)k * P.nc + cid) * 9;
                  for (int b2 = 0; b2 < 3; ++b2) g += pyr4(P, row, b2) * R[a * 3 + b2];
                } else {
                  g = pyr4(P, row, a);
                }
                gt[row * 3 + a] = g;
                mx = fmax(mx, fabs(g) / sqrt(P.Wu[3 * cid + a]));
              }
              gt[12 + row] = mx > 0.0 ? 1.0 / (mx * mx) : 0.0;
            }
          }
          for (int row = 0; row < 4; ++row) {
            double cf = 0.0;
            for (int a = 0; a < 3; ++a) cf += fric_G(c, k, sl, row, a) * ub[a];
            sk[ST_VF + 4 * sl + row] = fmin(cf, 0.0);
          }
        }
      }
    }
#if defined(__CUDA_ARCH__)
    vq = mq; vc = mc;
#else
    vq.v[l] = mq; vc.v[l] = mc;
#endif
  }
  S.nq = wmax(vq);
  double mi = 0.0;
  for (int i = 0; i < 9; ++i) mi = fmax(mi, fabs(c.xi[i]));
  S.dynrow = fmax(wmax(vc), mi);
  S.ye = vconst(0.0);
  S.kap = 0;
  S.fail = 0;
  S.n_pmm = S.n_polish = 0;
  S.pri = S.dua = S.npri = S.ndua = 0.0;
  wsync();
}

CMPC_F void write_solution(Ctx& c) {
  const Params& P = *c.prm;
  const int N = P.N;
  wsync();
  CMPC_LANES(l) {
    for (int e = l; e < (N + 1) * 9; e += 32)
      c.bt.X_out[(long)c.b * (N + 1) * 9 + e] = c.sol[(long)(e / 9) * SOL + e % 9];
    for (int e = l; e < N * P.nu; e += 32) {
      const int k = e / P.nu, j = e % P.nu, ct = j / 3, a = j % 3;
      const int mt = c.meta[k];
      const int ns = mt & 7;
      double v = 0.0;
      for (int sl = 0; sl < ns; ++sl)
        if (((mt >> (4 + 2 * sl)) & 3) == ct) v = c.sol[(long)k * SOL + SOL_U + 3 * sl + a];
      c.bt.U_out[(long)c.b * N * P.nu + e] = v;
    }
  }
  wsync();
}

// ---------------------------------------------------------------- the SCP loop of one instance
// scp_solver.py:118-179.  The linearisation point never moves (:129-130), so the stage records
// are built once; each SCP iteration re-solves the QP for the current (radius, weight).
CMPC_F void solve_instance(Ctx& c) {
  const Params& P = *c.prm;
  const Lane L = make_lane(P);
  Sv S;
  setup_instance(c, L, S);
  double radius = P.radius0, weight = P.omega0;
  int it = 0, success = 0, n_acc = 0, status = ST_OK, qp_total = 0, nf_total = 0, polished = 0;
  double snorm = 0.0, acc = 0.0;
  set_rho(P, S, P.rho0);
  while (it < P.max_scp && weight < P.omega_max && !(it != 0 && success && 0.0 < P.conv_thresh)) {
    success = 0;
    S.radius = radius;
    S.weight = weight;
    int qi = 0, nf = 0;
    const int solved = qp_solve(c, L, S, &qi, &nf, &polished);
    qp_total += qi;
    nf_total += nf;
    if (!solved) { status = S.fail ? ST_QP_NUMERIC : ST_QP_MAXITER; break; }
    double num = 0.0, den = 1.0;
    evaluate(c, L, &snorm, &num, &den);
    if (snorm < radius) {
      acc = num / den;
      if (acc > P.acc_rho1) {
        radius *= P.beta_fail;
      } else {
        write_solution(c);
        success = 1;
        ++n_acc;
        if (acc < P.acc_rho0) radius = fmin(P.beta_succ * radius, P.radius0);
      }
    } else {
      weight *= P.gamma_fail;
    }
    ++it;
  }
  // nothing accepted: hand back the last QP solution (n_accepted == 0 tells the caller; the
  // reference returns empty lists in that case)
  if (n_acc == 0 && status == ST_OK && it > 0) write_solution(c);
  CMPC_LANE0 {
    c.bt.scp_iters[c.b] = it;
    c.bt.status[c.b] = status;
    c.bt.n_accepted[c.b] = n_acc;
    c.bt.qp_iters[c.b] = qp_total;
    c.bt.n_factor[c.b] = nf_total;
    double* inf = c.bt.info + (long)c.b * INFO;
    inf[0] = snorm; inf[1] = acc; inf[2] = S.pri; inf[3] = S.dua;
    inf[4] = S.rho; inf[5] = radius; inf[6] = weight; inf[7] = (double)polished;
    inf[8] = (double)S.n_pmm; inf[9] = (double)S.n_polish; inf[10] = 0.0; inf[11] = 0.0;
  }
  wsync();
}

// bind the per-instance pointers
CMPC_HD void bind_instance(Ctx& c, const Params* prm, const Batch& bt, WarpMem* s, int b) {
  c.prm = prm; c.bt = bt; c.s = s; c.b = b;
  const int N = prm->N;
  const long plan = (long)b * bt.plan_stride;
  c.cpos = bt.cpos + plan * N * prm->nc * 3;
  c.cR = bt.cR ? bt.cR + plan * N * prm->nc * 9 : nullptr;
  c.cact = bt.cact + plan * N * prm->nc;
  c.Xr = bt.X_ref + (long)b * (N + 1) * 9;
  c.Ui = bt.U_init + (long)b * N * prm->nu;
  c.xi = bt.x_init + (long)b * 9;
  c.xf = bt.x_final + (long)b * 9;
  c.stg = bt.stg + (long)b * (N + 1) * SG;
  c.sta = bt.sta + (long)b * (N + 1) * ST;
  c.fac = bt.fac + (long)b * N * FAC;
  c.dvec = bt.dvec + (long)b * N * DVC;
  c.pm = bt.pm + (long)b * (N + 1) * PM;
  c.sol = bt.sol + (long)b * (N + 1) * SOL;
  c.gtab = bt.gtab ? bt.gtab + (long)b * N * MAXC * 16 : nullptr;
  c.meta = bt.meta + (long)b * (N + 1);
  c.pmask = bt.pmask + (long)b * (N + 1);
}

}  // namespace cmpc
