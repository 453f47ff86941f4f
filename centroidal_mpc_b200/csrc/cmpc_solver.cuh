// cmpc_solver.cuh — Riccati factorisation, ADMM sweeps, polish, trust-region loop.
// See cmpc_core.cuh for the execution model (warp-uniform driver code + lane-parallel phases)
// and DESIGN.md "device algorithm" for the mathematics.  oracle/device_model.py is the
// executable numpy specification of this file.
#pragma once
#include "cmpc_core.cuh"

namespace cmpc {

constexpr int MODE_ADMM = 0, MODE_POLISH = 1;

// ---------------------------------------------------------------- small lane-local helpers
CMPC_HD int popc4(int m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1); }
// contact of the j-th compact control (j/3-th set bit of the active mask)
CMPC_HD int contact_of(int mask, int j) {
  int want = j / 3;
  for (int c = 0; c < MAXC; ++c) {
    if ((mask >> c) & 1) {
      if (want == 0) return c;
      --want;
    }
  }
  return 0;
}

// prox of  omega*max(0, |v - kbar|_1 - r)  with weight rho (oracle/device_model.py prox_trust)
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
  // projection onto the L1 ball of radius r: sort descending (3 elements)
  double s[3] = {ab[0], ab[1], ab[2]};
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

// ---------------------------------------------------------------- friction rows G (4x3/contact)
// G = pyr4 * R^T, pyr4 = [[1,0,-k],[-1,0,-k],[0,1,-k],[0,-1,-k]], k = mu/sqrt2
// (utils.py:9-16, constraints.py:178-184).  Also the row equilibration factors e^2 under the
// variable scaling D_u = 1/sqrt(W_u).
CMPC_HD void fill_G_phase(Ctx& c, int k) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const double kf = P.mu * 0.70710678118654752440;
  CMPC_LANES(l) {
    if (l < 4 * P.nc) {
      int ct = l >> 2, row = l & 3;
      double pr[3] = {row == 0 ? 1.0 : (row == 1 ? -1.0 : 0.0), row == 2 ? 1.0 : (row == 3 ? -1.0 : 0.0), -kf};
      double g[3];
      if (P.identity_R) {
        g[0] = pr[0]; g[1] = pr[1]; g[2] = pr[2];
      } else {
        const double* R = c.cR + ((long)k * P.nc + ct) * 9;
        for (int a = 0; a < 3; ++a) g[a] = pr[0] * R[a * 3 + 0] + pr[1] * R[a * 3 + 1] + pr[2] * R[a * 3 + 2];
      }
      double mx = 0.0;
      for (int a = 0; a < 3; ++a) {
        s.G[ct][row * 3 + a] = g[a];
        mx = fmax(mx, fabs(g[a]) / sqrt(P.Wu[3 * ct + a]));
      }
      s.ef2[l] = mx > 0.0 ? 1.0 / (mx * mx) : 0.0;
    }
  }
}


// value of friction row (ct,row) at knot k for control vector u (full layout): G_row . f_ct
CMPC_HD double friction_row_value(const Ctx& c, int k, int ct, int row, const double* u) {
  const Params& P = *c.prm;
  const double kf = P.mu * 0.70710678118654752440;
  const double pr[3] = {row == 0 ? 1.0 : (row == 1 ? -1.0 : 0.0), row == 2 ? 1.0 : (row == 3 ? -1.0 : 0.0), -kf};
  double cf = 0.0;
  if (P.identity_R) {
    for (int a = 0; a < 3; ++a) cf += pr[a] * u[3 * ct + a];
  } else {
    const double* R = c.cR + ((long)k * P.nc + ct) * 9;
    for (int a = 0; a < 3; ++a) cf += (pr[0] * R[a * 3] + pr[1] * R[a * 3 + 1] + pr[2] * R[a * 3 + 2]) * u[3 * ct + a];
  }
  return cf;
}

// ---------------------------------------------------------------- per-knot penalty/linear terms
// Fills rrow/lrow (friction rows) and kM/kl (kappa block) for knot k from the loaded records
// s.stg / s.sta (ADMM) or the polish records.  One phase; caller syncs.
CMPC_HD void knot_terms_phase(Ctx& c, int k, int mode) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const int mask = (k < P.N) ? (int)s.stg[O_ACT] : 0;
  CMPC_LANES(l) {
    if (l < 16) {
      int ct = l >> 2;
      double rr = 0.0, lr = 0.0;
      if (ct < P.nc && ((mask >> ct) & 1)) {
        if (mode == MODE_ADMM) {
          rr = s.sc[SC_RHO] * s.ef2[l];
          lr = rr * fabs(s.sta[O_VF + l]);             // -(rho w - y) = rho |v|
        } else {
          int pm = c.pmask[k];
          if ((pm >> l) & 1) { rr = 1.0 / P.delta; lr = c.pol[(long)k * POL + l]; }
        }
      }
      s.rrow[l] = rr;
      s.lrow[l] = lr;
    } else if (l == 16) {
      double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, kl[3] = {0, 0, 0};
      if (k >= 1) {
        if (mode == MODE_ADMM) {
          double rk = s.sc[SC_RHOK], w[3];
          prox_trust(&s.sta[O_VK], &s.stg[O_KB], s.sc[SC_RADIUS], s.sc[SC_WEIGHT], rk, w);
          for (int i = 0; i < 3; ++i) { M[4 * i] = rk; kl[i] = -rk * (2.0 * w[i] - s.sta[O_VK + i]); }
        } else {
          int pm = c.pmask[k];
          int br = (pm >> 16) & 3;
          if (br != 0) {
            const double inv = 1.0 / P.delta;
            const double* yk = c.pol + (long)k * POL + 16;
            double sg[3];
            for (int i = 0; i < 3; ++i) {
              int code = (pm >> (18 + 2 * i)) & 3;
              sg[i] = code == 1 ? 1.0 : (code == 2 ? -1.0 : 0.0);
              if (code == 0) {   // pinned component: kappa_i = kbar_i
                M[4 * i] += inv;
                kl[i] -= inv * s.stg[O_KB + i] - yk[i];
              }
            }
            if (br == 1) {
              for (int i = 0; i < 3; ++i) kl[i] += s.sc[SC_WEIGHT] * sg[i];
            } else {             // surface: sg'(kappa - kbar) = radius
              double bb = s.sc[SC_RADIUS];
              for (int i = 0; i < 3; ++i) bb += sg[i] * s.stg[O_KB + i];
              for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) M[3 * i + j] += inv * sg[i] * sg[j];
                kl[i] -= sg[i] * (inv * bb - yk[3]);
              }
            }
          }
        }
      }
      for (int i = 0; i < 9; ++i) s.kM[i] = M[i];
      for (int i = 0; i < 3; ++i) s.kl[i] = kl[i];
    }
  }
}

// ---------------------------------------------------------------- Riccati factorisation
// Backward over k: H_uu = R + B'PB, H_ux = B'PA, Hinv = H_uu^-1 (in-place Gauss-Jordan),
// K = -Hinv H_ux, Pc = P c, P <- Q + A'PA + H_ux' K.   Writes fac[k] = {K, Hinv, Pc}.
// Returns 0 on success, 1 if a pivot was not positive.
CMPC_HD int factor(Ctx& c, int mode) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const int N = P.N;
  const double sg = (mode == MODE_ADMM) ? P.sigma : P.delta;
  const double rho_e = (mode == MODE_ADMM) ? s.sc[SC_RHOE] : s.sc[SC_RHOEP];
  // terminal knot: P = Q_N + rho_e I
  CMPC_COPY(s.stg, c.stg + (long)N * STG, STG);
  CMPC_COPY(s.sta, c.sta + (long)N * STA, STA);
  CMPC_SYNC();
  knot_terms_phase(c, N, mode);
  CMPC_SYNC();
  CMPC_LANES(l) {
    for (int e = l; e < 81; e += 32) {
      int i = e / 9, r = e % 9;
      double v = (i == r) ? (P.Wx[i] + sg + rho_e) : 0.0;
      if (i >= 6 && r >= 6) v += s.kM[(i - 6) * 3 + (r - 6)];
      s.P[e] = v;
    }
    if (l == 0) s.sc[15] = 0.0;   // failure flag
  }
  CMPC_SYNC();
  for (int k = N - 1; k >= 0; --k) {
    CMPC_COPY(s.stg, c.stg + (long)k * STG, STG);
    CMPC_COPY(s.sta, c.sta + (long)k * STA, STA);
    if (!P.identity_R) fill_G_phase(c, k);
    CMPC_SYNC();
    const int mask = (int)s.stg[O_ACT];
    const int na = 3 * popc4(mask);
    double* fk = c.fac + (long)k * FAC;
    knot_terms_phase(c, k, mode);
    // Pc = P c (c has entries 5..8 only);  PBt[j][i] = (P B)[i][j] = B_col(j) . P_row(i)
    CMPC_LANES(l) {
      if (l < 9) {
        double v = 0.0;
        for (int r = 5; r < 9; ++r) v += s.P[l * 9 + r] * s.stg[O_C + r];
        fk[O_PC + l] = v;
      }
      for (int e = l; e < na * 9; e += 32) {
        int j = e / 9, i = e % 9;
        int ct = contact_of(mask, j);
        s.PB[e] = BTv_elem(P, &s.stg[O_D + 3 * ct], &s.P[i * 9], j % 3);
      }
    }
    CMPC_SYNC();
    // Huu[j][l2] = R_jl + B_col(j) . PBt[l2] ;  Hux[j][i] = (A^T PBt[j])[i]
    CMPC_LANES(l) {
      for (int e = l; e < na * na; e += 32) {
        int j = e / na, j2 = e % na;
        int cj = contact_of(mask, j), c2 = contact_of(mask, j2);
        int aj = j % 3, a2 = j2 % 3;
        double v = BTv_elem(P, &s.stg[O_D + 3 * cj], &s.PB[j2 * 9], aj);
        if (cj == c2) {
          for (int row = 0; row < 4; ++row)
            v += s.rrow[4 * cj + row] * s.G[cj][row * 3 + aj] * s.G[cj][row * 3 + a2];
          if (j == j2) v += P.Wu[3 * cj + aj] + sg;
        }
        s.Huu[e] = v;
      }
      for (int e = l; e < na * 9; e += 32) {
        int j = e / 9, i = e % 9;
        s.PB[108 + e] = ATv_elem(P, &s.stg[O_S], &s.PB[j * 9], i);
      }
    }
    CMPC_SYNC();
    // in-place Gauss-Jordan inversion of the SPD matrix Huu (no pivoting needed)
    for (int pv = 0; pv < na; ++pv) {
      CMPC_LANES(l) {
        if (l < na) { s.prow[l] = s.Huu[pv * na + l]; s.pcol[l] = s.Huu[l * na + pv]; }
      }
      CMPC_SYNC();
      const double piv = s.prow[pv];
      if (!(piv > 0.0)) {
        CMPC_LANES(l) { if (l == 0) s.sc[15] = 1.0; }
      }
      const double ip = 1.0 / piv;
      CMPC_LANES(l) {
        for (int e = l; e < na * na; e += 32) {
          int i = e / na, j = e % na;
          double v;
          if (i == pv && j == pv) v = ip;
          else if (i == pv) v = s.prow[j] * ip;
          else if (j == pv) v = -s.pcol[i] * ip;
          else v = s.Huu[e] - s.pcol[i] * s.prow[j] * ip;
          s.Huu[e] = v;
        }
      }
      CMPC_SYNC();
    }
    // symmetrise Hinv, K = -Hinv Hux, write the factor record
    CMPC_LANES(l) {
      for (int e = l; e < na * 9; e += 32) {
        int j = e / 9, i = e % 9;
        double v = 0.0;
        for (int j2 = 0; j2 < na; ++j2)
          v -= 0.5 * (s.Huu[j * na + j2] + s.Huu[j2 * na + j]) * s.PB[108 + j2 * 9 + i];
        s.K[e] = v;
        fk[O_K + e] = v;
      }
      for (int e = l; e < na * na; e += 32) {
        int j = e / na, j2 = e % na;
        fk[O_HI + e] = 0.5 * (s.Huu[e] + s.Huu[j2 * na + j]);
      }
      // T = P A  (row i of T = A^T applied to row i of P)
      for (int e = l; e < 81; e += 32) {
        int i = e / 9, r = e % 9;
        s.T[e] = ATv_elem(P, &s.stg[O_S], &s.P[i * 9], r);
      }
    }
    CMPC_SYNC();
    // P <- Q_k + A^T T + Hux^T K   (upper triangle computed, mirrored)
    CMPC_LANES(l) {
      for (int e = l; e < 45; e += 32) {
        int i = 0, rem = e;
        while (rem >= 9 - i) { rem -= 9 - i; ++i; }
        int r = i + rem;
        double v = ATv_elem(P, &s.stg[O_S], &s.T[r], i, 9);
        for (int j = 0; j < na; ++j) v += s.PB[108 + j * 9 + i] * s.K[j * 9 + r];
        if (i == r) v += P.Wx[i] + sg;
        if (i >= 6) v += 0.5 * (s.kM[(i - 6) * 3 + (r - 6)] + s.kM[(r - 6) * 3 + (i - 6)]);
        s.P[i * 9 + r] = v;
        s.P[r * 9 + i] = v;
      }
    }
    CMPC_SYNC();
  }
  return s.sc[15] != 0.0;
}

// ---------------------------------------------------------------- backward sweep (linear term)
// p_N = qx_N;  g = p + Pc;  hu = ru + B'g;  d = -Hinv hu;  p = qx + A'g + K'hu.
CMPC_HD void backward_sweep(Ctx& c, int mode) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const int N = P.N;
  const double sg = (mode == MODE_ADMM) ? P.sigma : P.delta;
  const double rho_e = (mode == MODE_ADMM) ? s.sc[SC_RHOE] : s.sc[SC_RHOEP];
  const double* xf = c.bt.x_final + (long)c.b * 9;
  CMPC_COPY(s.stg, c.stg + (long)N * STG, STG);
  CMPC_COPY(s.sta, c.sta + (long)N * STA, STA);
  CMPC_SYNC();
  knot_terms_phase(c, N, mode);
  CMPC_SYNC();
  CMPC_LANES(l) {
    if (l < 9) {
      double v = s.stg[O_Q + l] - sg * s.sta[O_X + l] - (rho_e * xf[l] - s.ye[l]);
      if (l >= 6) v += s.kl[l - 6];
      s.p[l] = v;
    }
  }
  CMPC_SYNC();
  for (int k = N - 1; k >= 0; --k) {
    const double* fk = c.fac + (long)k * FAC;
    CMPC_COPY(s.stg, c.stg + (long)k * STG, STG);
    CMPC_COPY(s.sta, c.sta + (long)k * STA, STA);
    if (!P.identity_R) fill_G_phase(c, k);
    CMPC_SYNC();
    const int mask = (int)s.stg[O_ACT];
    const int na = 3 * popc4(mask);
    CMPC_COPY(s.K, fk + O_K, na * 9);
    CMPC_COPY(s.Huu, fk + O_HI, na * na);
    knot_terms_phase(c, k, mode);
    CMPC_LANES(l) {
      if (l >= 20 && l < 29) s.g[l - 20] = s.p[l - 20] + fk[O_PC + l - 20];
    }
    CMPC_SYNC();
    CMPC_LANES(l) {
      if (l < na) {
        int ct = contact_of(mask, l), a = l % 3;
        double v = -sg * s.sta[O_U + 3 * ct + a] + BTv_elem(P, &s.stg[O_D + 3 * ct], s.g, a);
        for (int row = 0; row < 4; ++row) v += s.G[ct][row * 3 + a] * s.lrow[4 * ct + row];
        s.hu[l] = v;
      }
    }
    CMPC_SYNC();
    CMPC_LANES(l) {
      if (l < na) {
        double v = 0.0;
        for (int j = 0; j < na; ++j) v -= s.Huu[l * na + j] * s.hu[j];
        c.dvec[(long)k * DVC + l] = v;
      } else if (l >= 16 && l < 25) {
        int i = l - 16;
        double v = s.stg[O_Q + i] - sg * s.sta[O_X + i] + ATv_elem(P, &s.stg[O_S], s.g, i);
        if (i >= 6) v += s.kl[i - 6];
        for (int j = 0; j < na; ++j) v += s.K[j * 9 + i] * s.hu[j];
        s.p[i] = v;
      }
    }
    CMPC_SYNC();
  }
}

// ---------------------------------------------------------------- forward sweep + local updates
// u~ = K x~ + d,  x~+ = A x~ + B u~ + c.  ADMM mode: relaxation, friction / kappa / terminal
// projections and dual updates, all knot-local.  Polish mode: plain assignment + multiplier
// updates of the proximal method of multipliers.
CMPC_HD void forward_sweep(Ctx& c, int mode) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const int N = P.N;
  const double al = (mode == MODE_ADMM) ? P.alpha : 1.0;
  const double* xi = c.bt.x_init + (long)c.b * 9;
  const double* xf = c.bt.x_final + (long)c.b * 9;
  const double inv = 1.0 / P.delta;
  double* xa = s.xk;
  double* xb = s.xn;
  CMPC_LANES(l) { if (l < 9) xa[l] = xi[l]; }
  CMPC_SYNC();
  for (int k = 0; k <= N; ++k) {
    double* sk = c.sta + (long)k * STA;
    CMPC_COPY(s.stg, c.stg + (long)k * STG, STG);
    CMPC_COPY(s.sta, sk, STA);
    if (k < N && !P.identity_R) fill_G_phase(c, k);
    CMPC_SYNC();
    const int mask = (k < N) ? (int)s.stg[O_ACT] : 0;
    const int na = 3 * popc4(mask);
    if (k < N) {
      CMPC_COPY(s.K, c.fac + (long)k * FAC + O_K, na * 9);
      CMPC_COPY(s.dd, c.dvec + (long)k * DVC, na);
      CMPC_SYNC();
    }
    // phase A: controls u~ (+ relaxed u), state part of knot k (x, kappa copy / multipliers)
    CMPC_LANES(l) {
      if (l < na) {
        double v = s.dd[l];
        for (int i = 0; i < 9; ++i) v += s.K[l * 9 + i] * xa[i];
        s.ut[l] = v;
        int ct = contact_of(mask, l), a = l % 3;
        sk[O_U + 3 * ct + a] = al * v + (1.0 - al) * s.sta[O_U + 3 * ct + a];
      } else if (l >= 16 && l < 25) {
        int i = l - 16;
        sk[O_X + i] = al * xa[i] + (1.0 - al) * s.sta[O_X + i];
      } else if (l == 25 && k >= 1) {
        if (mode == MODE_ADMM) {
          double rk = s.sc[SC_RHOK], w[3], a3[3];
          prox_trust(&s.sta[O_VK], &s.stg[O_KB], s.sc[SC_RADIUS], s.sc[SC_WEIGHT], rk, w);
          for (int i = 0; i < 3; ++i) a3[i] = al * xa[6 + i] + (1.0 - al) * w[i] + (s.sta[O_VK + i] - w[i]);
          for (int i = 0; i < 3; ++i) sk[O_VK + i] = a3[i];
        } else {
          int pm = c.pmask[k];
          int br = (pm >> 16) & 3;
          if (br != 0) {
            double* yk = c.pol + (long)k * POL + 16;
            double sgn[3], acc = -s.sc[SC_RADIUS];
            for (int i = 0; i < 3; ++i) {
              int code = (pm >> (18 + 2 * i)) & 3;
              sgn[i] = code == 1 ? 1.0 : (code == 2 ? -1.0 : 0.0);
              if (code == 0) yk[i] += inv * (xa[6 + i] - s.stg[O_KB + i]);
              acc += sgn[i] * (xa[6 + i] - s.stg[O_KB + i]);
            }
            if (br == 2) yk[3] += inv * acc;
          }
        }
      } else if (l == 26 && k == N) {
        // terminal equality: w_e = x_final;  y_e += rho_e (al x~ + (1-al) x_f - x_f)
        const double re = (mode == MODE_ADMM) ? s.sc[SC_RHOE] : s.sc[SC_RHOEP];
        for (int i = 0; i < 9; ++i) s.ye[i] += re * al * (xa[i] - xf[i]);
      }
    }
    CMPC_SYNC();
    if (k == N) break;
    // phase B: next state, friction rows
    CMPC_LANES(l) {
      if (l < 9) {
        double v = Av_elem(P, &s.stg[O_S], xa, l) + s.stg[O_C + l];
        for (int j = 0; j < na; ++j) {
          int ct = contact_of(mask, j);
          v += Bcol_elem(P, &s.stg[O_D + 3 * ct], j % 3, l) * s.ut[j];
        }
        xb[l] = v;
      } else if (l >= 16) {
        int r = l - 16, ct = r >> 2, row = r & 3;
        if (ct < P.nc && ((mask >> ct) & 1)) {
          int base = 3 * popc4(mask & ((1 << ct) - 1));
          double cf = 0.0;
          for (int a = 0; a < 3; ++a) cf += s.G[ct][row * 3 + a] * s.ut[base + a];
          if (mode == MODE_ADMM) {
            double vo = s.sta[O_VF + r];
            sk[O_VF + r] = al * cf + (1.0 - al) * fmin(vo, 0.0) + fmax(vo, 0.0);
          } else if ((c.pmask[k] >> r) & 1) {
            c.pol[(long)k * POL + r] += inv * cf;
          }
        }
      }
    }
    CMPC_SYNC();
    double* t = xa; xa = xb; xb = t;
  }
}

// ---------------------------------------------------------------- KKT residuals
// Primal: split rows (ADMM) or original-constraint violation (polish).  Dual: costate recursion
// lam_k = Wx x_k + q_k + cx_k + A_k' lam_{k+1} makes the x rows of the stationarity condition
// hold exactly; the u rows carry the dual residual.  Norms follow OSQP's definitions
// (unscaled, infinity norm).  Results in sc[SC_PRI], sc[SC_DUA], sc[SC_NPRI], sc[SC_NDUA].
CMPC_HD void residuals(Ctx& c, int mode) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const int N = P.N;
  const double* xi = c.bt.x_init + (long)c.b * 9;
  const double* xf = c.bt.x_final + (long)c.b * 9;
  // red[0]: pri, red[1]: dua, red[2]: max(|Az|,|w|), red[3]: max(|Px|,|A'y|,|q|)
  CMPC_LANES(l) {
    s.PB[32 * 0 + l] = 0.0; s.PB[32 * 1 + l] = 0.0; s.PB[32 * 2 + l] = 0.0; s.PB[32 * 3 + l] = 0.0;
  }
  CMPC_COPY(s.stg, c.stg + (long)N * STG, STG);
  CMPC_COPY(s.sta, c.sta + (long)N * STA, STA);
  CMPC_SYNC();
  double* la = s.lam;
  double* lb = s.g;
  for (int k = N; k >= 0; --k) {
    if (k < N) {
      CMPC_COPY(s.stg, c.stg + (long)k * STG, STG);
      CMPC_COPY(s.sta, c.sta + (long)k * STA, STA);
      if (!P.identity_R) fill_G_phase(c, k);
      CMPC_SYNC();
    }
    const int mask = (k < N) ? (int)s.stg[O_ACT] : 0;
    const int na = 3 * popc4(mask);
    // phase 1: multipliers per friction row -> lrow; kappa multiplier -> kl; primal residuals
    CMPC_LANES(l) {
      if (l < 16) {
        int ct = l >> 2, row = l & 3;
        double y = 0.0;
        if (k < N && ct < P.nc && ((mask >> ct) & 1)) {
          double cf = 0.0;
          for (int a = 0; a < 3; ++a) cf += s.G[ct][row * 3 + a] * s.sta[O_U + 3 * ct + a];
          if (mode == MODE_ADMM) {
            double v = s.sta[O_VF + l], w = fmin(v, 0.0);
            y = s.sc[SC_RHO] * s.ef2[l] * fmax(v, 0.0);
            s.PB[32 * 0 + l] = fmax(s.PB[32 * 0 + l], fabs(cf - w));
            s.PB[32 * 2 + l] = fmax(s.PB[32 * 2 + l], fmax(fabs(cf), fabs(w)));
          } else {
            if ((c.pmask[k] >> l) & 1) y = c.pol[(long)k * POL + l];
            s.PB[32 * 0 + l] = fmax(s.PB[32 * 0 + l], fmax(cf, 0.0));
            s.PB[32 * 2 + l] = fmax(s.PB[32 * 2 + l], fabs(cf));
          }
        }
        s.lrow[l] = y;
      } else if (l == 16) {
        double yk[3] = {0, 0, 0};
        if (k >= 1) {
          if (mode == MODE_ADMM) {
            double rk = s.sc[SC_RHOK], w[3];
            prox_trust(&s.sta[O_VK], &s.stg[O_KB], s.sc[SC_RADIUS], s.sc[SC_WEIGHT], rk, w);
            for (int i = 0; i < 3; ++i) {
              yk[i] = rk * (s.sta[O_VK + i] - w[i]);
              s.PB[32 * 0 + l] = fmax(s.PB[32 * 0 + l], fabs(s.sta[O_X + 6 + i] - w[i]));
              s.PB[32 * 2 + l] = fmax(s.PB[32 * 2 + l], fmax(fabs(s.sta[O_X + 6 + i]), fabs(w[i])));
            }
          } else {
            int pm = c.pmask[k];
            int br = (pm >> 16) & 3;
            if (br != 0) {
              const double* yp = c.pol + (long)k * POL + 16;
              for (int i = 0; i < 3; ++i) {
                int code = (pm >> (18 + 2 * i)) & 3;
                double sgn = code == 1 ? 1.0 : (code == 2 ? -1.0 : 0.0);
                if (code == 0) yk[i] = yp[i];
                else yk[i] = sgn * (br == 1 ? s.sc[SC_WEIGHT] : yp[3]);
              }
            }
            for (int i = 0; i < 3; ++i) s.PB[32 * 2 + l] = fmax(s.PB[32 * 2 + l], fabs(s.sta[O_X + 6 + i]));
          }
        }
        for (int i = 0; i < 3; ++i) s.kl[i] = yk[i];
      } else if (l == 17) {
        double m2 = 0.0;
        if (k < N) for (int i = 5; i < 9; ++i) m2 = fmax(m2, fabs(s.stg[O_C + i]));   // dynamics rows
        if (k == 0) for (int i = 0; i < 9; ++i) m2 = fmax(m2, fabs(xi[i]));            // initial rows
        if (k == N) {
          for (int i = 0; i < 9; ++i) {
            s.PB[32 * 0 + l] = fmax(s.PB[32 * 0 + l], fabs(s.sta[O_X + i] - xf[i]));
            m2 = fmax(m2, fmax(fabs(s.sta[O_X + i]), fabs(xf[i])));
          }
        }
        s.PB[32 * 2 + l] = fmax(s.PB[32 * 2 + l], m2);
      }
    }
    CMPC_SYNC();
    // phase 2: u rows (dual residual) and the costate recursion
    CMPC_LANES(l) {
      if (l < na) {
        int ct = contact_of(mask, l), a = l % 3;
        double gy = 0.0;
        for (int row = 0; row < 4; ++row) gy += s.G[ct][row * 3 + a] * s.lrow[4 * ct + row];
        double aty = gy + BTv_elem(P, &s.stg[O_D + 3 * ct], la, a);
        double px = P.Wu[3 * ct + a] * s.sta[O_U + 3 * ct + a];
        s.PB[32 * 1 + l] = fmax(s.PB[32 * 1 + l], fabs(px + aty));
        s.PB[32 * 3 + l] = fmax(s.PB[32 * 3 + l], fmax(fabs(px), fabs(aty)));
      } else if (l >= 16 && l < 25) {
        int i = l - 16;
        double px = P.Wx[i] * s.sta[O_X + i], q = s.stg[O_Q + i];
        double v = px + q + (i >= 6 ? s.kl[i - 6] : 0.0);
        if (k == N) v += s.ye[i];
        else v += ATv_elem(P, &s.stg[O_S], la, i);
        lb[i] = v;
        if (k >= 1) s.PB[32 * 3 + l] = fmax(s.PB[32 * 3 + l], fmax(fabs(px), fabs(px + q)));   // (A'y)_x = -(Px+q)_x
        s.PB[32 * 3 + l] = fmax(s.PB[32 * 3 + l], fabs(q));
      }
    }
    CMPC_SYNC();
    double* t = la; la = lb; lb = t;
  }
  CMPC_REDUCE_MAX(c, 0, SC_PRI);
  CMPC_REDUCE_MAX(c, 1, SC_DUA);
  CMPC_REDUCE_MAX(c, 2, SC_NPRI);
  CMPC_REDUCE_MAX(c, 3, SC_NDUA);
}

// ---------------------------------------------------------------- rho change: keep (w, y), move v
CMPC_HD void rescale_duals(Ctx& c, double rho_old, double rho_new, double rk_old) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const double ratio = rho_old / rho_new;
  CMPC_LANES(l) {
    for (int k = l; k <= P.N; k += 32) {
      double* sk = c.sta + (long)k * STA;
      if (k < P.N)
        for (int r = 0; r < 16; ++r) { double v = sk[O_VF + r]; sk[O_VF + r] = fmin(v, 0.0) + ratio * fmax(v, 0.0); }
      if (k >= 1) {
        double w[3];
        prox_trust(&sk[O_VK], &c.stg[(long)k * STG + O_KB], s.sc[SC_RADIUS], s.sc[SC_WEIGHT], rk_old, w);
        for (int i = 0; i < 3; ++i) sk[O_VK + i] = w[i] + ratio * (sk[O_VK + i] - w[i]);
      }
    }
  }
  CMPC_SYNC();
}

CMPC_HD void set_rho(Ctx& c, double rho) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  CMPC_LANES(l) {
    if (l == 0) {
      double wk = fmin(P.Wx[6], fmin(P.Wx[7], P.Wx[8]));
      double wm = 0.0;
      for (int i = 0; i < 9; ++i) wm = fmax(wm, P.Wx[i]);
      s.sc[SC_RHO] = rho;
      s.sc[SC_RHOK] = rho * P.rho_k_rel * wk;
      s.sc[SC_RHOE] = P.rho_e_rel * wm;
      s.sc[SC_RHOEP] = P.rho_e_pol_rel * wm;
    }
  }
  CMPC_SYNC();
}

// ---------------------------------------------------------------- ADMM driver
// returns 1 when the OSQP termination test passed ("solved"), 0 on max_iter / numeric failure
CMPC_HD int admm_solve(Ctx& c, int* iters_out, int* nfact_out) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  int nfact = 0, solved = 0, it = 0;
  if (factor(c, MODE_ADMM)) { *iters_out = 0; *nfact_out = 1; return 0; }
  ++nfact;
  for (it = 1; it <= P.max_iter; ++it) {
    backward_sweep(c, MODE_ADMM);
    forward_sweep(c, MODE_ADMM);
    if (it % P.check_every == 0) {
      residuals(c, MODE_ADMM);
      const double pri = s.sc[SC_PRI], dua = s.sc[SC_DUA], npri = s.sc[SC_NPRI], ndua = s.sc[SC_NDUA];
      if (pri <= P.eps_abs + P.eps_rel * npri && dua <= P.eps_abs + P.eps_rel * ndua) { solved = 1; break; }
      if (!(pri == pri) || !(dua == dua)) break;   // NaN
      if (P.adaptive_rho && it >= P.adapt_start) {
        const double rho = s.sc[SC_RHO];
        double est = rho * sqrt((pri / (npri + 1e-10)) / (dua / (ndua + 1e-10) + 1e-10));
        est = fmin(fmax(est, 1e-6), 1e6);
        if (est > rho * P.adapt_tol || est < rho / P.adapt_tol) {
          const double rk_old = s.sc[SC_RHOK];
          rescale_duals(c, rho, est, rk_old);
          set_rho(c, est);
          if (factor(c, MODE_ADMM)) break;
          ++nfact;
        }
      }
    }
  }
  *iters_out = it > P.max_iter ? P.max_iter : it;
  *nfact_out = nfact;
  return solved;
}

// ---------------------------------------------------------------- polish
// Active-set guess from the ADMM iterate, then the equality-constrained QP by a proximal
// method of multipliers with penalty 1/delta (OSQP: regularised KKT + iterative refinement).
// Accepts the polished point only if it improves the residuals (OSQP's rule).
CMPC_HD int polish(Ctx& c) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const int N = P.N;
  const double pri0 = s.sc[SC_PRI], dua0 = s.sc[SC_DUA];
  const double m0 = fmax(pri0 / (P.eps_abs + P.eps_rel * s.sc[SC_NPRI]), dua0 / (P.eps_abs + P.eps_rel * s.sc[SC_NDUA]));
  // keep the ADMM iterate; build the polish records
  CMPC_LANES(l) {
    for (long e = l; e < (long)(N + 1) * STA; e += 32) c.sta2[e] = c.sta[e];
    if (l < 9) s.Pc[l] = s.ye[l];
  }
  CMPC_SYNC();
  CMPC_LANES(l) {
    for (int k = l; k <= N; k += 32) {
      const double* sk = c.sta + (long)k * STA;
      const double* gk = c.stg + (long)k * STG;
      double* pk = c.pol + (long)k * POL;
      int pm = 0;
      for (int r = 0; r < POL; ++r) pk[r] = 0.0;
      if (k < N) {
        int mask = (int)gk[O_ACT];
        for (int r = 0; r < 16; ++r) {
          int ct = r >> 2;
          double v = sk[O_VF + r];
          if (ct < P.nc && ((mask >> ct) & 1) && v > 0.0) {   // -w < y  <=>  v > 0
            pm |= 1 << r;
            pk[r] = s.sc[SC_RHO] * s.ef2[r] * v;
          }
        }
      }
      if (k >= 1) {
        double rk = s.sc[SC_RHOK], w[3];
        int br = prox_trust(&sk[O_VK], &gk[O_KB], s.sc[SC_RADIUS], s.sc[SC_WEIGHT], rk, w);
        pm |= br << 16;
        if (br != 0) {
          double msum = 0.0;
          int nz = 0;
          for (int i = 0; i < 3; ++i) {
            double d = w[i] - gk[O_KB + i];
            double yk = rk * (sk[O_VK + i] - w[i]);
            int code = d > 0.0 ? 1 : (d < 0.0 ? 2 : 0);
            pm |= code << (18 + 2 * i);
            if (code == 0) pk[16 + i] = yk;
            else { msum += (code == 1 ? yk : -yk); ++nz; }
          }
          if (br == 2) pk[19] = nz ? msum / nz : 0.0;
        }
      }
      c.pmask[k] = pm;
    }
  }
  CMPC_SYNC();
  // Rounds of (factor, PMM sweeps); after each round the friction active set is corrected:
  // rows that came out violated are added, rows whose multiplier came out negative are dropped.
  // OSQP polishes once; the correction only matters when the ADMM iterate at eps = 1e-7 did not
  // identify the active set exactly (DESIGN.md "polish").
  int bad = 0;
  for (int round = 0; round < 1 + P.polish_rounds && !bad; ++round) {
    bad = factor(c, MODE_POLISH);
    if (bad) break;
    for (int r = 0; r < 1 + P.refine; ++r) {
      backward_sweep(c, MODE_POLISH);
      forward_sweep(c, MODE_POLISH);
    }
    if (round == P.polish_rounds) break;
    CMPC_LANES(l) {
      double changes = 0.0;
      for (int k = l; k < N; k += 32) {
        const double* sk = c.sta + (long)k * STA;
        double* pk = c.pol + (long)k * POL;
        const int mask = (int)c.stg[(long)k * STG + O_ACT];
        int pm = c.pmask[k];
        for (int r = 0; r < 16; ++r) {
          int ct = r >> 2;
          if (ct >= P.nc || !((mask >> ct) & 1)) continue;
          if ((pm >> r) & 1) {
            if (pk[r] < 0.0) { pm &= ~(1 << r); pk[r] = 0.0; changes += 1.0; }
          } else if (friction_row_value(c, k, ct, r & 3, &sk[O_U]) > 0.0) {
            pm |= 1 << r; pk[r] = 0.0; changes += 1.0;
          }
        }
        c.pmask[k] = pm;
      }
      s.PB[l] = changes;
    }
    CMPC_REDUCE_SUM(c, 0, SC_NUM);
    if (s.sc[SC_NUM] == 0.0) break;
  }
  if (!bad) residuals(c, MODE_POLISH);
  // Acceptance: OSQP keeps the polished point when both residuals improve.  Here the polished
  // dual residual carries the round-off of the 1e9 terminal penalty (~1e-6, three orders below
  // eps_dual), so "improve" is judged on the residuals normalised by their tolerances.
  const double pri = s.sc[SC_PRI], dua = s.sc[SC_DUA];
  const double m1 = fmax(pri / (P.eps_abs + P.eps_rel * s.sc[SC_NPRI]), dua / (P.eps_abs + P.eps_rel * s.sc[SC_NDUA]));
  const int ok = !bad && (m1 < m0) && (pri == pri) && (dua == dua);
  if (!ok) {
    CMPC_LANES(l) {
      for (long e = l; e < (long)(N + 1) * STA; e += 32) c.sta[e] = c.sta2[e];
      if (l < 9) s.ye[l] = s.Pc[l];
      if (l == 0) { s.sc[SC_PRI] = pri0; s.sc[SC_DUA] = dua0; }
    }
    CMPC_SYNC();
  } else {
    // keep the ADMM (v) part of the iterate so that a warm start of the next SCP iteration is
    // consistent; x and u are the polished ones
    CMPC_LANES(l) {
      for (int k = l; k <= N; k += 32) {
        double* sk = c.sta + (long)k * STA;
        const double* s2 = c.sta2 + (long)k * STA;
        for (int r = O_VF; r < STA; ++r) sk[r] = s2[r];
      }
      if (l < 9) s.ye[l] = s.Pc[l];
    }
    CMPC_SYNC();
  }
  return ok;
}

// ---------------------------------------------------------------- trust test and accuracy ratio
// sigma_max(X - Xbar) via the 9x9 Gram matrix + cyclic Jacobi (scp_solver.py:151: np.linalg.norm(.,2));
// rho = sum ||(f(x,u) - lin)[6:9]||^2 / sum ||lin||^2 (scp_solver.py:71-87).
CMPC_HD void evaluate(Ctx& c) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const int N = P.N;
  const double* Xr = c.bt.X_ref + (long)c.b * (N + 1) * 9;
  // Gram matrix G = D D^T, D = X - Xbar (upper triangle, 45 entries)
  CMPC_LANES(l) {
    for (int e = l; e < 45; e += 32) {
      int i = 0, rem = e;
      while (rem >= 9 - i) { rem -= 9 - i; ++i; }
      int r = i + rem;
      double acc = 0.0;
      for (int k = 0; k <= N; ++k) {
        double di = c.sta[(long)k * STA + O_X + i] - Xr[k * 9 + i];
        double dr = c.sta[(long)k * STA + O_X + r] - Xr[k * 9 + r];
        acc += di * dr;
      }
      s.T[i * 9 + r] = acc;
      s.T[r * 9 + i] = acc;
    }
    s.PB[32 * 0 + l] = 0.0;
    s.PB[32 * 1 + l] = 0.0;
  }
  CMPC_SYNC();
  // accuracy ratio, lanes over knots
  CMPC_LANES(l) {
    double num = 0.0, den = 0.0;
    for (int k = l; k < N; k += 32) {
      const double* sk = c.sta + (long)k * STA;
      const double* gk = c.stg + (long)k * STG;
      int mask = (int)gk[O_ACT];
      double lin[9], nl[9];
      for (int i = 0; i < 9; ++i) {
        double v = Av_elem(P, &gk[O_S], &sk[O_X], i) + gk[O_C + i];
        for (int ct = 0; ct < P.nc; ++ct)
          if ((mask >> ct) & 1)
            for (int a = 0; a < 3; ++a) v += Bcol_elem(P, &gk[O_D + 3 * ct], a, i) * sk[O_U + 3 * ct + a];
        lin[i] = v;
      }
      step_knot(P, &sk[O_X], &sk[O_U], c.cpos + (long)k * P.nc * 3, c.cact + (long)k * P.nc, nl);
      for (int i = 6; i < 9; ++i) num += (nl[i] - lin[i]) * (nl[i] - lin[i]);
      for (int i = 0; i < 9; ++i) den += lin[i] * lin[i];
    }
    s.PB[32 * 0 + l] = num;
    s.PB[32 * 1 + l] = den;
  }
  CMPC_REDUCE_SUM(c, 0, SC_NUM);
  CMPC_REDUCE_SUM(c, 1, SC_DEN);
  // largest eigenvalue of the Gram matrix: cyclic Jacobi on lane 0
  CMPC_LANES(l) {
    if (l == 0) {
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
      s.sc[SC_SNORM] = sqrt(mx);
    }
  }
  CMPC_SYNC();
}

// ---------------------------------------------------------------- per-instance setup
CMPC_HD void setup_instance(Ctx& c) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const int N = P.N;
  const double* Xr = c.bt.X_ref + (long)c.b * (N + 1) * 9;
  const double* Ui = c.bt.U_init + (long)c.b * N * P.nu;
  const double* xi = c.bt.x_init + (long)c.b * 9;
  const double* xf = c.bt.x_final + (long)c.b * 9;
  fill_G_phase(c, 0);
  CMPC_SYNC();
  // K1: linearise every knot (lanes over knots) and start the iterate at the linearisation point
  CMPC_LANES(l) {
    for (int k = l; k <= N; k += 32) {
      double* gk = c.stg + (long)k * STG;
      double* sk = c.sta + (long)k * STA;
      const int kk = k < N ? k : N - 1;
      linearize_knot(P, Xr + k * 9, Ui + kk * P.nu, c.cpos + (long)kk * P.nc * 3, c.cact + (long)kk * P.nc, k, gk);
      for (int r = 0; r < STA; ++r) sk[r] = 0.0;
      for (int i = 0; i < 9; ++i) sk[O_X + i] = (k == 0) ? xi[i] : Xr[k * 9 + i];
      for (int i = 0; i < 3; ++i) sk[O_VK + i] = Xr[k * 9 + 6 + i];
      if (k < N) {
        int mask = (int)gk[O_ACT];
        for (int ct = 0; ct < P.nc; ++ct) {
          if (!((mask >> ct) & 1)) continue;
          for (int a = 0; a < 3; ++a) sk[O_U + 3 * ct + a] = Ui[k * P.nu + 3 * ct + a];
          for (int row = 0; row < 4; ++row) {
            double cf = 0.0;
            if (P.identity_R) {
              for (int a = 0; a < 3; ++a) cf += s.G[ct][row * 3 + a] * Ui[k * P.nu + 3 * ct + a];
            } else {
              const double kf = P.mu * 0.70710678118654752440;
              const double* R = c.cR + ((long)k * P.nc + ct) * 9;
              double pr[3] = {row == 0 ? 1.0 : (row == 1 ? -1.0 : 0.0), row == 2 ? 1.0 : (row == 3 ? -1.0 : 0.0), -kf};
              for (int a = 0; a < 3; ++a)
                cf += (pr[0] * R[a * 3] + pr[1] * R[a * 3 + 1] + pr[2] * R[a * 3 + 2]) * Ui[k * P.nu + 3 * ct + a];
            }
            sk[O_VF + 4 * ct + row] = fmin(cf, 0.0);
          }
        }
      }
    }
    if (l < 9) s.ye[l] = 0.0;
    (void)xf;
  }
  CMPC_SYNC();
}

CMPC_HD void write_solution(Ctx& c) {
  const Params& P = *c.prm;
  const int N = P.N;
  CMPC_LANES(l) {
    for (int e = l; e < (N + 1) * 9; e += 32)
      c.bt.X_out[(long)c.b * (N + 1) * 9 + e] = c.sta[(long)(e / 9) * STA + O_X + e % 9];
    for (int e = l; e < N * P.nu; e += 32)
      c.bt.U_out[(long)c.b * N * P.nu + e] = c.sta[(long)(e / P.nu) * STA + O_U + e % P.nu];
  }
  CMPC_SYNC();
}

// ---------------------------------------------------------------- the SCP loop of one instance
// scp_solver.py:118-179.  The linearisation point never moves (:129-130), so the stage records
// are built once; each SCP iteration re-solves the QP for the current (radius, weight).
CMPC_HD void solve_instance(Ctx& c) {
  const Params& P = *c.prm;
  WarpMem& s = *c.s;
  const int N = P.N;
  setup_instance(c);
  double radius = P.radius0, weight = P.omega0;
  int it = 0, success = 0, n_acc = 0, status = ST_OK, qp_total = 0, nf_total = 0, polished = 0;
  double snorm = 0.0, acc = 0.0;
  while (it < P.max_scp && weight < P.omega_max && !(it != 0 && success && 0.0 < P.conv_thresh)) {
    success = 0;
    CMPC_LANES(l) {
      if (l == 0) { s.sc[SC_RADIUS] = radius; s.sc[SC_WEIGHT] = weight; }
    }
    CMPC_SYNC();
    if (it == 0) set_rho(c, P.rho0);
    int qi = 0, nf = 0;
    int solved = admm_solve(c, &qi, &nf);
    qp_total += qi;
    nf_total += nf;
    if (!solved) { status = (s.sc[15] != 0.0) ? ST_QP_NUMERIC : ST_QP_MAXITER; break; }
    polished = 0;
    if (P.polish) { polished = polish(c); ++nf_total; }
    evaluate(c);
    snorm = s.sc[SC_SNORM];
    if (snorm < radius) {
      acc = s.sc[SC_NUM] / s.sc[SC_DEN];
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
  if (n_acc == 0) write_solution(c);
  CMPC_LANES(l) {
    if (l == 0) {
      c.bt.scp_iters[c.b] = it;
      c.bt.status[c.b] = status;
      c.bt.n_accepted[c.b] = n_acc;
      c.bt.qp_iters[c.b] = qp_total;
      c.bt.n_factor[c.b] = nf_total;
      double* inf = c.bt.info + (long)c.b * 8;
      inf[0] = snorm; inf[1] = acc; inf[2] = s.sc[SC_PRI]; inf[3] = s.sc[SC_DUA];
      inf[4] = s.sc[SC_RHO]; inf[5] = radius; inf[6] = weight; inf[7] = (double)polished;
    }
  }
  CMPC_SYNC();
}

// bind the per-instance pointers
CMPC_HD void bind_instance(Ctx& c, const Params* prm, const Batch& bt, WarpMem* s, int b) {
  c.prm = prm; c.bt = bt; c.s = s; c.b = b;
  const int N = prm->N;
  const long plan = (long)b * bt.plan_stride;
  c.cpos = bt.cpos + plan * N * prm->nc * 3;
  c.cR = bt.cR ? bt.cR + plan * N * prm->nc * 9 : nullptr;
  c.cact = bt.cact + plan * N * prm->nc;
  c.stg = bt.stg + (long)b * (N + 1) * STG;
  c.sta = bt.sta + (long)b * (N + 1) * STA;
  c.sta2 = bt.sta2 + (long)b * (N + 1) * STA;
  c.fac = bt.fac + (long)b * N * FAC;
  c.dvec = bt.dvec + (long)b * N * DVC;
  c.pol = bt.pol + (long)b * (N + 1) * POL;
  c.pmask = bt.pmask + (long)b * (N + 1);
}

}  // namespace cmpc
