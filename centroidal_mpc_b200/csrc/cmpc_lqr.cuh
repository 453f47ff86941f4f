// cmpc_lqr.cuh — LQR feedback gains and covariance propagation along the nominal trajectory
// (SURVEY.md section 8 row f1).  Scalar per-knot functions, host/device like cmpc_core.cuh.
//
// What this replaces in the reference (paths relative to /root/reference):
//   src/centroidal_model.py:215-227   compute_lqr_feedback_gains: P = Q, two Riccati steps
//                                     P <- Q + A'PA - A'PB (R + B'PB)^-1 B'PA, K = -(R + B'PB)^-1 B'PA
//   src/centroidal_model.py:232       C = df/dp (contact-position Jacobian)
//   src/centroidal_model.py:234-238   Sigma_next = [A B] Sigma_xu [A B]' + C Cov_w C' + Cov_eta,
//                                     Sigma_xu = [[S, SK'],[KS, KSK']]  ==  (A+BK) S (A+BK)'
//   src/centroidal_model.py:284-285   LQR_gains[k] = K, Covs[k+1] = Sigma_next (Covs[0] = 0)
//   src/constraints.py:157-163,187-214 chance-constraint back-offs of the friction rows (stochastic mode)
#pragma once
#include "cmpc_core.cuh"

namespace cmpc {

// weights of the gain / covariance recursion: conf.Q, conf.R, conf.cov_w, conf.cov_white_noise
// (src/centroidal_model.py:34-35,41-42), dense row-major, R and cov_w with leading dimension nu
struct LqrWeights {
  double Q[NX * NX], R[MAXU * MAXU], cov_w[MAXU * MAXU], cov_eta[NX * NX];
};

// dense B (9 x nu, row-major, per-contact control layout) and the torque rows of C = df/dp
// (C[6:9, 3c:3c+3] = -dt [f_c]x for active contacts, zero elsewhere; stored as Ct[3][nu])
// Wrench model (WR): B in the reference's control order (six per foot) from the slot blocks of linearize_knot; the
// contact-position noise has three components per FOOT, so only the first 3 nf columns of Ct (and the leading
// 3 nf x 3 nf block of cov_w) are used.
CMPC_HD void dense_B_C(const Params& P, const double* x, const double* u, const double* cpos, const int* cact,
                       double* Bm, double* Ct, const double* cR = nullptr) {
  const int nu = P.nu;
  for (int i = 0; i < 9 * nu; ++i) Bm[i] = 0.0;
  for (int i = 0; i < 3 * nu; ++i) Ct[i] = 0.0;
  if (WR) {
    KnotLin L;
    linearize_knot(P, x, u, cpos, cact, 0, L, cR);
    const int ns = L.meta & 7;
    for (int sl = 0; sl < ns; ++sl) {
      const int pc = (L.meta >> (4 + 2 * sl)) & 3;
      for (int a = 0; a < 3; ++a) {
        const int col = uix(pc, a);
        if (!(pc & 1)) Bm[(3 + a) * nu + col] = P.dt;
        for (int j = 0; j < 3; ++j) Bm[(6 + j) * nu + col] = P.dt * L.d[9 * sl + 3 * a + j];
      }
    }
    for (int f = 0; f < P.nf; ++f) {
      if (!cact[f]) continue;
      const double* fb = u + 6 * f + 2;
      for (int a = 0; a < 3; ++a) {
        Ct[nxt3(a) * nu + 3 * f + a] = -P.dt * fb[prv3(a)];
        Ct[prv3(a) * nu + 3 * f + a] = P.dt * fb[nxt3(a)];
      }
    }
    return;
  }
  for (int c = 0; c < P.nc; ++c) {
    if (!cact[c]) continue;
    const double d[3] = {cpos[3 * c] - x[0], cpos[3 * c + 1] - x[1], cpos[3 * c + 2] - x[2]};
    for (int a = 0; a < 3; ++a) {
      double col[9];
      dense_Bcol(P, d, a, col);
      for (int i = 0; i < 9; ++i) Bm[i * nu + 3 * c + a] = col[i];
      // column a of -dt [f]x  =  -dt (f x e_a)  =  dt (e_a x f)
      Ct[nxt3(a) * nu + 3 * c + a] = -P.dt * u[3 * c + prv3(a)];
      Ct[prv3(a) * nu + 3 * c + a] = P.dt * u[3 * c + nxt3(a)];
    }
  }
}

// A (9x9), B (9 x nu), torque rows of C at one knot of the nominal trajectory
CMPC_HD void knot_ABC(const Params& P, const double* x, const double* u, const double* cpos, const int* cact,
                      double* A, double* Bm, double* Ct, const double* cR = nullptr) {
  KnotLin L;
  linearize_knot(P, x, u, cpos, cact, 0, L, cR);
  dense_A(P, L.S, A);
  dense_B_C(P, x, u, cpos, cact, Bm, Ct, cR);
}

// in-place Cholesky H = L L' (lower triangle, leading dimension ld); returns 0, or -1 when not SPD
CMPC_HD int chol_lower(double* H, int n, int ld) {
  for (int j = 0; j < n; ++j) {
    double s = H[j * ld + j];
    for (int l = 0; l < j; ++l) s = fma(-H[j * ld + l], H[j * ld + l], s);
    if (!(s > 0.0)) return -1;
    const double dj = sqrt(s);
    H[j * ld + j] = dj;
    for (int i = j + 1; i < n; ++i) {
      double t = H[i * ld + j];
      for (int l = 0; l < j; ++l) t = fma(-H[i * ld + l], H[j * ld + l], t);
      H[i * ld + j] = t / dj;
    }
  }
  return 0;
}
// solves L L' X = G for the 9 columns of G (n x 9, row-major), in place
CMPC_HD void chol_solve9(const double* L, int n, int ld, double* G) {
  for (int c = 0; c < 9; ++c) {
    for (int i = 0; i < n; ++i) {
      double t = G[i * 9 + c];
      for (int l = 0; l < i; ++l) t = fma(-L[i * ld + l], G[l * 9 + c], t);
      G[i * 9 + c] = t / L[i * ld + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double t = G[i * 9 + c];
      for (int l = i + 1; l < n; ++l) t = fma(-L[l * ld + i], G[l * 9 + c], t);
      G[i * 9 + c] = t / L[i * ld + i];
    }
  }
}

// G = (R + B'PB)^-1 B'PA  (nu x 9) and, when Pn != null, Pn = Q + A'PA - (B'PA)' G
CMPC_HD int riccati_step(const double* A, const double* Bm, int nu, const LqrWeights& W, const double* Pm,
                         double* G, double* Pn) {
  double PA[81], PB[9 * MAXU], H[MAXU * MAXU], BtPA[MAXU * 9];
  for (int i = 0; i < 9; ++i) {
    for (int j = 0; j < 9; ++j) {
      double s = 0.0;
      for (int l = 0; l < 9; ++l) s = fma(Pm[i * 9 + l], A[l * 9 + j], s);
      PA[i * 9 + j] = s;
    }
    for (int j = 0; j < nu; ++j) {
      double s = 0.0;
      for (int l = 0; l < 9; ++l) s = fma(Pm[i * 9 + l], Bm[l * nu + j], s);
      PB[i * nu + j] = s;
    }
  }
  for (int i = 0; i < nu; ++i) {
    for (int j = 0; j <= i; ++j) {
      double s = W.R[i * nu + j];
      for (int l = 0; l < 9; ++l) s = fma(Bm[l * nu + i], PB[l * nu + j], s);
      H[i * nu + j] = s;
    }
    for (int j = 0; j < 9; ++j) {
      double s = 0.0;
      for (int l = 0; l < 9; ++l) s = fma(Bm[l * nu + i], PA[l * 9 + j], s);
      BtPA[i * 9 + j] = G[i * 9 + j] = s;
    }
  }
  if (chol_lower(H, nu, nu)) return -1;
  chol_solve9(H, nu, nu, G);
  if (Pn) {
    for (int i = 0; i < 9; ++i)
      for (int j = 0; j < 9; ++j) {
        double s = W.Q[i * 9 + j];
        for (int l = 0; l < 9; ++l) s = fma(A[l * 9 + i], PA[l * 9 + j], s);
        for (int l = 0; l < nu; ++l) s = fma(-BtPA[l * 9 + i], G[l * 9 + j], s);
        Pn[i * 9 + j] = s;
      }
  }
  return 0;
}

// compute_lqr_feedback_gains (centroidal_model.py:215-227, niter = 2): K (nu x 9, row-major).
// Returns 0, or -1 when R + B'PB is not positive definite (K is then filled with NaN).
CMPC_HD int lqr_gain_knot(const double* A, const double* Bm, int nu, const LqrWeights& W, double* K) {
  double P0[81], P1[81];
  int rc = riccati_step(A, Bm, nu, W, W.Q, K, P0);
  if (!rc) rc = riccati_step(A, Bm, nu, W, P0, K, P1);
  if (!rc) rc = riccati_step(A, Bm, nu, W, P1, K, nullptr);
  for (int i = 0; i < nu * 9; ++i) K[i] = rc ? NAN : -K[i];
  return rc;
}

// Sigma_next = (A + B K) Sigma (A + B K)' + C cov_w C' + cov_eta   (centroidal_model.py:234-238)
CMPC_HD void cov_step_knot(const double* A, const double* Bm, const double* Ct, const double* K, int nu,
                           const LqrWeights& W, const double* Sg, double* Sn) {
  double M[81], T[81];
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) {
      double s = A[i * 9 + j];
      for (int l = 0; l < nu; ++l) s = fma(Bm[i * nu + l], K[l * 9 + j], s);
      M[i * 9 + j] = s;
    }
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) {
      double s = 0.0;
      for (int l = 0; l < 9; ++l) s = fma(M[i * 9 + l], Sg[l * 9 + j], s);
      T[i * 9 + j] = s;
    }
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) {
      double s = W.cov_eta[i * 9 + j];
      for (int l = 0; l < 9; ++l) s = fma(T[i * 9 + l], M[j * 9 + l], s);
      Sn[i * 9 + j] = s;
    }
  // C cov_w C' only touches the angular-momentum block
  double CW[3 * MAXU];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < nu; ++j) {
      double s = 0.0;
      for (int l = 0; l < nu; ++l) s = fma(Ct[i * nu + l], W.cov_w[l * nu + j], s);
      CW[i * nu + j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int l = 0; l < nu; ++l) s = fma(CW[i * nu + l], Ct[j * nu + l], s);
      Sn[(6 + i) * 9 + 6 + j] += s;
    }
}

// Friction-row upper bounds of one knot in stochastic mode (constraints.py:187-214 with the
// identically-zero covariance-gradient terms dropped):
//   ub[c][j] = - sum_u xi 2 G_ju sqrt((K_c Sigma_k K_c')_uu)   over G_ju > 1e-6, sqrt(.) > 1e-6,
// G = pyramid R_c' (utils.py:9-16), K_c = rows 3c..3c+2 of K_k; zero at k = 0 and for inactive contacts.
CMPC_HD void friction_backoff_knot(const Params& P, double xi, int k, const double* K, const double* Sg,
                                   const double* cR, const int* cact, double* ub) {
  for (int c = 0; c < P.nc; ++c) {
    for (int j = 0; j < 4; ++j) ub[4 * c + j] = 0.0;
    if (k == 0 || !cact[c]) continue;
    double sq[3];
    for (int u = 0; u < 3; ++u) {
      const double* Ku = K + (3 * c + u) * 9;
      double q = 0.0;
      for (int i = 0; i < 9; ++i) {
        double t = 0.0;
        for (int l = 0; l < 9; ++l) t = fma(Ku[l], Sg[l * 9 + i], t);
        q = fma(t, Ku[i], q);
      }
      sq[u] = sqrt(q);
    }
    for (int j = 0; j < 4; ++j) {
      // pyramid row j: (+-1, 0, -kf) for j < 2, (0, +-1, -kf) for j >= 2
      const double sgn = (j & 1) ? -1.0 : 1.0;
      const int ax = j >> 1;
      for (int u = 0; u < 3; ++u) {
        double G;
        if (cR) G = sgn * cR[9 * c + 3 * u + ax] - P.kf * cR[9 * c + 3 * u + 2];   // (pyr R')_ju = sum_a pyr_ja R_ua
        else G = (u == ax ? sgn : 0.0) - (u == 2 ? P.kf : 0.0);
        if (G > 1e-6 && sq[u] > 1e-6) ub[4 * c + j] -= xi * (2.0 * G * sq[u]);
        if (!(sq[u] == sq[u])) ub[4 * c + j] = sq[u];   // non-finite gains (R + B'PB not positive definite) must not pass as
                                                        // "no back-off": the NaN bound makes this instance's solve end with status != 0
      }
    }
  }
}

#if defined(__CUDACC__)
// LQR gains along the nominal trajectory (one thread per instance and knot): K [B][N][nu][9]
__global__ void cmpc_lqr_gains_kernel(const __grid_constant__ Params prm, const LqrWeights* __restrict__ W, int B,
                                      int shared_plan, const double* __restrict__ X, const double* __restrict__ U,
                                      const double* __restrict__ cpos, const double* __restrict__ cR,
                                      const int* __restrict__ cact, double* __restrict__ Kout) {
  const int N = prm.N, nu = prm.nu, nf = prm.nf;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)B * N) return;
  int b = (int)(t / N), k = (int)(t % N);
  const long plan = shared_plan ? 0 : b;
  double A[81], Bm[9 * MAXU], Ct[3 * MAXU], K[MAXU * 9];
  knot_ABC(prm, X + ((long)b * (N + 1) + k) * 9, U + ((long)b * N + k) * nu, cpos + (plan * N + k) * nf * 3,
           cact + (plan * N + k) * nf, A, Bm, Ct, cR ? cR + (plan * N + k) * nf * 9 : nullptr);
  lqr_gain_knot(A, Bm, nu, *W, K);
  for (int i = 0; i < nu * 9; ++i) Kout[t * nu * 9 + i] = K[i];
}

// covariance propagation (sequential in k; one thread per instance): covs [B][N+1][9][9], covs[b][0] = 0
__global__ void cmpc_covs_kernel(const __grid_constant__ Params prm, const LqrWeights* __restrict__ W, int B,
                                 int shared_plan, const double* __restrict__ X, const double* __restrict__ U,
                                 const double* __restrict__ cpos, const double* __restrict__ cR,
                                 const int* __restrict__ cact, const double* __restrict__ Kin, double* __restrict__ covs) {
  const int N = prm.N, nu = prm.nu, nf = prm.nf;
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const long plan = shared_plan ? 0 : b;
  double Sg[81], Sn[81], A[81], Bm[9 * MAXU], Ct[3 * MAXU];
  for (int i = 0; i < 81; ++i) { Sg[i] = 0.0; covs[(long)b * (N + 1) * 81 + i] = 0.0; }
  for (int k = 0; k < N; ++k) {
    knot_ABC(prm, X + ((long)b * (N + 1) + k) * 9, U + ((long)b * N + k) * nu, cpos + (plan * N + k) * nf * 3,
             cact + (plan * N + k) * nf, A, Bm, Ct, cR ? cR + (plan * N + k) * nf * 9 : nullptr);
    cov_step_knot(A, Bm, Ct, Kin + ((long)b * N + k) * nu * 9, nu, *W, Sg, Sn);
    for (int i = 0; i < 81; ++i) { Sg[i] = Sn[i]; covs[((long)b * (N + 1) + k + 1) * 81 + i] = Sn[i]; }
  }
}

// launches the two kernels; the weights travel through `scratch` (device memory, CMPC_LQR_SCRATCH_BYTES)
inline cudaError_t launch_lqr_covs(const Params& prm, const LqrWeights* w_host, int B, int shared_plan, const double* X,
                                   const double* U, const double* cpos, const double* cR, const int* cact, double* gains,
                                   double* covs, void* scratch, cudaStream_t st, long long* n_launches) {
  cudaError_t e = cudaMemcpyAsync(scratch, w_host, sizeof(LqrWeights), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  const LqrWeights* dW = (const LqrWeights*)scratch;
  const long total = (long)B * prm.N;
  cmpc_lqr_gains_kernel<<<(unsigned)((total + 63) / 64), 64, 0, st>>>(prm, dW, B, shared_plan, X, U, cpos, cR, cact, gains);
  ++*n_launches;
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (covs) {
    cmpc_covs_kernel<<<(unsigned)((B + 31) / 32), 32, 0, st>>>(prm, dW, B, shared_plan, X, U, cpos, cR, cact, gains, covs);
    ++*n_launches;
    e = cudaGetLastError();
  }
  return e;
}
#endif

}  // namespace cmpc
