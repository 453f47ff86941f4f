// cmpc_params.h — host-side translation of the C-ABI structs (include/cmpc.h) into the
// solver's Params.  Shared by the CUDA library and by the test-only host build (tests/emu).
#pragma once
#include "../../include/cmpc.h"
#include "cmpc_core.cuh"

namespace cmpc {

inline void default_qp_settings(cmpc_qp_settings* s) {
  s->eps_abs = 1e-7;            // scp_solver.py:63
  s->eps_rel = 1e-7;
  s->sigma = 1e-6;              // OSQP default (unused, see cmpc.h)
  s->alpha = 1.8;               // over-relaxation (OSQP's default is 1.6).  Swept on B200, 1.6 / 1.7 / 1.8 / 1.9: trot 4096 9.7 / 9.2 / 8.9-9.0 /
                                // 9.1 ms, bound 9.5 / 9.4 / 9.0 / 9.1, bolt 8192 12.1 / 12.4 / 11.9-12.2 / 11.5, talos 11.3 / 11.1 / 11.0 / 10.9;
                                // fewer polish rounds and a shorter straggler tail (at most 7 instead of 9 factorisations); the
                                // exception is pace mode A (no active rows at the solution): 3.2 ms up to 1.7, 4.1 ms from 1.75 on
  s->rho = 2.0;                 // initial penalty in equilibrated units
  s->delta = 1e-6;              // OSQP polish regularisation
  s->adaptive_rho_tolerance = 5.0;
  s->max_iter = 4000;           // OSQP default
  s->check_termination = 25;    // OSQP default; residuals are knot-local by-products of the forward sweep
  s->polish = 1;                // scp_solver.py:63
  s->polish_refine_iter = 3;    // OSQP default
  s->adaptive_rho = 1;
  s->adaptive_rho_start = 200;  // early residuals are transient: adapting on them hurts
  s->polish_active_set_rounds = 9;
  s->active_set_start = 8;      // first certified-polish attempt after 8 ADMM iterations (swept on B200: DESIGN.md section 6)
  s->active_set_step = 8;      // doubled after every failed attempt
  s->active_set_tol = 1e-9;
  s->warm_start_tol = 1e-7;
  s->warm_start = 0;
}

inline int fill_params(Params* p, const cmpc_dims* d, const cmpc_model* m, const cmpc_scp_params* scp,
                       const cmpc_qp_settings* qp, int identity_R) {
  if (d->N < 1 || d->nc < 1 || d->nc > MAXC || d->batch < 0) return -1;
  if (WR && 2 * d->nc > MAXC) return -1;
  cmpc_qp_settings dq;
  if (!qp) { default_qp_settings(&dq); qp = &dq; }
  memset(p, 0, sizeof(Params));
  p->N = d->N; p->nc = WR ? 2 * d->nc : d->nc; p->nu = 3 * p->nc; p->identity_R = identity_R;
  p->nf = d->nc;
  p->qs = WR ? 0.0 : 1.0;
  for (int i = 0; i < 4; ++i) p->foot_range[i] = WR ? m->foot_range[i] : 0.0;
  p->m = m->robot_mass; p->g = m->gravity_constant; p->dt = m->dt; p->mu = m->mu;
  p->kf = p->mu * 0.70710678118654752440;
  p->dt_m = p->dt / p->m;
  p->dtmg = p->dt * p->m * p->g;
  for (int i = 0; i < NX; ++i) p->Wx[i] = m->state_cost_weights[i];
  for (int i = 0; i < MAXU; ++i) p->Wu[i] = 1.0;
  for (int c = 0; c < p->nc; ++c)   // the solver's control order: three per (pseudo-)contact
    for (int a = 0; a < 3; ++a) p->Wu[3 * c + a] = m->control_cost_weights[uix(c, a)];
  for (int i = 0; i < NX; ++i) if (!(p->Wx[i] > 0.0)) return -2;
  for (int i = 0; i < p->nu; ++i) if (!(p->Wu[i] > 0.0)) return -2;
  // fast path: identity contact frames and the same control weights for every contact, so the
  // friction rows and their equilibration factors are the same for all contacts and knots
  int uniform = 1;
  for (int c = 1; c < p->nc; ++c)
    for (int a = 0; a < 3; ++a) if (p->Wu[3 * c + a] != p->Wu[a]) uniform = 0;
  p->fast = identity_R && uniform && !WR;
  if (WR) for (int i = 0; i < 4; ++i) if (!(p->foot_range[i] >= 0.0)) return -1;
  for (int row = 0; row < 4; ++row) {
    const double gx = (row < 2) ? 1.0 / sqrt(p->Wu[0]) : 0.0, gy = (row >= 2) ? 1.0 / sqrt(p->Wu[1]) : 0.0;
    const double gz = p->kf / sqrt(p->Wu[2]);
    const double mx = fmax(fmax(gx, gy), gz);
    p->e2[row] = 1.0 / (mx * mx);
  }
  p->alpha = qp->alpha; p->rho0 = qp->rho; p->eps_abs = qp->eps_abs;
  p->eps_rel = qp->eps_rel; p->delta = qp->delta; p->inv_delta = 1.0 / qp->delta; p->adapt_tol = qp->adaptive_rho_tolerance;
  p->rho_e_rel = 100.0; p->rho_k_rel = 1.0;
  // terminal-equality penalty while polishing (x max W_x).  Wrench model: active CoP / friction rows at the last
  // knots and x_N = x_final nearly over-determine the last controls; the multiplier iteration then contracts by
  // 1 / (1 + penalty x joint compliance) only when BOTH penalties are large (measured: 0.93 per sweep at
  // 1e7 / delta 1e-6, 0.08 at 1e9 / delta 1e-9)
  p->rho_e_pol_rel = WR ? 1e6 : 1e4;
  p->max_iter = qp->max_iter; p->check_every = qp->check_termination > 0 ? qp->check_termination : 5;
  p->polish = qp->polish; p->refine = qp->polish_refine_iter; p->adaptive_rho = qp->adaptive_rho;
  p->adapt_start = qp->adaptive_rho_start;
  p->as_rounds = qp->polish_active_set_rounds;
  p->as_start = qp->active_set_start; p->as_step = qp->active_set_step > 0 ? qp->active_set_step : 10;
  p->as_tol = qp->active_set_tol > 0.0 ? qp->active_set_tol : 1e-9;
  p->as_tol_loose = WR ? 100.0 * p->as_tol : p->as_tol;
  p->warm = qp->warm_start != 0 && qp->polish != 0;
  p->warm_tol = qp->warm_start_tol > 0.0 ? qp->warm_start_tol : 1e-7;
  if (scp) {
    p->radius0 = scp->trust_region_radius0; p->omega0 = scp->omega0; p->omega_max = scp->omega_max;
    p->acc_rho0 = scp->rho0; p->acc_rho1 = scp->rho1; p->beta_succ = scp->beta_succ;
    p->beta_fail = scp->beta_fail; p->gamma_fail = scp->gamma_fail;
    p->conv_thresh = scp->convergence_threshold; p->max_scp = scp->max_iterations;
  }
  return 0;
}

// workspace sizes for a batch: doubles of the knot records, ints of the slot table.  The records are
// allocated for the general layout of the problem's contact count (friction table included), so that a
// handle can switch between the nominal and the stochastic / rotated-contact path without reallocating.
struct WsSizes { long tiles, ws, nst, info; int rfields; };
inline WsSizes ws_sizes(int B, int N, int nc_in) {
  WsSizes w;
  const int nc = WR ? 2 * nc_in : nc_in;
  w.tiles = (B + TL - 1) / TL;
  w.rfields = rec_fields(nc, true);
  w.ws = w.tiles * (N + 1) * (long)(w.rfields * TL);
  w.nst = w.tiles * (N + 1);
  w.info = (long)B * INFO;
  return w;
}

}  // namespace cmpc
