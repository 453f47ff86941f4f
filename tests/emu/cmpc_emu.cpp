// cmpc_emu.cpp — TEST-ONLY host build of the device solver source (csrc/cmpc_tile.cuh).
//
// The solver is plain scalar code per MPC instance (one CUDA thread per instance, 32 instances
// per tile with lane-interleaved records).  Compiled with g++ the same functions run one lane
// after the other over the same interleaved layout; a lane's arithmetic does not depend on its
// neighbours and FMA contraction is explicit in both builds, so this library executes the same
// arithmetic in the same order as the CUDA kernel.  It exists so that the kernel logic can be
// unit-tested on a machine without a GPU (pytest -m "not gpu").
// It is NOT part of libcmpc_b200.so, exports different symbol names (cmpc_emu_*), and nothing
// in the product package loads it: the product path fails loudly without CUDA.
#include <stdlib.h>
#include <stdio.h>
#include <vector>
#if defined(CMPC_NL) && CMPC_NL > 1
#include <ucontext.h>
#endif

#include "../../centroidal_mpc_b200/csrc/cmpc_params.h"
#include "../../centroidal_mpc_b200/csrc/cmpc_tile.cuh"
#include "../../centroidal_mpc_b200/csrc/cmpc_lqr.cuh"

using namespace cmpc;

// one lane (t = instance lane of the tile, q = lane of the instance's team): what a CUDA thread of
// run_tile (cmpc_api.cu) does for its instance
template <bool FAST>
static void run_lane(const Params& prm, const Batch& bt, TileCtx T, int tile, int t, int q) {
  Inst I;
  Sv S;
  Drv D;
  bind_instance(I, prm, bt, tile * TL + t);
  I.sub = q;
  double mq = 0.0, mc = 0.0;
  int nconv = 0;
  setup_knots(prm, T, I, true, &mq, &mc, &nconv);
  setup_finish(I, S, mq, mc, nconv);
  team_sync(I);
  drv_init(prm, S, D);
  for (int op = advance(prm, S, D); op != OP_DONE; op = advance(prm, S, D))
    execute<FAST>(op, prm, T, I, bt, S, D, true, D.check != 0);
  write_stats(bt, I, S, D);
}

#if CMPC_NL > 1
// lock-step build: the NL lanes of a team are coroutines that switch at every team_sync
namespace {
ucontext_t g_main, g_ctx[NL];
int g_cur = 0;
bool g_done[NL];
struct LaneArgs { const Params* prm; const Batch* bt; const TileCtx* T; int tile, t, fast; } g_args;
void lane_entry(int q) {
  if (g_args.fast) run_lane<true>(*g_args.prm, *g_args.bt, *g_args.T, g_args.tile, g_args.t, q);
  else run_lane<false>(*g_args.prm, *g_args.bt, *g_args.T, g_args.tile, g_args.t, q);
  g_done[q] = true;
  swapcontext(&g_ctx[q], &g_main);
}
}  // namespace
void cmpc::cmpc_emu_yield() { swapcontext(&g_ctx[g_cur], &g_main); }
static int run_team(const Params& prm, const Batch& bt, const TileCtx& T, int tile, int t, bool fast) {
  static std::vector<char> stacks((size_t)NL << 20);
  g_args = LaneArgs{&prm, &bt, &T, tile, t, fast ? 1 : 0};
  for (int q = 0; q < NL; ++q) {
    getcontext(&g_ctx[q]);
    g_ctx[q].uc_stack.ss_sp = stacks.data() + ((size_t)q << 20);
    g_ctx[q].uc_stack.ss_size = (size_t)1 << 20;
    g_ctx[q].uc_link = &g_main;
    g_done[q] = false;
    makecontext(&g_ctx[q], (void (*)())lane_entry, 1, q);
  }
  for (;;) {   // one round = every lane runs up to its next team_sync
    int ndone = 0;
    for (int q = 0; q < NL; ++q) {
      g_cur = q;
      swapcontext(&g_main, &g_ctx[q]);
      ndone += g_done[q] ? 1 : 0;
    }
    if (ndone == NL) return 0;
    if (ndone != 0) { fprintf(stderr, "cmpc_emu: the lanes of a team disagree on the number of team_sync calls\n"); return -9; }
  }
}
#else
static int run_team(const Params& prm, const Batch& bt, const TileCtx& T, int tile, int t, bool fast) {
  if (fast) run_lane<true>(prm, bt, T, tile, t, 0);
  else run_lane<false>(prm, bt, T, tile, t, 0);
  return 0;
}
#endif

static int run_tile_host(const Params& prm, const Batch& bt, int tile, std::vector<double>& scratch) {
  TileCtx T;
  bind_tile(T, prm, bt, tile);
  T.scratch = scratch.data();
  for (int k = 0; k <= prm.N; ++k) T.nst[k] = 0;
  for (int t = 0; t < TL; ++t) {   // slots per knot of the tile: the record layout of the knot
    const int b = tile * TL + t;
    if (b >= bt.B) continue;
    Inst I;
    bind_instance(I, prm, bt, b);
    for (int k = 0; k < prm.N; ++k) {
      const int ns = active_slots(prm, I, k);
      if (ns > T.nst[k]) T.nst[k] = ns;
    }
  }
  for (int t = 0; t < TL; ++t) {
    if (tile * TL + t >= bt.B) continue;
    int rc = run_team(prm, bt, T, tile, t, prm.fast != 0);
    if (rc) return rc;
  }
  return 0;
}

// stochastic mode: friction-row upper bounds [B][N][nc][4] used by the following solves (null = nominal)
static const double* g_fub = nullptr;
extern "C" void cmpc_emu_set_friction_ub(const double* fub) { g_fub = fub; }

extern "C" int cmpc_emu_solve_scp(const cmpc_dims* dims, const cmpc_model* model, const cmpc_scp_params* scp,
                                  const cmpc_qp_settings* qp, const double* x_init, const double* x_final,
                                  const double* X_ref, const double* U_init, const double* contact_pos,
                                  const double* contact_R, const int32_t* contact_active, double* X_out,
                                  double* U_out, int32_t* scp_iters, int32_t* status, int32_t* n_accepted,
                                  int32_t* qp_iters, int32_t* n_factor, double* info) {
  Params prm;
  int rc = fill_params(&prm, dims, model, scp, qp, contact_R == nullptr);
  if (rc) return rc;
  if (g_fub) prm.fast = 0;
  const int B = dims->batch, N = dims->N;
  WsSizes w1 = ws_sizes(TL, N, dims->nc);   // tiles run one after the other: one tile of workspace
  WsSizes wb = ws_sizes(B, N, dims->nc);
  const int rfields = rec_fields(prm.nc, !prm.fast);   // prm.nc: contacts, or the pseudo-contacts of the wrench model
  const long tile_ws = (long)(N + 1) * rfields * TL;
  std::vector<double> ws(w1.ws), scratch((size_t)X_FAC_END * TL);
  std::vector<int> nst(w1.nst);
  for (int tile = 0; tile < wb.tiles; ++tile) {
    Batch bt;
    memset(&bt, 0, sizeof(bt));
    bt.B = B; bt.x_init = x_init; bt.x_final = x_final; bt.X_ref = X_ref; bt.U_init = U_init;
    bt.cpos = contact_pos; bt.cR = contact_R; bt.cact = contact_active;
    bt.plan_stride = dims->shared_plan ? 0 : 1;
    bt.fub = g_fub;
    bt.rfields = rfields;
    // workspace views shifted so that this tile lands on the single-tile buffers
    bt.ws = ws.data() - (long)tile * tile_ws;
    bt.nst = nst.data() - (long)tile * w1.nst;
    bt.X_out = X_out; bt.U_out = U_out; bt.scp_iters = scp_iters; bt.status = status;
    bt.n_accepted = n_accepted; bt.qp_iters = qp_iters; bt.n_factor = n_factor; bt.info = info;
    std::fill(ws.begin(), ws.end(), 0.0);
    rc = run_tile_host(prm, bt, tile, scratch);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int cmpc_emu_team_lanes(void) { return NL; }
extern "C" int cmpc_emu_record_doubles(void) { return REC_MAX; }

// host run of csrc/cmpc_lqr.cuh with the loop structure of cmpc_lqr_gains_kernel / cmpc_covs_kernel
extern "C" int cmpc_emu_lqr_covs(const cmpc_dims* dims, const cmpc_model* model, const cmpc_lqr_weights* w,
                                 const double* X, const double* U, const double* contact_pos,
                                 const int32_t* contact_active, double* gains, double* covs, const double* contact_R) {
  static_assert(sizeof(cmpc_lqr_weights) == sizeof(LqrWeights), "cmpc_lqr_weights layout");
  Params prm;
  int rc = fill_params(&prm, dims, model, nullptr, nullptr, 1);
  if (rc) return rc;
  const LqrWeights& W = *(const LqrWeights*)w;
  const int B = dims->batch, N = prm.N, nu = prm.nu, nc = prm.nf;   // rows of the contact arrays per knot
  for (int b = 0; b < B; ++b) {
    const long plan = dims->shared_plan ? 0 : b;
    double Sg[81] = {0}, Sn[81], A[81], Bm[9 * MAXU], Ct[3 * MAXU];
    if (covs) for (int i = 0; i < 81; ++i) covs[(long)b * (N + 1) * 81 + i] = 0.0;
    for (int k = 0; k < N; ++k) {
      double* K = gains + ((long)b * N + k) * nu * 9;
      knot_ABC(prm, X + ((long)b * (N + 1) + k) * 9, U + ((long)b * N + k) * nu, contact_pos + (plan * N + k) * nc * 3,
               (const int*)contact_active + (plan * N + k) * nc, A, Bm, Ct, contact_R ? contact_R + (plan * N + k) * nc * 9 : nullptr);
      lqr_gain_knot(A, Bm, nu, W, K);
      if (!covs) continue;
      cov_step_knot(A, Bm, Ct, K, nu, W, Sg, Sn);
      for (int i = 0; i < 81; ++i) { Sg[i] = Sn[i]; covs[((long)b * (N + 1) + k + 1) * 81 + i] = Sn[i]; }
    }
  }
  return 0;
}

extern "C" int cmpc_emu_friction_backoffs(const cmpc_dims* dims, const cmpc_model* model, double xi,
                                          const double* gains, const double* covs, const double* contact_R,
                                          const int32_t* contact_active, double* friction_ub) {
  Params prm;
  int rc = fill_params(&prm, dims, model, nullptr, nullptr, contact_R == nullptr);
  if (rc) return rc;
  const int B = dims->batch, N = prm.N, nu = prm.nu, nc = prm.nc;
  for (long t = 0; t < (long)B * N; ++t) {
    const int b = (int)(t / N), k = (int)(t % N);
    const long plan = dims->shared_plan ? 0 : b;
    friction_backoff_knot(prm, xi, k, gains + t * nu * 9, covs + ((long)b * (N + 1) + k) * 81,
                          contact_R ? contact_R + (plan * N + k) * nc * 9 : nullptr,
                          (const int*)contact_active + (plan * N + k) * nc, friction_ub + t * 4 * nc);
  }
  return 0;
}
