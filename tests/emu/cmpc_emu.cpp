// cmpc_emu.cpp — TEST-ONLY host build of the device solver source (csrc/cmpc_solver.cuh).
//
// The solver is written against the 32-lane vector abstraction of csrc/cmpc_simt.cuh; compiled
// with g++ every varying value is a 32-element array and every operation a loop over the lanes,
// so this library executes the same arithmetic in the same order as the CUDA kernel.  It exists
// so that the kernel logic can be unit-tested on a machine without a GPU (pytest -m "not gpu").
// It is NOT part of libcmpc_b200.so, exports different symbol names (cmpc_emu_*), and nothing
// in the product package loads it: the product path fails loudly without CUDA.
#include <stdlib.h>
#include <vector>

#include "../../centroidal_mpc_b200/csrc/cmpc_params.h"
#include "../../centroidal_mpc_b200/csrc/cmpc_solver.cuh"

using namespace cmpc;

extern "C" int cmpc_emu_solve_scp(const cmpc_dims* dims, const cmpc_model* model, const cmpc_scp_params* scp,
                                  const cmpc_qp_settings* qp, const double* x_init, const double* x_final,
                                  const double* X_ref, const double* U_init, const double* contact_pos,
                                  const double* contact_R, const int32_t* contact_active, double* X_out,
                                  double* U_out, int32_t* scp_iters, int32_t* status, int32_t* n_accepted,
                                  int32_t* qp_iters, int32_t* n_factor, double* info) {
  Params prm;
  int rc = fill_params(&prm, dims, model, scp, qp, contact_R == nullptr);
  if (rc) return rc;
  const int B = dims->batch, N = dims->N;
  WsSizes w = ws_sizes(1, N);   // instances run one after the other: one instance of workspace
  std::vector<double> stg(w.stg), sta(w.sta), fac(w.fac), dvec(w.dvec), pm(w.pm), sol(w.sol), gtab(w.gtab);
  std::vector<int> meta(w.meta), pmask(w.pmask);
  WarpMem* s = new WarpMem;
  for (int b = 0; b < B; ++b) {
    Batch bt;
    memset(&bt, 0, sizeof(bt));
    bt.B = B; bt.x_init = x_init; bt.x_final = x_final; bt.X_ref = X_ref; bt.U_init = U_init;
    bt.cpos = contact_pos; bt.cR = contact_R; bt.cact = contact_active;
    bt.plan_stride = dims->shared_plan ? 0 : 1;
    // workspace views shifted so that instance b lands on the single-instance buffers
    bt.stg = stg.data() - (long)b * (N + 1) * SG; bt.sta = sta.data() - (long)b * (N + 1) * ST;
    bt.fac = fac.data() - (long)b * N * FAC; bt.dvec = dvec.data() - (long)b * N * DVC;
    bt.pm = pm.data() - (long)b * (N + 1) * PM; bt.sol = sol.data() - (long)b * (N + 1) * SOL;
    bt.gtab = prm.fast ? nullptr : gtab.data() - (long)b * N * MAXC * 16;
    bt.meta = meta.data() - (long)b * (N + 1); bt.pmask = pmask.data() - (long)b * (N + 1);
    bt.X_out = X_out; bt.U_out = U_out; bt.scp_iters = scp_iters; bt.status = status;
    bt.n_accepted = n_accepted; bt.qp_iters = qp_iters; bt.n_factor = n_factor; bt.info = info;
    memset(s, 0, sizeof(WarpMem));
    Ctx c;
    bind_instance(c, &prm, bt, s, b);
    solve_instance(c);
  }
  delete s;
  return 0;
}

extern "C" int cmpc_emu_warpmem_bytes(void) { return (int)sizeof(WarpMem); }
