// cmpc_emu.cpp — TEST-ONLY host build of the device solver source (csrc/cmpc_solver.cuh).
//
// The solver is written as warp-uniform driver code with lane-parallel phases; compiled with
// g++ the phase macro becomes a loop over 32 lanes, so this library executes the same
// arithmetic in the same order as the CUDA kernel.  It exists so that the kernel logic can be
// unit-tested on a machine without a GPU (pytest -m "not gpu").  It is NOT part of
// libcmpc_b200.so, exports different symbol names (cmpc_emu_*), and nothing in the product
// package loads it: the product path fails loudly without CUDA.
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
  WsSizes w = ws_sizes(B, N);
  std::vector<double> stg(w.stg), sta(w.sta), sta2(w.sta2), fac(w.fac), dvec(w.dvec), pol(w.pol);
  std::vector<int> pmask(w.pmask);
  Batch bt;
  bt.B = B; bt.x_init = x_init; bt.x_final = x_final; bt.X_ref = X_ref; bt.U_init = U_init;
  bt.cpos = contact_pos; bt.cR = contact_R; bt.cact = contact_active;
  bt.plan_stride = dims->shared_plan ? 0 : 1;
  bt.stg = stg.data(); bt.sta = sta.data(); bt.sta2 = sta2.data(); bt.fac = fac.data();
  bt.dvec = dvec.data(); bt.pol = pol.data(); bt.pmask = pmask.data();
  bt.X_out = X_out; bt.U_out = U_out; bt.scp_iters = scp_iters; bt.status = status;
  bt.n_accepted = n_accepted; bt.qp_iters = qp_iters; bt.n_factor = n_factor; bt.info = info;
  WarpMem* s = new WarpMem;
  for (int b = 0; b < B; ++b) {
    memset(s, 0, sizeof(WarpMem));
    Ctx c;
    bind_instance(c, &prm, bt, s, b);
    solve_instance(c);
  }
  delete s;
  return 0;
}

extern "C" int cmpc_emu_warpmem_bytes(void) { return (int)sizeof(WarpMem); }
