// cmpc_launch.cuh — the SCP kernel and its launch, shared by the two compilations of the solver:
// cmpc_api.cu (point contacts, namespace cmpc) and cmpc_wrench.cu (CoP / wrench contact model, the same
// source compiled with CMPC_WRENCH=1 into namespace cmpc_wr).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <string>

#include "cmpc_params.h"
#include "cmpc_tile.cuh"

namespace cmpc {

// ------------------------------------------------------------------------------------------
// The SCP kernel: one warp per tile of TL = 32 / NL instances, NL lanes per instance (cmpc_core.cuh),
// one tile per CTA in a wave launch (launch_scp below; with a queue counter the CTAs are persistent and pull
// tiles, kept for comparison); a CTA is one warp, seven CTAs share an SM.  Each instance runs
// the driver state machine (cmpc_tile.cuh: advance(), replicated in the lanes of its team); the warp
// executes one whole-horizon operation at a time for the instances that asked for it, lowest operation
// code first, so that instances that are ahead wait at the later operations (evaluate, write) and the
// tile does those together.  Warps never synchronise with each other.
// ------------------------------------------------------------------------------------------
constexpr int THREADS = 32;
constexpr int BAR_BYTES = 128;  // 4 mbarriers ("full" per ring slot), then the byte-range table of the open stream (5 x 16 bytes)
inline long scp_smem_bytes(int N, bool gen) { return (long)tile_smem_fields(gen) * TL * 8 + BAR_BYTES + ((N + 1 + 15) & ~15); }

#if defined(CMPC_PROFILE)
// profiling build: cycles per operation kind summed over all warps (lane 0 of each warp)
// [0..9] per Op code (sweeps: whole op), [10] backward part of ADMM sweeps, [11] of PMM sweeps,
// [12] setup, [13] whole tile, [14] cycles waiting for bulk copies, [15] number of waits
__device__ unsigned long long g_prof[32];
#endif

template <bool FAST>
__device__ void run_tile(const Params& prm, const Batch& bt, int tile, TileCtx& T, unsigned char* nst_s) {
  const int lane = (int)(threadIdx.x & 31u);
  const int t = lane & (TL - 1), q = lane / TL;
  bind_tile(T, prm, bt, tile);
#if defined(CMPC_PROFILE)
  long long prof[32];
  for (int i = 0; i < 32; ++i) prof[i] = 0;
  T.prof = prof;
  const long long tile_t0 = clock64();
#endif
  const int b = tile * TL + t;
  const bool live = b < bt.B;
  Inst I;
  Sv S;
  Drv D;
  D.check = D.upd = 0;
  bind_instance(I, prm, bt, live ? b : bt.B - 1);   // lanes past the batch run along on a valid instance, without writing
  I.lane = t;
  I.sub = q;
  for (int k0 = 0; k0 <= prm.N; k0 += NL) {   // slots per knot of the tile: the record layout of the knot
    const int k = k0 + q;
    int ns = (live && k < prm.N) ? active_slots(prm, I, k) : 0;
#pragma unroll
    for (int m = 1; m < TL; m <<= 1) ns = max(ns, __shfl_xor_sync(0xffffffffu, ns, m));
    if (t == 0 && k <= prm.N) { T.nst[k] = ns; nst_s[k] = (unsigned char)ns; }
  }
  __syncwarp();
  {
    double mq = 0.0, mc = 0.0;
    int nconv = 0;
    setup_knots(prm, T, I, live, &mq, &mc, &nconv);
    setup_finish(I, S, mq, mc, nconv);
  }
  __syncwarp();
#if defined(CMPC_PROFILE)
  prof[12] += clock64() - tile_t0;
#endif
  int op = OP_DONE;
  if (live) {
    drv_init(prm, S, D);
    op = advance(prm, S, D);
  }
  for (;;) {
    const int sel = __reduce_min_sync(0xffffffffu, op);
    if (sel == OP_DONE) break;
    const bool on = op == sel;
    const bool anycheck = __any_sync(0xffffffffu, on && D.check);
    execute<FAST>(sel, prm, T, I, bt, S, D, on, anycheck);
    if (on) op = advance(prm, S, D);
    __syncwarp();
  }
  if (live) write_stats(bt, I, S, D);
#if defined(CMPC_PROFILE)
  prof[13] += clock64() - tile_t0;
  if (lane == 0)
    for (int i = 0; i < 32; ++i) atomicAdd(&g_prof[i], (unsigned long long)prof[i]);
#endif
}

template <bool FAST>
__global__ void __launch_bounds__(THREADS, 8)
cmpc_scp_kernel(const __grid_constant__ Params prm, const __grid_constant__ Batch bt, int* __restrict__ queue, int tile0,
                int tiles) {   // this launch solves the tiles [tile0, tiles): tile0 + blockIdx.x, or pulled from the queue counter
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const unsigned lane = threadIdx.x & 31u;
  constexpr long SCR_BYTES = (long)tile_smem_fields(!FAST) * TL * 8;
  TileCtx T;
  T.smem_sa = smem_addr(smem_raw);
  T.bars_sa = T.smem_sa + (unsigned)SCR_BYTES;
  T.phases = 0;
  unsigned char* nst_s = smem_raw + SCR_BYTES + BAR_BYTES;
  T.nst_sa = smem_addr(nst_s);
  if (lane == 0) {
    for (int d = 0; d < 4; ++d)   // "full" barriers of the ring slots (bulk-copy completion)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(T.bars_sa + 8u * d) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (!queue) {   // wave launch: one tile per CTA, every warp of the chip starts the same code together
    const int tile = tile0 + (int)blockIdx.x;
    if (tile < tiles) run_tile<FAST>(prm, bt, tile, T, nst_s);
    return;
  }
  for (;;) {
    int tile = 0;
    if (lane == 0) tile = tile0 + atomicAdd(queue, 1);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= tiles) break;
    run_tile<FAST>(prm, bt, tile, T, nst_s);
    __syncwarp();
  }
}


struct LaunchCfg { int num_sms; long smem_sm, smem_max; int* queue; };

// Fills the solver parameters and launches the tiles [tile0, tile1) of the batch on `st`.  Returns 0 or a
// negative error code with the text in *msg; *n_launches counts the kernels that went out.
inline int launch_scp(const cmpc_dims* dims, const cmpc_model* model, const cmpc_scp_params* scp, const cmpc_qp_settings* qp,
                      const Batch& bt_in, const LaunchCfg& cfg, int tile0, int tile1, cudaStream_t st, std::string* msg,
                      long long* n_launches) {
  Params prm;
  cmpc_qp_settings dq;
  if (!qp && (bt_in.fub || WR)) {   // upper bounds on the friction rows (stochastic mode, CoP box): the polish needs
    default_qp_settings(&dq);       // more multiplier sweeps and active-set rounds to certify
    dq.polish_refine_iter = 10;
    dq.polish_active_set_rounds = 19;
    if (WR) { dq.active_set_tol = 1e-11; dq.polish_refine_iter = 30; dq.delta = 1e-9; dq.alpha = 1.6; }   // multipliers of the order of the 900 N forces: 1e-9 leaves 6e-6 in X;
                                                                          // the multiplier sweeps stop as soon as the certificate holds
    // (the first polish attempt stays at 8 ADMM iterations also with back-offs: measured on B200, 4096 x N=100,
    //  first attempt after 20 / 16 / 12 / 8 iterations: trot 21.4 / 19.1 / 17.5 / 16.6 ms, bound 20.3 / 19.0 / 18.3 / 17.4 ms)
    qp = &dq;
  }
  int rc = fill_params(&prm, dims, model, scp, qp, bt_in.cR == nullptr);
  if (rc) { *msg = rc == -2 ? "cost weights must be positive" : "bad dims"; return rc; }
  Batch bt = bt_in;
  if (bt.fub) prm.fast = 0;   // upper bounds live in the general friction table
  bt.rfields = rec_fields(prm.nc, !prm.fast);
  const long smem = scp_smem_bytes(prm.N, !prm.fast);
  if (smem > cfg.smem_max) { *msg = "horizon too long for the shared-memory slot table"; return -3; }
  int per_sm = (int)(cfg.smem_sm / (smem + 1024));   // 1 KB per block is reserved by the driver
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;                        // registers: __launch_bounds__(32, 8)
  const int cap = cfg.num_sms * per_sm;
  // Waves: the tiles go out in back-to-back launches of at most one resident set (cap CTAs, one tile each).
  // The solver's code (0.77 MB, of which a knot loop touches 15-30 KB) is far larger than the 32 KB
  // instruction cache of an SM; warps that start together run the same operations at about the same time
  // and share the fetched lines.  A persistent grid pulling tiles from a queue lets the warps drift apart:
  // measured on B200 at 16384 instances, 37 x more instruction-cache misses (ncu gcc__cache_requests_type_
  // instruction_lookup_miss), 10 instead of 2 no_instruction stall cycles per issue and 62 ms instead of
  // 4 x 9.3 ms (profiles/r2_icache.md).  CMPC_PERSISTENT=1 selects the queue (for that comparison).
  static const bool persistent = [] { const char* e = getenv("CMPC_PERSISTENT"); return e && e[0] == '1'; }();
  const int total = tile1 - tile0;
  auto check = [&](cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    *msg = std::string(what) + ": " + cudaGetErrorString(e);
    return -100 - (int)e;
  };
  if (persistent) {
    if ((rc = check(cudaMemsetAsync(cfg.queue, 0, sizeof(int), st), "cudaMemsetAsync"))) return rc;
    const int blocks = total > cap ? cap : total;
    if (prm.fast) cmpc_scp_kernel<true><<<blocks, THREADS, smem, st>>>(prm, bt, cfg.queue, tile0, tile1);
    else cmpc_scp_kernel<false><<<blocks, THREADS, smem, st>>>(prm, bt, cfg.queue, tile0, tile1);
    ++*n_launches;
    return check(cudaGetLastError(), "cmpc_scp_kernel");
  }
  const int waves = (total + cap - 1) / cap;
  const int per = (total + waves - 1) / waves;       // equal waves: no short last one
  for (int w0 = tile0; w0 < tile1; w0 += per) {
    const int w1 = w0 + per < tile1 ? w0 + per : tile1;
    if (prm.fast) cmpc_scp_kernel<true><<<w1 - w0, THREADS, smem, st>>>(prm, bt, nullptr, w0, w1);
    else cmpc_scp_kernel<false><<<w1 - w0, THREADS, smem, st>>>(prm, bt, nullptr, w0, w1);
    ++*n_launches;
    if ((rc = check(cudaGetLastError(), "cmpc_scp_kernel"))) return rc;
  }
  return 0;
}

// the kernels' dynamic shared memory limit: the device maximum, so that handles with different horizons
// can be alive at the same time (the attribute is per function and device)
inline cudaError_t set_scp_smem_limit(int smem_optin) {
  cudaError_t e = cudaFuncSetAttribute(cmpc_scp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(cmpc_scp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
}

}  // namespace cmpc
