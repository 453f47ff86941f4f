// cmpc_api.cu — CUDA kernels and the C ABI (include/cmpc.h) of libcmpc_b200.so.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <string>

#include "cmpc_launch.cuh"
#include "cmpc_lqr.cuh"

using namespace cmpc;

// the second compilation of the solver (cmpc_wrench.cu): CoP / wrench contact model
struct WrSizes { long tiles, ws, nst, info, smem; };
int cmpc_wr_sizes(int B, int N, int feet, WrSizes* out);
int cmpc_wr_set_smem_limit(int smem_optin);
int cmpc_wr_launch_scp(const cmpc_dims* dims, const cmpc_model* model, const cmpc_scp_params* scp, const cmpc_qp_settings* qp,
                       const void* batch, size_t batch_bytes, const void* cfg, int tile0, int tile1, cudaStream_t st,
                       std::string* msg, long long* n_launches);
int cmpc_wr_lqr_covs(const cmpc_dims* dims, const cmpc_model* model, const cmpc_lqr_weights* w, const double* X, const double* U,
                     const double* contact_pos, const double* contact_R, const int32_t* contact_active, double* gains,
                     double* covs, void* scratch, cudaStream_t st, std::string* msg, long long* n_launches);
int cmpc_wr_linearize(const cmpc_dims* dims, const cmpc_model* model, const double* X, const double* U, const double* contact_pos,
                      const double* contact_R, const int32_t* contact_active, double* f, double* fx, double* fu, cudaStream_t st,
                      std::string* msg);

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess) return fail(-100 - (int)e_, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// compute_trajectory_data / integrate_dynamics_trajectory (one thread per instance and knot)
__global__ void cmpc_linearize_kernel(const __grid_constant__ Params prm, int B, int shared_plan,
                                      const double* __restrict__ X, const double* __restrict__ U,
                                      const double* __restrict__ cpos, const int* __restrict__ cact,
                                      double* __restrict__ f, double* __restrict__ fx, double* __restrict__ fu) {
  const int N = prm.N, nu = prm.nu, nc = prm.nc;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)B * N) return;
  int b = (int)(t / N), k = (int)(t % N);
  const double* x = X + ((long)b * (N + 1) + k) * 9;
  const double* u = U + ((long)b * N + k) * nu;
  const long plan = shared_plan ? 0 : b;
  const double* p = cpos + (plan * N + k) * nc * 3;
  const int* a = cact + (plan * N + k) * nc;
  double xn[9];
  step_knot(prm, x, u, p, a, xn);
  for (int i = 0; i < 9; ++i) f[t * 9 + i] = xn[i];
  if (!fx) return;
  KnotLin L;
  linearize_knot(prm, x, u, p, a, 0, L);
  const int mt = L.meta;
  double A[81];
  dense_A(prm, L.S, A);
  for (int i = 0; i < 81; ++i) fx[t * 81 + i] = A[i];
  for (int i = 0; i < 9 * nu; ++i) fu[t * 9 * nu + i] = 0.0;
  const int ns = mt & 7;
  for (int sl = 0; sl < ns; ++sl) {
    const int ct = (mt >> (4 + 2 * sl)) & 3;
    for (int ax = 0; ax < 3; ++ax) {
      double col[9];
      dense_Bcol(prm, &L.d[3 * sl], ax, col);
      for (int i = 0; i < 9; ++i) fu[(t * 9 + i) * nu + 3 * ct + ax] = col[i];
    }
  }
}

// friction-row upper bounds in stochastic mode (one thread per instance and knot): ub [B][N][nc][4]
__global__ void cmpc_backoff_kernel(const __grid_constant__ Params prm, double xi, int B, int shared_plan,
                                    const double* __restrict__ gains, const double* __restrict__ covs,
                                    const double* __restrict__ cR, const int* __restrict__ cact,
                                    double* __restrict__ ub) {
  const int N = prm.N, nu = prm.nu, nc = prm.nc;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)B * N) return;
  int b = (int)(t / N), k = (int)(t % N);
  const long plan = shared_plan ? 0 : b;
  double o[4 * MAXC];
  friction_backoff_knot(prm, xi, k, gains + t * nu * 9, covs + ((long)b * (N + 1) + k) * 81,
                        cR ? cR + (plan * N + k) * nc * 9 : nullptr, cact + (plan * N + k) * nc, o);
  for (int i = 0; i < 4 * nc; ++i) ub[t * 4 * nc + i] = o[i];
}

// DFMA micro-benchmark: 8 independent FMA chains per thread
__global__ void cmpc_dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(long)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace

struct cmpc_handle_s {
  cmpc_dims dims;
  cmpc_model model;
  int have_problem;
  int device;
  int num_sms;
  Batch bt;          // device pointers
  int tiles;
  long smem_max, smem_sm, wr_smem;
  cudaStream_t last_stream;
  void* ws;          // one allocation
  long ws_bytes;
  int* queue;
  int *d_nacc, *d_qpit, *d_nfac;
  // staging for the host entry point
  void *d_in, *d_out, *h_pin;
  long in_bytes, out_bytes;
  // host entry point: the batch is cut into up to MAX_CHUNKS tile-aligned chunks, each with its own
  // stream (H2D of chunk c+1 and D2H of chunk c-1 overlap the solve of chunk c)
  cudaStream_t cs[8];
  cudaEvent_t ev_small, ev_up[8], ev_done[8];
  int have_streams;
};
static const int MAX_CHUNKS = 8;

extern "C" {

const char* cmpc_last_error(void) { return g_err.c_str(); }
const char* cmpc_version(void) { return "cmpc_b200 0.2.0 (sm_100a)"; }
#ifndef CMPC_BUILD_ID
#define CMPC_BUILD_ID "unstamped"
#endif
const char* cmpc_build_id(void) { return CMPC_BUILD_ID; }
int64_t cmpc_launch_count(void) { return g_launches.load(); }
void cmpc_default_qp_settings(cmpc_qp_settings* s) { default_qp_settings(s); }

// every error path after the allocation releases the handle
static int create_fail(cmpc_handle h, int code, const std::string& msg) {
  if (h) {
    if (h->ws) cudaFree(h->ws);
    free(h);
  }
  return fail(code, msg);
}
#define CREATE_TRY(expr)                                                                       \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess) return create_fail(h, -100 - (int)e_, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

int cmpc_create(const cmpc_dims* dims, cmpc_handle* out) {
  if (!dims || !out) return fail(-1, "null argument");
  if (dims->N < 1 || dims->nc < 1 || dims->nc > MAXC || dims->batch < 1) return fail(-1, "bad dims");
  const bool wrench = dims->contact_model == CMPC_CONTACT_WRENCH;
  if (dims->contact_model != CMPC_CONTACT_POINT && !wrench) return fail(-1, "unknown contact model");
  if (wrench && 2 * dims->nc > MAXC) return fail(-1, "the wrench contact model takes at most two feet");
  cmpc_handle h = (cmpc_handle)calloc(1, sizeof(cmpc_handle_s));
  if (!h) return fail(-1, "out of host memory");
  h->dims = *dims;
  CREATE_TRY(cudaGetDevice(&h->device));
  CREATE_TRY(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
  const int B = dims->batch, N = dims->N;
  WsSizes w = ws_sizes(B, N, dims->nc);
  long smem_need = scp_smem_bytes(N, true);
  if (wrench) {
    WrSizes z;
    cmpc_wr_sizes(B, N, dims->nc, &z);
    w.tiles = z.tiles; w.ws = z.ws; w.nst = z.nst; w.info = z.info;
    smem_need = z.smem;
  }
  h->wr_smem = smem_need;
  const long nd = w.ws + w.info;
  const long ni = w.nst + 3L * B + 64;
  h->ws_bytes = nd * 8 + ni * 4;
  CREATE_TRY(cudaMalloc(&h->ws, h->ws_bytes));
  double* d = (double*)h->ws;
  h->bt.ws = d; d += w.ws;        // 256-byte aligned records first
  h->bt.info = d; d += w.info;
  int* ip = (int*)d;
  h->bt.nst = ip; ip += w.nst;
  h->d_nacc = ip; ip += B;
  h->d_qpit = ip; ip += B;
  h->d_nfac = ip; ip += B;
  h->queue = ip;
  h->tiles = (int)w.tiles;
  int smem_optin = 0, smem_sm = 0;
  CREATE_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  CREATE_TRY(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, h->device));
  h->smem_max = smem_optin;
  h->smem_sm = smem_sm;
  if (smem_need > h->smem_max) return create_fail(h, -3, "horizon too long for the shared-memory slot table");
  // the attribute is per function and device: always the device maximum, so that handles with different
  // horizons can be alive at the same time
  CREATE_TRY(set_scp_smem_limit(smem_optin));
  if (wrench) CREATE_TRY((cudaError_t)cmpc_wr_set_smem_limit(smem_optin));
  h->bt.B = B;
  h->bt.plan_stride = dims->shared_plan ? 0 : 1;
  *out = h;
  return 0;
}

int64_t cmpc_workspace_bytes(cmpc_handle h) { return h ? h->ws_bytes : 0; }

int cmpc_destroy(cmpc_handle h) {
  if (!h) return 0;
  cudaFree(h->ws);
  if (h->d_in) cudaFree(h->d_in);
  if (h->d_out) cudaFree(h->d_out);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  if (h->have_streams) {
    for (int c = 0; c < MAX_CHUNKS; ++c) { cudaStreamDestroy(h->cs[c]); cudaEventDestroy(h->ev_up[c]); cudaEventDestroy(h->ev_done[c]); }
    cudaEventDestroy(h->ev_small);
  }
  free(h);
  return 0;
}

int cmpc_set_problem(cmpc_handle h, const cmpc_model* model, const double* x_init, const double* x_final,
                     const double* X_ref, const double* U_init, const double* contact_pos,
                     const double* contact_R, const int32_t* contact_active) {
  if (!h || !model || !x_init || !x_final || !X_ref || !U_init || !contact_pos || !contact_active)
    return fail(-1, "null argument");
  if (h->dims.contact_model == CMPC_CONTACT_WRENCH && !contact_R) return fail(-1, "the wrench contact model needs contact_R");
  h->model = *model;
  h->bt.x_init = x_init; h->bt.x_final = x_final; h->bt.X_ref = X_ref; h->bt.U_init = U_init;
  h->bt.cpos = contact_pos; h->bt.cR = contact_R; h->bt.cact = (const int*)contact_active;
  h->have_problem = 1;
  return 0;
}

int cmpc_set_friction_ub(cmpc_handle h, const double* friction_ub) {
  if (!h) return fail(-1, "null argument");
  if (friction_ub && h->dims.contact_model == CMPC_CONTACT_WRENCH) return fail(-1, "friction back-offs are not available for the wrench contact model");
  h->bt.fub = friction_ub;
  return 0;
}

// launches the solver for the tiles [tile0, tile1) of the batch `bt_in` on `st` (cmpc_launch.cuh); the wrench
// contact model runs the second compilation of the solver (cmpc_wrench.cu)
static int launch_tiles(cmpc_handle h, const Batch& bt_in, const cmpc_model* model, const cmpc_scp_params* scp,
                        const cmpc_qp_settings* qp, double* X_out, double* U_out, int32_t* scp_iters, int32_t* status,
                        int32_t* n_accepted, int tile0, int tile1, int slot, cudaStream_t st) {
  Batch bt = bt_in;
  bt.X_out = X_out; bt.U_out = U_out; bt.scp_iters = (int*)scp_iters; bt.status = (int*)status;
  bt.n_accepted = n_accepted ? (int*)n_accepted : h->d_nacc;
  bt.qp_iters = h->d_qpit; bt.n_factor = h->d_nfac;
  LaunchCfg cfg{h->num_sms, h->smem_sm, h->smem_max, h->queue + slot};
  std::string msg;
  long long nl = 0;
  const int rc = h->dims.contact_model == CMPC_CONTACT_WRENCH
                     ? cmpc_wr_launch_scp(&h->dims, model, scp, qp, &bt, sizeof(Batch), &cfg, tile0, tile1, st, &msg, &nl)
                     : launch_scp(&h->dims, model, scp, qp, bt, cfg, tile0, tile1, st, &msg, &nl);
  g_launches.fetch_add(nl);
  return rc ? fail(rc, msg) : 0;
}

int cmpc_solve_scp(cmpc_handle h, const cmpc_scp_params* scp, const cmpc_qp_settings* qp, double* X_out,
                   double* U_out, int32_t* scp_iters, int32_t* status, int32_t* n_accepted, void* stream) {
  if (!h || !scp || !X_out || !U_out || !scp_iters || !status) return fail(-1, "null argument");
  if (!h->have_problem) return fail(-2, "cmpc_set_problem has not been called");
  h->last_stream = (cudaStream_t)stream;
  return launch_tiles(h, h->bt, &h->model, scp, qp, X_out, U_out, scp_iters, status, n_accepted, 0, h->tiles, 0,
                      (cudaStream_t)stream);
}

int cmpc_get_stats(cmpc_handle h, int32_t* qp_iters, int32_t* n_factor, double* info, void* stream) {
  if (!h) return fail(-1, "null handle");
  const int B = h->dims.batch;
  cudaStream_t st = (cudaStream_t)stream;
  if (qp_iters) CUDA_TRY(cudaMemcpyAsync(qp_iters, h->d_qpit, B * sizeof(int), cudaMemcpyDeviceToDevice, st));
  if (n_factor) CUDA_TRY(cudaMemcpyAsync(n_factor, h->d_nfac, B * sizeof(int), cudaMemcpyDeviceToDevice, st));
  if (info) CUDA_TRY(cudaMemcpyAsync(info, h->bt.info, (long)B * INFO * sizeof(double), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int cmpc_solve_scp_host(cmpc_handle h, const cmpc_model* model, const cmpc_scp_params* scp,
                        const cmpc_qp_settings* qp, const double* x_init, const double* x_final,
                        const double* X_ref, const double* U_init, const double* contact_pos,
                        const double* contact_R, const int32_t* contact_active, double* X_out,
                        double* U_out, int32_t* scp_iters, int32_t* status, int32_t* n_accepted) {
  if (!h || !model || !scp) return fail(-1, "null argument");
  if (!x_init || !x_final || !X_ref || !U_init || !contact_pos || !contact_active || !X_out || !U_out || !scp_iters || !status)
    return fail(-1, "null argument");
  const bool wrench = h->dims.contact_model == CMPC_CONTACT_WRENCH;   // nc feet, six controls each
  if (wrench && !contact_R) return fail(-1, "the wrench contact model needs contact_R");
  const int B = h->dims.batch, N = h->dims.N, nc = h->dims.nc, nu = (wrench ? 6 : 3) * nc;
  const long Bp = h->dims.shared_plan ? 1 : B;
  const long n_xi = (long)B * 9, n_X = (long)B * (N + 1) * 9, n_U = (long)B * N * nu;
  const long n_cp = Bp * N * nc * 3, n_cR = contact_R ? Bp * N * nc * 9 : 0, n_ca = Bp * N * nc;
  const long in_d = 2 * n_xi + n_X + n_U + n_cp + n_cR;
  const long in_bytes = in_d * 8 + n_ca * 4;
  const long out_bytes = (n_X + n_U) * 8 + 3L * B * 4;
  if (!h->d_in || h->in_bytes < in_bytes) {
    if (h->d_in) cudaFree(h->d_in);
    h->d_in = nullptr;
    CUDA_TRY(cudaMalloc(&h->d_in, in_bytes));
    h->in_bytes = in_bytes;
  }
  if (!h->d_out || h->out_bytes < out_bytes) {
    if (h->d_out) cudaFree(h->d_out);
    h->d_out = nullptr;
    CUDA_TRY(cudaMalloc(&h->d_out, out_bytes));
    h->out_bytes = out_bytes;
  }
  double* d = (double*)h->d_in;
  double *dxi = d, *dxf = d + n_xi, *dX = d + 2 * n_xi, *dU = dX + n_X, *dcp = dU + n_U, *dcR = dcp + n_cp;
  int* dca = (int*)(dcR + n_cR);
  if (!h->have_streams) {
    for (int c = 0; c < MAX_CHUNKS; ++c) CUDA_TRY(cudaStreamCreate(&h->cs[c]));   // blocking: ordered after earlier work on the null stream
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_small, cudaEventDisableTiming));
    for (int c = 0; c < MAX_CHUNKS; ++c) {
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_up[c], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_done[c], cudaEventDisableTiming));
    }
    h->have_streams = 1;
  }
  // Small inputs and the contact plan first, on chunk 0's stream; the other chunks wait for them.
  cudaStream_t s0 = h->cs[0];
  CUDA_TRY(cudaMemcpyAsync(dxi, x_init, n_xi * 8, cudaMemcpyHostToDevice, s0));
  CUDA_TRY(cudaMemcpyAsync(dxf, x_final, n_xi * 8, cudaMemcpyHostToDevice, s0));
  CUDA_TRY(cudaMemcpyAsync(dcp, contact_pos, n_cp * 8, cudaMemcpyHostToDevice, s0));
  if (contact_R) CUDA_TRY(cudaMemcpyAsync(dcR, contact_R, n_cR * 8, cudaMemcpyHostToDevice, s0));
  CUDA_TRY(cudaMemcpyAsync(dca, contact_active, n_ca * 4, cudaMemcpyHostToDevice, s0));
  CUDA_TRY(cudaEventRecord(h->ev_small, s0));
  // a private view of the batch over the staging buffers: the problem bound with cmpc_set_problem (and the
  // friction upper bounds, which are device data of the caller) stays as it is
  Batch bt = h->bt;
  bt.x_init = dxi; bt.x_final = dxf; bt.X_ref = dX; bt.U_init = dU;
  bt.cpos = dcp; bt.cR = contact_R ? dcR : nullptr; bt.cact = dca;
  double* oX = (double*)h->d_out;
  double* oU = oX + n_X;
  int* oi = (int*)(oU + n_U);
  int *oit = oi, *ost = oi + B, *ona = oi + 2 * B;
  // Page-locked (mapped) result buffers: the kernel writes the solution straight into host memory at a
  // tile's write-back, so that the transfer of the tiles that finish early hides behind the slower ones
  // and no device-to-host copy is left at the end.  Pageable buffers go through the staging buffer.
  auto mapped = [](const void* p, void** dev) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
    *dev = at.devicePointer;
    return true;
  };
  void *mX = nullptr, *mU = nullptr, *mi = nullptr, *ms = nullptr, *mn = nullptr;
  static const int env_zc = [] { const char* e = getenv("CMPC_HOST_ZEROCOPY"); return e ? atoi(e) : 1; }();
  static const int env_chunks = [] { const char* e = getenv("CMPC_HOST_CHUNKS"); return e ? atoi(e) : 0; }();
  const bool zero_copy = env_zc && mapped(X_out, &mX) && mapped(U_out, &mU) && mapped(scp_iters, &mi) && mapped(status, &ms) &&
                         (!n_accepted || mapped(n_accepted, &mn));
  if (zero_copy) {
    oX = (double*)mX; oU = (double*)mU; oit = (int*)mi; ost = (int*)ms; ona = n_accepted ? (int*)mn : oi + 2 * B;
  }
  // Chunks of whole tiles.  ALL kernels go to one stream, one wave after the other; the trajectories of chunk
  // c+1 are uploaded on a second stream while chunk c is being solved, and, without mapped buffers, the results
  // of chunk c-1 go down on a third.  (Kernels of different chunks must not overlap: a chunk on its own stream
  // starts in the tail of the previous one, its warps drift apart and the instruction caches thrash -- measured
  // at 16 384 instances: 60 ms with one stream per chunk, 44 ms with a single upload and one stream.)
  // One chunk per resident set of tiles (a "wave", cmpc_launch.cuh), at most MAX_CHUNKS: the warps of a wave must
  // start together (they share instruction fetches), so a batch that fits one wave is ONE upload and ONE launch --
  // measured on B200, 4096 instances: 11.2 ms with 1 chunk, 12.1 / 13.0 / 13.4 ms with 2 / 4 / 8 chunks
  // (scripts/e2e_chunks.py).  CMPC_HOST_CHUNKS overrides the chunk count (for that comparison).
  const int tiles = h->tiles;
  long smem = h->wr_smem;            // shared memory per CTA of the path this batch takes -> CTAs per SM -> resident set
  if (!wrench) {
    Params prm;
    if (fill_params(&prm, &h->dims, model, scp, qp, contact_R == nullptr)) return fail(-1, "bad dims or weights");
    smem = scp_smem_bytes(N, !(prm.fast && !h->bt.fub));
  }
  int per_sm = (int)(h->smem_sm / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  const int cap = h->num_sms * per_sm;
  int chunks = (tiles + cap - 1) / cap;
  if (chunks > MAX_CHUNKS) chunks = MAX_CHUNKS;
  if (env_chunks >= 1 && env_chunks <= MAX_CHUNKS && env_chunks <= tiles) chunks = env_chunks;
  const int per = (tiles + chunks - 1) / chunks;
  int rc = 0;
  for (int c = 0; c < chunks; ++c) {
    const int t0 = c * per, t1 = (c + 1) * per < tiles ? (c + 1) * per : tiles;
    if (t0 >= t1) break;
    const long b0 = (long)t0 * TL, b1 = (long)t1 * TL < B ? (long)t1 * TL : B, nb = b1 - b0;
    cudaStream_t s_kern = h->cs[0], s_up = chunks > 1 ? h->cs[1] : h->cs[0], s_down = h->cs[2];
    if (c == 0 && chunks > 1) CUDA_TRY(cudaStreamWaitEvent(s_up, h->ev_small, 0));   // (staging buffers: after the previous call's work)
    const long xo = b0 * (N + 1) * 9, uo = b0 * N * nu;
    CUDA_TRY(cudaMemcpyAsync(dX + xo, X_ref + xo, nb * (N + 1) * 9 * 8, cudaMemcpyHostToDevice, s_up));
    CUDA_TRY(cudaMemcpyAsync(dU + uo, U_init + uo, nb * N * nu * 8, cudaMemcpyHostToDevice, s_up));
    if (chunks > 1) {
      CUDA_TRY(cudaEventRecord(h->ev_up[c], s_up));
      CUDA_TRY(cudaStreamWaitEvent(s_kern, h->ev_up[c], 0));
    }
    rc = launch_tiles(h, bt, model, scp, qp, oX, oU, oit, ost, ona, t0, t1, 0, s_kern);
    if (rc) break;
    if (zero_copy) continue;
    CUDA_TRY(cudaEventRecord(h->ev_done[c], s_kern));
    CUDA_TRY(cudaStreamWaitEvent(s_down, h->ev_done[c], 0));
    CUDA_TRY(cudaMemcpyAsync(X_out + xo, oX + xo, nb * (N + 1) * 9 * 8, cudaMemcpyDeviceToHost, s_down));
    CUDA_TRY(cudaMemcpyAsync(U_out + uo, oU + uo, nb * N * nu * 8, cudaMemcpyDeviceToHost, s_down));
    CUDA_TRY(cudaMemcpyAsync(scp_iters + b0, oit + b0, nb * 4, cudaMemcpyDeviceToHost, s_down));
    CUDA_TRY(cudaMemcpyAsync(status + b0, ost + b0, nb * 4, cudaMemcpyDeviceToHost, s_down));
    if (n_accepted) CUDA_TRY(cudaMemcpyAsync(n_accepted + b0, ona + b0, nb * 4, cudaMemcpyDeviceToHost, s_down));
  }
  for (int c = 0; c < 3; ++c) CUDA_TRY(cudaStreamSynchronize(h->cs[c]));
  return rc;
}

static int lin_common(const cmpc_dims* dims, const cmpc_model* model, const double* X, const double* U,
                      const double* contact_pos, const int32_t* contact_active, double* f, double* fx,
                      double* fu, void* stream) {
  if (!dims || !model || !X || !U || !contact_pos || !contact_active || !f) return fail(-1, "null argument");
  if (dims->contact_model != CMPC_CONTACT_POINT) return fail(-1, "point-contact entry: use cmpc_linearize_wrench for the wrench model");
  Params prm;
  int rc = fill_params(&prm, dims, model, nullptr, nullptr, 1);
  if (rc) return fail(rc, "bad dims or weights");
  long total = (long)dims->batch * dims->N;
  int threads = 128;
  long blocks = (total + threads - 1) / threads;
  cmpc_linearize_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      prm, dims->batch, dims->shared_plan, X, U, contact_pos, (const int*)contact_active, f, fx, fu);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int cmpc_linearize(const cmpc_dims* dims, const cmpc_model* model, const double* X, const double* U,
                   const double* contact_pos, const int32_t* contact_active, double* f, double* fx,
                   double* fu, void* stream) {
  if (!fx || !fu) return fail(-1, "null argument");
  return lin_common(dims, model, X, U, contact_pos, contact_active, f, fx, fu, stream);
}

int cmpc_rollout(const cmpc_dims* dims, const cmpc_model* model, const double* X, const double* U,
                 const double* contact_pos, const int32_t* contact_active, double* f, void* stream) {
  return lin_common(dims, model, X, U, contact_pos, contact_active, f, nullptr, nullptr, stream);
}

int cmpc_linearize_wrench(const cmpc_dims* dims, const cmpc_model* model, const double* X, const double* U,
                          const double* contact_pos, const double* contact_R, const int32_t* contact_active, double* f,
                          double* fx, double* fu, void* stream) {
  if (!dims || !model || !X || !U || !contact_pos || !contact_R || !contact_active || !f) return fail(-1, "null argument");
  if (dims->contact_model != CMPC_CONTACT_WRENCH) return fail(-1, "cmpc_linearize_wrench needs dims.contact_model = CMPC_CONTACT_WRENCH");
  if ((fx == nullptr) != (fu == nullptr)) return fail(-1, "fx and fu go together");
  std::string msg;
  const int rc = cmpc_wr_linearize(dims, model, X, U, contact_pos, contact_R, contact_active, f, fx, fu, (cudaStream_t)stream, &msg);
  g_launches.fetch_add(1);
  return rc ? fail(rc, msg) : 0;
}

int cmpc_lqr_covs(const cmpc_dims* dims, const cmpc_model* model, const cmpc_lqr_weights* w, const double* X,
                  const double* U, const double* contact_pos, const int32_t* contact_active, double* gains,
                  double* covs, void* scratch, void* stream) {
  if (!dims || !model || !w || !X || !U || !contact_pos || !contact_active || !gains || !scratch)
    return fail(-1, "null argument");
  if (dims->contact_model != CMPC_CONTACT_POINT) return fail(-1, "point-contact entry: use cmpc_lqr_covs_wrench for the wrench model");
  static_assert(sizeof(cmpc_lqr_weights) == sizeof(LqrWeights), "cmpc_lqr_weights layout");
  Params prm;
  int rc = fill_params(&prm, dims, model, nullptr, nullptr, 1);
  if (rc) return fail(rc, "bad dims or weights");
  const int nu = prm.nu;
  for (int i = 0; i < nu; ++i)
    if (!(w->R[i * nu + i] > 0.0)) return fail(-2, "R must have a positive diagonal");
  long long nl = 0;
  const cudaError_t e = launch_lqr_covs(prm, (const LqrWeights*)w, dims->batch, dims->shared_plan, X, U, contact_pos, nullptr,
                                        (const int*)contact_active, gains, covs, scratch, (cudaStream_t)stream, &nl);
  g_launches.fetch_add(nl);
  CUDA_TRY(e);
  return 0;
}

int cmpc_lqr_covs_wrench(const cmpc_dims* dims, const cmpc_model* model, const cmpc_lqr_weights* w, const double* X,
                         const double* U, const double* contact_pos, const double* contact_R, const int32_t* contact_active,
                         double* gains, double* covs, void* scratch, void* stream) {
  if (!dims || !model || !w || !X || !U || !contact_pos || !contact_R || !contact_active || !gains || !scratch)
    return fail(-1, "null argument");
  if (dims->contact_model != CMPC_CONTACT_WRENCH) return fail(-1, "cmpc_lqr_covs_wrench needs dims.contact_model = CMPC_CONTACT_WRENCH");
  std::string msg;
  long long nl = 0;
  const int rc = cmpc_wr_lqr_covs(dims, model, w, X, U, contact_pos, contact_R, contact_active, gains, covs, scratch,
                                  (cudaStream_t)stream, &msg, &nl);
  g_launches.fetch_add(nl);
  return rc ? fail(rc, msg) : 0;
}

int cmpc_friction_backoffs(const cmpc_dims* dims, const cmpc_model* model, double xi, const double* gains,
                           const double* covs, const double* contact_R, const int32_t* contact_active,
                           double* friction_ub, void* stream) {
  if (!dims || !model || !gains || !covs || !contact_active || !friction_ub) return fail(-1, "null argument");
  Params prm;
  int rc = fill_params(&prm, dims, model, nullptr, nullptr, contact_R == nullptr);
  if (rc) return fail(rc, "bad dims or weights");
  long total = (long)dims->batch * dims->N;
  cmpc_backoff_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      prm, xi, dims->batch, dims->shared_plan, gains, covs, contact_R, (const int*)contact_active, friction_ub);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// profiling build only (-DCMPC_PROFILE): reads and clears the cycle counters; returns -1 in a normal build
// ---- result buffers in the memory of another GPU of the node (include/cmpc.h "multi-GPU")
int cmpc_peer_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64) {
  if (bytes <= 0 || !dev_ptr || !handle64) return fail(-1, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, (size_t)bytes));          // a plain allocation of its own: IPC handles name whole allocations
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) { cudaFree(p); return fail(-100 - (int)e, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
  memcpy(handle64, &hd, 64);
  *dev_ptr = p;
  return 0;
}

int cmpc_peer_open(const unsigned char* handle64, void** dev_ptr) {
  if (!handle64 || !dev_ptr) return fail(-1, "null argument");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int cmpc_peer_close(void* dev_ptr) {
  if (dev_ptr) CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

int cmpc_peer_free(void* dev_ptr) {
  if (dev_ptr) CUDA_TRY(cudaFree(dev_ptr));
  return 0;
}

int cmpc_debug_profile(double* out16) {
#if defined(CMPC_PROFILE)
  unsigned long long h[32], z[32] = {0};
  CUDA_TRY(cudaMemcpyFromSymbol(h, g_prof, sizeof(h)));
  CUDA_TRY(cudaMemcpyToSymbol(g_prof, z, sizeof(z)));
  for (int i = 0; i < 32; ++i) out16[i] = (double)h[i];
  return 0;
#else
  (void)out16;
  return -1;
#endif
}

int cmpc_fp64_peak(double* tflops, double* ms_out) {
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int threads = 256, blocks = sms * 8, iters = 1 << 16;
  double* out = nullptr;
  CUDA_TRY(cudaMalloc(&out, (long)threads * blocks * 8));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  cmpc_dfma_kernel<<<blocks, threads>>>(out, iters / 16);   // warm-up
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CUDA_TRY(cudaEventRecord(e0));
    cmpc_dfma_kernel<<<blocks, threads>>>(out, iters);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  g_launches.fetch_add(4);
  double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
  if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return 0;
}

}  // extern "C"
