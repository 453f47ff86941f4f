// cmpc_wrench.cu — second compilation of the solver source for the CoP / wrench contact model (TALOS:
// /root/reference/src/centroidal_model.py:204-208, src/constraints.py:111-145, src/optimizer.py:48-64).
// The same headers as cmpc_api.cu, compiled with CMPC_WRENCH=1 into namespace cmpc_wr, so that the
// point-contact kernels stay exactly as they are.  cmpc_api.cu calls the entry points below for handles
// created with dims.contact_model == CMPC_CONTACT_WRENCH.
#define CMPC_WRENCH 1
#define cmpc cmpc_wr
#include "cmpc_launch.cuh"
#include "cmpc_lqr.cuh"
#undef cmpc

struct WrSizes { long tiles, ws, nst, info, smem; };

int cmpc_wr_sizes(int B, int N, int feet, WrSizes* out) {
  const cmpc_wr::WsSizes w = cmpc_wr::ws_sizes(B, N, feet);
  out->tiles = w.tiles; out->ws = w.ws; out->nst = w.nst; out->info = w.info;
  out->smem = cmpc_wr::scp_smem_bytes(N, true);
  return 0;
}

int cmpc_wr_set_smem_limit(int smem_optin) { return (int)cmpc_wr::set_scp_smem_limit(smem_optin); }

int cmpc_wr_launch_scp(const cmpc_dims* dims, const cmpc_model* model, const cmpc_scp_params* scp, const cmpc_qp_settings* qp,
                       const void* batch, size_t batch_bytes, const void* cfg, int tile0, int tile1, cudaStream_t st,
                       std::string* msg, long long* n_launches) {
  if (batch_bytes != sizeof(cmpc_wr::Batch)) { *msg = "batch layout mismatch between the two solver builds"; return -1; }
  return cmpc_wr::launch_scp(dims, model, scp, qp, *static_cast<const cmpc_wr::Batch*>(batch),
                             *static_cast<const cmpc_wr::LaunchCfg*>(cfg), tile0, tile1, st, msg, n_launches);
}

// compute_trajectory_data / integrate_dynamics_trajectory for the wrench model (one thread per instance and
// knot): f [B][N][9], A [B][N][9][9], B [B][N][9][6 feet] in the reference's control order
__global__ void cmpc_wr_linearize_kernel(const __grid_constant__ cmpc_wr::Params prm, int B, int shared_plan,
                                         const double* __restrict__ X, const double* __restrict__ U,
                                         const double* __restrict__ cpos, const double* __restrict__ cR,
                                         const int* __restrict__ cact, double* __restrict__ f, double* __restrict__ fx,
                                         double* __restrict__ fu) {
  using namespace cmpc_wr;
  const int N = prm.N, nu = prm.nu, nf = prm.nf;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)B * N) return;
  int b = (int)(t / N), k = (int)(t % N);
  const double* x = X + ((long)b * (N + 1) + k) * 9;
  const double* u = U + ((long)b * N + k) * nu;
  const long plan = shared_plan ? 0 : b;
  const double* p = cpos + (plan * N + k) * nf * 3;
  const double* R = cR + (plan * N + k) * nf * 9;
  const int* a = cact + (plan * N + k) * nf;
  double xn[9];
  step_knot(prm, x, u, p, a, xn, R);
  for (int i = 0; i < 9; ++i) f[t * 9 + i] = xn[i];
  if (!fx) return;
  KnotLin L;
  linearize_knot(prm, x, u, p, a, 0, L, R);
  double A[81];
  dense_A(prm, L.S, A);
  for (int i = 0; i < 81; ++i) fx[t * 81 + i] = A[i];
  for (int i = 0; i < 9 * nu; ++i) fu[t * 9 * nu + i] = 0.0;
  const int ns = L.meta & 7;
  for (int sl = 0; sl < ns; ++sl) {
    const int ct = (L.meta >> (4 + 2 * sl)) & 3;
    for (int ax = 0; ax < 3; ++ax) {
      const int col = uix(ct, ax);
      if (!(ct & 1)) fu[(t * 9 + 3 + ax) * nu + col] = prm.dt;
      for (int j = 0; j < 3; ++j) fu[(t * 9 + 6 + j) * nu + col] = prm.dt * L.d[9 * sl + 3 * ax + j];
    }
  }
}

int cmpc_wr_linearize(const cmpc_dims* dims, const cmpc_model* model, const double* X, const double* U, const double* contact_pos,
                      const double* contact_R, const int32_t* contact_active, double* f, double* fx, double* fu, cudaStream_t st,
                      std::string* msg) {
  cmpc_wr::Params prm;
  int rc = cmpc_wr::fill_params(&prm, dims, model, nullptr, nullptr, 0);
  if (rc) { *msg = "bad dims or weights"; return rc; }
  const long total = (long)dims->batch * dims->N;
  cmpc_wr_linearize_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(prm, dims->batch, dims->shared_plan, X, U, contact_pos,
                                                                            contact_R, (const int*)contact_active, f, fx, fu);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *msg = cudaGetErrorString(e); return -100 - (int)e; }
  return 0;
}

// LQR gains / covariances along (X, U) for the wrench model (compute_trajectory_data: LQR_gains, Covs;
// /root/reference/src/centroidal_model.py:215-227,233-238,284-285 with the TALOS Jacobians)
int cmpc_wr_lqr_covs(const cmpc_dims* dims, const cmpc_model* model, const cmpc_lqr_weights* w, const double* X, const double* U,
                     const double* contact_pos, const double* contact_R, const int32_t* contact_active, double* gains,
                     double* covs, void* scratch, cudaStream_t st, std::string* msg, long long* n_launches) {
  static_assert(sizeof(cmpc_lqr_weights) == sizeof(cmpc_wr::LqrWeights), "cmpc_lqr_weights layout");
  cmpc_wr::Params prm;
  int rc = cmpc_wr::fill_params(&prm, dims, model, nullptr, nullptr, 0);
  if (rc) { *msg = "bad dims or weights"; return rc; }
  for (int i = 0; i < prm.nu; ++i)
    if (!(w->R[i * prm.nu + i] > 0.0)) { *msg = "R must have a positive diagonal"; return -2; }
  const cudaError_t e = cmpc_wr::launch_lqr_covs(prm, (const cmpc_wr::LqrWeights*)w, dims->batch, dims->shared_plan, X, U, contact_pos,
                                                 contact_R, (const int*)contact_active, gains, covs, scratch, st, n_launches);
  if (e != cudaSuccess) { *msg = cudaGetErrorString(e); return -100 - (int)e; }
  return 0;
}
