// micro-benchmarks (one warp): dependent DFMA chain, DADD, LDS->use, STS->syncwarp->LDS, shfl, division
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n) {
  __shared__ double sm[1024];
  int lane = threadIdx.x;
  sm[lane] = lane * 1e-3; sm[lane + 32] = 1.0;
  __syncwarp();
  double a = lane * 1e-9, m = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = fma(a, m, c);
  long long t1 = clock64();
  cyc[0] = (t1 - t0);
  double b0 = a, b1 = a + 1, b2 = a + 2, b3 = a + 3;
  t0 = clock64();
  for (int i = 0; i < n; ++i) { b0 = fma(b0, m, c); b1 = fma(b1, m, c); b2 = fma(b2, m, c); b3 = fma(b3, m, c); }
  t1 = clock64();
  cyc[1] = (t1 - t0);
  a += b0 + b1 + b2 + b3;
  t0 = clock64();
  for (int i = 0; i < n; ++i) a = a + c;
  t1 = clock64();
  cyc[2] = t1 - t0;
  // LDS dependent chain: index from loaded value
  unsigned sa = (unsigned)__cvta_generic_to_shared(sm);
  int idx = lane;
  t0 = clock64();
  for (int i = 0; i < n; ++i) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa + idx * 8)); idx = (idx + (int)v) & 31; }
  t1 = clock64();
  cyc[3] = t1 - t0;
  a += idx;
  // STS -> syncwarp -> LDS (neighbour) -> DADD chain
  t0 = clock64();
  for (int i = 0; i < n; ++i) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(sa + lane * 8), "d"(a) : "memory");
    __syncwarp();
    double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa + ((lane + 1) & 31) * 8) : "memory");
    __syncwarp();
    a = a * 0.5 + v * 0.25;
  }
  t1 = clock64();
  cyc[4] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) a = __shfl_xor_sync(0xffffffffu, a, 4) * 0.5 + 1e-9;
  t1 = clock64();
  cyc[5] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) a = 1.0 / (a + 1.5);
  t1 = clock64();
  cyc[6] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) { __syncwarp(); }
  t1 = clock64();
  cyc[7] = t1 - t0;
  out[lane] = a;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 256); cudaMalloc(&cyc, 64);
  const int n = 4096;
  k<<<1, 32>>>(out, cyc, n); k<<<1, 32>>>(out, cyc, n);
  long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
  const char* nm[8] = {"DFMA dependent", "4 DFMA chains (per iter)", "DADD dependent", "LDS dependent", "STS+sync+LDS+sync+2flop", "SHFL(f64)+DFMA", "1.0/x", "syncwarp"};
  for (int i = 0; i < 8; ++i) printf("%-28s %.1f cycles\n", nm[i], (double)h[i] / n);
  return 0;
}
