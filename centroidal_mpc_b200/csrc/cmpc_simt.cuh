// cmpc_simt.cuh — the warp as a 32-wide vector machine, written once for two builds.
//
// The solver (cmpc_solver.cuh) is warp-per-instance SIMT code: every value is either UNIFORM
// over the warp (plain int/double) or VARYING per lane (vd / vi / vb).  Lanes exchange data with
// register shuffles and a small per-warp shared-memory scratch; control flow branches only on
// uniform values, lane-dependent choices are selects and predicated loads/stores.
//
//   * CUDA build (sm_100a): vd = double, vi = int, vb = bool, one thread per lane; shfl() is
//     __shfl_sync, wsync() is __syncwarp().  This is the product.
//   * Host build (tests/emu only): vd/vi/vb are 32-element arrays and every operation loops over
//     the lanes, so the identical source runs the identical arithmetic on a CPU.  It exists to
//     unit-test the kernel logic on a machine without a GPU and is never part of libcmpc_b200.so.
//
// All floating-point contractions are written explicitly (vfma) and both builds disable
// automatic contraction, so the two builds agree bit for bit except for sqrt/division-free
// library differences (there are none on the hot path).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace cmpc {

#if defined(__CUDA_ARCH__)
// =================================================================================== device
#define CMPC_F __device__ __forceinline__
#define CMPC_FULL 0xffffffffu

typedef double vd;
typedef int vi;
typedef bool vb;

CMPC_F vi lane_id() { return (int)(threadIdx.x & 31u); }
CMPC_F vd vconst(double x) { return x; }
CMPC_F vi viconst(int x) { return x; }
CMPC_F vd shfl(vd v, vi src) { return __shfl_sync(CMPC_FULL, v, src); }
CMPC_F vi shfl(vi v, vi src) { return __shfl_sync(CMPC_FULL, v, src); }
CMPC_F double uni(vd v, int lane) { return __shfl_sync(CMPC_FULL, v, lane); }
CMPC_F int uni(vi v, int lane) { return __shfl_sync(CMPC_FULL, v, lane); }
CMPC_F vd sel(vb c, vd a, vd b) { return c ? a : b; }
CMPC_F vi seli(vb c, vi a, vi b) { return c ? a : b; }
CMPC_F vd ldif(vb c, const double* p, vi i) { return c ? p[i] : 0.0; }
CMPC_F vi ldifi(vb c, const int* p, vi i) { return c ? p[i] : 0; }
CMPC_F void stif(vb c, double* p, vi i, vd v) { if (c) p[i] = v; }
CMPC_F void stifi(vb c, int* p, vi i, vi v) { if (c) p[i] = v; }
CMPC_F vd vfma(vd a, vd b, vd c) { return fma(a, b, c); }
CMPC_F vd vabs(vd a) { return fabs(a); }
CMPC_F vd vmin(vd a, vd b) { return fmin(a, b); }
CMPC_F vd vmax(vd a, vd b) { return fmax(a, b); }
CMPC_F vd vsqrt(vd a) { return sqrt(a); }
CMPC_F vd vfromi(vi a) { return (double)a; }
CMPC_F bool vany(vb c) { return __any_sync(CMPC_FULL, c) != 0; }
CMPC_F unsigned vballot(vb c) { return __ballot_sync(CMPC_FULL, c); }
CMPC_F void wsync() { __syncwarp(); }
CMPC_F double wmax(vd v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(CMPC_FULL, v, o));
  return v;
}
CMPC_F double wmin(vd v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(CMPC_FULL, v, o));
  return v;
}
CMPC_F double wsum(vd v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CMPC_FULL, v, o);
  return v;
}
CMPC_F int wsumi(vi v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CMPC_FULL, v, o);
  return v;
}

#else
// ===================================================================================== host
#define CMPC_F inline
#define CMPC_VL for (int l_ = 0; l_ < 32; ++l_)

struct vb {
  bool v[32];
};
struct vi {
  int v[32];
  vi() {}
  vi(int x) { CMPC_VL v[l_] = x; }
};
struct vd {
  double v[32];
  vd() {}
  vd(double x) { CMPC_VL v[l_] = x; }
};

#define CMPC_BIN(T, R, op)                                                      \
  inline R operator op(const T& a, const T& b) { R r; CMPC_VL r.v[l_] = a.v[l_] op b.v[l_]; return r; }
CMPC_BIN(vd, vd, +) CMPC_BIN(vd, vd, -) CMPC_BIN(vd, vd, *) CMPC_BIN(vd, vd, /)
CMPC_BIN(vd, vb, <) CMPC_BIN(vd, vb, >) CMPC_BIN(vd, vb, <=) CMPC_BIN(vd, vb, >=)
CMPC_BIN(vi, vi, +) CMPC_BIN(vi, vi, -) CMPC_BIN(vi, vi, *) CMPC_BIN(vi, vi, /) CMPC_BIN(vi, vi, %)
CMPC_BIN(vi, vi, &) CMPC_BIN(vi, vi, |) CMPC_BIN(vi, vi, >>) CMPC_BIN(vi, vi, <<)
CMPC_BIN(vi, vb, <) CMPC_BIN(vi, vb, >) CMPC_BIN(vi, vb, <=) CMPC_BIN(vi, vb, >=) CMPC_BIN(vi, vb, ==) CMPC_BIN(vi, vb, !=)
CMPC_BIN(vb, vb, &&) CMPC_BIN(vb, vb, ||)
#undef CMPC_BIN
inline vb operator!(const vb& a) { vb r; CMPC_VL r.v[l_] = !a.v[l_]; return r; }
inline vd operator-(const vd& a) { vd r; CMPC_VL r.v[l_] = -a.v[l_]; return r; }
inline vd& operator+=(vd& a, const vd& b) { CMPC_VL a.v[l_] += b.v[l_]; return a; }
inline vd& operator-=(vd& a, const vd& b) { CMPC_VL a.v[l_] -= b.v[l_]; return a; }
inline vd& operator*=(vd& a, const vd& b) { CMPC_VL a.v[l_] *= b.v[l_]; return a; }
inline vi& operator+=(vi& a, const vi& b) { CMPC_VL a.v[l_] += b.v[l_]; return a; }

CMPC_F vi lane_id() { vi r; CMPC_VL r.v[l_] = l_; return r; }
CMPC_F vd vconst(double x) { return vd(x); }
CMPC_F vi viconst(int x) { return vi(x); }
CMPC_F vd shfl(const vd& v, const vi& src) { vd r; CMPC_VL r.v[l_] = v.v[src.v[l_] & 31]; return r; }
CMPC_F vi shfl(const vi& v, const vi& src) { vi r; CMPC_VL r.v[l_] = v.v[src.v[l_] & 31]; return r; }
CMPC_F double uni(const vd& v, int lane) { return v.v[lane]; }
CMPC_F int uni(const vi& v, int lane) { return v.v[lane]; }
CMPC_F vd sel(const vb& c, const vd& a, const vd& b) { vd r; CMPC_VL r.v[l_] = c.v[l_] ? a.v[l_] : b.v[l_]; return r; }
CMPC_F vi seli(const vb& c, const vi& a, const vi& b) { vi r; CMPC_VL r.v[l_] = c.v[l_] ? a.v[l_] : b.v[l_]; return r; }
CMPC_F vd ldif(const vb& c, const double* p, const vi& i) { vd r; CMPC_VL r.v[l_] = c.v[l_] ? p[i.v[l_]] : 0.0; return r; }
CMPC_F vi ldifi(const vb& c, const int* p, const vi& i) { vi r; CMPC_VL r.v[l_] = c.v[l_] ? p[i.v[l_]] : 0; return r; }
CMPC_F void stif(const vb& c, double* p, const vi& i, const vd& v) { CMPC_VL if (c.v[l_]) p[i.v[l_]] = v.v[l_]; }
CMPC_F void stifi(const vb& c, int* p, const vi& i, const vi& v) { CMPC_VL if (c.v[l_]) p[i.v[l_]] = v.v[l_]; }
CMPC_F vd vfma(const vd& a, const vd& b, const vd& c) { vd r; CMPC_VL r.v[l_] = fma(a.v[l_], b.v[l_], c.v[l_]); return r; }
CMPC_F vd vabs(const vd& a) { vd r; CMPC_VL r.v[l_] = fabs(a.v[l_]); return r; }
CMPC_F vd vmin(const vd& a, const vd& b) { vd r; CMPC_VL r.v[l_] = fmin(a.v[l_], b.v[l_]); return r; }
CMPC_F vd vmax(const vd& a, const vd& b) { vd r; CMPC_VL r.v[l_] = fmax(a.v[l_], b.v[l_]); return r; }
CMPC_F vd vsqrt(const vd& a) { vd r; CMPC_VL r.v[l_] = sqrt(a.v[l_]); return r; }
CMPC_F vd vfromi(const vi& a) { vd r; CMPC_VL r.v[l_] = (double)a.v[l_]; return r; }
CMPC_F bool vany(const vb& c) { bool r = false; CMPC_VL r = r || c.v[l_]; return r; }
CMPC_F unsigned vballot(const vb& c) { unsigned r = 0; CMPC_VL if (c.v[l_]) r |= 1u << l_; return r; }
CMPC_F void wsync() {}
// the butterfly order of the device reductions, so that sums round identically
CMPC_F double wmax(vd v) {
  for (int o = 16; o > 0; o >>= 1) { vd t = v; CMPC_VL v.v[l_] = fmax(t.v[l_], t.v[l_ ^ o]); }
  return v.v[0];
}
CMPC_F double wmin(vd v) {
  for (int o = 16; o > 0; o >>= 1) { vd t = v; CMPC_VL v.v[l_] = fmin(t.v[l_], t.v[l_ ^ o]); }
  return v.v[0];
}
CMPC_F double wsum(vd v) {
  for (int o = 16; o > 0; o >>= 1) { vd t = v; CMPC_VL v.v[l_] = t.v[l_] + t.v[l_ ^ o]; }
  return v.v[0];
}
CMPC_F int wsumi(vi v) {
  for (int o = 16; o > 0; o >>= 1) { vi t = v; CMPC_VL v.v[l_] = t.v[l_] + t.v[l_ ^ o]; }
  return v.v[0];
}
#endif

// mixed uniform/varying helpers common to both builds
CMPC_F vd vsel0(vb c, vd a) { return sel(c, a, vconst(0.0)); }

}  // namespace cmpc
