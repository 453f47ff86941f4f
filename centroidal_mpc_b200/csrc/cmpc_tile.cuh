// cmpc_tile.cuh — the SCP solver of one MPC instance, written as plain scalar code that runs
// once per LANE: Riccati factorisation, ADMM / multiplier-method sweeps, certified active-set
// polish, trust-region loop.  32 instances (a tile) share a warp; each lane walks its own
// column of the lane-interleaved knot records (cmpc_core.cuh), so every load and store of the
// warp is one coalesced 256-byte access and no lane ever talks to another.  DESIGN.md "device
// algorithm" has the mathematics; oracle/device_model.py is the executable numpy specification.
//
// Control flow: every lane owns a small state machine (advance()) that names the next whole-
// horizon operation it needs (factorise, sweep, build active set, evaluate, ...).  The warp
// executes one operation at a time for the lanes that asked for it (run_tile in cmpc_api.cu),
// so lanes whose QPs need more iterations or polish rounds do not change what the others
// compute.  The host build (tests/emu) runs the same state machine one lane at a time: the
// arithmetic of a lane does not depend on its neighbours, so both builds are bit-identical
// (FMA contraction is explicit, automatic contraction is off in both).
#pragma once
#include "cmpc_core.cuh"

namespace cmpc {

constexpr int MODE_ADMM = 0, MODE_PMM = 1;
// forward-sweep kinds
constexpr int FW_ADMM = 0, FW_ADMM_CHECK = 1, FW_PMM = 2, FW_COPY = 4;

#define CMPC_R(p, f) (p)[(f) * TL]
// Staged data (what a KnotStream hands out) and per-lane scratch live in shared memory on the
// device and are addressed by their 32-bit shared-space byte address through explicit
// ld/st.shared (the compiler cannot see the address space through the ring bookkeeping, and
// generic accesses cost a long-scoreboard round trip); in the host build they are plain pointers.
// A StagedPtr addresses field BASE of the current knot for this lane (BASE = first field the
// operation stages; a constexpr in the scope of every user of CMPC_S).
#if defined(__CUDACC__)
typedef unsigned StagedPtr;
typedef unsigned ScratchPtr;
CMPC_HD double staged_ld(StagedPtr p, int idx) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(p + (unsigned)idx * 8u));
  return v;
}
template <int BASE>
CMPC_HD int staged_meta(StagedPtr p, int lane, int which) {   // which: 0 meta word, 1 active-set word
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(p - (unsigned)lane * 8u + (unsigned)((R_META - BASE) * TL * 8 + (which * TL + lane) * 4)));
  return v;
}
CMPC_HD double sc_ld(ScratchPtr t, int idx) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(t + (unsigned)idx * 8u));
  return v;
}
CMPC_HD void sc_st(ScratchPtr t, int idx, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(t + (unsigned)idx * 8u), "d"(v) : "memory");
}
#else
typedef const double* StagedPtr;
typedef double* ScratchPtr;
CMPC_HD double staged_ld(StagedPtr p, int idx) { return p[idx]; }
template <int BASE>
CMPC_HD int staged_meta(StagedPtr p, int lane, int which) {
  return (reinterpret_cast<const int*>(p - lane + (R_META - BASE) * TL) + lane)[which * TL];
}
CMPC_HD double sc_ld(ScratchPtr t, int idx) { return t[idx]; }
CMPC_HD void sc_st(ScratchPtr t, int idx, double v) { t[idx] = v; }
#endif
#define CMPC_S(p, f) staged_ld(p, ((f) - BASE) * TL)

struct TileCtx {
  const Params* prm;
  double* ws;      // tile workspace [N+1][REC][TL]
  double* gt;      // tile friction table [N][GT][TL], null on the fast path
  int* nst;        // [N+1] slots per knot (tile maximum; 0 at the terminal knot), global memory
#if defined(__CUDACC__)
  // the warp's shared-memory ring (RING_DEPTH slots of R_STAGED fields), its mbarriers and a
  // shared copy of nst; all warp-uniform
  double* ring;
  unsigned ring_sa, bars_sa, phases;
  int tile;
  const unsigned char* nst_s;
  CMPC_HD int ns(int k) const { return nst_s[k]; }
#else
  CMPC_HD int ns(int k) const { return nst[k]; }
#endif
};

struct Inst {
  int b, lane;
  const double *Xr, *Ui, *xi, *xf, *cpos, *cR, *fub;
  const int* cact;
};

CMPC_HD double* rec_of(const TileCtx& T, const Inst& I, int k) {
  return T.ws + (long)k * (REC * TL) + I.lane;
}
CMPC_HD int* meta_of(const TileCtx& T, const Inst& I, int k) {
  return reinterpret_cast<int*>(T.ws + (long)k * (REC * TL) + R_META * TL) + I.lane;   // [0] meta, [TL] active set
}
CMPC_HD double* gt_of(const TileCtx& T, const Inst& I, int k) { return T.gt ? T.gt + (long)k * (GT * TL) + I.lane : nullptr; }

// ---------------------------------------------------------------- knot stream
// Walks the knots of a tile in one direction and hands out a pointer through which the fields
// of the requested segments of the current knot can be read (CMPC_S(r, field)).
// Device: a ring of RING_DEPTH shared-memory slots filled by cp.async.bulk (one elected lane,
// one bulk copy per segment, completion on an mbarrier), so the HBM latency of knot k+3 hides
// behind the arithmetic of knots k..k+2 and every operand read is a shared-memory read.
// Host build: the pointer is the global record itself.
CMPC_HD int seg_start(int seg) {
  return seg == SEG_A ? R_PC : seg == SEG_B ? R_K : seg == SEG_C ? R_META : seg == SEG_D ? R_VK : seg == SEG_E ? R_DV : R_YK;
}
CMPC_HD int seg_len(int seg, int ns) {
  const int na = 3 * ns;
  return seg == SEG_A ? 9 + na * (na + 1) / 2 : seg == SEG_B ? 9 * na : seg == SEG_C ? 16 + na
       : seg == SEG_D ? 3 + 4 * ns : seg == SEG_E ? na : 4 + 4 * ns;
}

#if defined(__CUDACC__)
CMPC_HD unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// commands to the producer warp other than a stream (the descriptor's count field)
constexpr int CMD_EXIT = -1, CMD_SETUP = -2, CMD_WRITE = -3;
// bounded wait on an mbarrier phase: a bulk copy that never lands traps instead of hanging
CMPC_HD void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spins = 0; !done; ++spins) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spins > (1u << 24)) asm volatile("trap;");
  }
}

// solver warp: hand the producer warp the odd knots of a per-knot operation of this tile (mask:
// the lanes that take part), and wait for it to finish
CMPC_HD void helper_fork(const TileCtx& T, int cmd, unsigned mask);
CMPC_HD void helper_join() { asm volatile("bar.sync 2, 64;" ::: "memory"); }

struct KnotStream {
  // copies of the tile's ring description (kept by value so that they live in registers)
  unsigned ring_sa, bars_sa, phases;
  int lane, slot_f, s_wait;

  // knots k_first, k_first + dir, ... (count of them); fields below base_field are not staged.
  // The stream is handed to the CTA's producer warp: descriptor in shared memory, named barrier 1.
  CMPC_HD void open(TileCtx& Tc, const Inst& I, int segments, int base_field, int k_first, int count, int direction) {
    ring_sa = Tc.ring_sa; bars_sa = Tc.bars_sa; phases = Tc.phases;
    lane = I.lane; slot_f = R_STAGED - base_field; s_wait = 0;
    // earlier generic-proxy writes of this warp (records written by the previous operation) must
    // be visible to the async proxy before the bulk copies read them
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    if ((threadIdx.x & 31u) == 0) {   // (lane is the instance's lane, which wraps for tiles narrower than a warp)
      const unsigned c = bars_sa + 32u;
      asm volatile("st.shared.u64 [%0], %1;" ::"r"(c), "l"(Tc.ws) : "memory");
      asm volatile("st.shared.s32 [%0], %1;" ::"r"(c + 8u), "r"(direction) : "memory");
      asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(c + 16u), "r"(segments), "r"(base_field), "r"(k_first), "r"(count) : "memory");
    }
    asm volatile("bar.arrive 1, 64;" ::: "memory");
  }
  CMPC_HD StagedPtr acquire() {
    const unsigned bar = bars_sa + 8u * s_wait;
    mbar_wait(bar, (phases >> s_wait) & 1u);
    phases ^= 1u << s_wait;
    return ring_sa + (unsigned)((s_wait * slot_f) * TL + lane) * 8u;
  }
  CMPC_HD void release() {
    __syncwarp();          // every lane is done reading the slot
    if ((threadIdx.x & 31u) == 0)   // tell the producer warp that the slot is free
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars_sa + 8u * (RING_DEPTH + s_wait)) : "memory");
    s_wait = (s_wait + 1 == RING_DEPTH) ? 0 : s_wait + 1;
  }
  CMPC_HD void close(TileCtx& Tc) { Tc.phases = phases; }   // every issued copy has been consumed
};

// The producer warp of a CTA: waits for a stream descriptor (named barrier 1), then issues the
// bulk copies of the stream's knots as ring slots fall free ("empty" mbarriers, one arrival per
// release of the solver warp), and goes back to waiting.  count < 0 ends it.
CMPC_FN void setup_knots(const Params& P, const TileCtx& T, const Inst& I, int k0, int kstep, double* mq_out, double* mc_out);
CMPC_FN void write_solution_knots(const Params& P, const TileCtx& T, const Inst& I, double* X_out, double* U_out, int k0, int kstep);
CMPC_HD void bind_instance(Inst& I, const Params& P, const Batch& bt, int b);
CMPC_HD void bind_tile(TileCtx& T, const Params& P, const Batch& bt, int tile);

CMPC_HD void producer_warp(const Params& P, const Batch& bt, unsigned ring_sa, unsigned bars_sa, const unsigned char* nst_s) {
  const int lane = (int)(threadIdx.x & 31u);
  unsigned eph = 0xffffffffu;   // parity to wait for per "empty" barrier: a fresh barrier passes parity 1
  for (;;) {
    asm volatile("bar.sync 1, 64;" ::: "memory");
    const unsigned c = bars_sa + 32u;
    const double* ws;
    int segs, base_f, k0, count, dir;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(ws) : "r"(c) : "memory");
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(dir) : "r"(c + 8u) : "memory");
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(segs), "=r"(base_f), "=r"(k0), "=r"(count) : "r"(c + 16u) : "memory");
    if (count == CMD_EXIT) break;
    if (count == CMD_SETUP || count == CMD_WRITE) {   // half of a per-knot operation: the odd knots
      const int tile = segs;
      const unsigned mask = (unsigned)base_f;
      TileCtx T;
      bind_tile(T, P, bt, tile);
      Inst I;
      I.lane = lane;
      if ((mask >> lane) & 1u) {
        bind_instance(I, P, bt, tile * TL + lane);
        if (count == CMD_SETUP) {
          double mq, mc;
          setup_knots(P, T, I, 1, 2, &mq, &mc);
          sc_st(ring_sa + (unsigned)lane * 8u, 0, mq);      // the ring is idle during setup
          sc_st(ring_sa + (unsigned)lane * 8u, TL, mc);
        } else {
          write_solution_knots(P, T, I, bt.X_out, bt.U_out, 1, 2);
        }
      }
      asm volatile("bar.sync 2, 64;" ::: "memory");         // join
      continue;
    }
    const int slot_f = R_STAGED - base_f;
    auto issue = [&](int i, int s) {   // knot number i of the stream into slot s
      mbar_wait(bars_sa + 8u * (RING_DEPTH + s), (eph >> s) & 1u);
      eph ^= 1u << s;
      if (lane == 0) {
        const int k = k0 + i * dir;
        const int ns = nst_s[k];
        const unsigned bar = bars_sa + 8u * s;
        const unsigned dst0 = ring_sa + (unsigned)(s * slot_f) * (TL * 8);
        const double* src0 = ws + (long)k * (REC * TL);
        unsigned bytes = 0;
#pragma unroll
        for (int sg = 1; sg <= SEG_F; sg <<= 1)
          if (segs & sg) bytes += (unsigned)seg_len(sg, ns) * (TL * 8);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#pragma unroll
        for (int sg = 1; sg <= SEG_F; sg <<= 1) {
          if (!(segs & sg)) continue;
          const unsigned len = (unsigned)seg_len(sg, ns) * (TL * 8);
          if (len == 0) continue;
          const int st = seg_start(sg);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst0 + (unsigned)(st - base_f) * (TL * 8)), "l"(src0 + st * TL), "r"(len), "r"(bar) : "memory");
        }
      }
    };
    for (int i = 0; i < RING_DEPTH && i < count; ++i) issue(i, i);
    int s = 0;
    for (int i = 0; i < count; ++i) {
      if (i + RING_DEPTH < count) issue(i + RING_DEPTH, s);
      s = (s + 1 == RING_DEPTH) ? 0 : s + 1;
    }
  }
}
#else
struct KnotStream {
  const double* ws;
  int lane, dir, k, base_f;
  CMPC_HD void open(TileCtx& T, const Inst& I, int, int base_field, int k_first, int, int direction) {
    ws = T.ws; lane = I.lane; dir = direction; k = k_first; base_f = base_field;
  }
  CMPC_HD StagedPtr acquire() const { return ws + (long)k * (REC * TL) + base_f * TL + lane; }
  CMPC_HD void release() { k += dir; }
  CMPC_HD void close(TileCtx&) {}
};
#endif

// Solver scalars of one lane.
struct Sv {
  double rho, rhok, rhoe, rhoep, radius, weight;
  double tau;                      // weight / rhok: threshold of the trust-region prox
  double pri, dua, npri, ndua;     // last residuals
  double nq, dynrow;               // constant parts of the residual norms
  int kap;                         // multiplier method: some knot has trust-region rows
  int fail;                        // a pivot was not positive
  int n_pmm, n_polish;             // statistics: multiplier-method sweeps, polish attempts
  double ye[9];                    // multiplier of x_N = x_final
};

// ---------------------------------------------------------------- friction rows
// G = pyr4 * R^T, pyr4 = [[1,0,-k],[-1,0,-k],[0,1,-k],[0,-1,-k]], k = mu/sqrt2
// (utils.py:9-16, constraints.py:178-184); e2 = row equilibration under D_u = 1/sqrt(W_u).
// Fast path (identity R, same W_u for all contacts): constants; else the per-knot table.
CMPC_HD double pyr4(const Params& P, int row, int a) {
  if (a == 2) return -P.kf;
  if (a == 0) return row == 0 ? 1.0 : (row == 1 ? -1.0 : 0.0);
  return row == 2 ? 1.0 : (row == 3 ? -1.0 : 0.0);
}

template <bool FAST> struct Fric;
template <> struct Fric<true> {
  double kf, ea, eb;
  CMPC_HD void load(const Params& P, const double*, int) { kf = P.kf; ea = P.e2[0]; eb = P.e2[2]; }
  CMPC_HD double e2(int r) const { return r < 2 ? ea : eb; }
  CMPC_HD double ub(int) const { return 0.0; }
  CMPC_HD double G(int r, int a) const {
    if (a == 2) return -kf;
    if (a == 0) return r == 0 ? 1.0 : (r == 1 ? -1.0 : 0.0);
    return r == 2 ? 1.0 : (r == 3 ? -1.0 : 0.0);
  }
  CMPC_HD void rows(const double* u, double* cf) const {          // cf = G u
    cf[0] = fma(-kf, u[2], u[0]);
    cf[1] = fma(-kf, u[2], -u[0]);
    cf[2] = fma(-kf, u[2], u[1]);
    cf[3] = fma(-kf, u[2], -u[1]);
  }
  CMPC_HD void trans(const double* t, double* o) const {          // o = G' t
    o[0] = t[0] - t[1];
    o[1] = t[2] - t[3];
    o[2] = -kf * ((t[0] + t[1]) + (t[2] + t[3]));
  }
};
template <> struct Fric<false> {
  double g[12], e[4], b[4];
  CMPC_HD void load(const Params&, const double* gt, int s) {
#pragma unroll
    for (int i = 0; i < 12; ++i) g[i] = CMPC_R(gt, s * GS + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) { e[i] = CMPC_R(gt, s * GS + 12 + i); b[i] = CMPC_R(gt, s * GS + 16 + i); }
  }
  CMPC_HD double e2(int r) const { return e[r]; }
  CMPC_HD double ub(int r) const { return b[r]; }
  CMPC_HD double G(int r, int a) const { return g[r * 3 + a]; }
  // cf = G u - ub: the rows are carried SHIFTED by their upper bound (stochastic mode: the
  // chance-constraint back-offs, constraints.py:187-214; zero otherwise), so that w = min(v, 0),
  // y ~ max(v, 0), "active <=> cf = 0" and "violated <=> cf > 0" hold unchanged for the shifted values
  CMPC_HD void rows(const double* u, double* cf) const {
#pragma unroll
    for (int r = 0; r < 4; ++r) cf[r] = fma(g[r * 3 + 2], u[2], fma(g[r * 3 + 1], u[1], g[r * 3] * u[0])) - b[r];
  }
  CMPC_HD void trans(const double* t, double* o) const {
#pragma unroll
    for (int a = 0; a < 3; ++a) o[a] = fma(g[9 + a], t[3], fma(g[6 + a], t[2], fma(g[3 + a], t[1], g[a] * t[0])));
  }
};

// ---------------------------------------------------------------- trust-region prox
// argmin_v omega*max(0, |v - kbar|_1 - r) + rho/2 |v - a|^2  (oracle/device_model.py prox_trust)
// branch 0: inside the L1 ball; 1: outside after soft-thresholding; 2: on the surface.
// tau = omega / rho is passed in (no division on the device in the per-knot path: a double
// division is a subroutine call in SASS and spills every live register around it)
CMPC_HD int prox_trust(const double* a, const double* kbar, double r, double tau_in, double* w) {
  double b[3], ab[3];
  double s1 = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) { b[i] = a[i] - kbar[i]; ab[i] = fabs(b[i]); s1 += ab[i]; }
  if (s1 <= r) { for (int i = 0; i < 3; ++i) w[i] = a[i]; return 0; }
  double tau = tau_in, s2 = 0.0, d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { d[i] = fmax(ab[i] - tau, 0.0); s2 += d[i]; }
  if (s2 >= r) {
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = kbar[i] + (b[i] < 0.0 ? -d[i] : d[i]);
    return 1;
  }
  double s[3] = {ab[0], ab[1], ab[2]};   // projection onto the L1 ball: sort descending
  if (s[0] < s[1]) { double t = s[0]; s[0] = s[1]; s[1] = t; }
  if (s[1] < s[2]) { double t = s[1]; s[1] = s[2]; s[2] = t; }
  if (s[0] < s[1]) { double t = s[0]; s[0] = s[1]; s[1] = t; }
  double css = 0.0;
  tau = 0.0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    css += s[j];
    double t = (css - r) * (j == 0 ? 1.0 : (j == 1 ? 0.5 : 1.0 / 3.0));
    if (s[j] - t > 0.0) tau = t;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double di = fmax(ab[i] - tau, 0.0);
    w[i] = kbar[i] + (b[i] < 0.0 ? -di : di);
  }
  return 2;
}
CMPC_HD void prox_kappa(const Sv& S, const double* v, const double* kbar, double* w) {
  prox_trust(v, kbar, S.radius, S.tau, w);
}

// ---------------------------------------------------------------- multiplier-method kappa rows
// Penalty block kM (3x3) and linear term kl (3) of knot k from the active-set word pm and the
// multipliers yk (only reached when S.kap).  Codes per component: 0 pinned to kbar, 1 sign +,
// 2 sign -; branch (bits 16..17): 1 linear penalty, 2 surface row.
CMPC_HD void pmm_kappa_terms(const Params& P, const Sv& S, int pm, const double* kbar, const double* yk, double* M,
                             double* kl) {
#pragma unroll
  for (int i = 0; i < 9; ++i) M[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) kl[i] = 0.0;
  const int br = (pm >> 16) & 3;
  if (br == 0) return;
  const double inv = P.inv_delta;
  double sg[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int code = (pm >> (18 + 2 * i)) & 3;
    sg[i] = code == 1 ? 1.0 : (code == 2 ? -1.0 : 0.0);
    if (code == 0) {   // pinned component: kappa_i = kbar_i
      M[4 * i] += inv;
      kl[i] -= inv * kbar[i] - yk[i];
    }
  }
  if (br == 1) {
#pragma unroll
    for (int i = 0; i < 3; ++i) kl[i] += S.weight * sg[i];
  } else {             // surface: sg'(kappa - kbar) = radius
    double bb = S.radius;
#pragma unroll
    for (int i = 0; i < 3; ++i) bb += sg[i] * kbar[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) M[3 * i + j] += inv * sg[i] * sg[j];
      kl[i] -= sg[i] * (inv * bb - yk[3]);
    }
  }
}

CMPC_HD void set_rho(const Params& P, double rho, double* rho_out, double* rhok_out, double* rhoe, double* rhoep) {
  const double wk = fmin(P.Wx[6], fmin(P.Wx[7], P.Wx[8]));
  double wm = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) wm = fmax(wm, P.Wx[i]);
  *rho_out = rho;
  *rhok_out = rho * P.rho_k_rel * wk;
  if (rhoe) *rhoe = P.rho_e_rel * wm;
  if (rhoep) *rhoep = P.rho_e_pol_rel * wm;
}

// ---------------------------------------------------------------- Riccati factorisation
// Backward over k.  The symmetric tableau  T = [[Huu, Hux],[Hux', Q + A'PA]]  (Huu = R + B'PB,
// Hux = B'PA; lower triangle, u indices first) is swept on its na control pivots (SPD, no
// pivoting):  T -> [[-Huu^-1, .],[Hux' Huu^-1, Q + A'PA - Hux' Huu^-1 Hux]], which delivers Hinv,
// K = -Huu^-1 Hux and P_k in one pass and keeps all three exactly symmetric.  Pc = P c.
// ADMM mode: R = W_u + rho G'E2G, Q = W_x + rho_k I (kappa); multiplier mode: active friction
// rows and kappa rows carry the penalty 1/delta.
// The tableau lives in per-lane scratch tb (stride TS: shared memory on the device, a local
// array in the host build); P (packed lower triangle) stays in registers across the knots.
CMPC_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }
CMPC_HD constexpr int trs(int i, int j) { return i >= j ? tri(i, j) : tri(j, i); }

template <int NS, int MODE, bool FAST, int TS>
CMPC_HD void factor_knot(const Params& P, Sv& S, StagedPtr r, double* w, int lane, const double* gt, int k,
                         double* Pm, ScratchPtr tb, bool on) {
  constexpr int BASE = R_META;
  constexpr int NA = 3 * NS, n = NA + 9;
  const double inv = P.inv_delta;
  const int mt = staged_meta<BASE>(r, lane, 0), nsl = mt & 7;
  const int pm = (MODE == MODE_PMM) ? staged_meta<BASE>(r, lane, 1) : 0;
  const double S3[3] = {CMPC_S(r, R_S), CMPC_S(r, R_S + 1), CMPC_S(r, R_S + 2)};
  const double ck[3] = {CMPC_S(r, R_CK), CMPC_S(r, R_CK + 1), CMPC_S(r, R_CK + 2)};
  // Pc = P c
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    double pc = Pm[trs(i, 5)] * P.dtmg;
    pc = fma(Pm[trs(i, 6)], ck[0], pc);
    pc = fma(Pm[trs(i, 7)], ck[1], pc);
    pc = fma(Pm[trs(i, 8)], ck[2], pc);
    if (on) CMPC_R(w, R_PC + i) = pc;
  }
  // control part of the tableau, one slot at a time:  W = P B (rows 3..8 kept), Hux = W'A, Huu = R + B'W
  double Wm[6][NA > 0 ? NA : 1];   // rows 3..8 of P B
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const double dts = s < nsl ? P.dt : 0.0;
    const double ds[3] = {CMPC_S(r, R_D + 3 * s), CMPC_S(r, R_D + 3 * s + 1), CMPC_S(r, R_D + 3 * s + 2)};
    Fric<FAST> fr;
    fr.load(P, gt, s);
    double rr[4];
#pragma unroll
    for (int row = 0; row < 4; ++row) {
      if (MODE == MODE_ADMM) rr[row] = S.rho * fr.e2(row);
      else rr[row] = ((pm >> (4 * s + row)) & 1) ? inv : 0.0;
    }
    const int cid = (s < nsl) ? ((mt >> (4 + 2 * s)) & 3) : 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int a1 = nxt3(a), a2 = prv3(a), j = 3 * s + a;
      // (P B)[i][(s,a)] = dt_s (P[i][3+a] + P[i][6+a1] d[a2] - P[i][6+a2] d[a1])
      double wc[9];
#pragma unroll
      for (int i = 0; i < 9; ++i)
        wc[i] = dts * fma(Pm[trs(i, 6 + a1)], ds[a2], fma(-Pm[trs(i, 6 + a2)], ds[a1], Pm[trs(i, 3 + a)]));
#pragma unroll
      for (int i = 0; i < 6; ++i) Wm[i][j] = wc[3 + i];
      // Hux[j][q] = (W'A)[j][q]
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int q1 = nxt3(q), q2 = prv3(q);
        sc_st(tb, (tri(NA + q, j)) * TS, fma(P.dt, fma(wc[6 + q1], S3[q2], -(wc[6 + q2] * S3[q1])), wc[q]));
        sc_st(tb, (tri(NA + 3 + q, j)) * TS, fma(P.dt_m, wc[q], wc[3 + q]));
        sc_st(tb, (tri(NA + 6 + q, j)) * TS, wc[6 + q]);
      }
      // Huu[j][l], l <= j:  row (s,a) of B' v = dt_s (v[3+a] + v[6+a1] d[a2] - v[6+a2] d[a1])
#pragma unroll
      for (int l = 0; l <= j; ++l) {
        double v = dts * fma(Wm[3 + a1][l], ds[a2], fma(-Wm[3 + a2][l], ds[a1], Wm[a][l]));
        if (l >= 3 * s) {   // R block of the slot: W_u + G' diag(rr) G
          const int b2 = l - 3 * s;
          double radd = (b2 == a) ? (FAST ? P.Wu[a] : P.Wu[3 * cid + a]) : 0.0;
#pragma unroll
          for (int row = 0; row < 4; ++row) radd = fma(rr[row] * fr.G(row, a), fr.G(row, b2), radd);
          v += radd;
        }
        sc_st(tb, (tri(j, l)) * TS, v);
      }
    }
  }
  // Q + A'(PA), lower triangle.  This state block of the tableau (the future P_k) never goes to
  // the scratch: it is built, swept and handed on in registers.
  double Pn[45];
  {
    double kM[9], kl[3];
    const bool kap = MODE == MODE_PMM && S.kap && k >= 1;
    if (kap) {
      const double kb[3] = {CMPC_S(r, R_XB + 6), CMPC_S(r, R_XB + 7), CMPC_S(r, R_XB + 8)};
      const double yk[4] = {CMPC_S(r, R_YK), CMPC_S(r, R_YK + 1), CMPC_S(r, R_YK + 2), CMPC_S(r, R_YK + 3)};
      pmm_kappa_terms(P, S, pm, kb, yk, kM, kl);
    }
    double PA[9][9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {     // (P [S]x)[i][q] = P[i][6+q1] S[q2] - P[i][6+q2] S[q1]
        const int q1 = nxt3(q), q2 = prv3(q);
        PA[i][q] = fma(P.dt, fma(Pm[trs(i, 6 + q1)], S3[q2], -(Pm[trs(i, 6 + q2)] * S3[q1])), Pm[trs(i, q)]);
        PA[i][3 + q] = fma(P.dt_m, Pm[trs(i, q)], Pm[trs(i, 3 + q)]);
        PA[i][6 + q] = Pm[trs(i, 6 + q)];
      }
    }
#pragma unroll
    for (int rr2 = 0; rr2 < 9; ++rr2) {
      const int g3 = rr2 / 3, a = rr2 - 3 * g3, a1 = nxt3(a), a2 = prv3(a);
#pragma unroll
      for (int c = 0; c <= rr2; ++c) {
        double v = PA[rr2][c];
        if (g3 == 0) v = fma(P.dt, fma(PA[6 + a1][c], S3[a2], -(PA[6 + a2][c] * S3[a1])), v);
        else if (g3 == 1) v = fma(P.dt_m, PA[a][c], v);
        if (c == rr2) {
          v += P.Wx[rr2];
          if (MODE == MODE_ADMM && k >= 1 && rr2 >= 6) v += S.rhok;
        }
        if (MODE == MODE_PMM && c >= 6 && kap) v += kM[3 * (rr2 - 6) + (c - 6)];
        Pn[tri(rr2, c)] = v;
      }
    }
  }
  // sweep the control pivots.  Row / column pv of the generic update is garbage and is
  // overwritten afterwards, so that the update itself has no pv-dependent addressing.
  for (int pv = 0; pv < NA; ++pv) {
    const int tpv = pv * (pv + 1) / 2;
    const double piv = sc_ld(tb, ((tpv + pv)) * TS);
    if (!(piv > 0.0) && on) S.fail = 1;
    const double ip = 1.0 / piv;
    double c[n], bc[n];
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const int idx = i < pv ? tpv + i : i * (i + 1) / 2 + pv;
      c[i] = sc_ld(tb, (idx) * TS);
      bc[i] = c[i] * ip;
    }
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        if (j >= NA) Pn[tri(i - NA, j - NA)] = fma(-bc[i], c[j], Pn[tri(i - NA, j - NA)]);   // state block: registers
        else sc_st(tb, tri(i, j) * TS, fma(-bc[i], c[j], sc_ld(tb, tri(i, j) * TS)));
      }
    }
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const int idx = i < pv ? tpv + i : i * (i + 1) / 2 + pv;
      sc_st(tb, (idx) * TS, (i == pv) ? -ip : bc[i]);
    }
  }
  // factor record and P_k
#pragma unroll
  for (int j = 0; j < NA; ++j) {
#pragma unroll
    for (int l = 0; l <= j; ++l) { const double v = -sc_ld(tb, (tri(j, l)) * TS); if (on) CMPC_R(w, R_HI + tri(j, l)) = v; }
#pragma unroll
    for (int i = 0; i < 9; ++i) { const double v = -sc_ld(tb, (tri(NA + i, j)) * TS); if (on) CMPC_R(w, R_K + 9 * j + i) = v; }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) Pm[tri(i, j)] = Pn[tri(i, j)];
  }
}

template <int MODE, bool FAST>
CMPC_FN void factor_op(const Params& P_in, TileCtx& T, const Inst& I_in, Sv& S_in, bool on) {
  // local copies: the reference arguments live in the caller's frame, which every generic store
  // of the operation could alias (the compiler would reload them after each one)
  const Params P = P_in;
  const Inst I = I_in;
  Sv S = S_in;
  constexpr int BASE = R_META;
  const int N = P.N;
  const double rho_e = (MODE == MODE_ADMM) ? S.rhoe : S.rhoep;
  KnotStream ks;
  ks.open(T, I, SEG_C | (MODE == MODE_PMM ? SEG_F : 0), R_META, N, N + 1, -1);
#if defined(__CUDACC__)
  constexpr int TS = TL;
  const ScratchPtr tb = T.ring_sa + (unsigned)(RING_DEPTH * (R_STAGED - R_META) * TL + I.lane) * 8u;   // behind the stream's slots
#else
  constexpr int TS = 1;
  double tbl[231];
  const ScratchPtr tb = tbl;
#endif
  double Pm[45];
  {
    const StagedPtr r = ks.acquire();
    if (on) {
#pragma unroll
      for (int i = 0; i < 45; ++i) Pm[i] = 0.0;
#pragma unroll
      for (int i = 0; i < 9; ++i) Pm[tri(i, i)] = P.Wx[i] + rho_e;
      if (MODE == MODE_ADMM) {
#pragma unroll
        for (int i = 6; i < 9; ++i) Pm[tri(i, i)] += S.rhok;
      } else if (S.kap) {
        double kM[9], kl[3];
        const double kb[3] = {CMPC_S(r, R_XB + 6), CMPC_S(r, R_XB + 7), CMPC_S(r, R_XB + 8)};
        const double yk[4] = {CMPC_S(r, R_YK), CMPC_S(r, R_YK + 1), CMPC_S(r, R_YK + 2), CMPC_S(r, R_YK + 3)};
        pmm_kappa_terms(P, S, staged_meta<BASE>(r, I.lane, 1), kb, yk, kM, kl);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) Pm[tri(6 + i, 6 + j)] += kM[3 * i + j];
      }
    }
    ks.release();
  }
  // runs of equal slot count: one specialisation of the knot step per inner loop, P in registers
#define CMPC_FAC_RUN(NS_)                                                             \
  do {                                                                                \
    const StagedPtr r = ks.acquire();                                                 \
    if (on) factor_knot<NS_, MODE, FAST, TS>(P, S, r, rec_of(T, I, k), I.lane, gt_of(T, I, k), k, Pm, tb, on); \
    ks.release();                                                                     \
    --k;                                                                              \
  } while (k >= 0 && T.ns(k) == NS_)
  for (int k = N - 1; k >= 0;) {
    switch (T.ns(k)) {
      case 0: CMPC_FAC_RUN(0); break;
      case 1: CMPC_FAC_RUN(1); break;
      case 2: CMPC_FAC_RUN(2); break;
      case 3: CMPC_FAC_RUN(3); break;
      default: CMPC_FAC_RUN(4); break;
    }
  }
#undef CMPC_FAC_RUN
  ks.close(T);
  if (S.fail) S_in.fail = 1;
}

// ---------------------------------------------------------------- backward sweep (linear term)
// p_N = qx_N;  g = p + Pc;  hu = ru + B'g;  d = -Hinv hu;  p = qx + A'g + K'hu.
// qx = -Wx xbar (+ kappa / terminal penalty terms), ru = friction penalty terms.
// r: staged read pointer, w: the knot's record in global memory (writes).
template <int MODE>
CMPC_HD void kappa_linear_term(const Params& P, const Sv& S, StagedPtr r, int pm, double* p) {
  constexpr int BASE = 0;
  const double kb[3] = {CMPC_S(r, R_XB + 6), CMPC_S(r, R_XB + 7), CMPC_S(r, R_XB + 8)};
  if (MODE == MODE_ADMM) {
    const double vk[3] = {CMPC_S(r, R_VK), CMPC_S(r, R_VK + 1), CMPC_S(r, R_VK + 2)};
    double w[3];
    prox_kappa(S, vk, kb, w);
#pragma unroll
    for (int a = 0; a < 3; ++a) p[6 + a] += -S.rhok * (w[a] + w[a] - vk[a]);
  } else if (S.kap) {
    const double yk[4] = {CMPC_S(r, R_YK), CMPC_S(r, R_YK + 1), CMPC_S(r, R_YK + 2), CMPC_S(r, R_YK + 3)};
    double kM[9], kl[3];
    pmm_kappa_terms(P, S, pm, kb, yk, kM, kl);
#pragma unroll
    for (int a = 0; a < 3; ++a) p[6 + a] += kl[a];
  }
}

template <int NS, int MODE, bool FAST>
CMPC_HD void bwd_knot(const Params& P, const Sv& S, StagedPtr r, double* w, int lane, const double* gt, int k, double* p) {
  constexpr int BASE = 0;
  constexpr int NA = 3 * NS;
  const int nsl = staged_meta<BASE>(r, lane, 0) & 7;
  const int pm = (MODE == MODE_PMM) ? staged_meta<BASE>(r, lane, 1) : 0;
  double g[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) g[i] = p[i] + CMPC_S(r, R_PC + i);
  double hu[NA > 0 ? NA : 1];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const double dts = s < nsl ? P.dt : 0.0;
    const double ds[3] = {CMPC_S(r, R_D + 3 * s), CMPC_S(r, R_D + 3 * s + 1), CMPC_S(r, R_D + 3 * s + 2)};
    Fric<FAST> fr;
    fr.load(P, gt, s);
    double t[4], o[3];
#pragma unroll
    for (int row = 0; row < 4; ++row) {
      if (MODE == MODE_ADMM) {
        t[row] = S.rho * fr.e2(row) * fabs(CMPC_S(r, R_VF + 4 * s + row));   // -(rho e2 w - y) = rho e2 |v|
        if (!FAST) t[row] = fma(-S.rho * fr.e2(row), fr.ub(row), t[row]);     // unshifted w = min(v, 0) + ub
      } else {
        const double y = CMPC_S(r, R_YF + 4 * s + row);
        t[row] = ((pm >> (4 * s + row)) & 1) ? (FAST ? y : fma(-P.inv_delta, fr.ub(row), y)) : 0.0;
      }
    }
    fr.trans(t, o);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int a1 = nxt3(a), a2 = prv3(a);
      hu[3 * s + a] = dts * fma(g[6 + a1], ds[a2], fma(-g[6 + a2], ds[a1], g[3 + a])) + o[a];
    }
  }
  // d = -Hinv hu (packed symmetric), K'hu
  double acc[NA > 0 ? NA : 1], kh[9];
#pragma unroll
  for (int j = 0; j < NA; ++j) acc[j] = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) kh[i] = 0.0;
#pragma unroll
  for (int j = 0; j < NA; ++j) {
#pragma unroll
    for (int l = 0; l <= j; ++l) {
      const double h = CMPC_S(r, R_HI + j * (j + 1) / 2 + l);
      acc[j] = fma(h, hu[l], acc[j]);
      if (l < j) acc[l] = fma(h, hu[j], acc[l]);
    }
  }
#pragma unroll
  for (int j = 0; j < NA; ++j) CMPC_R(w, R_DV + j) = -acc[j];
#pragma unroll
  for (int j = 0; j < NA; ++j) {
#pragma unroll
    for (int i = 0; i < 9; ++i) kh[i] = fma(CMPC_S(r, R_K + 9 * j + i), hu[j], kh[i]);
  }
  // p = qx + A'g + K'hu
  const double S3[3] = {CMPC_S(r, R_S), CMPC_S(r, R_S + 1), CMPC_S(r, R_S + 2)};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int a1 = nxt3(a), a2 = prv3(a);
    p[a] = fma(P.dt, fma(g[6 + a1], S3[a2], -(g[6 + a2] * S3[a1])), g[a]) + kh[a] - P.Wx[a] * CMPC_S(r, R_XB + a);
    p[3 + a] = fma(P.dt_m, g[a], g[3 + a]) + kh[3 + a] - P.Wx[3 + a] * CMPC_S(r, R_XB + 3 + a);
    p[6 + a] = g[6 + a] + kh[6 + a] - P.Wx[6 + a] * CMPC_S(r, R_XB + 6 + a);
  }
  if (k >= 1) kappa_linear_term<MODE>(P, S, r, pm, p);
}

template <int MODE, bool FAST>
CMPC_FN void backward_op(const Params& P_in, TileCtx& T, const Inst& I_in, const Sv& S_in, bool on) {
  const Params P = P_in;
  const Inst I = I_in;
  const Sv S = S_in;
  constexpr int BASE = 0;
  const int N = P.N;
  const double rho_e = (MODE == MODE_ADMM) ? S.rhoe : S.rhoep;
  KnotStream ks;
  ks.open(T, I, SEG_A | SEG_B | SEG_C | (MODE == MODE_ADMM ? SEG_D : SEG_F), 0, N, N + 1, -1);
  double p[9];
  {
    const StagedPtr r = ks.acquire();
    if (on) {
#pragma unroll
      for (int i = 0; i < 9; ++i) p[i] = -(P.Wx[i] * CMPC_S(r, R_XB + i)) - (rho_e * I.xf[i] - S.ye[i]);
      kappa_linear_term<MODE>(P, S, r, staged_meta<BASE>(r, I.lane, 1), p);
    }
    ks.release();
  }
  // the knots are walked in runs of equal slot count, so that the inner loop is one
  // specialisation of the knot step and the carried vector p stays in registers
#define CMPC_BWD_RUN(NS_)                                                             \
  do {                                                                                \
    const StagedPtr r = ks.acquire();                                                 \
    if (on) bwd_knot<NS_, MODE, FAST>(P, S, r, rec_of(T, I, k), I.lane, gt_of(T, I, k), k, p); \
    ks.release();                                                                     \
    --k;                                                                              \
  } while (k >= 0 && T.ns(k) == NS_)
  for (int k = N - 1; k >= 0;) {
    switch (T.ns(k)) {
      case 0: CMPC_BWD_RUN(0); break;
      case 1: CMPC_BWD_RUN(1); break;
      case 2: CMPC_BWD_RUN(2); break;
      case 3: CMPC_BWD_RUN(3); break;
      default: CMPC_BWD_RUN(4); break;
    }
  }
#undef CMPC_BWD_RUN
  ks.close(T);
}

// ---------------------------------------------------------------- forward sweep + local updates
// u~ = K x~ + d,  x~+ = A x~ + B u~ + c.
// ADMM kinds: relaxation, friction / kappa / terminal updates of (w, y) stored as v, all
// knot-local; with CHECK also OSQP's residuals (oracle/device_model.py iterate()).
// Multiplier kind: y += (1/delta) row on the active rows, solution record, primal residual,
// and (upd) the corrected friction active set: violated rows join, rows with a negative
// multiplier leave; the number of changes is accumulated in nchg.
// COPY: read-only LQR roll-out of the ADMM iterate into the solution record.
struct Res { double pri, dua, npri, ndua; };

template <int KIND>
CMPC_HD void fwd_state(const Params& P, Sv& S, Res& R, StagedPtr r, double* w, const Inst& I, int k, const double* x) {
  constexpr int BASE = R_K;
  constexpr bool ADMM = KIND == FW_ADMM || KIND == FW_ADMM_CHECK, CHK = KIND == FW_ADMM_CHECK;
  constexpr bool PMMK = KIND == FW_PMM, COPY = KIND == FW_COPY;
  const int N = P.N;
  const double al = ADMM ? P.alpha : 1.0;
  if (PMMK || COPY) {
#pragma unroll
    for (int i = 0; i < 9; ++i) CMPC_R(w, R_X + i) = x[i];
  }
  double rdx[3] = {0.0, 0.0, 0.0};
  if (k >= 1) {
    if (ADMM) {
      const double kb[3] = {CMPC_S(r, R_XB + 6), CMPC_S(r, R_XB + 7), CMPC_S(r, R_XB + 8)};
      const double vk[3] = {CMPC_S(r, R_VK), CMPC_S(r, R_VK + 1), CMPC_S(r, R_VK + 2)};
      double wk[3], vn[3];
      prox_kappa(S, vk, kb, wk);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        vn[a] = fma(al, x[6 + a], fma(1.0 - al, wk[a], vk[a] - wk[a]));
        CMPC_R(w, R_VK + a) = vn[a];
      }
      if (CHK) {
        double wn[3];
        prox_kappa(S, vn, kb, wn);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          R.pri = fmax(R.pri, fabs(x[6 + a] - wn[a]));
          R.npri = fmax(R.npri, fmax(fabs(x[6 + a]), fabs(wn[a])));
          rdx[a] = S.rhok * ((vn[a] - wn[a]) - (vk[a] - wk[a]) - (x[6 + a] - wk[a]));
        }
      }
    } else if (PMMK) {
      if (S.kap) {
        const int pm = staged_meta<BASE>(r, I.lane, 1);
        const int br = (pm >> 16) & 3;
        if (br != 0) {
          const double inv = P.inv_delta;
          double accv = -S.radius;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int code = (pm >> (18 + 2 * i)) & 3;
            const double sgn = code == 1 ? 1.0 : (code == 2 ? -1.0 : 0.0);
            const double dk = x[6 + i] - CMPC_S(r, R_XB + 6 + i);
            if (code == 0) CMPC_R(w, R_YK + i) = CMPC_S(r, R_YK + i) + inv * dk;
            accv += sgn * dk;
          }
          if (br == 2) CMPC_R(w, R_YK + 3) = CMPC_S(r, R_YK + 3) + inv * accv;
        }
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) R.npri = fmax(R.npri, fabs(x[6 + a]));
    }
  }
  if (k == N && !COPY) {   // terminal equality
    const double re = ADMM ? S.rhoe : S.rhoep;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double dx = x[i] - I.xf[i];
      S.ye[i] = S.ye[i] + (re * al) * dx;
      if (CHK || PMMK) {
        R.pri = fmax(R.pri, fabs(dx));
        R.npri = fmax(R.npri, fmax(fabs(x[i]), fabs(I.xf[i])));
        // no stationarity term: the certificate uses y_e + rho_e (x_N - x_f) for these equality
        // rows, the multiplier of the x-update itself (any value is admissible on an equality)
      }
    }
  }
  if (CHK && k >= 1) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double Px = P.Wx[i] * x[i];
      const double rd = i >= 6 ? rdx[i - 6] : 0.0;
      const double aty = rd - Px + P.Wx[i] * CMPC_S(r, R_XB + i);   // (A'y)_x = r_d - P x - q,  q = -Wx xbar
      R.dua = fmax(R.dua, fabs(rd));
      R.ndua = fmax(R.ndua, fmax(fabs(Px), fabs(aty)));
    }
  }
}

template <int NS, int KIND, bool FAST>
CMPC_HD void fwd_knot(const Params& P, const Sv& S, Res& R, StagedPtr r, double* w, int* imw, int lane,
                      const double* gt, double* x, bool upd, int& nchg) {
  constexpr int BASE = R_K;
  constexpr int NA = 3 * NS;
  constexpr bool ADMM = KIND == FW_ADMM || KIND == FW_ADMM_CHECK, CHK = KIND == FW_ADMM_CHECK;
  constexpr bool PMMK = KIND == FW_PMM, COPY = KIND == FW_COPY;
  const double al = ADMM ? P.alpha : 1.0;
  const double inv = P.inv_delta;
  const double tolc = P.as_tol * (1.0 + S.npri);   // row-violation threshold (npri of the previous sweep)
  // controls u~ = K x + d
  double u[NA > 0 ? NA : 1];
#pragma unroll
  for (int j = 0; j < NA; ++j) {
    double t = CMPC_S(r, R_DV + j);
#pragma unroll
    for (int i = 0; i < 9; ++i) t = fma(CMPC_S(r, R_K + 9 * j + i), x[i], t);
    u[j] = t;
  }
  if (PMMK || COPY) {
#pragma unroll
    for (int j = 0; j < NA; ++j) CMPC_R(w, R_U + j) = u[j];
  }
  // next state
  double sF[3] = {0.0, 0.0, 0.0}, sT[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const double ds[3] = {CMPC_S(r, R_D + 3 * s), CMPC_S(r, R_D + 3 * s + 1), CMPC_S(r, R_D + 3 * s + 2)};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int a1 = nxt3(a), a2 = prv3(a);
      sF[a] = sF[a] + u[3 * s + a];
      sT[a] = sT[a] + fma(ds[a1], u[3 * s + a2], -(ds[a2] * u[3 * s + a1]));
    }
  }
  double xn[9];
  {
    const double S3[3] = {CMPC_S(r, R_S), CMPC_S(r, R_S + 1), CMPC_S(r, R_S + 2)};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int a1 = nxt3(a), a2 = prv3(a);
      xn[a] = fma(P.dt_m, x[3 + a], x[a]);
      xn[3 + a] = x[3 + a] + fma(P.dt, sF[a], a == 2 ? P.dtmg : 0.0);
      xn[6 + a] = fma(P.dt * S3[a1], x[a2], fma(-P.dt * S3[a2], x[a1], x[6 + a])) + fma(P.dt, sT[a], CMPC_S(r, R_CK + a));
    }
  }
  // friction rows of the knot
  if (!COPY) {
    const int mt = staged_meta<BASE>(r, lane, 0);
    const int pm = PMMK ? staged_meta<BASE>(r, lane, 1) : 0;
    int newpm = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      Fric<FAST> fr;
      fr.load(P, gt, s);
      double cf[4];
      fr.rows(u + 3 * s, cf);
      if (ADMM) {
        double dl[4];
#pragma unroll
        for (int row = 0; row < 4; ++row) {
          const double v = CMPC_S(r, R_VF + 4 * s + row);
          const double w0 = fmin(v, 0.0), y0 = fmax(v, 0.0);
          const double vn = fma(al, cf[row], fma(1.0 - al, w0, y0));
          CMPC_R(w, R_VF + 4 * s + row) = vn;
          if (CHK) {
            const double wn = fmin(vn, 0.0);
            R.pri = fmax(R.pri, fabs(cf[row] - wn));
            R.npri = fmax(R.npri, FAST ? fmax(fabs(cf[row]), fabs(wn))
                                       : fmax(fabs(cf[row] + fr.ub(row)), fabs(wn + fr.ub(row))));
            dl[row] = S.rho * fr.e2(row) * (fmax(vn, 0.0) - y0 - cf[row] + w0);
          }
        }
        if (CHK) {   // u rows of the stationarity residual: G' delta; norms of P u and A'y
          double rdu[3];
          fr.trans(dl, rdu);
          const int cid = (s < (mt & 7)) ? ((mt >> (4 + 2 * s)) & 3) : 0;
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const double Pu = (FAST ? P.Wu[a] : P.Wu[3 * cid + a]) * u[3 * s + a];
            R.dua = fmax(R.dua, fabs(rdu[a]));
            R.ndua = fmax(R.ndua, fmax(fabs(Pu), fabs(rdu[a] - Pu)));
          }
        }
      } else {
#pragma unroll
        for (int row = 0; row < 4; ++row) {
          const int bit = 4 * s + row;
          const bool on = (pm >> bit) & 1;
          const double yn = fma(inv, cf[row], on ? CMPC_S(r, R_YF + bit) : 0.0);
          R.pri = fmax(R.pri, on ? fabs(cf[row]) : fmax(cf[row], 0.0));
          R.npri = fmax(R.npri, FAST ? fabs(cf[row]) : fabs(cf[row] + fr.ub(row)));
          if (upd) {
            const bool keep = on && !(yn < 0.0);
            const bool join = !on && (cf[row] > tolc);
            const bool nb = keep || join;
            nchg += (nb != on) ? 1 : 0;
            CMPC_R(w, R_YF + bit) = keep ? yn : 0.0;
            newpm |= (nb ? 1 : 0) << bit;
          } else if (on) {
            CMPC_R(w, R_YF + bit) = yn;
          }
        }
      }
    }
    if (PMMK && upd) imw[TL] = (pm & ~0xffff) | newpm;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) x[i] = xn[i];
}

// commit: the lane wants the residuals of this sweep (a lane that did not ask for a check may
// ride along in the CHECK kind when a neighbour did; its iterate update is the same arithmetic)
template <int KIND, bool FAST>
CMPC_FN void forward_op(const Params& P_in, TileCtx& T, const Inst& I_in, Sv& S_in, bool on, bool commit, bool upd, int* changes) {
  const Params P = P_in;
  const Inst I = I_in;
  Sv S = S_in;
  constexpr bool CHK = KIND == FW_ADMM_CHECK, PMMK = KIND == FW_PMM;
  const int N = P.N;
  KnotStream ks;
  ks.open(T, I, SEG_B | SEG_C | SEG_E | (KIND == FW_PMM ? SEG_F : (KIND == FW_COPY ? 0 : SEG_D)), R_K, 0, N + 1, 1);
  Res R;
  R.pri = R.dua = R.npri = R.ndua = 0.0;
  int nchg = 0;
  double x[9];
  if (on) {
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = I.xi[i];
  }
  // runs of equal slot count: one specialisation of the knot step per inner loop, x in registers
#define CMPC_FWD_RUN(NS_)                                                             \
  do {                                                                                \
    const StagedPtr r = ks.acquire();                                                 \
    if (on) {                                                                         \
      double* w = rec_of(T, I, k);                                                    \
      fwd_state<KIND>(P, S, R, r, w, I, k, x);                                        \
      fwd_knot<NS_, KIND, FAST>(P, S, R, r, w, meta_of(T, I, k), I.lane, gt_of(T, I, k), x, upd, nchg); \
    }                                                                                 \
    ks.release();                                                                     \
    ++k;                                                                              \
  } while (k < N && T.ns(k) == NS_)
  for (int k = 0; k < N;) {
    switch (T.ns(k)) {
      case 0: CMPC_FWD_RUN(0); break;
      case 1: CMPC_FWD_RUN(1); break;
      case 2: CMPC_FWD_RUN(2); break;
      case 3: CMPC_FWD_RUN(3); break;
      default: CMPC_FWD_RUN(4); break;
    }
  }
#undef CMPC_FWD_RUN
  {   // terminal knot
    const StagedPtr r = ks.acquire();
    if (on) fwd_state<KIND>(P, S, R, r, rec_of(T, I, N), I, N, x);
    ks.release();
  }
  ks.close(T);
  if (on && commit && (CHK || PMMK)) {
    S.pri = R.pri;
    S.npri = fmax(R.npri, S.dynrow);
    if (CHK) {
      S.dua = R.dua;
      S.ndua = fmax(R.ndua, S.nq);
    }
  }
  if (on && changes) *changes = nchg;
  if (on) S_in = S;
}

// ---------------------------------------------------------------- rho change: keep (w, y), move v
CMPC_FN void rescale_op(const Params& P, const TileCtx& T, const Inst& I, const Sv& S, double rho_new, double rhok_new) {
  const double ratio = S.rho / rho_new;
  for (int k = 0; k <= P.N; ++k) {
    double* r = rec_of(T, I, k);
    if (k < P.N) {
      const int nr = 4 * T.ns(k);
      for (int j = 0; j < nr; ++j) {
        const double v = CMPC_R(r, R_VF + j);
        CMPC_R(r, R_VF + j) = fma(ratio, fmax(v, 0.0), fmin(v, 0.0));
      }
    }
    if (k >= 1) {
      const double kb[3] = {CMPC_R(r, R_XB + 6), CMPC_R(r, R_XB + 7), CMPC_R(r, R_XB + 8)};
      const double vk[3] = {CMPC_R(r, R_VK), CMPC_R(r, R_VK + 1), CMPC_R(r, R_VK + 2)};
      double w[3];
      prox_kappa(S, vk, kb, w);
#pragma unroll
      for (int a = 0; a < 3; ++a) CMPC_R(r, R_VK + a) = fma(S.rhok / rhok_new, vk[a] - w[a], w[a]);
    }
  }
}

// ---------------------------------------------------------------- active set of the polish
// Friction row active iff its multiplier is positive (OSQP's rule -w < y <=> v > 0); trust-
// region rows by the branch the prox took.  Sets *kap when some knot has trust-region rows.
template <bool FAST>
CMPC_FN void build_active_set_op(const Params& P_in, TileCtx& T, const Inst& I_in, const Sv& S_in, bool on, int* kap_out) {
  const Params P = P_in;
  const Inst I = I_in;
  const Sv S = S_in;
  constexpr int BASE = R_META;
  const int N = P.N;
  int kap = 0;
  KnotStream ks;
  ks.open(T, I, SEG_C | SEG_D, R_META, 0, N + 1, 1);
  for (int k = 0; k <= N; ++k) {
    const StagedPtr r = ks.acquire();
    if (on) {
      double* w = rec_of(T, I, k);
      int pm = 0;
      if (k < N) {
        const int ns = T.ns(k);
        const double* gt = gt_of(T, I, k);
        for (int s = 0; s < ns; ++s) {
          Fric<FAST> fr;
          fr.load(P, gt, s);
#pragma unroll
          for (int row = 0; row < 4; ++row) {
            const double v = CMPC_S(r, R_VF + 4 * s + row);
            const bool act = v > 0.0;
            if (act) pm |= 1 << (4 * s + row);
            CMPC_R(w, R_YF + 4 * s + row) = act ? S.rho * fr.e2(row) * v : 0.0;
          }
        }
      }
      if (k >= 1) {
        const double kb[3] = {CMPC_S(r, R_XB + 6), CMPC_S(r, R_XB + 7), CMPC_S(r, R_XB + 8)};
        const double a3[3] = {CMPC_S(r, R_VK), CMPC_S(r, R_VK + 1), CMPC_S(r, R_VK + 2)};
        double wk[3], yk4[4] = {0.0, 0.0, 0.0, 0.0};
        const int br = prox_trust(a3, kb, S.radius, S.tau, wk);
        if (br != 0) {
          kap = 1;
          pm |= br << 16;
          double msum = 0.0;
          int nz = 0;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const double d = wk[i] - kb[i];
            const double yk = S.rhok * (a3[i] - wk[i]);
            const int code = d > 0.0 ? 1 : (d < 0.0 ? 2 : 0);
            pm |= code << (18 + 2 * i);
            if (code == 0) yk4[i] = yk;
            else { msum += (code == 1 ? yk : -yk); ++nz; }
          }
          if (br == 2) yk4[3] = nz ? msum / nz : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) CMPC_R(w, R_YK + i) = yk4[i];
      }
      meta_of(T, I, k)[TL] = pm;
    }
    ks.release();
  }
  ks.close(T);
  if (on) *kap_out = kap;
}

// ---------------------------------------------------------------- trust test and accuracy ratio
// sigma_max(X - Xbar) via the 9x9 Gram matrix + cyclic Jacobi (scp_solver.py:151: np.linalg.norm(.,2));
// rho = sum ||(f(x,u) - lin)[6:9]||^2 / sum ||lin||^2 (scp_solver.py:71-87).
CMPC_FN void evaluate_op(const Params& P, const TileCtx& T, const Inst& I, double* snorm, double* num_out, double* den_out) {
  const int N = P.N;
  double A[81];
  for (int i = 0; i < 81; ++i) A[i] = 0.0;
  double num = 0.0, den = 0.0;
  for (int k = 0; k <= N; ++k) {
    const double* r = rec_of(T, I, k);
    double x[9], dx[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { x[i] = CMPC_R(r, R_X + i); dx[i] = x[i] - I.Xr[k * 9 + i]; }
#pragma unroll
    for (int i = 0; i < 9; ++i)
      for (int j = i; j < 9; ++j) A[i * 9 + j] = fma(dx[i], dx[j], A[i * 9 + j]);
    if (k == N) break;
    const int mt = meta_of(T, I, k)[0];
    const int ns = mt & 7;
    double u[MAXU];
#pragma unroll
    for (int i = 0; i < MAXU; ++i) u[i] = 0.0;
    double F[3] = {0, 0, 0}, Tq[3] = {0, 0, 0};
    for (int sl = 0; sl < ns; ++sl) {
      const int cid = (mt >> (4 + 2 * sl)) & 3;
      const double ds[3] = {CMPC_R(r, R_D + 3 * sl), CMPC_R(r, R_D + 3 * sl + 1), CMPC_R(r, R_D + 3 * sl + 2)};
      const double us[3] = {CMPC_R(r, R_U + 3 * sl), CMPC_R(r, R_U + 3 * sl + 1), CMPC_R(r, R_U + 3 * sl + 2)};
      double t[3];
      cross3(ds, us, t);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        u[3 * cid + a] = us[a];
        F[a] += us[a];
        Tq[a] += t[a];
      }
    }
    // lin = A x + B u + c with the structured A, B
    const double S3[3] = {CMPC_R(r, R_S), CMPC_R(r, R_S + 1), CMPC_R(r, R_S + 2)};
    double lin[9], nl[9], Sxc[3];
    cross3(S3, x, Sxc);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lin[a] = x[a] + P.dt_m * x[3 + a];
      lin[3 + a] = x[3 + a] + P.dt * F[a] + (a == 2 ? P.dtmg : 0.0);
      lin[6 + a] = x[6 + a] + P.dt * Sxc[a] + P.dt * Tq[a] + CMPC_R(r, R_CK + a);
    }
    step_knot(P, x, u, I.cpos + (long)k * P.nc * 3, I.cact + (long)k * P.nc, nl);
#pragma unroll
    for (int i = 6; i < 9; ++i) num += (nl[i] - lin[i]) * (nl[i] - lin[i]);
#pragma unroll
    for (int i = 0; i < 9; ++i) den += lin[i] * lin[i];
  }
  *num_out = num;
  *den_out = den;
#pragma unroll
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < i; ++j) A[i * 9 + j] = A[j * 9 + i];
  // largest eigenvalue of the Gram matrix: cyclic Jacobi
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < 9; ++i)
      for (int j = i + 1; j < 9; ++j) off += A[i * 9 + j] * A[i * 9 + j];
    double dg = 0.0;
    for (int i = 0; i < 9; ++i) dg += A[i * 9 + i] * A[i * 9 + i];
    if (off <= 1e-30 * dg || off == 0.0) break;
    for (int p = 0; p < 8; ++p) {
      for (int q = p + 1; q < 9; ++q) {
        double apq = A[p * 9 + q];
        if (apq == 0.0) continue;
        double th = (A[q * 9 + q] - A[p * 9 + p]) / (2.0 * apq);
        double t = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
        double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        for (int rr = 0; rr < 9; ++rr) {
          double arp = A[rr * 9 + p], arq = A[rr * 9 + q];
          A[rr * 9 + p] = cs * arp - sn * arq;
          A[rr * 9 + q] = sn * arp + cs * arq;
        }
        for (int rr = 0; rr < 9; ++rr) {
          double apr = A[p * 9 + rr], aqr = A[q * 9 + rr];
          A[p * 9 + rr] = cs * apr - sn * aqr;
          A[q * 9 + rr] = sn * apr + cs * aqr;
        }
      }
    }
  }
  double mx = 0.0;
  for (int i = 0; i < 9; ++i) mx = fmax(mx, A[i * 9 + i]);
  *snorm = sqrt(mx);
}

// ---------------------------------------------------------------- per-instance setup
// K1 for every knot, friction table when not on the fast path, start of the iterate at the
// linearisation point, constant parts of the residual norms.
// setup_knots handles the knots k0, k0 + kstep, ... and returns the partial maxima through mq / mc
// (device: the solver warp takes the even knots, the producer warp the odd ones).
CMPC_FN void setup_knots(const Params& P, const TileCtx& T, const Inst& I, int k0, int kstep, double* mq_out, double* mc_out) {
  const int N = P.N;
  double mq = 0.0, mc = 0.0;
  for (int k = k0; k <= N; k += kstep) {
    double* r = rec_of(T, I, k);
    int* im = meta_of(T, I, k);
    const int kk = k < N ? k : N - 1;
    const double* xb = I.Xr + k * 9;
    KnotLin L;
    linearize_knot(P, xb, I.Ui + kk * P.nu, I.cpos + (long)kk * P.nc * 3, I.cact + (long)kk * P.nc, k == N, L);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      CMPC_R(r, R_XB + i) = xb[i];
      mq = fmax(mq, fabs(P.Wx[i] * xb[i]));
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { CMPC_R(r, R_S + a) = L.S[a]; CMPC_R(r, R_CK + a) = L.ck[a]; }
#pragma unroll
    for (int j = 0; j < MAXU; ++j) { CMPC_R(r, R_D + j) = L.d[j]; CMPC_R(r, R_DV + j) = 0.0; CMPC_R(r, R_U + j) = 0.0; }
#pragma unroll
    for (int j = 0; j < 16; ++j) { CMPC_R(r, R_VF + j) = 0.0; CMPC_R(r, R_YF + j) = 0.0; }
#pragma unroll
    for (int j = 0; j < 4; ++j) CMPC_R(r, R_YK + j) = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) CMPC_R(r, R_VK + a) = xb[6 + a];
    im[0] = L.meta;
    im[TL] = 0;
    if (k < N) {
      mc = fmax(mc, fabs(P.dtmg));
#pragma unroll
      for (int a = 0; a < 3; ++a) mc = fmax(mc, fabs(L.ck[a]));
      const int ns = L.meta & 7;
      double* gt = gt_of(T, I, k);
#pragma unroll
      for (int sl = 0; sl < MAXC; ++sl) {
        if (sl >= ns && !gt) continue;   // padded slot on the fast path: nothing to write
        const int cid = sl < ns ? ((L.meta >> (4 + 2 * sl)) & 3) : 0;
        double G[12];
#pragma unroll
        for (int row = 0; row < 4; ++row) {
          double mx = 0.0;
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            double g = pyr4(P, row, a);
            if (sl < ns && I.cR) {
              const double* Rm = I.cR + ((long)k * P.nc + cid) * 9;
              g = 0.0;
#pragma unroll
              for (int b2 = 0; b2 < 3; ++b2) g += pyr4(P, row, b2) * Rm[a * 3 + b2];
            }
            G[row * 3 + a] = g;
            if (gt) mx = fmax(mx, fabs(g) / sqrt(P.Wu[3 * cid + a]));
          }
          if (gt) {
#pragma unroll
            for (int a = 0; a < 3; ++a) CMPC_R(gt, sl * GS + row * 3 + a) = G[row * 3 + a];
            CMPC_R(gt, sl * GS + 12 + row) = mx > 0.0 ? 1.0 / (mx * mx) : 0.0;
            CMPC_R(gt, sl * GS + 16 + row) = (sl < ns && I.fub) ? I.fub[((long)k * P.nc + cid) * 4 + row] : 0.0;
          }
        }
        if (sl < ns) {
          const double* ub = I.Ui + k * P.nu + 3 * cid;
#pragma unroll
          for (int row = 0; row < 4; ++row) {
            double cf = 0.0;
#pragma unroll
            for (int a = 0; a < 3; ++a) cf += G[row * 3 + a] * ub[a];
            if (I.fub) cf -= I.fub[((long)k * P.nc + cid) * 4 + row];
            CMPC_R(r, R_VF + 4 * sl + row) = fmin(cf, 0.0);
          }
        }
      }
    }
  }
  *mq_out = mq;
  *mc_out = mc;
}
CMPC_FN void setup_finish(const Inst& I, Sv& S, double mq, double mc) {
  S.nq = mq;
  double mi = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) mi = fmax(mi, fabs(I.xi[i]));
  S.dynrow = fmax(mc, mi);
#pragma unroll
  for (int i = 0; i < 9; ++i) S.ye[i] = 0.0;
  S.kap = 0;
  S.fail = 0;
  S.n_pmm = S.n_polish = 0;
  S.pri = S.dua = S.npri = S.ndua = 0.0;
}

// knots k0, k0 + kstep, ... (device: even knots on the solver warp, odd ones on the producer warp)
CMPC_FN void write_solution_knots(const Params& P, const TileCtx& T, const Inst& I, double* X_out, double* U_out, int k0, int kstep) {
  const int N = P.N;
  double* Xo = X_out + (long)I.b * (N + 1) * 9;
  double* Uo = U_out + (long)I.b * N * P.nu;
  for (int k = k0; k <= N; k += kstep) {
    const double* r = rec_of(T, I, k);
#pragma unroll
    for (int i = 0; i < 9; ++i) Xo[k * 9 + i] = CMPC_R(r, R_X + i);
    if (k == N) break;
    const int mt = meta_of(T, I, k)[0];
    const int ns = mt & 7;
    for (int j = 0; j < P.nu; ++j) Uo[k * P.nu + j] = 0.0;
    for (int sl = 0; sl < ns; ++sl) {
      const int cid = (mt >> (4 + 2 * sl)) & 3;
#pragma unroll
      for (int a = 0; a < 3; ++a) Uo[k * P.nu + 3 * cid + a] = CMPC_R(r, R_U + 3 * sl + a);
    }
  }
}

// ---------------------------------------------------------------- the per-lane driver
// scp_solver.py:118-179 with the QP solve (ADMM interleaved with certified active-set polishes)
// flattened into a state machine: advance() runs the scalar decisions of a lane until the lane
// needs a whole-horizon operation and returns its code; the caller executes it and calls
// advance() again.  The linearisation point never moves (:129-130), so the stage data are built
// once; each SCP iteration re-solves the QP for the current (radius, weight).
enum Op {
  OP_FACTOR_ADMM = 0, OP_SWEEP_ADMM, OP_BUILD_AS, OP_FACTOR_PMM, OP_SWEEP_PMM, OP_RESCALE, OP_COPY_SOL,
  OP_EVAL, OP_WRITE, OP_DONE
};
enum Pc {
  PC_SCP_TOP = 0, PC_AFTER_FACTOR0, PC_LOOP_NEXT, PC_AFTER_SWEEP, PC_AFTER_BUILD, PC_ROUND_TOP, PC_AFTER_FACTOR_PMM,
  PC_AFTER_PMM0, PC_SW_TOP, PC_AFTER_PMM1, PC_ROUND_CHECK, PC_POLISH_END, PC_NO_POLISH, PC_AFTER_REFACTOR, PC_ADAPT,
  PC_AFTER_RESCALE, PC_AFTER_ADAPT_FACTOR, PC_QP_END, PC_QP_DONE, PC_AFTER_EVAL, PC_FINISH, PC_END
};

struct Drv {
  int pc;
  // SCP loop
  int it_scp, success, n_acc, status, qp_total, nf_total, polished;
  double radius, weight, snorm, acc, num, den;
  // QP solve
  int it, next_as, as_step, nfact, solved, check, term;
  double pri0, dua0, npri0, ndua0;
  // polish
  int round, sw, prev_chg, chg, certified, upd;
  double ye_keep[9];
  double rho_new, rhok_new;
};

CMPC_HD void drv_init(const Params& P, Sv& S, Drv& D) {
  D.pc = PC_SCP_TOP;
  D.it_scp = D.success = D.n_acc = D.qp_total = D.nf_total = D.polished = 0;
  D.status = ST_OK;
  D.radius = P.radius0;
  D.weight = P.omega0;
  D.snorm = D.acc = D.num = 0.0;
  D.den = 1.0;
  D.it = D.nfact = D.solved = D.check = D.term = 0;
  D.next_as = -1;
  D.as_step = P.as_step;
  D.pri0 = D.dua0 = D.npri0 = D.ndua0 = 0.0;
  D.round = D.sw = D.chg = D.certified = D.upd = 0;
  D.prev_chg = 1 << 30;
#pragma unroll
  for (int i = 0; i < 9; ++i) D.ye_keep[i] = 0.0;
  D.rho_new = D.rhok_new = 0.0;
  set_rho(P, P.rho0, &S.rho, &S.rhok, &S.rhoe, &S.rhoep);
  S.radius = D.radius;
  S.weight = D.weight;
  S.tau = S.weight / S.rhok;
}

CMPC_FN int advance(const Params& P, Sv& S, Drv& D) {
  for (;;) {
    switch (D.pc) {
      case PC_SCP_TOP:
        if (!(D.it_scp < P.max_scp && D.weight < P.omega_max && !(D.it_scp != 0 && D.success && 0.0 < P.conv_thresh))) {
          D.pc = PC_FINISH;
          break;
        }
        D.success = 0;
        S.radius = D.radius;
        S.weight = D.weight;
        S.tau = S.weight / S.rhok;
        D.nfact = 0; D.solved = 0; D.polished = 0; D.it = 0;
        S.fail = 0;
        D.pc = PC_AFTER_FACTOR0;
        return OP_FACTOR_ADMM;
      case PC_AFTER_FACTOR0:
        ++D.nfact;
        if (S.fail) { D.pc = PC_QP_DONE; break; }
        D.next_as = (P.polish && P.as_start > 0) ? P.as_start : -1;
        D.as_step = P.as_step;
        D.pc = PC_LOOP_NEXT;
        break;
      case PC_LOOP_NEXT:
        ++D.it;
        if (D.it > P.max_iter) { D.pc = PC_QP_END; break; }
        D.check = (D.it % P.check_every == 0) || (D.it == D.next_as);
        D.pc = PC_AFTER_SWEEP;
        return OP_SWEEP_ADMM;
      case PC_AFTER_SWEEP:
        if (!D.check) { D.pc = PC_LOOP_NEXT; break; }
        if (!(S.pri == S.pri) || !(S.dua == S.dua)) { D.pc = PC_QP_END; break; }   // NaN
        D.term = S.pri <= P.eps_abs + P.eps_rel * S.npri && S.dua <= P.eps_abs + P.eps_rel * S.ndua;
        if (D.term || D.it == D.next_as) {
          D.pri0 = S.pri; D.dua0 = S.dua; D.npri0 = S.npri; D.ndua0 = S.ndua;
          if (P.polish) {
#pragma unroll
            for (int i = 0; i < 9; ++i) D.ye_keep[i] = S.ye[i];
            D.pc = PC_AFTER_BUILD;
            return OP_BUILD_AS;
          }
          D.pc = PC_NO_POLISH;
          break;
        }
        D.pc = PC_ADAPT;
        break;
      // ---- certified active-set polish: solve the equality-constrained QP of the guessed active
      // set by the method of multipliers (penalty 1/delta; OSQP: regularised KKT + iterative
      // refinement), correct the friction active set, repeat at most 1 + rounds times.
      case PC_AFTER_BUILD:
        ++S.n_polish;
        D.certified = 0;
        D.prev_chg = 1 << 30;
        D.round = 0;
        D.pc = PC_ROUND_TOP;
        break;
      case PC_ROUND_TOP:
        S.fail = 0;
        D.pc = PC_AFTER_FACTOR_PMM;
        return OP_FACTOR_PMM;
      case PC_AFTER_FACTOR_PMM:
        ++D.nfact;
        if (S.fail) { D.pc = PC_POLISH_END; break; }
        D.chg = 0;
        D.upd = 1;   // the first multiplier sweep already corrects the active set
        D.pc = PC_AFTER_PMM0;
        return OP_SWEEP_PMM;
      case PC_AFTER_PMM0:
        ++S.n_pmm;
        D.sw = 0;
        if (D.chg) { D.pc = PC_ROUND_CHECK; break; }   // rows changed: refactor right away
        D.pc = PC_SW_TOP;
        break;
      case PC_SW_TOP:
        if (D.sw < 1 + P.refine) {
          ++S.n_pmm;
          D.upd = 1;
          D.pc = PC_AFTER_PMM1;
          return OP_SWEEP_PMM;
        }
        D.pc = PC_ROUND_CHECK;
        break;
      case PC_AFTER_PMM1:
        if (D.chg || S.pri <= P.as_tol * (1.0 + S.npri)) { D.pc = PC_ROUND_CHECK; break; }
        ++D.sw;
        D.pc = PC_SW_TOP;
        break;
      case PC_ROUND_CHECK:
        if (!(S.pri == S.pri)) { D.pc = PC_POLISH_END; break; }   // NaN
        if (D.chg == 0) {
          D.certified = S.pri <= P.as_tol * (1.0 + S.npri);   // absolute + relative, the form of OSQP's test
          D.pc = PC_POLISH_END;
          break;
        }
        D.prev_chg = D.chg;
        ++D.round;
        D.pc = D.round > P.as_rounds ? PC_POLISH_END : PC_ROUND_TOP;
        break;
      case PC_POLISH_END: {
        S.kap = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) S.ye[i] = D.ye_keep[i];   // the ADMM multiplier comes back
        // OSQP's rule for an uncertified polish after normal termination: keep it if it improves
        const double m0 = fmax(D.pri0 / (P.eps_abs + P.eps_rel * D.npri0), D.dua0 / (P.eps_abs + P.eps_rel * D.ndua0));
        const double m1 = S.pri / (P.eps_abs + P.eps_rel * S.npri);
        if (D.certified || (D.term && !S.fail && m1 < m0)) {
          D.solved = 1;
          D.polished = 1;
          S.dua = 0.0;
          S.ndua = D.ndua0;
          D.pc = PC_QP_END;
          break;
        }
        S.pri = D.pri0; S.dua = D.dua0; S.npri = D.npri0; S.ndua = D.ndua0;
        S.fail = 0;
        D.pc = PC_NO_POLISH;
        break;
      }
      case PC_NO_POLISH:
        if (D.term) { D.solved = 1; D.pc = PC_QP_END; break; }
        D.next_as = D.it + D.as_step;
        D.as_step *= 2;
        D.pc = PC_AFTER_REFACTOR;   // the polish overwrote the factor records
        return OP_FACTOR_ADMM;
      case PC_AFTER_REFACTOR:
        ++D.nfact;
        if (S.fail) { D.pc = PC_QP_END; break; }
        D.pc = PC_ADAPT;
        break;
      case PC_ADAPT:
        if (P.adaptive_rho && D.it >= P.adapt_start && D.it % P.check_every == 0) {
          double est = S.rho * sqrt((S.pri / (S.npri + 1e-10)) / (S.dua / (S.ndua + 1e-10) + 1e-10));
          est = fmin(fmax(est, 1e-6), 1e6);
          if (est > S.rho * P.adapt_tol || est < S.rho / P.adapt_tol) {
            set_rho(P, est, &D.rho_new, &D.rhok_new, nullptr, nullptr);
            D.pc = PC_AFTER_RESCALE;
            return OP_RESCALE;
          }
        }
        D.pc = PC_LOOP_NEXT;
        break;
      case PC_AFTER_RESCALE:
        S.rho = D.rho_new;
        S.rhok = D.rhok_new;
        S.tau = S.weight / S.rhok;
        D.pc = PC_AFTER_ADAPT_FACTOR;
        return OP_FACTOR_ADMM;
      case PC_AFTER_ADAPT_FACTOR:
        ++D.nfact;
        if (S.fail) { D.pc = PC_QP_END; break; }
        D.pc = PC_LOOP_NEXT;
        break;
      case PC_QP_END:
        D.pc = PC_QP_DONE;
        if (D.solved && !D.polished) return OP_COPY_SOL;   // unpolished answer: one more x-update, read-only
        break;
      case PC_QP_DONE:
        D.qp_total += D.it > P.max_iter ? P.max_iter : D.it;
        D.nf_total += D.nfact;
        if (!D.solved) {
          D.status = S.fail ? ST_QP_NUMERIC : ST_QP_MAXITER;
          D.pc = PC_FINISH;
          break;
        }
        D.pc = PC_AFTER_EVAL;
        return OP_EVAL;
      case PC_AFTER_EVAL: {
        int write = 0;
        if (D.snorm < D.radius) {
          D.acc = D.num / D.den;
          if (D.acc > P.acc_rho1) {
            D.radius *= P.beta_fail;
          } else {
            write = 1;
            D.success = 1;
            ++D.n_acc;
            if (D.acc < P.acc_rho0) D.radius = fmin(P.beta_succ * D.radius, P.radius0);
          }
        } else {
          D.weight *= P.gamma_fail;
        }
        ++D.it_scp;
        D.pc = PC_SCP_TOP;
        if (write) return OP_WRITE;
        break;
      }
      case PC_FINISH:
        // nothing accepted: hand back the last QP solution (n_accepted == 0 tells the caller; the
        // reference returns empty lists in that case)
        D.pc = PC_END;
        if (D.n_acc == 0 && D.status == ST_OK && D.it_scp > 0) return OP_WRITE;
        break;
      default:
        return OP_DONE;
    }
  }
}

// Executes one operation of the tile.  Every lane of the warp calls it with the same op (the
// streamed operations are warp-collective); `on` says whether this lane takes part.
// anycheck: some participating lane wants residuals from this ADMM sweep.
template <bool FAST>
CMPC_FN void execute(int op, const Params& P, TileCtx& T, const Inst& I, const Batch& bt, Sv& S, Drv& D, bool on, bool anycheck) {
  switch (op) {
    case OP_FACTOR_ADMM: factor_op<MODE_ADMM, FAST>(P, T, I, S, on); break;
    case OP_SWEEP_ADMM:
      backward_op<MODE_ADMM, FAST>(P, T, I, S, on);
      if (anycheck) forward_op<FW_ADMM_CHECK, FAST>(P, T, I, S, on, on && D.check, false, nullptr);
      else forward_op<FW_ADMM, FAST>(P, T, I, S, on, false, false, nullptr);
      break;
    case OP_BUILD_AS: build_active_set_op<FAST>(P, T, I, S, on, &S.kap); break;
    case OP_FACTOR_PMM: factor_op<MODE_PMM, FAST>(P, T, I, S, on); break;
    case OP_SWEEP_PMM:
      backward_op<MODE_PMM, FAST>(P, T, I, S, on);
      forward_op<FW_PMM, FAST>(P, T, I, S, on, true, on && D.upd, &D.chg);
      break;
    case OP_RESCALE: if (on) rescale_op(P, T, I, S, D.rho_new, D.rhok_new); break;
    case OP_COPY_SOL:
      backward_op<MODE_ADMM, FAST>(P, T, I, S, on);
      forward_op<FW_COPY, FAST>(P, T, I, S, on, false, false, nullptr);
      break;
    case OP_EVAL: if (on) evaluate_op(P, T, I, &D.snorm, &D.num, &D.den); break;
    case OP_WRITE:
#if defined(__CUDACC__)
      helper_fork(T, CMD_WRITE, __ballot_sync(0xffffffffu, on));
      if (on) write_solution_knots(P, T, I, bt.X_out, bt.U_out, 0, 2);
      helper_join();
#else
      if (on) write_solution_knots(P, T, I, bt.X_out, bt.U_out, 0, 1);
#endif
      break;
    default: break;
  }
}

CMPC_FN void write_stats(const Batch& bt, const Inst& I, const Sv& S, const Drv& D) {
  bt.scp_iters[I.b] = D.it_scp;
  bt.status[I.b] = D.status;
  bt.n_accepted[I.b] = D.n_acc;
  bt.qp_iters[I.b] = D.qp_total;
  bt.n_factor[I.b] = D.nf_total;
  double* inf = bt.info + (long)I.b * INFO;
  inf[0] = D.snorm; inf[1] = D.acc; inf[2] = S.pri; inf[3] = S.dua;
  inf[4] = S.rho; inf[5] = D.radius; inf[6] = D.weight; inf[7] = (double)D.polished;
  inf[8] = (double)S.n_pmm; inf[9] = (double)S.n_polish; inf[10] = 0.0; inf[11] = 0.0;
}

// bind the per-instance input pointers
CMPC_HD void bind_instance(Inst& I, const Params& P, const Batch& bt, int b) {
  const int N = P.N;
  const long plan = (long)b * bt.plan_stride;
  I.b = b;
  I.lane = b & (TL - 1);
  I.cpos = bt.cpos + plan * N * P.nc * 3;
  I.cR = bt.cR ? bt.cR + plan * N * P.nc * 9 : nullptr;
  I.fub = bt.fub ? bt.fub + (long)b * N * P.nc * 4 : nullptr;
  I.cact = bt.cact + plan * N * P.nc;
  I.Xr = bt.X_ref + (long)b * (N + 1) * 9;
  I.Ui = bt.U_init + (long)b * N * P.nu;
  I.xi = bt.x_init + (long)b * 9;
  I.xf = bt.x_final + (long)b * 9;
}
#if defined(__CUDACC__)
CMPC_HD void helper_fork(const TileCtx& T, int cmd, unsigned mask) {
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncwarp();
  if ((threadIdx.x & 31u) == 0)
    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(T.bars_sa + 48u), "r"(T.tile), "r"((int)mask), "r"(0), "r"(cmd) : "memory");
  asm volatile("bar.arrive 1, 64;" ::: "memory");
}
#endif
CMPC_HD void bind_tile(TileCtx& T, const Params& P, const Batch& bt, int tile) {
  T.prm = &P;
#if defined(__CUDACC__)
  T.tile = tile;
#endif
  T.ws = bt.ws + (long)tile * (P.N + 1) * (REC * TL);
  T.gt = bt.gtab ? bt.gtab + (long)tile * P.N * (GT * TL) : nullptr;
  T.nst = bt.nst + (long)tile * (P.N + 1);
}

}  // namespace cmpc
