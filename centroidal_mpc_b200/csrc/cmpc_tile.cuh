// cmpc_tile.cuh — the SCP solver of one MPC instance: Riccati factorisation, ADMM / multiplier-
// method sweeps, certified active-set polish, trust-region loop.
//
// Execution model (cmpc_core.cuh): NL lanes of a warp form the TEAM of one instance, TL = 32 / NL
// instances (a tile) share the warp.  Every per-knot operation is written "owner computes": lane q
// of the team owns the rows q, q + NL, ... of the knot's small dense systems, vectors that every
// lane needs (u, hu, p, the pivot column) go through a few shared-memory words per instance and a
// __syncwarp.  An output element is produced by exactly one lane with a fixed operation order and
// the only cross-lane reductions are max / or / integer sums, so the results do not depend on NL:
// the host build (tests/emu) runs NL = 1, a lock-step host build runs NL = 8 on coroutines, and
// both are bit-identical to the GPU (FMA contraction is explicit, automatic contraction is off).
// Scalars of the driver state machine are replicated in the lanes of a team.
//
// Control flow: every instance owns a small state machine (advance()) that names the next whole-
// horizon operation it needs (factorise, sweep, build active set, evaluate, ...).  The warp
// executes one operation at a time for the instances that asked for it (run_tile in cmpc_api.cu).
// DESIGN.md "device algorithm" has the mathematics; oracle/device_model.py is the numpy model.
#pragma once
#include "cmpc_core.cuh"

namespace cmpc {

constexpr int MODE_ADMM = 0, MODE_PMM = 1;
// forward-sweep kinds
constexpr int FW_ADMM = 0, FW_ADMM_CHECK = 1, FW_PMM = 2, FW_COPY = 4;

// ---------------------------------------------------------------- staged ranges of the streamed operations
// A streamed operation walks the knots; per knot it stages one or two contiguous field ranges of the
// record into a ring slot (one cp.async.bulk each).  so() maps a record field to its slot position.
enum StreamKind { SK_FAC_ADMM = 0, SK_FAC_PMM, SK_BWD_ADMM, SK_BWD_PMM, SK_FWD_ADMM, SK_FWD_PMM, SK_FWD_COPY };
struct Ranges { int s1, e1, s2, e2; };
CMPC_CX Ranges ranges_of(int sk, int ns, bool gen) {
  const Lay L = lay_of(ns, gen);
  return sk == SK_FAC_ADMM ? Ranges{L.meta, L.vk, 0, 0}
       : sk == SK_FAC_PMM  ? Ranges{L.meta, L.vk, L.yk, L.x}
       : sk == SK_BWD_ADMM ? Ranges{0, L.dv, 0, 0}
       : sk == SK_BWD_PMM  ? Ranges{0, L.vk, L.yk, L.x}
       : sk == SK_FWD_ADMM ? Ranges{L.kt, L.yk, 0, 0}
       : sk == SK_FWD_PMM  ? Ranges{L.kt, L.vk, L.dv, L.x}
                           : Ranges{L.kt, L.vk, L.dv, L.yk};
}
CMPC_CX int slot_fields(int sk, bool gen) {
  const Ranges R = ranges_of(sk, MAXC, gen);
  return (R.e1 - R.s1) + (R.e2 - R.s2);
}
CMPC_CX int imax(int a, int b) { return a > b ? a : b; }
CMPC_CX int imin(int a, int b) { return a < b ? a : b; }
// per-instance scratch words (exchange vectors of the sweeps, tableau data of the factorisation)
constexpr int X_PX = 0;            // p, double-buffered   [2][9]
constexpr int X_UX = 18;           // u / hu               [12]
constexpr int X_KX = 30;           // kappa terms          [12]  (kl 3 | kM 9)
constexpr int X_DX = 42;           // friction residual terms of a CHECK sweep [16]
constexpr int X_RX = 58;           // team reductions of the host lock-step build [NL]
constexpr int X_COMMON = 58 + 8;
constexpr int X_CXS = 2 * 12 + 9;  // one pivot column: control entries twice (rotated window), state entries
constexpr int X_CX = X_COMMON;     // factorisation: pivot column, double-buffered [2][X_CXS]
constexpr int X_PS = X_CX + 2 * X_CXS;   // P (full square 9 x 9)
constexpr int X_Y = X_PS + 81;     // Y = P [B A]  (9 x (na + 9))
constexpr int X_FAC_END = X_Y + 9 * 21;
// shared memory of a tile (in field rows of TL doubles): common scratch, then a region that holds the
// ring of a sweep or the factorisation scratch plus its small ring
// (measured on B200 with 3 and 4 slots: at 4096 instances the larger ring costs a resident CTA per SM -- two waves,
// 13.5 instead of 9.7 ms; at 1024 instances, where occupancy does not matter, 5.86 against 5.80 ms: the bulk copies
// are not what the knot steps wait for)
#ifndef CMPC_RING_SLOTS
#define CMPC_RING_SLOTS 2
#endif
CMPC_CX int ring_fields(bool gen) { return CMPC_RING_SLOTS * imax(slot_fields(SK_BWD_ADMM, gen), slot_fields(SK_BWD_PMM, gen)); }
CMPC_CX int ring_depth(int sk, bool gen) {
  const int avail = (sk == SK_FAC_ADMM || sk == SK_FAC_PMM) ? ring_fields(gen) - (X_FAC_END - X_COMMON) : ring_fields(gen);
  return imin(4, imax(2, avail / slot_fields(sk, gen)));
}
CMPC_CX int ring_base(int sk) { return (sk == SK_FAC_ADMM || sk == SK_FAC_PMM) ? X_FAC_END : X_COMMON; }
CMPC_CX int tile_smem_fields(bool gen) { return X_COMMON + ring_fields(gen); }
static_assert(X_FAC_END - X_COMMON + 2 * slot_fields(SK_FAC_PMM, true) <= ring_fields(true), "factor scratch");
static_assert(X_FAC_END - X_COMMON + 2 * slot_fields(SK_FAC_PMM, false) <= ring_fields(false), "factor scratch");

#if defined(CMPC_PROFILE) && defined(__CUDACC__)
#define CMPC_PROF_PTR(T) (T).prof
#else
#define CMPC_PROF_PTR(T) nullptr
#endif

// ---------------------------------------------------------------- memory spaces
// Device: staged data and scratch live in shared memory and are addressed by 32-bit shared-space byte
// addresses through explicit ld/st.shared (generic accesses cost a long-scoreboard round trip).
// Host build: plain pointers; "staged" data is the global record itself.
#if defined(__CUDACC__)
typedef unsigned StagedPtr;
typedef unsigned ScratchPtr;
CMPC_HD double sp_ld(unsigned p, int e) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(p + (unsigned)e * (TL * 8u)));
  return v;
}
CMPC_HD void sp_st(unsigned p, int e, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(p + (unsigned)e * (TL * 8u)), "d"(v) : "memory");
}
CMPC_HD int sp_ldi(unsigned p, int e, int t, int which) {   // int32 pair field: [t] word 0, [TL + t] word 1
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(p - (unsigned)t * 8u + (unsigned)e * (TL * 8u) + (unsigned)(which * TL + t) * 4u));
  return v;
}
// pre-scaled offsets of a lane's own rows: formed once per run of knots and made opaque to the compiler, which
// would otherwise rematerialise the index arithmetic (a third of the instructions of a knot) in every knot
typedef unsigned SOff;
CMPC_HD SOff s_off(int fields) { return (unsigned)fields * (TL * 8u); }
CMPC_HD unsigned sp_at(unsigned p, SOff o) { return p + o; }
#define CMPC_OPAQUE(x) asm volatile("" : "+r"(x))
#else
typedef const double* StagedPtr;
typedef double* ScratchPtr;
typedef long SOff;
CMPC_HD SOff s_off(int fields) { return (long)fields * TL; }
CMPC_HD const double* sp_at(const double* p, SOff o) { return p + o; }
CMPC_HD double* sp_at(double* p, SOff o) { return p + o; }
#define CMPC_OPAQUE(x)
CMPC_HD double sp_ld(const double* p, int e) { return p[(long)e * TL]; }
CMPC_HD void sp_st(double* p, int e, double v) { p[(long)e * TL] = v; }
CMPC_HD int sp_ldi(const double* p, int e, int t, int which) {
  return (reinterpret_cast<const int*>(p - t + (long)e * TL))[which * TL + t];
}
#endif
#define CMPC_R(p, f) (p)[(long)(f) * TL]

// ---------------------------------------------------------------- team primitives
struct Inst {
  int b, lane, sub;   // batch index, instance lane t within the tile, lane q within the team
  const double *Xr, *Ui, *xi, *xf, *cpos, *cR, *fub;
  const int* cact;
};
#if defined(__CUDACC__)
CMPC_HD void team_sync(const Inst&) { if (NL > 1) __syncwarp(); }
CMPC_HD double team_max(const Inst&, ScratchPtr, double v) {
#pragma unroll
  for (int m = TL; m < 32; m <<= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, m));
  return v;
}
CMPC_HD int team_or(const Inst&, ScratchPtr, int v) {
#pragma unroll
  for (int m = TL; m < 32; m <<= 1) v |= __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
CMPC_HD int team_sum(const Inst&, ScratchPtr, int v) {
#pragma unroll
  for (int m = TL; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
#elif CMPC_NL == 1
CMPC_HD void team_sync(const Inst&) {}
CMPC_HD double team_max(const Inst&, ScratchPtr, double v) { return v; }
CMPC_HD int team_or(const Inst&, ScratchPtr, int v) { return v; }
CMPC_HD int team_sum(const Inst&, ScratchPtr, int v) { return v; }
#else
// lock-step host build: the NL lanes of a team are coroutines; cmpc_emu_yield() returns when every
// lane of the team has reached its own yield (tests/emu/cmpc_emu.cpp)
void cmpc_emu_yield();
CMPC_HD void team_sync(const Inst&) { cmpc_emu_yield(); }
CMPC_HD double team_max(const Inst& I, ScratchPtr xs, double v) {
  sp_st(xs, X_RX + I.sub, v);
  cmpc_emu_yield();
  double m = sp_ld(xs, X_RX);
  for (int q = 1; q < NL; ++q) m = fmax(m, sp_ld(xs, X_RX + q));
  cmpc_emu_yield();
  return m;
}
CMPC_HD int team_or(const Inst& I, ScratchPtr xs, int v) {
  sp_st(xs, X_RX + I.sub, (double)(unsigned)v);
  cmpc_emu_yield();
  unsigned m = 0;
  for (int q = 0; q < NL; ++q) m |= (unsigned)sp_ld(xs, X_RX + q);
  cmpc_emu_yield();
  return (int)m;
}
CMPC_HD int team_sum(const Inst& I, ScratchPtr xs, int v) {
  sp_st(xs, X_RX + I.sub, (double)v);
  cmpc_emu_yield();
  int m = 0;
  for (int q = 0; q < NL; ++q) m += (int)sp_ld(xs, X_RX + q);
  cmpc_emu_yield();
  return m;
}
#endif
CMPC_HD int sub_of(const Inst& I) { return NL == 1 ? 0 : (I.sub & (NL - 1)); }   // the mask tells the compiler the range: dead row branches go
// true when the predicate holds for every instance of the tile (device: the whole warp; host build: the
// instances of a tile run one after the other, each for itself)
CMPC_HD bool tile_all(bool v) {
#if defined(__CUDACC__)
  return __all_sync(0xffffffffu, v);
#else
  return v;
#endif
}

struct TileCtx {
  double* ws;      // tile workspace [N+1][rfields][TL]
  long rstride;    // doubles per knot record (rfields * TL)
  int* nst;        // [N+1] slots per knot (tile maximum; 0 at the terminal knot), global memory
  int gen;         // general friction table present
#if defined(__CUDACC__)
  unsigned smem_sa, bars_sa, phases;   // the warp's shared memory (scratch, ring), the stream's range table; warp-uniform
  int tile;
#if defined(CMPC_PROFILE)
  long long* prof;                     // cycle counters of a profiling build (scripts/prof_cycles.py)
#endif
  unsigned nst_sa;                     // shared copy of nst (bytes), shared-space address
  CMPC_HD int ns(int k) const {
    int v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(nst_sa + (unsigned)k));
    return v;
  }
#else
  double* scratch;                     // [X_FAC_END][TL]
  CMPC_HD int ns(int k) const { return nst[k]; }
#endif
};

CMPC_HD double* rec_of(const TileCtx& T, const Inst& I, int k) { return T.ws + (long)k * T.rstride + I.lane; }
CMPC_HD int* meta_of(const TileCtx& T, const Inst& I, int k, int f_meta) {
  return reinterpret_cast<int*>(T.ws + (long)k * T.rstride + (long)f_meta * TL) + I.lane;   // [0] meta, [TL] active set
}
CMPC_HD ScratchPtr scratch_of(const TileCtx& T, const Inst& I) {
#if defined(__CUDACC__)
  return T.smem_sa + (unsigned)I.lane * 8u;
#else
  return T.scratch + I.lane;
#endif
}

// ---------------------------------------------------------------- knot stream
// Walks the knots of a tile in one direction and hands out a pointer through which the staged fields of
// the current knot can be read.  Device: a ring of DEPTH shared-memory slots filled asynchronously, so that
// the HBM latency hides behind the arithmetic of the knots in between and every operand read is a
// shared-memory read.  Host build: the pointer is the global record itself.
#if defined(__CUDACC__)
CMPC_HD unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// bounded wait on an mbarrier phase: returns false when the bulk copy never lands
CMPC_HD bool mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spins = 0; !done; ++spins) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spins > (1u << 22)) return false;
  }
  return true;
}

// Device: a ring of DEPTH shared-memory slots filled by cp.async.bulk: lane 0 of the warp issues the copy
// of knot k + DEPTH (one bulk copy per field range, byte ranges from a small table per slot count) when the
// warp releases knot k; completion on the slot's mbarrier.
// Measured alternatives (B200, 4096 x N=100): an L2 prefetch (cp.async.bulk.prefetch.L2) four
// knots ahead of the ring: 14.0 instead of 11.7 ms per batch; a non-blocking mbarrier.test_wait of the next
// slot early in the knot: the instruction holds the warp for its whole latency; all 32 lanes copying with
// cp.async (16 bytes per lane) instead of the bulk copy: 13.8 instead of 12.0 ms.
template <int SK, bool GEN>
struct KnotStream {
  static constexpr int DEPTH = ring_depth(SK, GEN);
  static constexpr int SLOTF = slot_fields(SK, GEN);
  unsigned ring_sa, bars_sa, phases, nst_sa, tab_sa;
  unsigned no1, nb1, no2, nb2;   // byte ranges of the next knot to issue (fetched by peek(), off the critical path)
  int lane, s_wait, k_issue, n_left, dir, s_issue, lost;
  const double* ws;
  long rstride;
#if defined(CMPC_PROFILE)
  long long* prof;
#endif

  CMPC_HD int ns_at(int k) const {
    int v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(nst_sa + (unsigned)k));
    return v;
  }
  // The stream state (k_issue, n_left, s_issue) is warp-uniform: every lane keeps it, every lane looks the byte
  // ranges up (a broadcast load, issued early by peek() so that it overlaps the knot's arithmetic), and only the
  // asynchronous instructions themselves are predicated on lane 0 -- no divergent region with a chain of
  // dependent shared-memory loads while 31 lanes wait (the profiling build showed acquire + release at 40 % of a
  // backward knot step).
  CMPC_HD void lookup() {   // byte ranges of knot k_issue
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(no1), "=r"(nb1), "=r"(no2), "=r"(nb2)
                 : "r"(tab_sa + 16u * (unsigned)ns_at(n_left > 0 ? k_issue : 0)));
  }
  CMPC_HD void issue() {   // knot k_issue into slot s_issue (ranges from lookup())
    const unsigned bar = bars_sa + 8u * s_issue;
    const unsigned dst = ring_sa + (unsigned)(s_issue * SLOTF) * (TL * 8u);
    const char* src = reinterpret_cast<const char*>(ws + (long)k_issue * rstride);
    const unsigned first = (threadIdx.x & 31u) == 0 ? 1u : 0u;
    asm volatile(
        "{\n"
        ".reg .pred p, p2;\n"
        "setp.ne.u32 p, %7, 0;\n"
        "setp.ne.and.u32 p2, %6, 0, p;\n"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3], %4, [%0];\n"
        "@p2 cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%8], [%5], %6, [%0];\n"
        "}" ::"r"(bar), "r"(nb1 + nb2), "r"(dst), "l"(src + no1), "r"(nb1), "l"(src + no2), "r"(nb2), "r"(first), "r"(dst + nb1)
        : "memory");
    k_issue += dir;
    --n_left;
    s_issue = (s_issue + 1 == DEPTH) ? 0 : s_issue + 1;
  }
  // knots k_first, k_first + dir, ... (count of them)
  CMPC_HD void open(TileCtx& Tc, const Inst& I, int k_first, int count, int direction) {
    ring_sa = Tc.smem_sa + (unsigned)ring_base(SK) * (TL * 8u);
    bars_sa = Tc.bars_sa; phases = Tc.phases;
    lane = I.lane; s_wait = 0; s_issue = 0; k_issue = k_first; n_left = count; dir = direction; lost = 0;
    ws = Tc.ws; rstride = Tc.rstride; nst_sa = Tc.nst_sa; tab_sa = Tc.bars_sa + 32u;
#if defined(CMPC_PROFILE)
    prof = Tc.prof;
#endif
    // earlier generic-proxy writes of this warp (records written by the previous operation) must be
    // visible to the async proxy before the bulk copies read them
    asm volatile("fence.proxy.async;" ::: "memory");
    const unsigned wl = threadIdx.x & 31u;
    if (wl <= MAXC) {   // byte ranges of this operation per slot count: {offset 1, bytes 1, offset 2, bytes 2}
      const Ranges R = ranges_of(SK, (int)wl, GEN);
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(tab_sa + 16u * wl), "r"((unsigned)R.s1 * (TL * 8u)),
                   "r"((unsigned)(R.e1 - R.s1) * (TL * 8u)), "r"((unsigned)R.s2 * (TL * 8u)), "r"((unsigned)(R.e2 - R.s2) * (TL * 8u)) : "memory");
    }
    __syncwarp();
    for (int i = 0; i < DEPTH && n_left > 0; ++i) { lookup(); issue(); }
  }
  CMPC_HD void peek() { lookup(); }   // called right after acquire(): the ranges of the knot the next release() issues
  CMPC_HD StagedPtr acquire() {
    const unsigned bar = bars_sa + 8u * s_wait;
#if defined(CMPC_PROFILE)
    const long long t0_ = clock64();
#endif
    if (!mbar_wait(bar, (phases >> s_wait) & 1u)) lost = 1;
#if defined(CMPC_PROFILE)
    prof[14] += clock64() - t0_;
    prof[15] += 1;
#endif
    phases ^= 1u << s_wait;
    return ring_sa + (unsigned)((s_wait * SLOTF) * TL + lane) * 8u;
  }
  CMPC_HD void release() {
    __syncwarp();          // every lane is done reading the slot
    if (n_left > 0) issue();
    s_wait = (s_wait + 1 == DEPTH) ? 0 : s_wait + 1;
  }
  CMPC_HD int close(TileCtx& Tc) { Tc.phases = phases; return lost; }   // every issued copy has been consumed
};
template <int SK, int NS, bool GEN>
CMPC_CX int so_of(int f) {
  const Ranges R = ranges_of(SK, NS, GEN);
  return f < R.e1 ? f - R.s1 : f - R.s2 + (R.e1 - R.s1);
}
#else
template <int SK, bool GEN>
struct KnotStream {
  const double* ws;
  long rstride;
  int lane, dir, k;
  CMPC_HD void open(TileCtx& T, const Inst& I, int k_first, int, int direction) {
    ws = T.ws; rstride = T.rstride; lane = I.lane; dir = direction; k = k_first;
  }
  CMPC_HD void peek() {}
  CMPC_HD StagedPtr acquire() const { return ws + (long)k * rstride + lane; }
  CMPC_HD void release() { k += dir; }
  CMPC_HD int close(TileCtx&) { return 0; }
};
template <int SK, int NS, bool GEN>
CMPC_CX int so_of(int f) { return f; }
#endif
// staged field f (+ runtime offset o) of the current knot; SK, NS, GEN are constants of the enclosing scope
#define CMPC_S(r, f) sp_ld(r, so_of<SK, NS, !FAST>(f))
#define CMPC_SO(r, f, o) sp_ld(r, so_of<SK, NS, !FAST>(f) + (o))
#define CMPC_SI(r, f, which) sp_ldi(r, so_of<SK, NS, !FAST>(f), I.lane, which)

// Solver scalars of one instance (replicated in the lanes of its team).
struct Sv {
  double rho, rhok, rhoe, rhoep, radius, weight;
  double tau;                      // weight / rhok: threshold of the trust-region prox
  double pri, dua, npri, ndua;     // last residuals
  double nq, dynrow;               // constant parts of the residual norms
  int kap;                         // multiplier method: some knot has trust-region rows
  int fail;                        // a pivot was not positive
  int kbad;                        // polish: the trust-region rows of some knot contradict the branch they were given
  int lost;                        // a bulk copy never landed (device)
  int nconv;                       // all-zero warm start: the reference's convergence() is NaN, never below the threshold
  int badin;                       // a non-finite input (reference point, warm start, friction bound): no solve, status QP_NUMERIC
  int n_pmm, n_polish;             // statistics: multiplier-method sweeps, polish attempts
  double ye[9];                    // multiplier of x_N = x_final
};

// ---------------------------------------------------------------- friction rows
// G = pyr4 * R^T, pyr4 = [[1,0,-k],[-1,0,-k],[0,1,-k],[0,-1,-k]], k = mu/sqrt2
// (utils.py:9-16, constraints.py:178-184); e2 = row equilibration under D_u = 1/sqrt(W_u).
// Fast path (identity R, same W_u for all contacts): constants; else the staged table (segment G).
// The rows are carried SHIFTED by their upper bound (stochastic mode: the chance-constraint back-offs,
// constraints.py:187-214; zero otherwise): cf = G u - ub, so that w = min(v, 0), y ~ max(v, 0),
// "active <=> cf = 0" and "violated <=> cf > 0" hold unchanged for the shifted values.
CMPC_HD double pyr4(const Params& P, int row, int a) {
  if (a == 2) return -P.kf;
  if (a == 0) return row == 0 ? 1.0 : (row == 1 ? -1.0 : 0.0);
  return row == 2 ? 1.0 : (row == 3 ? -1.0 : 0.0);
}
struct FRow { double gx, gy, gz, e2, ub; };   // one pyramid row of a slot
#define CMPC_FROW(fr, r, s, row)                                                       \
  do {                                                                                 \
    if (FAST) {                                                                        \
      fr.gx = (row) == 0 ? 1.0 : ((row) == 1 ? -1.0 : 0.0);                            \
      fr.gy = (row) == 2 ? 1.0 : ((row) == 3 ? -1.0 : 0.0);                            \
      fr.gz = -P.kf; fr.e2 = (row) < 2 ? P.e2[0] : P.e2[2]; fr.ub = 0.0;               \
    } else {                                                                           \
      const int o_ = (s) * GS;                                                         \
      fr.gx = CMPC_SO(r, L.g, o_ + (row) * 3); fr.gy = CMPC_SO(r, L.g, o_ + (row) * 3 + 1); \
      fr.gz = CMPC_SO(r, L.g, o_ + (row) * 3 + 2);                                     \
      fr.e2 = CMPC_SO(r, L.g, o_ + 12 + (row)); fr.ub = CMPC_SO(r, L.g, o_ + 16 + (row)); \
    }                                                                                  \
  } while (0)
// column a of G for the four rows of slot s (G' t needs it)
#define CMPC_GCOL(gc, r, s, a)                                                         \
  do {                                                                                 \
    if (FAST) {                                                                        \
      gc[0] = (a) == 0 ? 1.0 : ((a) == 2 ? -P.kf : 0.0); gc[1] = (a) == 0 ? -1.0 : ((a) == 2 ? -P.kf : 0.0); \
      gc[2] = (a) == 1 ? 1.0 : ((a) == 2 ? -P.kf : 0.0); gc[3] = (a) == 1 ? -1.0 : ((a) == 2 ? -P.kf : 0.0); \
    } else {                                                                           \
      _Pragma("unroll") for (int row_ = 0; row_ < 4; ++row_) gc[row_] = CMPC_SO(r, L.g, (s) * GS + row_ * 3 + (a)); \
    }                                                                                  \
  } while (0)

// ---------------------------------------------------------------- trust-region prox
// argmin_v omega*max(0, |v - kbar|_1 - r) + rho/2 |v - a|^2  (oracle/device_model.py prox_trust)
// branch 0: inside the L1 ball; 1: outside after soft-thresholding; 2: on the surface.
// tau = omega / rho is passed in (no division on the device in the per-knot path)
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
// entry i of a 9-vector of kernel parameters with a run-time index (a select chain: a dynamically
// indexed copy of the parameter block would live in local memory)
CMPC_HD double pick9(const double* w, int i) {
  double v = w[0];
#pragma unroll
  for (int c = 1; c < 9; ++c) v = (i == c) ? w[c] : v;
  return v;
}
CMPC_HD double pick3(double a, double b, double c, int i) { return i == 0 ? a : (i == 1 ? b : c); }

// ---------------------------------------------------------------- Riccati factorisation
// Backward over k.  The symmetric tableau  T = [[Huu, Hux],[Hux', Q + A'PA]]  (Huu = R + B'PB, Hux = B'PA;
// n = na + 9 rows, controls first) is swept on its na control pivots (SPD, no pivoting):
//   T -> [[-Huu^-1, Huu^-1 Hux],[Hux' Huu^-1, Q + A'PA - Hux' Huu^-1 Hux]],
// which delivers Hn = -Huu^-1, Kt = Hux' Huu^-1 = -K' and P_k in one pass.  Pc = P c.
// ADMM mode: R = W_u + rho G'E2G, Q = W_x + rho_k I (kappa); multiplier mode: active friction rows and
// kappa rows carry the penalty 1/delta.
// Team work split: lane q owns the tableau rows q, q + NL, ... in registers: the control columns of every
// row (they are the factor record [Hn; Kt]) and the state columns of the state rows; the state columns of
// the control rows are the mirror image of the control columns of the state rows and are never formed.
template <int NS, int MODE, bool FAST, int SK>
CMPC_HD void factor_knot(const Params& P, Sv& S, StagedPtr r, double* w, const Inst& I, int k, ScratchPtr xs, bool on, long long* prof_) {
  constexpr Lay L = lay_of(NS, !FAST);
  constexpr int NA = 3 * NS, n = NA + 9, NAP = L.nap;
  constexpr int RT = (n + NL - 1) / NL;   // tableau rows per lane
  constexpr int ST = (9 + NL - 1) / NL;   // state rows per lane
  const int q = sub_of(I);
  const double inv = P.inv_delta;
  const int mt = CMPC_SI(r, L.meta, 0), nsl = mt & 7;
  const int pm = (MODE == MODE_PMM) ? CMPC_SI(r, L.meta, 1) : 0;
  const double S3[3] = {CMPC_S(r, L.s), CMPC_S(r, L.s + 1), CMPC_S(r, L.s + 2)};
  const double ck[3] = {CMPC_S(r, L.ck), CMPC_S(r, L.ck + 1), CMPC_S(r, L.ck + 2)};
  // wrench model: B_k[:, slot] = dt [0; sg I; M] with the slot's 3x3 block M (column a at 3a) and sg = 1 for a
  // force slot, 0 for a wrench slot (odd pseudo-contact id); point contacts: M = [d]x, sg = 1, never formed
  double ds[NS > 0 ? NS : 1][WR ? 9 : 3], dts[NS > 0 ? NS : 1], sg[NS > 0 ? NS : 1];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    dts[s] = s < nsl ? P.dt : 0.0;
    sg[s] = (WR && ((mt >> (4 + 2 * s)) & 1)) ? 0.0 : 1.0;
#pragma unroll
    for (int a = 0; a < (WR ? 9 : 3); ++a) ds[s][a] = CMPC_S(r, L.d + (WR ? 9 : 3) * s + a);
  }
  const bool kap = MODE == MODE_PMM && S.kap && k >= 1;
  if (kap) {   // kappa penalty block of the multiplier method -> scratch (every lane writes the same values)
    const double kb[3] = {CMPC_S(r, L.xb + 6), CMPC_S(r, L.xb + 7), CMPC_S(r, L.xb + 8)};
    const double yk[4] = {CMPC_S(r, L.yk), CMPC_S(r, L.yk + 1), CMPC_S(r, L.yk + 2), CMPC_S(r, L.yk + 3)};
    double kM[9], kl[3];
    pmm_kappa_terms(P, S, pm, kb, yk, kM, kl);
#pragma unroll
    for (int i = 0; i < 9; ++i) sp_st(xs, X_KX + 3 + i, kM[i]);
  }
#if defined(CMPC_PROFILE) && defined(__CUDACC__)
#define CMPC_FK(i) do { const long long t_ = clock64(); prof_[22 + (i)] += t_ - ck_; ck_ = t_; } while (0)
  long long ck_ = clock64();
#else
#define CMPC_FK(i)
#endif
  // ---- phase 1: rows of Y = P [B A] and Pc = P c, by the owners of the state rows
#pragma unroll
  for (int t = 0; t < ST; ++t) {
    const int i = q + NL * t;
    if (i < 9) {
      double Pr[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) Pr[c] = sp_ld(xs, X_PS + i * 9 + c);
      double pc = Pr[5] * P.dtmg;
      pc = fma(Pr[6], ck[0], pc);
      pc = fma(Pr[7], ck[1], pc);
      pc = fma(Pr[8], ck[2], pc);
      if (on) CMPC_R(w, L.pc + i) = pc;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {   // (P B)[i][(s,a)] = dt_s (P[i][3+a] + P[i][6+a1] d[a2] - P[i][6+a2] d[a1])
          const int a1 = nxt3(a), a2 = prv3(a);
          if (WR) sp_st(xs, X_Y + i * n + 3 * s + a, dts[s] * fma(Pr[8], ds[s][3 * a + 2], fma(Pr[7], ds[s][3 * a + 1], fma(Pr[6], ds[s][3 * a], sg[s] * Pr[3 + a]))));
          else
          sp_st(xs, X_Y + i * n + 3 * s + a, dts[s] * fma(Pr[6 + a1], ds[s][a2], fma(-Pr[6 + a2], ds[s][a1], Pr[3 + a])));
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {     // (P A)[i][:]
        const int c1 = nxt3(c), c2 = prv3(c);
        sp_st(xs, X_Y + i * n + NA + c, fma(P.dt, fma(Pr[6 + c1], S3[c2], -(Pr[6 + c2] * S3[c1])), Pr[c]));
        sp_st(xs, X_Y + i * n + NA + 3 + c, fma(P.dt_m, Pr[c], Pr[3 + c]));
        sp_st(xs, X_Y + i * n + NA + 6 + c, Pr[6 + c]);
      }
    }
  }
  team_sync(I);
  CMPC_FK(0);
  // ---- phase 2: the lane's tableau rows in registers: control columns by slot, state columns
  double Tf[RT][NA > 0 ? NA : 1], Ts[RT][9];
#pragma unroll
  for (int t = 0; t < RT; ++t) {
    const int rr = q + NL * t;
#pragma unroll
    for (int c = 0; c < 9; ++c) Ts[t][c] = 0.0;
#pragma unroll
    for (int e = 0; e < NA; ++e) Tf[t][e] = 0.0;
    if (rr < NA) {   // control row (s, a):  B' (P B)  + R block
      const int s = rr / 3, a = rr - 3 * s, a1 = a == 2 ? 0 : a + 1, a2 = a == 0 ? 2 : a - 1;
      const double dtr = s < nsl ? P.dt : 0.0;
      const double dA = WR ? 0.0 : CMPC_SO(r, L.d, 3 * s + a2), dB = WR ? 0.0 : CMPC_SO(r, L.d, 3 * s + a1);
      const double m0 = WR ? CMPC_SO(r, L.d, 9 * s + 3 * a) : 0.0, m1 = WR ? CMPC_SO(r, L.d, 9 * s + 3 * a + 1) : 0.0;
      const double m2 = WR ? CMPC_SO(r, L.d, 9 * s + 3 * a + 2) : 0.0;
      const double sgr = (WR && s < nsl && ((mt >> (4 + 2 * s)) & 1)) ? 0.0 : 1.0;
      // R block of the slot: W_u + G' diag(rr) G, columns 3s .. 3s+2
      double gc[3][4], rw[4];
#pragma unroll
      for (int b2 = 0; b2 < 3; ++b2) CMPC_GCOL(gc[b2], r, s, b2);
#pragma unroll
      for (int row = 0; row < 4; ++row) {
        if (MODE == MODE_ADMM) { FRow fr; CMPC_FROW(fr, r, s, row); rw[row] = S.rho * fr.e2; }
        else rw[row] = ((pm >> (4 * s + row)) & 1) ? inv : 0.0;
      }
      const double ga[4] = {pick3(gc[0][0], gc[1][0], gc[2][0], a), pick3(gc[0][1], gc[1][1], gc[2][1], a),
                            pick3(gc[0][2], gc[1][2], gc[2][2], a), pick3(gc[0][3], gc[1][3], gc[2][3], a)};
      const int cid = (s < nsl) ? ((mt >> (4 + 2 * s)) & 3) : 0;
      double radd[3];
#pragma unroll
      for (int b2 = 0; b2 < 3; ++b2) {
        double v = 0.0;
        if (b2 == a) {
          if (FAST) v = pick3(P.Wu[0], P.Wu[1], P.Wu[2], a);
          else {
            double wsel = P.Wu[b2];
#pragma unroll
            for (int c = 1; c < MAXC; ++c) wsel = (cid == c) ? P.Wu[3 * c + b2] : wsel;
            v = wsel;
          }
        }
#pragma unroll
        for (int row = 0; row < 4; ++row) v = fma(rw[row] * ga[row], gc[b2][row], v);
        radd[b2] = v;
      }
#pragma unroll
      for (int e = 0; e < NA; ++e) {
        double v;
        if (WR) v = dtr * fma(sp_ld(xs, X_Y + 8 * n + e), m2, fma(sp_ld(xs, X_Y + 7 * n + e), m1, fma(sp_ld(xs, X_Y + 6 * n + e), m0, sgr * sp_ld(xs, X_Y + (3 + a) * n + e))));
        else v = dtr * fma(sp_ld(xs, X_Y + (6 + a1) * n + e), dA, fma(-sp_ld(xs, X_Y + (6 + a2) * n + e), dB, sp_ld(xs, X_Y + (3 + a) * n + e)));
        if (e / 3 == s) v += radd[e % 3];
        Tf[t][e] = v;
      }
    } else if (rr < n) {   // state row i:  Hux' | Q + A'(P A)
      const int i = rr - NA;
      const int g3 = i / 3, a = i - 3 * g3, a1 = a == 2 ? 0 : a + 1, a2 = a == 0 ? 2 : a - 1;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int b2 = 0; b2 < 3; ++b2) {   // Hux[(s,b2)][i]
          const int b1 = nxt3(b2), bp = prv3(b2);
          if (WR) Tf[t][3 * s + b2] = dts[s] * fma(sp_ld(xs, X_Y + 8 * n + NA + i), ds[s][3 * b2 + 2], fma(sp_ld(xs, X_Y + 7 * n + NA + i), ds[s][3 * b2 + 1],
                                      fma(sp_ld(xs, X_Y + 6 * n + NA + i), ds[s][3 * b2], sg[s] * sp_ld(xs, X_Y + (3 + b2) * n + NA + i))));
          else
          Tf[t][3 * s + b2] = dts[s] * fma(sp_ld(xs, X_Y + (6 + b1) * n + NA + i), ds[s][bp],
                                      fma(-sp_ld(xs, X_Y + (6 + bp) * n + NA + i), ds[s][b1], sp_ld(xs, X_Y + (3 + b2) * n + NA + i)));
        }
      }
      // (A'X)[i][c] = X[i][c] + mul (X[rA][c] sA - X[rB][c] sB):  rows 0..2: dt [S]x', rows 3..5: dt/m, rows 6..8: -
      const double mul = g3 == 0 ? P.dt : (g3 == 1 ? P.dt_m : 0.0);
      const int rA = g3 == 0 ? 6 + a1 : a, rB = g3 == 0 ? 6 + a2 : a;
      const double sA = g3 == 0 ? pick3(S3[0], S3[1], S3[2], a2) : 1.0, sB = g3 == 0 ? pick3(S3[0], S3[1], S3[2], a1) : 0.0;
      const double wxi = pick9(P.Wx, i);
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        double v = sp_ld(xs, X_Y + i * n + NA + c);
        v = fma(mul, fma(sp_ld(xs, X_Y + rA * n + NA + c), sA, -(sp_ld(xs, X_Y + rB * n + NA + c) * sB)), v);
        if (c == i) {
          v += wxi;
          if (MODE == MODE_ADMM && k >= 1 && i >= 6) v += S.rhok;
        }
        if (MODE == MODE_PMM && c >= 6 && kap && i >= 6) v += sp_ld(xs, X_KX + 3 + 3 * (i - 6) + (c - 6));
        Ts[t][c] = v;
      }
    }
  }
  CMPC_FK(1);
  // ---- phase 3: sweep the control pivots, one per trip of a rolled loop: the body is small enough to stay in
  // the instruction cache.  The pivot column always sits in register column 0: every trip writes the updated
  // row back rotated by one column (the values are fresh, so the rotation costs no moves); after na trips the
  // order is the original one.  Per pivot the owners publish their entry of the pivot column (double-buffered;
  // the control part twice, back to back, so that the rotated window pv .. pv + na - 1 is contiguous), everybody
  // reads the column -- it is the pivot row as well -- and updates its rows.
#pragma unroll 1
  for (int pv = 0; pv < NA; ++pv) {
    const int cb = X_CX + (pv & 1) * X_CXS;
#pragma unroll
    for (int t = 0; t < RT; ++t) {
      const int rr = q + NL * t;
      if (rr < n) sp_st(xs, cb + NA + rr, Tf[t][0]);    // control rows: second copy; state rows: 2 na + i
      if (rr < NA) sp_st(xs, cb + rr, Tf[t][0]);
    }
    team_sync(I);
    const ScratchPtr xw = sp_at(xs, s_off(cb + pv));
    double c[NA > 0 ? NA : 1], c9[9];
#pragma unroll
    for (int j = 0; j < NA; ++j) c[j] = sp_ld(xw, j);
#pragma unroll
    for (int cc = 0; cc < 9; ++cc) c9[cc] = sp_ld(xs, cb + 2 * NA + cc);
    const double piv = c[0];
    if (!(piv > 0.0) && on) S.fail = 1;
    const double ip = 1.0 / piv;
#pragma unroll
    for (int t = 0; t < RT; ++t) {
      const int rr = q + NL * t;
      const double bc = Tf[t][0] * ip;
      double nw[NA > 0 ? NA : 1];
      if (rr == pv) {   // the pivot row: scaled column, -1/pivot on the diagonal
        nw[0] = -ip;
#pragma unroll
        for (int j = 1; j < NA; ++j) nw[j] = c[j] * ip;
      } else {
        nw[0] = bc;
#pragma unroll
        for (int j = 1; j < NA; ++j) nw[j] = fma(-bc, c[j], Tf[t][j]);
        if (NL * t + NL - 1 >= NA) {   // state columns (state rows only)
#pragma unroll
          for (int cc = 0; cc < 9; ++cc) Ts[t][cc] = fma(-bc, c9[cc], Ts[t][cc]);
        }
      }
#pragma unroll
      for (int j = 0; j + 1 < NA; ++j) Tf[t][j] = nw[j + 1];
      Tf[t][NA > 0 ? NA - 1 : 0] = nw[0];
    }
  }
  CMPC_FK(2);
  // ---- factor record [Hn; Kt] and P_k (lower triangle, mirrored: P stays exactly symmetric)
#pragma unroll
  for (int t = 0; t < RT; ++t) {
    const int rr = q + NL * t;
    if (rr < n) {
      if (on) {
#pragma unroll
        for (int l = 0; l < NA; ++l) CMPC_R(w, L.hn + rr * NAP + l) = Tf[t][l];
      }
      if (rr >= NA) {
        const int i = rr - NA;
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          if (c <= i) {
            sp_st(xs, X_PS + i * 9 + c, Ts[t][c]);
            sp_st(xs, X_PS + c * 9 + i, Ts[t][c]);
          }
        }
      }
    }
  }
  team_sync(I);
  CMPC_FK(3);
#undef CMPC_FK
}

template <int MODE, bool FAST>
CMPC_OP void factor_op(const Params& P_in, TileCtx& T, const Inst& I_in, Sv& S_in, bool on) {
  // local copies: the reference arguments live in the caller's frame, which every generic store
  // of the operation could alias (the compiler would reload them after each one)
  const Params P = P_in;
  const Inst I = I_in;
  Sv S = S_in;
  constexpr int SK = MODE == MODE_ADMM ? SK_FAC_ADMM : SK_FAC_PMM;
  constexpr int NS = 0;
  constexpr Lay L = lay_of(0, !FAST);
  const int N = P.N;
  const double rho_e = (MODE == MODE_ADMM) ? S.rhoe : S.rhoep;
  const ScratchPtr xs = scratch_of(T, I);
  KnotStream<SK, !FAST> ks;
  ks.open(T, I, N, N + 1, -1);
  {   // P_N (every lane of the team writes the same values)
    const StagedPtr r = ks.acquire();
    ks.peek();
    double Pm[81];
#pragma unroll
    for (int i = 0; i < 81; ++i) Pm[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) Pm[10 * i] = P.Wx[i] + rho_e;
    if (MODE == MODE_ADMM) {
#pragma unroll
      for (int i = 6; i < 9; ++i) Pm[10 * i] += S.rhok;
    } else if (S.kap) {
      double kM[9], kl[3];
      const double kb[3] = {CMPC_S(r, L.xb + 6), CMPC_S(r, L.xb + 7), CMPC_S(r, L.xb + 8)};
      const double yk[4] = {CMPC_S(r, L.yk), CMPC_S(r, L.yk + 1), CMPC_S(r, L.yk + 2), CMPC_S(r, L.yk + 3)};
      pmm_kappa_terms(P, S, CMPC_SI(r, L.meta, 1), kb, yk, kM, kl);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Pm[(6 + i) * 9 + 6 + j] += (j <= i) ? kM[3 * i + j] : kM[3 * j + i];
    }
#pragma unroll
    for (int i = 0; i < 81; ++i) sp_st(xs, X_PS + i, Pm[i]);
    ks.release();
    team_sync(I);
  }
  // runs of equal slot count: one specialisation of the knot step per inner loop
#define CMPC_FAC_RUN(NS_)                                                             \
  do {                                                                                \
    const StagedPtr r = ks.acquire();                                                 \
    ks.peek();                                                                        \
    factor_knot<NS_, MODE, FAST, SK>(P, S, r, rec_of(T, I, k), I, k, xs, on, CMPC_PROF_PTR(T)); \
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
  if (ks.close(T)) S_in.lost = 1;
  if (S.fail) S_in.fail = 1;
}

// ---------------------------------------------------------------- backward sweep (linear term)
// p_N = qx_N;  g = p + Pc;  hu = ru + B'g;  d = Hn hu;  p = qx + A'g + K'hu  (the record holds Kt = -K').
// qx = -Wx xbar (+ kappa / terminal penalty terms), ru = friction penalty terms.
// Team work split: phase 1 the owners of the control rows publish hu; phase 2 the rows of [Hn; K'] hu
// (na + 9 row tasks) by their owners, who write d to the record and the new p to the scratch.
// kl[3]: linear term of the kappa rows (replicated in every lane)
template <int NS, int MODE, bool FAST, int SK>
CMPC_HD void kappa_linear_term(const Params& P, const Sv& S, StagedPtr r, const Inst& I, int pm, double* kl) {
  constexpr Lay L = lay_of(NS, !FAST);
  const double kb[3] = {CMPC_S(r, L.xb + 6), CMPC_S(r, L.xb + 7), CMPC_S(r, L.xb + 8)};
#pragma unroll
  for (int a = 0; a < 3; ++a) kl[a] = 0.0;
  if (MODE == MODE_ADMM) {
    const double vk[3] = {CMPC_S(r, L.vk), CMPC_S(r, L.vk + 1), CMPC_S(r, L.vk + 2)};
    double w[3];
    prox_kappa(S, vk, kb, w);
#pragma unroll
    for (int a = 0; a < 3; ++a) kl[a] = -S.rhok * (w[a] + w[a] - vk[a]);
  } else if (S.kap) {
    const double yk[4] = {CMPC_S(r, L.yk), CMPC_S(r, L.yk + 1), CMPC_S(r, L.yk + 2), CMPC_S(r, L.yk + 3)};
    double kM[9];
    pmm_kappa_terms(P, S, pm, kb, yk, kM, kl);
  }
}

// One run of knots with the same slot count.  Everything that depends only on the lane (which rows it
// owns, their kind, their friction constants) is formed once, before the knots; the knot step itself is
// branch-free: rows past the end are clamped to a valid row and only their stores are predicated, so that
// the independent dot-product chains of the lane's rows interleave.
template <int NS, int MODE, bool FAST, int SK, class KS>
CMPC_HD void bwd_run(const Params& P, const Sv& S, const TileCtx& T, const Inst& I, KS& ks, ScratchPtr xs, int& k_io, int& buf_io, bool on) {
  constexpr Lay L = lay_of(NS, !FAST);
  constexpr int NA = 3 * NS, n = NA + 9, NAP = L.nap;
  constexpr int CT = (NA + NL - 1) / NL, RT = (n + NL - 1) / NL;
  constexpr int CT1 = CT > 0 ? CT : 1;
  const int q = sub_of(I);
  // the lane's control rows j = q + NL t  (hu_j): pre-scaled offsets of everything the row touches
  int cs4[CT1];
  bool cok[CT1];
  double cg[CT1][4];
  SOff o_vf[CT1], o_p0[CT1], o_pA[CT1], o_pB[CT1], o_c0[CT1], o_cA[CT1], o_cB[CT1], o_dA[CT1], o_dB[CT1], o_g[CT1], o_gs[CT1], o_ux[CT1];
#pragma unroll
  for (int t = 0; t < CT; ++t) {
    const int j = q + NL * t;
    cok[t] = j < NA;
    const int jc = cok[t] ? j : 0;
    const int s = jc / 3, a = jc - 3 * s, a1 = a == 2 ? 0 : a + 1, a2 = a == 0 ? 2 : a - 1;
    cs4[t] = 4 * s;
    cg[t][0] = a == 0 ? 1.0 : (a == 2 ? -P.kf : 0.0); cg[t][1] = a == 0 ? -1.0 : (a == 2 ? -P.kf : 0.0);
    cg[t][2] = a == 1 ? 1.0 : (a == 2 ? -P.kf : 0.0); cg[t][3] = a == 1 ? -1.0 : (a == 2 ? -P.kf : 0.0);
    o_vf[t] = s_off(so_of<SK, NS, !FAST>(MODE == MODE_ADMM ? L.vf : L.yf) + 4 * s);
    o_p0[t] = s_off(3 + a); o_pA[t] = s_off(6 + a1); o_pB[t] = s_off(6 + a2);
    o_c0[t] = s_off(so_of<SK, NS, !FAST>(L.pc) + 3 + a); o_cA[t] = s_off(so_of<SK, NS, !FAST>(L.pc) + 6 + a1);
    o_cB[t] = s_off(so_of<SK, NS, !FAST>(L.pc) + 6 + a2);
    o_dA[t] = s_off(so_of<SK, NS, !FAST>(L.d) + 3 * s + a2); o_dB[t] = s_off(so_of<SK, NS, !FAST>(L.d) + 3 * s + a1);
    if (WR) { o_dA[t] = s_off(so_of<SK, NS, !FAST>(L.d) + 9 * s + 3 * a); o_dB[t] = s_off(so_of<SK, NS, !FAST>(L.d) + 9 * s + 3 * a + 1); }   // column a of M
    o_g[t] = s_off(so_of<SK, NS, !FAST>(L.g) + s * GS + a);   // G[0][a] of the slot (rows 3 apart)
    o_gs[t] = s_off(so_of<SK, NS, !FAST>(L.g) + s * GS);      // the slot's table: e2 at 12 + row, ub at 16 + row
    o_ux[t] = s_off(X_UX + jc);
    CMPC_OPAQUE(cs4[t]); CMPC_OPAQUE(o_vf[t]); CMPC_OPAQUE(o_p0[t]); CMPC_OPAQUE(o_pA[t]); CMPC_OPAQUE(o_pB[t]);
    CMPC_OPAQUE(o_c0[t]); CMPC_OPAQUE(o_cA[t]); CMPC_OPAQUE(o_cB[t]); CMPC_OPAQUE(o_dA[t]); CMPC_OPAQUE(o_dB[t]);
    CMPC_OPAQUE(o_g[t]); CMPC_OPAQUE(o_gs[t]); CMPC_OPAQUE(o_ux[t]);
  }
  // the lane's rows rr = q + NL t of [Hn; Kt] hu: control rows give d, state rows the new p
  int rrow[RT];
  bool rok[RT], rst[RT], rg0[RT], rkx[RT];
  double rmul[RT], rwx[RT];
  SOff o_m[RT], o_pi[RT], o_qA[RT], o_qB[RT], o_ei[RT], o_eA[RT], o_eB[RT], o_sA[RT], o_sB[RT], o_xb[RT], o_kx[RT];
#pragma unroll
  for (int t = 0; t < RT; ++t) {
    const int rr = q + NL * t;
    rok[t] = rr < n;
    rrow[t] = rok[t] ? rr : n - 1;
    rst[t] = rrow[t] >= NA;
    const int i = rst[t] ? rrow[t] - NA : 0;
    const int g3 = i / 3, a = i - 3 * g3, a1 = a == 2 ? 0 : a + 1, a2 = a == 0 ? 2 : a - 1;
    const int rA = g3 == 0 ? 6 + a1 : a, rB = g3 == 0 ? 6 + a2 : a;
    rg0[t] = g3 == 0;
    rkx[t] = i >= 6;
    rmul[t] = g3 == 0 ? P.dt : (g3 == 1 ? P.dt_m : 0.0);   // (A'g)_i = g_i + mul (g[rA] sA - g[rB] sB)
    rwx[t] = pick9(P.Wx, i);
    if (WR) rwx[t] *= P.qs;
    o_m[t] = s_off(so_of<SK, NS, !FAST>(L.hn) + rrow[t] * NAP);
    o_pi[t] = s_off(i); o_qA[t] = s_off(rA); o_qB[t] = s_off(rB);
    o_ei[t] = s_off(so_of<SK, NS, !FAST>(L.pc) + i); o_eA[t] = s_off(so_of<SK, NS, !FAST>(L.pc) + rA);
    o_eB[t] = s_off(so_of<SK, NS, !FAST>(L.pc) + rB);
    o_sA[t] = s_off(so_of<SK, NS, !FAST>(L.s) + a2); o_sB[t] = s_off(so_of<SK, NS, !FAST>(L.s) + a1);
    o_xb[t] = s_off(so_of<SK, NS, !FAST>(L.xb) + i);
    o_kx[t] = s_off(X_KX + (i >= 6 ? i - 6 : 0));
    CMPC_OPAQUE(rrow[t]); CMPC_OPAQUE(o_m[t]); CMPC_OPAQUE(o_pi[t]); CMPC_OPAQUE(o_qA[t]); CMPC_OPAQUE(o_qB[t]);
    CMPC_OPAQUE(o_ei[t]); CMPC_OPAQUE(o_eA[t]); CMPC_OPAQUE(o_eB[t]); CMPC_OPAQUE(o_sA[t]); CMPC_OPAQUE(o_sB[t]);
    CMPC_OPAQUE(o_xb[t]); CMPC_OPAQUE(o_kx[t]);
  }
  int k = k_io, buf = buf_io;
#if defined(CMPC_PROFILE) && defined(__CUDACC__)
#define CMPC_CK(i) do { const long long t_ = clock64(); T.prof[16 + (i)] += t_ - ck_; ck_ = t_; } while (0)
  long long ck_ = clock64();
#else
#define CMPC_CK(i)
#endif
  // Every phase is written loads -> arithmetic -> stores: the shared-memory accesses keep their program
  // order (a store may alias a later load as far as the compiler can tell), so a store in the middle of a
  // phase would serialise the rows of the lane.
  do {
    const StagedPtr r = ks.acquire();
    ks.peek();
    CMPC_CK(0);
    double* w = rec_of(T, I, k);
    const ScratchPtr pin = sp_at(xs, s_off(X_PX + buf * 9)), pout = sp_at(xs, s_off(X_PX + (buf ^ 1) * 9));
    // ---- phase 1 loads
    const int mt_b = CMPC_SI(r, L.meta, 0);
    const int nsl4 = 4 * (mt_b & 7);
    const int pm = (MODE == MODE_PMM) ? CMPC_SI(r, L.meta, 1) : 0;
    const double kb[3] = {CMPC_S(r, L.xb + 6), CMPC_S(r, L.xb + 7), CMPC_S(r, L.xb + 8)};
    double g6w[3] = {0.0, 0.0, 0.0};   // wrench model: g[6..8] = p + Pc, every control row needs all three
    if (WR) {
#pragma unroll
      for (int j = 0; j < 3; ++j) g6w[j] = sp_ld(pin, 6 + j) + CMPC_S(r, L.pc + 6 + j);
    }
    double kin[4] = {0.0, 0.0, 0.0, 0.0};   // ADMM: vk[3]; multiplier mode: yk[4]
    if (MODE == MODE_ADMM) {
#pragma unroll
      for (int a = 0; a < 3; ++a) kin[a] = CMPC_S(r, L.vk + a);
    } else {
#pragma unroll
      for (int a = 0; a < 4; ++a) kin[a] = CMPC_S(r, L.yk + a);
    }
    double fv[CT1][4], fg[CT1][4], fe[CT1][4], fu[CT1][4];
    double p0[CT1], pA[CT1], pB[CT1], c0[CT1], cA[CT1], cB[CT1], dA[CT1], dB[CT1];
#pragma unroll
    for (int t = 0; t < CT; ++t) {
      const StagedPtr rv = sp_at(r, o_vf[t]), rg = sp_at(r, o_g[t]);
#pragma unroll
      for (int row = 0; row < 4; ++row) {
        fv[t][row] = sp_ld(rv, row);
        fg[t][row] = cg[t][row];
        fe[t][row] = row < 2 ? P.e2[0] : P.e2[2];
        fu[t][row] = 0.0;
        if (!FAST) {
          fg[t][row] = sp_ld(rg, row * 3);
          fe[t][row] = sp_ld(sp_at(r, o_gs[t]), 12 + row);
          fu[t][row] = sp_ld(sp_at(r, o_gs[t]), 16 + row);
        }
      }
      p0[t] = sp_ld(sp_at(pin, o_p0[t]), 0); c0[t] = sp_ld(sp_at(r, o_c0[t]), 0);
      pA[t] = sp_ld(sp_at(pin, o_pA[t]), 0); cA[t] = sp_ld(sp_at(r, o_cA[t]), 0);
      pB[t] = sp_ld(sp_at(pin, o_pB[t]), 0); cB[t] = sp_ld(sp_at(r, o_cB[t]), 0);
      dA[t] = sp_ld(sp_at(r, o_dA[t]), 0);
      dB[t] = sp_ld(sp_at(r, o_dB[t]), 0);
      if (WR) pA[t] = sp_ld(sp_at(r, o_dB[t]), 1);   // third entry of the column of M
    }
    // ---- phase 1 arithmetic: linear term of the kappa rows (replicated), hu of the lane's control rows
    double kl[3] = {0.0, 0.0, 0.0};
    if (k >= 1) {
      if (MODE == MODE_ADMM) {
        double wv[3];
        prox_kappa(S, kin, kb, wv);
#pragma unroll
        for (int a = 0; a < 3; ++a) kl[a] = -S.rhok * (wv[a] + wv[a] - kin[a]);
      } else if (S.kap) {
        double kM[9];
        pmm_kappa_terms(P, S, pm, kb, kin, kM, kl);
      }
    }
    double huo[CT1];
#pragma unroll
    for (int t = 0; t < CT; ++t) {
      const double dtr = cs4[t] < nsl4 ? P.dt : 0.0;
      double tt[4];
#pragma unroll
      for (int row = 0; row < 4; ++row) {
        if (MODE == MODE_ADMM) {
          tt[row] = S.rho * fe[t][row] * fabs(fv[t][row]);                      // -(rho e2 w - y) = rho e2 |v|
          if (!FAST) tt[row] = fma(-S.rho * fe[t][row], fu[t][row], tt[row]);   // unshifted w = min(v, 0) + ub
        } else {
          tt[row] = ((pm >> (cs4[t] + row)) & 1) ? (FAST ? fv[t][row] : fma(-P.inv_delta, fu[t][row], fv[t][row])) : 0.0;
        }
      }
      const double o = fma(fg[t][3], tt[3], fma(fg[t][2], tt[2], fma(fg[t][1], tt[1], fg[t][0] * tt[0])));
      const double g0 = p0[t] + c0[t], gA = pA[t] + cA[t], gB = pB[t] + cB[t];
      if (WR) {
        const double sgm = ((mt_b >> (4 + (cs4[t] >> 1))) & 1) ? 0.0 : 1.0;
        huo[t] = dtr * fma(g6w[2], pA[t], fma(g6w[1], dB[t], fma(g6w[0], dA[t], sgm * g0))) + o;
      } else
      huo[t] = dtr * fma(gA, dA[t], fma(-gB, dB[t], g0)) + o;
    }
    CMPC_CK(1);
    // ---- phase 1 stores
#pragma unroll
    for (int a = 0; a < 3; ++a) sp_st(xs, X_KX + a, kl[a]);
#pragma unroll
    for (int t = 0; t < CT; ++t)
      if (cok[t]) sp_st(sp_at(xs, o_ux[t]), 0, huo[t]);
    team_sync(I);
    CMPC_CK(2);
    // ---- phase 2 loads: hu, the lane's rows of [Hn; Kt], the operands of the state rows
    double hu[NA > 0 ? NA : 1], Mr[RT][NA > 0 ? NA : 1];
#pragma unroll
    for (int l = 0; l < NA; ++l) hu[l] = sp_ld(xs, X_UX + l);
#pragma unroll
    for (int t = 0; t < RT; ++t) {
      const StagedPtr rm = sp_at(r, o_m[t]);
#pragma unroll
      for (int l = 0; l < NA; ++l) Mr[t][l] = sp_ld(rm, l);
    }
    double qi[RT], qA[RT], qB[RT], ei[RT], eA[RT], eB[RT], sA[RT], sB[RT], xbi[RT], kx[RT];
#pragma unroll
    for (int t = 0; t < RT; ++t) {
      if (NL * t + NL - 1 >= NA) {   // this row can be a state row
        qi[t] = sp_ld(sp_at(pin, o_pi[t]), 0); ei[t] = sp_ld(sp_at(r, o_ei[t]), 0);
        qA[t] = sp_ld(sp_at(pin, o_qA[t]), 0); eA[t] = sp_ld(sp_at(r, o_eA[t]), 0);
        qB[t] = sp_ld(sp_at(pin, o_qB[t]), 0); eB[t] = sp_ld(sp_at(r, o_eB[t]), 0);
        sA[t] = sp_ld(sp_at(r, o_sA[t]), 0);
        sB[t] = sp_ld(sp_at(r, o_sB[t]), 0);
        xbi[t] = sp_ld(sp_at(r, o_xb[t]), 0);
        kx[t] = sp_ld(sp_at(xs, o_kx[t]), 0);
      }
    }
    // ---- phase 2 arithmetic (two partial sums per row: shorter dependent chains)
    double acc[RT], pn[RT];
#pragma unroll
    for (int t = 0; t < RT; ++t) {
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int l = 0; l < NA; l += 2) {
        acc0 = fma(Mr[t][l], hu[l], acc0);
        if (l + 1 < NA) acc1 = fma(Mr[t][l + 1], hu[l + 1], acc1);
      }
      acc[t] = acc0 + acc1;
      pn[t] = 0.0;
      if (NL * t + NL - 1 >= NA) {   // p_i = qx_i + (A'g)_i + (K'hu)_i, Kt = -K'
        const double gi = qi[t] + ei[t], gA = qA[t] + eA[t], gB = qB[t] + eB[t];
        const double a_ = rg0[t] ? sA[t] : 1.0, b_ = rg0[t] ? sB[t] : 0.0;
        double pi = fma(rmul[t], fma(gA, a_, -(gB * b_)), gi) - acc[t] - rwx[t] * xbi[t];
        if (k >= 1 && rkx[t]) pi += kx[t];
        pn[t] = pi;
      }
    }
    // ---- phase 2 stores
#pragma unroll
    for (int t = 0; t < RT; ++t) {
      if (NL * t < NA) {             // this row can be a control row: d_j
        if (rok[t] && !rst[t] && on) CMPC_R(w, L.dv + rrow[t]) = acc[t];
      }
      if (NL * t + NL - 1 >= NA) {
        if (rok[t] && rst[t]) sp_st(sp_at(pout, o_pi[t]), 0, pn[t]);
      }
    }
    team_sync(I);
    CMPC_CK(3);
    buf ^= 1;
    ks.release();
    --k;
    CMPC_CK(4);
  } while (k >= 0 && T.ns(k) == NS);
  CMPC_CK(5);
#undef CMPC_CK
  k_io = k;
  buf_io = buf;
}

template <int MODE, bool FAST>
CMPC_OP void backward_op(const Params& P_in, TileCtx& T, const Inst& I_in, Sv& S_in, bool on) {
  const Params P = P_in;
  const Inst I = I_in;
  const Sv S = S_in;
  constexpr int SK = MODE == MODE_ADMM ? SK_BWD_ADMM : SK_BWD_PMM;
  const int N = P.N;
  const double rho_e = (MODE == MODE_ADMM) ? S.rhoe : S.rhoep;
  const ScratchPtr xs = scratch_of(T, I);
  KnotStream<SK, !FAST> ks;
  ks.open(T, I, N, N + 1, -1);
  int buf = 0;
  {   // p_N (every lane of the team writes the same values)
    constexpr int NS = 0;
    constexpr Lay L = lay_of(0, !FAST);
    const StagedPtr r = ks.acquire();
    ks.peek();
    double kl[3];
    kappa_linear_term<0, MODE, FAST, SK>(P, S, r, I, CMPC_SI(r, L.meta, 1), kl);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      double p = -((WR ? P.qs * P.Wx[i] : P.Wx[i]) * CMPC_S(r, L.xb + i)) - (rho_e * I.xf[i] - S.ye[i]);
      if (i >= 6) p += kl[i - 6];
      sp_st(xs, X_PX + i, p);
    }
    ks.release();
    team_sync(I);
  }
  for (int k = N - 1; k >= 0;) {   // runs of equal slot count: one specialisation of the knot step per run
    switch (T.ns(k)) {
      case 0: bwd_run<0, MODE, FAST, SK>(P, S, T, I, ks, xs, k, buf, on); break;
      case 1: bwd_run<1, MODE, FAST, SK>(P, S, T, I, ks, xs, k, buf, on); break;
      case 2: bwd_run<2, MODE, FAST, SK>(P, S, T, I, ks, xs, k, buf, on); break;
      case 3: bwd_run<3, MODE, FAST, SK>(P, S, T, I, ks, xs, k, buf, on); break;
      default: bwd_run<4, MODE, FAST, SK>(P, S, T, I, ks, xs, k, buf, on); break;
    }
  }
  if (ks.close(T)) S_in.lost = 1;
}

// ---------------------------------------------------------------- forward sweep + local updates
// u~ = K x~ + d,  x~+ = A x~ + B u~ + c.
// ADMM kinds: relaxation, friction / kappa / terminal updates of (w, y) stored as v, all
// knot-local; with CHECK also OSQP's residuals (oracle/device_model.py iterate()).
// Multiplier kind: y += (1/delta) row on the active rows, solution record, primal residual,
// and (upd) the corrected friction active set: violated rows join, rows with a negative
// multiplier leave; the number of changes is accumulated in nchg.
// COPY: read-only LQR roll-out of the ADMM iterate into the solution record.
// Team work split: x~ is replicated; phase 1 the owners of the control rows publish u~; phase 2 every
// lane forms x~+, the friction rows (4 per slot) are updated by their owners.
struct Res { double pri, dua, npri, ndua; int kbad; };

template <int NS, int KIND, bool FAST, int SK>
CMPC_HD void fwd_state(const Params& P, Sv& S, Res& R, StagedPtr r, double* w, const Inst& I, int k, const double* x) {
  constexpr Lay L = lay_of(NS, !FAST);
  constexpr bool ADMM = KIND == FW_ADMM || KIND == FW_ADMM_CHECK, CHK = KIND == FW_ADMM_CHECK;
  constexpr bool PMMK = KIND == FW_PMM, COPY = KIND == FW_COPY;
  const int N = P.N;
  const double al = ADMM ? P.alpha : 1.0;
  if (PMMK || COPY) {
#pragma unroll
    for (int i = 0; i < 9; ++i) CMPC_R(w, L.x + i) = x[i];
  }
  // the kappa rows (a read-modify-write of the record) belong to lane 0 of the team; its residual
  // terms reach the other lanes through the team maxima at the end of the sweep
  const bool lead = sub_of(I) == 0;
  double rdx[3] = {0.0, 0.0, 0.0};
  if (k >= 1 && lead) {
    if (ADMM) {
      const double kb[3] = {CMPC_S(r, L.xb + 6), CMPC_S(r, L.xb + 7), CMPC_S(r, L.xb + 8)};
      const double vk[3] = {CMPC_S(r, L.vk), CMPC_S(r, L.vk + 1), CMPC_S(r, L.vk + 2)};
      double wk[3], vn[3];
      prox_kappa(S, vk, kb, wk);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        vn[a] = fma(al, x[6 + a], fma(1.0 - al, wk[a], vk[a] - wk[a]));
        CMPC_R(w, L.vk + a) = vn[a];
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
      // The trust-region rows enter the polish by the branch the prox took in the ADMM iterate (inside the
      // L1 ball / linear penalty / on the surface).  Unlike the friction rows they are not corrected inside
      // a polish attempt, so the point must CONFIRM the branch: inside -> |dkappa|_1 <= r; linear ->
      // |dkappa|_1 >= r with the assumed signs; surface -> multiplier in [0, omega].  Otherwise the attempt
      // is void (R.kbad) and ADMM goes on to a better guess.
      const int pm = S.kap ? CMPC_SI(r, L.meta, 1) : 0;
      const int br = (pm >> 16) & 3;
      const double inv = P.inv_delta;
      const double ktol = 1e-9 * (1.0 + S.radius);
      double accv = -S.radius, dk1 = 0.0;
      bool flip = false;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int code = (pm >> (18 + 2 * i)) & 3;
        const double sgn = code == 1 ? 1.0 : (code == 2 ? -1.0 : 0.0);
        const double dk = x[6 + i] - CMPC_S(r, L.xb + 6 + i);
        dk1 += fabs(dk);
        if (br != 0) {
          if (code == 0) {   // pinned component: an equality row of the polish, part of its primal residual
            CMPC_R(w, L.yk + i) = CMPC_S(r, L.yk + i) + inv * dk;
            R.pri = fmax(R.pri, fabs(dk));
          } else if (sgn * dk < -ktol) flip = true;
          accv += sgn * dk;
        }
      }
      if (br == 0) {
        if (dk1 > S.radius + ktol) R.kbad = 1;
      } else if (br == 1) {
        if (dk1 < S.radius - ktol || flip) R.kbad = 1;
      } else {
        const double ys = CMPC_S(r, L.yk + 3) + inv * accv;
        CMPC_R(w, L.yk + 3) = ys;
        R.pri = fmax(R.pri, fabs(accv));   // the surface row |dkappa|_1 = r is an equality row of the polish too
        if (flip || ys < -1e-9 * (1.0 + S.weight) || ys > S.weight * (1.0 + 1e-9)) R.kbad = 1;
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
  if (CHK && k >= 1 && lead) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double Px = P.Wx[i] * x[i];
      const double rd = i >= 6 ? rdx[i - 6] : 0.0;
      const double aty = rd - Px + (WR ? P.qs * P.Wx[i] : P.Wx[i]) * CMPC_S(r, L.xb + i);   // (A'y)_x = r_d - P x - q,  q = -Wx xbar
      R.dua = fmax(R.dua, fabs(rd));
      R.ndua = fmax(R.ndua, fmax(fabs(Px), fabs(aty)));
    }
  }
}

template <int NS, int KIND, bool FAST, int SK, class KS>
CMPC_HD void fwd_run(const Params& P, Sv& S, Res& R, const TileCtx& T, const Inst& I, KS& ks, ScratchPtr xs, double* x,
                     int& k_io, bool on, bool upd, int& nchg) {
  constexpr Lay L = lay_of(NS, !FAST);
  constexpr int NA = 3 * NS, NAP = L.nap;
  constexpr int CT = (NA + NL - 1) / NL, FT = (4 * NS + NL - 1) / NL;
  constexpr bool ADMM = KIND == FW_ADMM || KIND == FW_ADMM_CHECK, CHK = KIND == FW_ADMM_CHECK;
  constexpr bool PMMK = KIND == FW_PMM, COPY = KIND == FW_COPY;
  const int q = sub_of(I);
  const int N = P.N;
  const double al = ADMM ? P.alpha : 1.0;
  const double inv = P.inv_delta;
  // the lane's control rows j = q + NL t (u_j) and friction rows bit = q + NL t (4 per slot): pre-scaled offsets
  constexpr int CT1 = CT > 0 ? CT : 1, FT1 = FT > 0 ? FT : 1;
  int cj[CT1], cs[CT1], ca[CT1];
  bool cok[CT1];
  double cg[CT1][4];
  SOff o_dv[CT1], o_kt[CT1], o_ux[CT1], o_gc[CT1], o_dx[CT1];
#pragma unroll
  for (int t = 0; t < CT; ++t) {
    const int j = q + NL * t;
    cok[t] = j < NA;
    cj[t] = cok[t] ? j : 0;
    cs[t] = cj[t] / 3;
    ca[t] = cj[t] - 3 * cs[t];
    cg[t][0] = ca[t] == 0 ? 1.0 : (ca[t] == 2 ? -P.kf : 0.0); cg[t][1] = ca[t] == 0 ? -1.0 : (ca[t] == 2 ? -P.kf : 0.0);
    cg[t][2] = ca[t] == 1 ? 1.0 : (ca[t] == 2 ? -P.kf : 0.0); cg[t][3] = ca[t] == 1 ? -1.0 : (ca[t] == 2 ? -P.kf : 0.0);
    o_dv[t] = s_off(so_of<SK, NS, !FAST>(L.dv) + cj[t]);
    o_kt[t] = s_off(so_of<SK, NS, !FAST>(L.kt) + cj[t]);
    o_ux[t] = s_off(X_UX + cj[t]);
    o_gc[t] = s_off(so_of<SK, NS, !FAST>(L.g) + cs[t] * GS + ca[t]);
    o_dx[t] = s_off(X_DX + 4 * cs[t]);
    CMPC_OPAQUE(cj[t]); CMPC_OPAQUE(cs[t]); CMPC_OPAQUE(ca[t]);
    CMPC_OPAQUE(o_dv[t]); CMPC_OPAQUE(o_kt[t]); CMPC_OPAQUE(o_ux[t]); CMPC_OPAQUE(o_gc[t]); CMPC_OPAQUE(o_dx[t]);
  }
  int fb[FT1];
  bool fok[FT1];
  double fgx[FT1], fgy[FT1], fe2[FT1];
  SOff o_fv[FT1], o_fu[FT1], o_fg[FT1], o_fe[FT1], o_fd[FT1];
#pragma unroll
  for (int t = 0; t < FT; ++t) {
    const int bit = q + NL * t;
    fok[t] = bit < 4 * NS;
    fb[t] = fok[t] ? bit : 0;
    const int s = fb[t] >> 2, row = fb[t] & 3;
    fgx[t] = row == 0 ? 1.0 : (row == 1 ? -1.0 : 0.0);
    fgy[t] = row == 2 ? 1.0 : (row == 3 ? -1.0 : 0.0);
    fe2[t] = row < 2 ? P.e2[0] : P.e2[2];
    o_fv[t] = s_off(so_of<SK, NS, !FAST>((KIND == FW_ADMM || KIND == FW_ADMM_CHECK) ? L.vf : L.yf) + fb[t]);
    o_fu[t] = s_off(X_UX + 3 * s);
    o_fg[t] = s_off(so_of<SK, NS, !FAST>(L.g) + s * GS + row * 3);
    o_fe[t] = s_off(so_of<SK, NS, !FAST>(L.g) + s * GS + 12 + row);   // ub 4 further
    o_fd[t] = s_off(X_DX + fb[t]);
    CMPC_OPAQUE(fb[t]); CMPC_OPAQUE(o_fv[t]); CMPC_OPAQUE(o_fu[t]); CMPC_OPAQUE(o_fg[t]); CMPC_OPAQUE(o_fe[t]); CMPC_OPAQUE(o_fd[t]);
  }
  int k = k_io;
  do {
    const StagedPtr r = ks.acquire();
    ks.peek();
    double* w = rec_of(T, I, k);
    int* imw = meta_of(T, I, k, L.meta);
    // ---- phase 1 loads (issued before the state part, whose latency they overlap)
    double dvj[CT > 0 ? CT : 1], ktc[CT > 0 ? CT : 1][9];
#pragma unroll
    for (int t = 0; t < CT; ++t) {
      dvj[t] = sp_ld(sp_at(r, o_dv[t]), 0);
      const StagedPtr rk = sp_at(r, o_kt[t]);
#pragma unroll
      for (int i = 0; i < 9; ++i) ktc[t][i] = sp_ld(rk, i * NAP);
    }
    if (on) fwd_state<NS, KIND, FAST, SK>(P, S, R, r, w, I, k, x);
    const double tolc = P.as_tol * (1.0 + S.npri);   // row-violation threshold (npri of the previous sweep)
    // ---- phase 1: controls u~ = K x + d of the lane's rows (Kt = -K'; two partial sums per row)
    double uo[CT > 0 ? CT : 1];
#pragma unroll
    for (int t = 0; t < CT; ++t) {
      double v0 = dvj[t], v1 = 0.0;
#pragma unroll
      for (int i = 0; i < 9; i += 2) {
        v0 = fma(-ktc[t][i], x[i], v0);
        if (i + 1 < 9) v1 = fma(-ktc[t][i + 1], x[i + 1], v1);
      }
      uo[t] = v0 + v1;
    }
#pragma unroll
    for (int t = 0; t < CT; ++t) {
      if (cok[t]) {
        sp_st(sp_at(xs, o_ux[t]), 0, uo[t]);
        if ((PMMK || COPY) && on) CMPC_R(w, L.u + cj[t]) = uo[t];
      }
    }
    team_sync(I);
    // ---- phase 2 loads: u~, stage data, the lane's friction rows
    double u[NA > 0 ? NA : 1], dsl[NS > 0 ? NS : 1][WR ? 9 : 3];
#pragma unroll
    for (int j = 0; j < NA; ++j) u[j] = sp_ld(xs, X_UX + j);
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int a = 0; a < (WR ? 9 : 3); ++a) dsl[s][a] = CMPC_S(r, L.d + (WR ? 9 : 3) * s + a);
    const int mt_f = WR ? CMPC_SI(r, L.meta, 0) : 0;
    const double S3[3] = {CMPC_S(r, L.s), CMPC_S(r, L.s + 1), CMPC_S(r, L.s + 2)};
    const double ckv[3] = {CMPC_S(r, L.ck), CMPC_S(r, L.ck + 1), CMPC_S(r, L.ck + 2)};
    double fvv[FT > 0 ? FT : 1], fu0[FT > 0 ? FT : 1], fu1[FT > 0 ? FT : 1], fu2[FT > 0 ? FT : 1];
    double tgx[FT > 0 ? FT : 1], tgy[FT > 0 ? FT : 1], tgz[FT > 0 ? FT : 1], te2[FT > 0 ? FT : 1], tub[FT > 0 ? FT : 1];
    if (!COPY) {
#pragma unroll
      for (int t = 0; t < FT; ++t) {
        const ScratchPtr xu = sp_at(xs, o_fu[t]);
        fvv[t] = sp_ld(sp_at(r, o_fv[t]), 0);
        fu0[t] = sp_ld(xu, 0); fu1[t] = sp_ld(xu, 1); fu2[t] = sp_ld(xu, 2);
        tgx[t] = fgx[t]; tgy[t] = fgy[t]; tgz[t] = -P.kf; te2[t] = fe2[t]; tub[t] = 0.0;
        if (!FAST) {
          const StagedPtr rg = sp_at(r, o_fg[t]), re = sp_at(r, o_fe[t]);
          tgx[t] = sp_ld(rg, 0); tgy[t] = sp_ld(rg, 1); tgz[t] = sp_ld(rg, 2);
          te2[t] = sp_ld(re, 0); tub[t] = sp_ld(re, 4);
        }
      }
    }
    // ---- phase 2: next state (replicated), friction rows of the lane
    double sF[3] = {0.0, 0.0, 0.0}, sT[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int s = 0; s < NS; ++s) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int a1 = nxt3(a), a2 = prv3(a);
        if (WR) {   // padded slots carry M = 0 and u = 0
          if (!((mt_f >> (4 + 2 * s)) & 1)) sF[a] = sF[a] + u[3 * s + a];
          sT[a] = sT[a] + fma(dsl[s][6 + a], u[3 * s + 2], fma(dsl[s][3 + a], u[3 * s + 1], dsl[s][a] * u[3 * s]));
          continue;
        }
        sF[a] = sF[a] + u[3 * s + a];
        sT[a] = sT[a] + fma(dsl[s][a1], u[3 * s + a2], -(dsl[s][a2] * u[3 * s + a1]));
      }
    }
    double xn[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int a1 = nxt3(a), a2 = prv3(a);
      xn[a] = fma(P.dt_m, x[3 + a], x[a]);
      xn[3 + a] = x[3 + a] + fma(P.dt, sF[a], a == 2 ? P.dtmg : 0.0);
      xn[6 + a] = fma(P.dt * S3[a1], x[a2], fma(-P.dt * S3[a2], x[a1], x[6 + a])) + fma(P.dt, sT[a], ckv[a]);
    }
    if (!COPY) {
      const int mt = CMPC_SI(r, L.meta, 0);
      const int pm = PMMK ? CMPC_SI(r, L.meta, 1) : 0;
      int newpm = 0;
#pragma unroll
      for (int t = 0; t < FT; ++t) {
        const int bit = fb[t];
        const double e2 = te2[t], ub = tub[t];
        double cf = fma(tgz[t], fu2[t], fma(tgy[t], fu1[t], tgx[t] * fu0[t]));
        if (!FAST) cf -= ub;
        if (ADMM) {
          const double v = fvv[t];
          const double w0 = fmin(v, 0.0), y0 = fmax(v, 0.0);
          const double vn = fma(al, cf, fma(1.0 - al, w0, y0));
          if (fok[t] && on) CMPC_R(w, L.vf + bit) = vn;
          if (CHK && fok[t]) {
            const double wn = fmin(vn, 0.0);
            R.pri = fmax(R.pri, fabs(cf - wn));
            R.npri = fmax(R.npri, FAST ? fmax(fabs(cf), fabs(wn)) : fmax(fabs(cf + ub), fabs(wn + ub)));
            sp_st(sp_at(xs, o_fd[t]), 0, S.rho * e2 * (fmax(vn, 0.0) - y0 - cf + w0));
          }
        } else if (fok[t]) {
          const bool act = (pm >> bit) & 1;
          const double yn = fma(inv, cf, act ? fvv[t] : 0.0);
          R.pri = fmax(R.pri, act ? fabs(cf) : fmax(cf, 0.0));
          R.npri = fmax(R.npri, FAST ? fabs(cf) : fabs(cf + ub));
          if (upd) {
            const bool keep = act && !(yn < 0.0);
            const bool join = !act && (cf > tolc);
            const bool nb = keep || join;
            nchg += (nb != act) ? 1 : 0;
            if (on) CMPC_R(w, L.yf + bit) = keep ? yn : 0.0;
            newpm |= (nb ? 1 : 0) << bit;
          } else if (act) {
            if (on) CMPC_R(w, L.yf + bit) = yn;
          }
        }
      }
      if (PMMK) {   // (warp-collective: every lane takes part, also those whose instance does not update)
        newpm = team_or(I, xs, newpm);
        if (upd && on) imw[TL] = (pm & ~0xffff) | newpm;
      }
      if (CHK) {   // u rows of the stationarity residual: G' delta; norms of P u and A'y
        team_sync(I);
#pragma unroll
        for (int t = 0; t < CT; ++t) {
          if (cok[t]) {
            const int s = cs[t], a = ca[t];
            double gc[4];
#pragma unroll
            for (int row = 0; row < 4; ++row) gc[row] = FAST ? cg[t][row] : sp_ld(sp_at(r, o_gc[t]), row * 3);
            const ScratchPtr xd = sp_at(xs, o_dx[t]);
            const double rdu = fma(gc[3], sp_ld(xd, 3), fma(gc[2], sp_ld(xd, 2), fma(gc[1], sp_ld(xd, 1), gc[0] * sp_ld(xd, 0))));
            const int cid = (s < (mt & 7)) ? ((mt >> (4 + 2 * s)) & 3) : 0;
            double wsel = pick3(P.Wu[0], P.Wu[1], P.Wu[2], a);
            if (!FAST) {
#pragma unroll
              for (int c = 1; c < MAXC; ++c) wsel = (cid == c) ? pick3(P.Wu[3 * c], P.Wu[3 * c + 1], P.Wu[3 * c + 2], a) : wsel;
            }
            const double Pu = wsel * uo[t];
            R.dua = fmax(R.dua, fabs(rdu));
            R.ndua = fmax(R.ndua, fmax(fabs(Pu), fabs(rdu - Pu)));
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = xn[i];
    team_sync(I);   // the exchange words are rewritten by the next knot
    ks.release();
    ++k;
  } while (k < N && T.ns(k) == NS);
  k_io = k;
}

// commit: the instance wants the residuals of this sweep (an instance that did not ask for a check may
// ride along in the CHECK kind when a neighbour did; its iterate update is the same arithmetic)
template <int KIND, bool FAST>
CMPC_OP void forward_op(const Params& P_in, TileCtx& T, const Inst& I_in, Sv& S_in, bool on, bool commit, bool upd, int* changes) {
  const Params P = P_in;
  const Inst I = I_in;
  Sv S = S_in;
  constexpr bool CHK = KIND == FW_ADMM_CHECK, PMMK = KIND == FW_PMM;
  constexpr int SK = KIND == FW_PMM ? SK_FWD_PMM : (KIND == FW_COPY ? SK_FWD_COPY : SK_FWD_ADMM);
  const int N = P.N;
  const ScratchPtr xs = scratch_of(T, I);
  KnotStream<SK, !FAST> ks;
  ks.open(T, I, 0, N + 1, 1);
  Res R;
  R.pri = R.dua = R.npri = R.ndua = 0.0;
  R.kbad = 0;
  int nchg = 0;
  double x[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) x[i] = I.xi[i];
  for (int k = 0; k < N;) {   // runs of equal slot count: one specialisation of the knot step per run, x in registers
    switch (T.ns(k)) {
      case 0: fwd_run<0, KIND, FAST, SK>(P, S, R, T, I, ks, xs, x, k, on, upd, nchg); break;
      case 1: fwd_run<1, KIND, FAST, SK>(P, S, R, T, I, ks, xs, x, k, on, upd, nchg); break;
      case 2: fwd_run<2, KIND, FAST, SK>(P, S, R, T, I, ks, xs, x, k, on, upd, nchg); break;
      case 3: fwd_run<3, KIND, FAST, SK>(P, S, R, T, I, ks, xs, x, k, on, upd, nchg); break;
      default: fwd_run<4, KIND, FAST, SK>(P, S, R, T, I, ks, xs, x, k, on, upd, nchg); break;
    }
  }
  {   // terminal knot
    const StagedPtr r = ks.acquire();
    ks.peek();
    if (on) fwd_state<0, KIND, FAST, SK>(P, S, R, r, rec_of(T, I, N), I, N, x);
    ks.release();
  }
  if (ks.close(T)) S.lost = 1;
  if (CHK || PMMK) {   // the lanes of a team hold partial maxima / counts of their own rows
    R.pri = team_max(I, xs, R.pri);
    R.npri = team_max(I, xs, R.npri);
    if (CHK) {
      R.dua = team_max(I, xs, R.dua);
      R.ndua = team_max(I, xs, R.ndua);
    }
    if (PMMK) {
      nchg = team_sum(I, xs, nchg);
      R.kbad = team_or(I, xs, R.kbad);
    }
  }
  if (on && commit && (CHK || PMMK)) {
    if (PMMK) S.kbad |= R.kbad;
    S.pri = R.pri;
    S.npri = fmax(R.npri, S.dynrow);
    if (CHK) {
      S.dua = R.dua;
      S.ndua = fmax(R.ndua, S.nq);
    }
  }
  if (on && changes) *changes = nchg;
  if (on) S_in = S;
  else if (S.lost) S_in.lost = 1;
}

// ---------------------------------------------------------------- per-knot operations
// Knots are independent here: lane q of the team takes the knots q, q + NL, ... and works on the
// global record directly.
// rho change: keep (w, y), move v
CMPC_OP void rescale_op(const Params& P, const TileCtx& T, const Inst& I, const Sv& S, double rho_new, double rhok_new) {
  const double ratio = S.rho / rho_new;
  for (int k = sub_of(I); k <= P.N; k += NL) {
    double* r = rec_of(T, I, k);
    const Lay L = lay_of(T.ns(k), T.gen != 0);
    if (k < P.N) {
      const int nr = 4 * T.ns(k);
      for (int j = 0; j < nr; ++j) {
        const double v = CMPC_R(r, L.vf + j);
        CMPC_R(r, L.vf + j) = fma(ratio, fmax(v, 0.0), fmin(v, 0.0));
      }
    }
    if (k >= 1) {
      const double kb[3] = {CMPC_R(r, L.xb + 6), CMPC_R(r, L.xb + 7), CMPC_R(r, L.xb + 8)};
      const double vk[3] = {CMPC_R(r, L.vk), CMPC_R(r, L.vk + 1), CMPC_R(r, L.vk + 2)};
      double w[3];
      prox_kappa(S, vk, kb, w);
#pragma unroll
      for (int a = 0; a < 3; ++a) CMPC_R(r, L.vk + a) = fma(S.rhok / rhok_new, vk[a] - w[a], w[a]);
    }
  }
}

// active set of the polish: friction row active iff its multiplier is positive (OSQP's rule
// -w < y <=> v > 0); trust-region rows by the branch the prox took.  Sets *kap when some knot has
// trust-region rows.
CMPC_OP void build_active_set_op(const Params& P, const TileCtx& T, const Inst& I, const Sv& S, bool on, int* kap_out) {
  const int N = P.N;
  int kap = 0;
  if (on) {
    for (int k = sub_of(I); k <= N; k += NL) {
      double* w = rec_of(T, I, k);
      const int ns = T.ns(k);
      const Lay L = lay_of(ns, T.gen != 0);
      int pm = 0;
      if (k < N) {
        for (int s = 0; s < ns; ++s) {
#pragma unroll
          for (int row = 0; row < 4; ++row) {
            const double e2 = T.gen ? CMPC_R(w, L.g + s * GS + 12 + row) : (row < 2 ? P.e2[0] : P.e2[2]);
            const double v = CMPC_R(w, L.vf + 4 * s + row);
            const bool act = v > 0.0;
            if (act) pm |= 1 << (4 * s + row);
            CMPC_R(w, L.yf + 4 * s + row) = act ? S.rho * e2 * v : 0.0;
          }
        }
      }
      if (k >= 1) {
        const double kb[3] = {CMPC_R(w, L.xb + 6), CMPC_R(w, L.xb + 7), CMPC_R(w, L.xb + 8)};
        const double a3[3] = {CMPC_R(w, L.vk), CMPC_R(w, L.vk + 1), CMPC_R(w, L.vk + 2)};
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
        for (int i = 0; i < 4; ++i) CMPC_R(w, L.yk + i) = yk4[i];
      }
      meta_of(T, I, k, L.meta)[TL] = pm;
    }
  }
  kap = team_or(I, scratch_of(T, I), kap);
  if (on) *kap_out = kap;
}

// ---------------------------------------------------------------- trust test and accuracy ratio
// sigma_max(X - Xbar) via the 9x9 Gram matrix + cyclic Jacobi (scp_solver.py:151: np.linalg.norm(.,2));
// rho = sum_k ||(f(x,u) - lin)[6:9]||^2 / sum_k ||lin||^2 (scp_solver.py:71-87).
// Team work split: the per-knot terms of the two sums by the lanes q, q + NL, ... (parked in the knot
// records, then added up in knot order by every lane); the Gram matrix by rows (a lane walks the whole
// horizon for its rows); the Jacobi rotations on the matrix in the scratch, rows / columns by their owners.
CMPC_OP void evaluate_op(const Params& P, const TileCtx& T, const Inst& I, bool on, double* snorm, double* num_out, double* den_out) {
  const int N = P.N;
  const int q = sub_of(I);
  const ScratchPtr xs = scratch_of(T, I);
  constexpr int GT_ = (9 + NL - 1) / NL;
  for (int k = q; on && k < N; k += NL) {   // accuracy-ratio terms of the knots q, q + NL, ...
    double* r = rec_of(T, I, k);
    const Lay L = lay_of(T.ns(k), T.gen != 0);
    double x[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = CMPC_R(r, L.x + i);
    const int mt = meta_of(T, I, k, L.meta)[0];
    const int ns = mt & 7;
    double u[MAXU];
#pragma unroll
    for (int i = 0; i < MAXU; ++i) u[i] = 0.0;
    double F[3] = {0, 0, 0}, Tq[3] = {0, 0, 0};
    for (int sl = 0; sl < ns; ++sl) {
      const int cid = (mt >> (4 + 2 * sl)) & 3;
      const double us[3] = {CMPC_R(r, L.u + 3 * sl), CMPC_R(r, L.u + 3 * sl + 1), CMPC_R(r, L.u + 3 * sl + 2)};
      if (WR) {
        for (int a = 0; a < 3; ++a) {
          u[uix(cid, a)] = us[a];
          if (!(cid & 1)) F[a] += us[a];
          Tq[a] += CMPC_R(r, L.d + 9 * sl + 6 + a) * us[2] + CMPC_R(r, L.d + 9 * sl + 3 + a) * us[1] + CMPC_R(r, L.d + 9 * sl + a) * us[0];
        }
        continue;
      }
      const double ds[3] = {CMPC_R(r, L.d + 3 * sl), CMPC_R(r, L.d + 3 * sl + 1), CMPC_R(r, L.d + 3 * sl + 2)};
      double t[3];
      cross3(ds, us, t);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int c = 0; c < MAXC; ++c) if (c == cid) u[3 * c + a] = us[a];
        F[a] += us[a];
        Tq[a] += t[a];
      }
    }
    // lin = A x + B u + c with the structured A, B
    const double S3[3] = {CMPC_R(r, L.s), CMPC_R(r, L.s + 1), CMPC_R(r, L.s + 2)};
    double lin[9], nl[9], Sxc[3];
    cross3(S3, x, Sxc);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lin[a] = x[a] + P.dt_m * x[3 + a];
      lin[3 + a] = x[3 + a] + P.dt * F[a] + (a == 2 ? P.dtmg : 0.0);
      lin[6 + a] = x[6 + a] + P.dt * Sxc[a] + P.dt * Tq[a] + CMPC_R(r, L.ck + a);
    }
    step_knot(P, x, u, I.cpos + (long)k * P.nf * 3, I.cact + (long)k * P.nf, nl, I.cR ? I.cR + (long)k * P.nf * 9 : nullptr);
    double nk = 0.0, dk = 0.0;
#pragma unroll
    for (int i = 6; i < 9; ++i) nk += (nl[i] - lin[i]) * (nl[i] - lin[i]);
#pragma unroll
    for (int i = 0; i < 9; ++i) dk += lin[i] * lin[i];
    CMPC_R(r, L.yk) = nk;        // (the multiplier fields are free here: every polish starts by rebuilding them)
    CMPC_R(r, L.yk + 1) = dk;
  }
  // Gram matrix of X - Xbar: the lane's rows i = q + NL t, the whole horizon
  double G[GT_][9];
#pragma unroll
  for (int t = 0; t < GT_; ++t)
#pragma unroll
    for (int j = 0; j < 9; ++j) G[t][j] = 0.0;
  for (int k = 0; on && k <= N; ++k) {
    const double* r = rec_of(T, I, k);
    const Lay L = lay_of(T.ns(k), T.gen != 0);
    double dx[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) dx[j] = CMPC_R(r, L.x + j) - I.Xr[k * 9 + j];
#pragma unroll
    for (int t = 0; t < GT_; ++t) {
      const int i = q + NL * t < 9 ? q + NL * t : 8;
      const double di = CMPC_R(r, L.x + i) - I.Xr[k * 9 + i];
#pragma unroll
      for (int j = 0; j < 9; ++j) G[t][j] = fma(di, dx[j], G[t][j]);
    }
  }
#pragma unroll
  for (int t = 0; t < GT_; ++t) {
    const int i = q + NL * t;
    if (i < 9) {
#pragma unroll
      for (int j = 0; j < 9; ++j) sp_st(xs, X_PS + i * 9 + j, G[t][j]);
    }
  }
  team_sync(I);
  double num = 0.0, den = 0.0;
  for (int k = 0; on && k < N; ++k) {   // the two sums in knot order
    const double* r = rec_of(T, I, k);
    const Lay L = lay_of(T.ns(k), T.gen != 0);
    num += CMPC_R(r, L.yk);
    den += CMPC_R(r, L.yk + 1);
  }
  *num_out = num;
  *den_out = den;
  // largest eigenvalue of the Gram matrix: cyclic Jacobi on the scratch copy.  The team_sync points must be
  // the same for every instance of the tile: an instance that has converged (or meets a zero off-diagonal
  // entry) goes through the remaining rotations with the identity, which leaves its matrix bit for bit.
  bool conv = !on;   // an instance that does not take part only keeps the synchronisation points
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < 9; ++i) {
      for (int j = i + 1; j < 9; ++j) { const double v = sp_ld(xs, X_PS + i * 9 + j); off += v * v; }
      const double d = sp_ld(xs, X_PS + i * 10);
      dg += d * d;
    }
    if (off <= 1e-30 * dg || off == 0.0 || !(off == off)) conv = true;
    if (tile_all(conv)) break;
    for (int p = 0; p < 8; ++p) {
      for (int qq = p + 1; qq < 9; ++qq) {
        const double apq = sp_ld(xs, X_PS + p * 9 + qq);
        const bool skip = conv || apq == 0.0 || !(apq == apq);
        const double th = (sp_ld(xs, X_PS + qq * 10) - sp_ld(xs, X_PS + p * 10)) / (2.0 * (skip ? 1.0 : apq));
        const double t = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
        const double cs = skip ? 1.0 : 1.0 / sqrt(t * t + 1.0), sn = skip ? 0.0 : t * cs;
        team_sync(I);   // everybody has read the pivot entries before rows p, q change
#pragma unroll
        for (int tt = 0; tt < GT_; ++tt) {   // columns p, q of the lane's rows
          const int rr = q + NL * tt;
          if (rr < 9 && !skip) {
            const double arp = sp_ld(xs, X_PS + rr * 9 + p), arq = sp_ld(xs, X_PS + rr * 9 + qq);
            sp_st(xs, X_PS + rr * 9 + p, cs * arp - sn * arq);
            sp_st(xs, X_PS + rr * 9 + qq, sn * arp + cs * arq);
          }
        }
        team_sync(I);
#pragma unroll
        for (int tt = 0; tt < GT_; ++tt) {   // rows p, q of the lane's columns
          const int rr = q + NL * tt;
          if (rr < 9 && !skip) {
            const double apr = sp_ld(xs, X_PS + p * 9 + rr), aqr = sp_ld(xs, X_PS + qq * 9 + rr);
            sp_st(xs, X_PS + p * 9 + rr, cs * apr - sn * aqr);
            sp_st(xs, X_PS + qq * 9 + rr, sn * apr + cs * aqr);
          }
        }
        team_sync(I);
      }
    }
  }
  double mx = 0.0;
  for (int i = 0; i < 9; ++i) mx = fmax(mx, sp_ld(xs, X_PS + i * 10));
  *snorm = sqrt(mx);
  team_sync(I);
}

// ---------------------------------------------------------------- per-instance setup
// Pass 1 (setup_slots): active contacts per knot, so that the tile-uniform slot count (the record
// layout of the knot) is known.  Pass 2 (setup_knots): K1 for every knot, friction table when not on
// the fast path, start of the iterate at the linearisation point, constant parts of the residual norms.
// Lane q of the team takes the knots q, q + NL, ...; the partial maxima come back through mq / mc.
CMPC_HD int active_slots(const Params& P, const Inst& I, int k) {
  if (k >= P.N) return 0;
  int ns = 0;
  for (int c = 0; c < P.nf; ++c) ns += I.cact[(long)k * P.nf + c] ? (WR ? 2 : 1) : 0;   // wrench model: force + wrench slot per foot
  return ns;
}
CMPC_OP void setup_knots(const Params& P, const TileCtx& T, const Inst& I, bool live, double* mq_out, double* mc_out, int* nconv_out) {
  const int N = P.N;
  double mq = 0.0, mc = 0.0;
  int nzx = 0, nzu = 0;
  for (int k = sub_of(I); live && k <= N; k += NL) {
    double* r = rec_of(T, I, k);
    const int nst = T.ns(k);
    const Lay L = lay_of(nst, T.gen != 0);
    int* im = meta_of(T, I, k, L.meta);
    const int kk = k < N ? k : N - 1;
    const double* xb = I.Xr + k * 9;
    KnotLin Lk;
    linearize_knot(P, xb, I.Ui + kk * P.nu, I.cpos + (long)kk * P.nf * 3, I.cact + (long)kk * P.nf, k == N, Lk,
                   I.cR ? I.cR + (long)kk * P.nf * 9 : nullptr);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      CMPC_R(r, L.xb + i) = xb[i];
      CMPC_R(r, L.x + i) = 0.0;
      mq = fmax(mq, fabs((WR ? P.qs * P.Wx[i] : P.Wx[i]) * xb[i]));
      nzx |= xb[i] != 0.0;
    }
    if (k < N)
      for (int j = 0; j < P.nu; ++j) nzu |= I.Ui[k * P.nu + j] != 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) { CMPC_R(r, L.s + a) = Lk.S[a]; CMPC_R(r, L.ck + a) = Lk.ck[a]; }
    for (int j = 0; j < L.na; ++j) { CMPC_R(r, L.dv + j) = 0.0; CMPC_R(r, L.u + j) = 0.0; }
    for (int j = 0; j < (WR ? 3 : 1) * L.na; ++j) CMPC_R(r, L.d + j) = Lk.d[j];
    for (int j = 0; j < 4 * nst; ++j) { CMPC_R(r, L.vf + j) = 0.0; CMPC_R(r, L.yf + j) = 0.0; }
#pragma unroll
    for (int j = 0; j < 4; ++j) CMPC_R(r, L.yk + j) = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) CMPC_R(r, L.vk + a) = xb[6 + a];
    im[0] = Lk.meta;
    im[TL] = 0;
    if (k < N) {
      mc = fmax(mc, fabs(P.dtmg));
#pragma unroll
      for (int a = 0; a < 3; ++a) mc = fmax(mc, fabs(Lk.ck[a]));
      const int ns = Lk.meta & 7;
      for (int sl = 0; sl < nst; ++sl) {
        if (sl >= ns && !T.gen) continue;   // padded slot on the fast path: nothing to write
        const int cid = sl < ns ? ((Lk.meta >> (4 + 2 * sl)) & 3) : 0;
        double G[12];
#pragma unroll
        for (int row = 0; row < 4; ++row) {
          double mx = 0.0;
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            double g = pyr4(P, row, a);
            if (WR && (cid & 1)) {   // wrench slot (cop_x, cop_y, tau_z): the CoP box, constraints.py:111-145
              g = (a == (row >> 1)) ? ((row & 1) ? -1.0 : 1.0) : 0.0;
            } else if (sl < ns && I.cR) {
              const double* Rm = I.cR + ((long)k * P.nf + (WR ? cid >> 1 : cid)) * 9;
              g = 0.0;
#pragma unroll
              for (int b2 = 0; b2 < 3; ++b2) g += pyr4(P, row, b2) * Rm[a * 3 + b2];
            }
            G[row * 3 + a] = g;
            if (T.gen) mx = fmax(mx, fabs(g) / sqrt(P.Wu[3 * cid + a]));
          }
          if (T.gen) {
#pragma unroll
            for (int a = 0; a < 3; ++a) CMPC_R(r, L.g + sl * GS + row * 3 + a) = G[row * 3 + a];
            CMPC_R(r, L.g + sl * GS + 12 + row) = mx > 0.0 ? 1.0 / (mx * mx) : 0.0;
            CMPC_R(r, L.g + sl * GS + 16 + row) = (WR && (cid & 1)) ? P.foot_range[row] : ((sl < ns && I.fub) ? I.fub[((long)k * P.nc + cid) * 4 + row] : 0.0);
          }
        }
        if (sl < ns) {
          const double* Uk = I.Ui + k * P.nu;
          const double ub[3] = {Uk[uix(cid, 0)], Uk[uix(cid, 1)], Uk[uix(cid, 2)]};
#pragma unroll
          for (int row = 0; row < 4; ++row) {
            double cf = 0.0, scale = 1.0;
#pragma unroll
            for (int a = 0; a < 3; ++a) { cf += G[row * 3 + a] * ub[a]; scale += fabs(G[row * 3 + a] * ub[a]); }
            const double cf0 = cf;
            if (WR && (cid & 1)) cf -= P.foot_range[row];
            else if (I.fub) cf -= I.fub[((long)k * P.nc + cid) * 4 + row];
            // (warm start: a row the warm start sits on starts in the active set -- "active <=> v > 0")
            CMPC_R(r, L.vf + 4 * sl + row) = (P.warm && cf >= -P.warm_tol * (scale + fabs(cf0 - cf))) ? 1e-300 : fmin(cf, 0.0);
          }
        }
      }
    }
  }
  const ScratchPtr xs = scratch_of(T, I);
  *mq_out = team_max(I, xs, mq);
  *mc_out = team_max(I, xs, mc);
  const int nz = team_or(I, xs, nzx | (nzu << 1));
  // non-finite inputs: the maxima above swallow NaN (fmax), so test the data itself
  int bad = 0;
  for (int k = sub_of(I); live && k <= N; k += NL) {
    for (int i = 0; i < 9; ++i) bad |= !(fabs(I.Xr[k * 9 + i]) <= 1.79e308);
    if (k < N) {
      for (int j = 0; j < P.nu; ++j) bad |= !(fabs(I.Ui[k * P.nu + j]) <= 1.79e308);
      if (I.fub)
        for (int j = 0; j < 4 * P.nc; ++j) bad |= !(fabs(I.fub[(long)k * P.nc * 4 + j]) <= 1.79e308);
    }
  }
  for (int i = 0; i < 9; ++i) bad |= !(fabs(I.xi[i]) <= 1.79e308) || !(fabs(I.xf[i]) <= 1.79e308);
  bad = team_or(I, xs, bad);
  *nconv_out = (nz != 3 ? 1 : 0) | (bad ? 2 : 0);
}
CMPC_FN void setup_finish(const Inst& I, Sv& S, double mq, double mc, int nconv) {
  S.nq = mq;
  S.nconv = nconv & 1;
  S.badin = (nconv >> 1) & 1;
  double mi = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) mi = fmax(mi, fabs(I.xi[i]));
  S.dynrow = fmax(mc, mi);
#pragma unroll
  for (int i = 0; i < 9; ++i) S.ye[i] = 0.0;
  S.kap = 0;
  S.fail = 0;
  S.kbad = 0;
  S.lost = 0;
  S.n_pmm = S.n_polish = 0;
  S.pri = S.dua = S.npri = S.ndua = 0.0;
}

CMPC_OP void write_solution_knots(const Params& P, const TileCtx& T, const Inst& I, double* X_out, double* U_out) {
  // The lanes of a team write consecutive elements of the instance's output arrays (8 consecutive doubles
  // per store instruction and instance): the outputs may be mapped host memory (cmpc_solve_scp_host with
  // page-locked buffers), where scattered 8-byte writes would be one PCIe transaction each.
  const int N = P.N, nu = P.nu;
  double* Xo = X_out + (long)I.b * (N + 1) * 9;
  double* Uo = U_out + (long)I.b * N * nu;
  for (int e = sub_of(I); e < (N + 1) * 9; e += NL) {
    const int k = e / 9, i = e - 9 * k;
    const Lay L = lay_of(T.ns(k), T.gen != 0);
    Xo[e] = CMPC_R(rec_of(T, I, k), L.x + i);
  }
  for (int e = sub_of(I); e < N * nu; e += NL) {
    const int k = e / nu, j = e - k * nu;
    int c = j / 3, a = j - 3 * c;
    if (WR) {   // (cop_x, cop_y, fx, fy, fz, tau_z) per foot -> pseudo-contact and axis
      const int ft = j / 6, w = j - 6 * ft;
      c = 2 * ft + ((w >= 2 && w <= 4) ? 0 : 1);
      a = (w >= 2 && w <= 4) ? w - 2 : (w == 5 ? 2 : w);
    }
    const Lay L = lay_of(T.ns(k), T.gen != 0);
    const double* r = rec_of(T, I, k);
    const int mt = meta_of(T, I, k, L.meta)[0];
    const int ns = mt & 7;
    double v = 0.0;   // inactive contacts carry exact zeros
    for (int sl = 0; sl < ns; ++sl)
      if (((mt >> (4 + 2 * sl)) & 3) == c) v = CMPC_R(r, L.u + 3 * sl + a);
    Uo[e] = v;
  }
}

// ---------------------------------------------------------------- the per-instance driver
// scp_solver.py:118-179 with the QP solve (ADMM interleaved with certified active-set polishes)
// flattened into a state machine: advance() runs the scalar decisions of an instance until it
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
  int round, sw, prev_chg, chg, certified, upd, stag;
  double prev_pri;
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
  D.round = D.sw = D.chg = D.certified = D.upd = D.stag = 0;
  D.prev_pri = 0.0;
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
    if (S.lost) {   // a bulk copy of this tile never landed: report it per instance and stop
      D.status = ST_DEVICE;
      return OP_DONE;
    }
    switch (D.pc) {
      case PC_SCP_TOP:
        if (S.badin) {   // NaN / Inf in the problem data: the arithmetic below would swallow it (fmin / fmax) and "solve"
          D.status = ST_QP_NUMERIC;
          D.pc = PC_END;
          break;
        }
        // convergence() compares the warm start with itself (scp_solver.py:90-93): 0 < threshold whenever
        // both norms are nonzero, 0/0 = NaN (never converged) when one of them vanishes (S.nconv)
        if (!(D.it_scp < P.max_scp && D.weight < P.omega_max && !(D.it_scp != 0 && D.success && !S.nconv && 0.0 < P.conv_thresh))) {
          D.pc = PC_FINISH;
          break;
        }
        D.success = 0;
        S.radius = D.radius;
        S.weight = D.weight;
        S.tau = S.weight / S.rhok;
        D.nfact = 0; D.solved = 0; D.polished = 0; D.it = 0;
        S.fail = 0;
        if (P.warm && D.it_scp == 0) {   // warm start: polish on the active set of the warm start first, no ADMM iteration
          D.term = 0;
          D.pri0 = D.dua0 = D.npri0 = D.ndua0 = 0.0;
          D.next_as = -1;
          D.as_step = P.as_step;
#pragma unroll
          for (int i = 0; i < 9; ++i) D.ye_keep[i] = S.ye[i];
          D.pc = PC_AFTER_BUILD;
          return OP_BUILD_AS;
        }
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
        S.kbad = 0;
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
        D.stag = 0;
        D.prev_pri = S.pri;
        if (S.kbad) { D.pc = PC_POLISH_END; break; }   // the trust-region branches were guessed wrongly: back to ADMM
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
#if !defined(__CUDACC__) && defined(CMPC_EMU_TRACE)
        fprintf(stderr, "  pmm sweep: round %d sw %d chg %d pri %.3e\n", D.round, D.sw, D.chg, S.pri);
#endif
        if (S.kbad) { D.pc = PC_POLISH_END; break; }
        if (D.chg || S.pri <= P.as_tol * (1.0 + S.npri)) { D.pc = PC_ROUND_CHECK; break; }
        // stiff penalties (wrench model): the multiplier iteration contracts fast and then sits on its rounding
        // floor; a residual that stopped halving below the loose tolerance is as good as it gets
        if (P.as_tol_loose > P.as_tol && S.pri > 0.5 * D.prev_pri && S.pri <= P.as_tol_loose * (1.0 + S.npri)) {
          D.stag = 1;
          D.pc = PC_ROUND_CHECK;
          break;
        }
        D.prev_pri = S.pri;
        ++D.sw;
        D.pc = PC_SW_TOP;
        break;
      case PC_ROUND_CHECK:
        if (!(S.pri == S.pri)) { D.pc = PC_POLISH_END; break; }   // NaN
        if (D.chg == 0) {
          D.certified = S.pri <= (D.stag ? P.as_tol_loose : P.as_tol) * (1.0 + S.npri);   // absolute + relative, the form of OSQP's test
          D.pc = PC_POLISH_END;
          break;
        }
        D.prev_chg = D.chg;
        ++D.round;
        D.pc = D.round > P.as_rounds ? PC_POLISH_END : PC_ROUND_TOP;
        break;
      case PC_POLISH_END: {
#if !defined(__CUDACC__) && defined(CMPC_EMU_TRACE)
        fprintf(stderr, "polish end: it %d round %d sw %d chg %d kbad %d fail %d pri %.3e npri %.3e certified %d\n", D.it, D.round, D.sw,
                D.chg, S.kbad, S.fail, S.pri, S.npri, D.certified);
#endif
        S.kap = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) S.ye[i] = D.ye_keep[i];   // the ADMM multiplier comes back
        // OSQP's rule for an uncertified polish after normal termination: keep it if it improves
        const double m0 = fmax(D.pri0 / (P.eps_abs + P.eps_rel * D.npri0), D.dua0 / (P.eps_abs + P.eps_rel * D.ndua0));
        const double m1 = S.pri / (P.eps_abs + P.eps_rel * S.npri);
        if (!S.kbad && (D.certified || (D.term && !S.fail && m1 < m0))) {
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
// streamed operations are warp-collective); `on` says whether this lane's instance takes part.
// anycheck: some participating instance wants residuals from this ADMM sweep.
template <bool FAST>
CMPC_FN void execute(int op, const Params& P, TileCtx& T, const Inst& I, const Batch& bt, Sv& S, Drv& D, bool on, bool anycheck) {
#if defined(CMPC_PROFILE) && defined(__CUDACC__)
  const long long t0_ = clock64();
#endif
  switch (op) {
    case OP_FACTOR_ADMM: factor_op<MODE_ADMM, FAST>(P, T, I, S, on); break;
    case OP_SWEEP_ADMM:
      backward_op<MODE_ADMM, FAST>(P, T, I, S, on);
#if defined(CMPC_PROFILE) && defined(__CUDACC__)
      T.prof[10] += clock64() - t0_;
#endif
      if (anycheck) forward_op<FW_ADMM_CHECK, FAST>(P, T, I, S, on, on && D.check, false, nullptr);
      else forward_op<FW_ADMM, FAST>(P, T, I, S, on, false, false, nullptr);
      break;
    case OP_BUILD_AS: build_active_set_op(P, T, I, S, on, &S.kap); break;
    case OP_FACTOR_PMM: factor_op<MODE_PMM, FAST>(P, T, I, S, on); break;
    case OP_SWEEP_PMM:
      backward_op<MODE_PMM, FAST>(P, T, I, S, on);
#if defined(CMPC_PROFILE) && defined(__CUDACC__)
      T.prof[11] += clock64() - t0_;
#endif
      forward_op<FW_PMM, FAST>(P, T, I, S, on, true, on && D.upd, &D.chg);
      break;
    case OP_RESCALE: if (on) rescale_op(P, T, I, S, D.rho_new, D.rhok_new); break;
    case OP_COPY_SOL:
      backward_op<MODE_ADMM, FAST>(P, T, I, S, on);
      forward_op<FW_COPY, FAST>(P, T, I, S, on, false, false, nullptr);
      break;
    case OP_EVAL: {
      double sn_ = 0.0, nu_ = 0.0, de_ = 1.0;
      evaluate_op(P, T, I, on, &sn_, &nu_, &de_);
      if (on) { D.snorm = sn_; D.num = nu_; D.den = de_; }
      break;
    }
    case OP_WRITE: if (on) write_solution_knots(P, T, I, bt.X_out, bt.U_out); break;
    default: break;
  }
  team_sync(I);
#if defined(CMPC_PROFILE) && defined(__CUDACC__)
  T.prof[op] += clock64() - t0_;
#endif
}

CMPC_FN void write_stats(const Batch& bt, const Inst& I, const Sv& S, const Drv& D) {
  if (sub_of(I) != 0) return;
  bt.scp_iters[I.b] = D.it_scp;
  bt.status[I.b] = D.status;
  bt.n_accepted[I.b] = D.n_acc;
  bt.qp_iters[I.b] = D.qp_total;
  bt.n_factor[I.b] = D.nf_total;
  double* inf = bt.info + (long)I.b * INFO;
  inf[0] = D.snorm; inf[1] = D.acc; inf[2] = S.pri; inf[3] = S.dua;
  inf[4] = S.rho; inf[5] = D.radius; inf[6] = D.weight; inf[7] = (double)D.polished;
  inf[8] = (double)S.n_pmm; inf[9] = (double)S.n_polish; inf[10] = (double)D.certified; inf[11] = 0.0;
}

// bind the per-instance input pointers
CMPC_HD void bind_instance(Inst& I, const Params& P, const Batch& bt, int b) {
  const int N = P.N;
  const long plan = (long)b * bt.plan_stride;
  I.b = b;
  I.lane = b & (TL - 1);
  I.cpos = bt.cpos + plan * N * P.nf * 3;
  I.cR = bt.cR ? bt.cR + plan * N * P.nf * 9 : nullptr;
  I.fub = bt.fub ? bt.fub + (long)b * N * P.nc * 4 : nullptr;
  I.cact = bt.cact + plan * N * P.nf;
  I.Xr = bt.X_ref + (long)b * (N + 1) * 9;
  I.Ui = bt.U_init + (long)b * N * P.nu;
  I.xi = bt.x_init + (long)b * 9;
  I.xf = bt.x_final + (long)b * 9;
}
CMPC_HD void bind_tile(TileCtx& T, const Params& P, const Batch& bt, int tile) {
#if defined(__CUDACC__)
  T.tile = tile;
#endif
  T.gen = P.fast ? 0 : 1;
  T.rstride = (long)bt.rfields * TL;
  T.ws = bt.ws + (long)tile * (P.N + 1) * T.rstride;
  T.nst = bt.nst + (long)tile * (P.N + 1);
}

}  // namespace cmpc
