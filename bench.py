#!/usr/bin/env python
"""bench.py — SCP-MPC solves/sec on the BASELINE.json headline workload.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one batched solve_scp over a synthetic batch: solo12 trot, horizon N=100, 4096
independent MPC instances per GPU (weak scaling; instances are independent, so ranks share
nothing on the solve path and one gather collects the solutions; the gather of step i runs on a side
stream under the solve of step i+1).  Other BASELINE.json configurations:
  --workload {solo12_trot,solo12_pace,solo12_bound,bolt,talos}  --mode {A,B}  --batch B  --scaling {weak,strong}
(strong: --batch is the GLOBAL batch, split contiguously over the ranks; e.g. the bolt configuration is
--workload bolt --batch 8192 --scaling strong --gpus 8, the pace one --workload solo12_pace --mode A --batch 1024).

  value   whole-job solves/s with the problem data resident in HBM (CUDA events, max over ranks)
  e2e     the same through the C-ABI host entry point cmpc_solve_scp_host: pinned host buffers,
          H2D of the inputs and D2H of the solutions inside the timed region
  roofline  dominant kernel cmpc_scp_kernel against the measured HBM copy bandwidth
          (MEASURED_PEAKS.json), algorithmic bytes per launch as defined in DESIGN.md
  cpu_baseline  the CPU oracle (a restatement of the reference: its dependencies jax/osqp/
          pinocchio are not installable here) timed on the box's host cores on a bounded sample

--impl reference times that CPU oracle as the reference arm (oracle/ is executed only there and
in the cpu_baseline leg, as the thing compared against, never as the product path).
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "solo12_trot"
HORIZON = 100
BATCH_PER_GPU = 4096
METRIC = "SCP-MPC solves/sec (solo12 trot N=100, batch 4096)"


# ------------------------------------------------------------------------------------------ CPU arm
def _oracle_one(arg):
    b, WORKLOAD = arg
    import numpy as np  # noqa: F401
    from centroidal_mpc_b200 import synthetic
    from centroidal_mpc_b200.src.centroidal_model import Centroidal_model
    from oracle import scp
    conf = synthetic.load_conf(WORKLOAD, N=HORIZON)
    model = Centroidal_model(conf, centroidal_traj=synthetic.reference_trajectory(conf, b))
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        sol = scp.solve_scp(model.problem_arrays(), conf.scp_params)
    return 0 if sol is False else sol["iterations"]


def cpu_oracle_throughput(n_instances, workers, workload=WORKLOAD):
    """solves/s of the CPU oracle on ``n_instances`` instances of the workload with a process
    pool of ``workers`` (one single-threaded solve per process)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        pool.map(_oracle_one, [(b, workload) for b in range(workers)])          # warm-up: imports, first-touch
        t0 = time.time()
        pool.map(_oracle_one, [(b, workload) for b in range(n_instances)])
        dt = time.time() - t0
    return n_instances / dt, dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    workers = max(1, min(cores, 64))
    per_step = workers * 16
    for _ in range(args.warmup and 1):
        cpu_oracle_throughput(workers, workers, args.workload)
    vals, times = [], []
    for _ in range(args.steps):
        v, dt = cpu_oracle_throughput(per_step, workers, args.workload)
        vals.append(v)
        times.append(dt)
    value = float(sum(vals) / len(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s N=%d" % (args.workload, HORIZON), "batch_per_step": per_step,
                   "note": "CPU restatement of the reference (jax/osqp/pinocchio unavailable offline)"},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": workers, "kind": "port",
                         "sample": "%d instances per step, one single-threaded oracle solve per process" % per_step},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Median SM clock / throttle reasons of the samples taken in [t0, t1] (host clock); when the
        timed region is shorter than the sampling period, the samples next to it."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for ts, r in self.rows if t0 is None or (t0 <= ts <= t1 + 0.05)]
        if not rows and self.rows and t0 is not None:
            near = sorted(self.rows, key=lambda tr: abs(tr[0] - 0.5 * (t0 + t1)))[:3]
            rows = [r for _, r in near]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_work(batch, stats):
    """HBM bytes and FP64 flops the algorithm must move / execute for one launch (DESIGN.md
    "roofline").  Bytes: per knot exactly the record fields each operation stages (the bulk-copy
    segments of csrc/cmpc_core.cuh) plus the fields it writes, 8 bytes each, times the
    operations each INSTANCE asked for (ADMM iterations, multiplier-method sweeps,
    factorisations: the solver's own statistics), plus compulsory I/O.  Work a tile repeats
    because a neighbouring lane needs another polish round is not algorithmic and is not counted."""
    import numpy as np
    ns = batch.contact_active[0].sum(axis=1).astype(float)   # active contacts per knot (shared plan)
    wrench = getattr(batch, "wrench", False)
    if wrench:
        ns = 2 * ns                                            # a foot is a force slot and a wrench slot
    na = 3 * ns
    N = batch.N
    # record ranges of csrc/cmpc_core.cuh (lay_of) / cmpc_tile.cuh (ranges_of): A Pc[9]; M Hn[na*na] + Kt[9 na];
    # C meta + xbar + S + ck + d; D vk + vf; E dv; F yk + yf
    gen = 20 * ns if (wrench or batch.contact_R is not None) else 0 * ns   # general friction table (frames / CoP box)
    segA, segHn, segKt, segC, segD, segE, segF = 9 + 0 * na, na * na, 9 * na, 16 + (3 if wrench else 1) * na + gen, 3 + 4 * ns, na, 4 + 4 * ns
    term = 9 + 16 + 3                                         # terminal knot: Pc slot + stage data + kappa copy
    bwd_admm = (segA + segHn + segKt + segC + segD + na).sum() + term          # [A .. D] + writes d_k
    fwd_admm = (segKt + segC + segD + segE + 3 + 4 * ns).sum() + term + 3      # [Kt .. E] + writes vk, vf
    bwd_pmm = (segA + segHn + segKt + segC + segF + na).sum() + 9 + 16 + 4
    fwd_pmm = (segKt + segC + segE + segF + 9 + na + 4 * ns).sum() + 16 + 4 + 9   # + writes x, u, yf
    fac = (segC + 9 + na * na + 9 * na).sum() + 16                             # reads stage, writes Pc, Hn, Kt
    sweep_bytes = 8.0 * float(bwd_admm + fwd_admm)
    pmm_bytes = 8.0 * float(bwd_pmm + fwd_pmm)
    factor_bytes = 8.0 * float(fac)
    sweep_flops = 2.0 * float((na * na + 9 * na + 9 * na + 14 * ns + 60).sum())
    factor_flops = 2.0 * float((na * (na + 9) * (na + 10) / 2 + 27 * na + 6 * na * (na + 1) / 2 + 350).sum())
    admm = float(stats["qp_iters"].sum())
    pmm = float(stats["info"][:, 8].sum())
    nfac = float(stats["n_factor"].sum())
    io = batch.input_bytes() / batch.B + ((N + 1) * 9 + N * batch.nu) * 8 + 12
    setup = 8.0 * float((segC + segD + segF + 12).sum())                   # records initialised once per solve
    total_bytes = sweep_bytes * admm + pmm_bytes * pmm + factor_bytes * nfac + (io + setup) * batch.B
    total_flops = sweep_flops * (admm + pmm) + factor_flops * nfac
    return float(total_bytes), float(total_flops), sweep_bytes


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from centroidal_mpc_b200 import _lib as L
    from centroidal_mpc_b200 import parallel, synthetic
    from centroidal_mpc_b200.device import BatchSolver
    import __graft_entry__ as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SCP hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g.build()
    lib = L.load()

    workload = args.workload
    conf = synthetic.load_conf(workload, N=HORIZON)
    if args.scaling == "strong":   # --batch is the global batch: contiguous split, ragged shards allowed
        lo, hi = parallel.shard_range(args.batch, rank, world)
        B, first, Bglobal = hi - lo, lo, args.batch
    else:                          # weak scaling: every rank owns --batch instances; instance ids are global
        B, first, Bglobal = args.batch, rank * args.batch, args.batch * world
    batch = synthetic.make_batch(conf, B, mode=args.mode, first=first)
    for name in ("x_init", "x_final", "X_ref", "U_init", "contact_pos", "contact_active", "contact_R"):
        if getattr(batch, name) is None:
            continue
        t = torch.from_numpy(getattr(batch, name)).pin_memory()
        setattr(batch, name, t.numpy())
        batch.__dict__.setdefault("_pinned", []).append(t)
    solver = BatchSolver(batch)
    out_host = dict(X=torch.empty((B, HORIZON + 1, 9), dtype=torch.float64).pin_memory(),
                    U=torch.empty((B, HORIZON, batch.nu), dtype=torch.float64).pin_memory(),
                    scp_iters=torch.empty(B, dtype=torch.int32).pin_memory(),
                    status=torch.empty(B, dtype=torch.int32).pin_memory(),
                    n_accepted=torch.empty(B, dtype=torch.int32).pin_memory())
    out_np = {k: v.numpy() for k, v in out_host.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The single end-of-batch collective runs on a side stream: the gather of step i overlaps the solve of
    # step i+1 (the solver writes into alternating result buffers, so the gather reads a finished one).
    # Multi-GPU result path.  "peer" (default): every rank's kernel writes its shard straight into a buffer on rank 0's
    # GPU over NVLink at write-back (parallel.PeerResults) -- no gather kernel competes with the next solve for the SMs.
    # "nccl": the single end-of-batch gather on a side stream (falls back to it when the peer mapping is refused).
    peer = None
    if world > 1 and args.gather == "peer":
        try:
            peer = parallel.PeerResults(Bglobal, HORIZON, batch.nu, dist, dst=0, slots=2)
        except Exception as e:   # noqa: BLE001  (reported in the JSON line)
            peer, args.gather = None, "nccl (peer mapping failed: %s)" % str(e)[:80]
        ok = torch.tensor([1 if peer is not None else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0 and peer is not None:
            peer.close()
            peer, args.gather = None, "nccl (peer mapping failed on another rank)"
    side = torch.cuda.Stream() if world > 1 and peer is None else None
    gathered = [None]
    res_bufs = [(solver.X, solver.U, solver.ints), (torch.zeros_like(solver.X), torch.zeros_like(solver.U), torch.zeros_like(solver.ints))]
    parity = [0]

    def step_device():
        if peer is not None:
            solver.solve(conf.scp_params, out=peer.pointers(parity[0], first))
            parity[0] ^= 1
            return
        if world > 1:
            solver.X, solver.U, solver.ints = res_bufs[parity[0]]
        solver.solve(conf.scp_params)
        if world > 1:
            done = torch.cuda.Event()
            done.record()
            X, U, ints = res_bufs[parity[0]]
            with torch.cuda.stream(side):
                side.wait_event(done)
                gathered[0] = parallel.gather_solutions(dict(X=X, U=U, ints=ints[:3]), Bglobal, dist, dst=0)
            parity[0] ^= 1

    def drain():
        if world > 1 and peer is None:
            torch.cuda.current_stream().wait_stream(side)

    sampler = ClockSampler(local_rank)
    sampler.start()                    # nvidia-smi needs ~100 ms to deliver its first sample
    for _ in range(max(args.warmup, 3)):
        step_device()
    drain()
    barrier()
    launches0 = lib.cmpc_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_host0 = time.time()
    t_all0 = torch.cuda.Event(enable_timing=True)
    t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for e0, e1 in ev:
        e0.record()
        step_device()
        e1.record()
    drain()          # the last gather belongs to the timed region
    t_all1.record()
    barrier()
    clocks = sampler.stop(t_host0, time.time())
    launches = lib.cmpc_launch_count() - launches0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    total_ms = t_all0.elapsed_time(t_all1)
    stats = solver.stats()
    if peer is not None:      # the whole job's solutions sit in rank 0's buffer (every rank has synchronised: barrier above)
        view = peer.tensors(parity[0] ^ 1)
        res = None if view is None else {k: v.cpu().numpy() for k, v in view.items()}
    else:
        res = solver.results()

    # end-to-end through the host entry point of the C ABI
    for _ in range(2):
        solver.solve_host(conf.scp_params, out=out_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solver.solve_host(conf.scp_params, out=out_np)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    solver.upload(batch)   # rebind the device copies

    tm = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(tm[0]), float(tm[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = Bglobal * args.steps / (total_ms * 1e-3)
    e2e = Bglobal * args.steps / (e2e_ms * 1e-3)
    kernel_ms = float(np.mean(step_ms))            # the step is one launch of cmpc_scp_kernel (+ gather)
    alg_bytes, flops, per_iter_bytes = algorithmic_work(batch, stats)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    # DRAM bytes of one launch from the committed ncu --set full capture of this kernel; the file is stamped
    # with the build id (hash of the sources) of the library it was captured with and is ignored for any other
    # build or workload
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if traffic.get("build_id") != L.load().cmpc_build_id().decode() or traffic.get("workload") != "%s N=%d B=%d" % (workload, HORIZON, B):
            traffic = {"dram_bytes_per_launch": None,
                       "source": "profiles/r2_traffic.json was captured with another build or workload (%s); not used" % traffic.get("workload")}
    except (OSError, ValueError):
        pass
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    tf, ms = __import__("ctypes").c_double(), __import__("ctypes").c_double()
    lib.cmpc_fp64_peak(__import__("ctypes").byref(tf), __import__("ctypes").byref(ms))

    cores = host_cores()
    workers = max(1, min(cores, 64))
    cpu_v, cpu_dt = cpu_oracle_throughput(workers * 48, workers, workload) if not args.no_cpu_baseline else (None, None)

    srt = sorted(step_ms)
    line = {
        "metric": METRIC if (workload == WORKLOAD and args.batch == BATCH_PER_GPU and args.scaling == "weak")
                  else "SCP-MPC solves/sec (%s N=%d, batch %d %s)" % (workload, HORIZON, args.batch, args.scaling),
        "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s N=%d" % (workload, HORIZON), "batch_per_gpu": B, "global_batch": Bglobal,
                   "mode": "B (independent reference trajectories, rng(1000+b))" if args.mode == "B"
                           else "A (one shared reference, perturbed initial states, rng(1000+b))",
                   "parallelism": "instances sharded across %d GPU(s), no collective on the solve path, %s" % (
                       world, "single GPU" if world == 1 else
                       ("results written by the kernels straight into rank 0's buffer over NVLink peer memory (no gather)"
                        if peer is not None else "one NCCL gather (side stream, under the next step's solve): " + args.gather)),
                   "l2": "solver workspace %.2f GB per GPU (>> 126 MB L2) is streamed by every sweep; no flush needed"
                         % (lib.cmpc_workspace_bytes(solver.handle) / 1e9)},
        "latency_ms_p50": srt[len(srt) // 2],
        "e2e": {"value": e2e, "unit": "solves/s", "h2d_bytes_per_step": batch.input_bytes(),
                "d2h_bytes_per_step": batch.output_bytes()},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs,
                     "traffic": None if traffic is None else traffic.get("dram_bytes_per_launch"),
                     "traffic_source": None if traffic is None else traffic.get("source"),
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                     "kernel": "cmpc_scp_kernel", "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_sweep_pair_per_solve": per_iter_bytes,
                     "note": "teams of 8 lanes per instance, 4 instances per warp, 7 warps per SM at batch 4096: latency-bound on the per-knot dependent chain (DESIGN.md section 6)",
                     "fp64": {"achieved_tflops": flops / (kernel_ms * 1e-3) / 1e12, "measured_peak_tflops": tf.value,
                              "frac": flops / (kernel_ms * 1e-3) / 1e12 / max(tf.value, 1e-9)}},
        "solver_stats": {"admm_iters_mean": float(stats["qp_iters"].mean()), "admm_iters_max": int(stats["qp_iters"].max()),
                         "factorisations_mean": float(stats["n_factor"].mean()),
                         # a warp works for the four instances of its tile until the last one is done, and a wave ends with
                         # its slowest tile: the per-tile maximum and the batch maximum next to the mean
                         "factorisations_tile_max_mean": float(np.pad(stats["n_factor"], (0, (-len(stats["n_factor"])) % 4)).reshape(-1, 4).max(1).mean()),
                         "factorisations_max": int(stats["n_factor"].max()),
                         "multiplier_sweeps_mean": float(stats["info"][:, 8].mean()),
                         "polish_attempts_mean": float(stats["info"][:, 9].mean()),
                         "scp_iters_mean": float(res["scp_iters"].mean()), "failed": int((res["status"] != 0).sum()),
                         "accepted": int((res["n_accepted"] > 0).sum()), "polished_frac": float(stats["info"][:, 7].mean())},
        "cpu_baseline": None if cpu_v is None else {
            "value": cpu_v, "unit": "solves/s", "cores": workers, "kind": "port",
            "sample": "%d instances of the same workload, one single-threaded oracle solve per process, %.1f s" % (workers * 48, cpu_dt)},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="instances per GPU (weak) or in total (strong)")
    ap.add_argument("--workload", default=WORKLOAD, choices=["solo12_trot", "solo12_pace", "solo12_bound", "bolt", "talos"])
    ap.add_argument("--mode", default="B", choices=["A", "B"], help="B: independent references; A: perturbed initial states")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--gather", default="peer", help="multi-GPU result path: peer (kernels write into rank 0's buffer) or nccl")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: library chatter (NCCL banner, ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_gpu_arm(args)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    lines = [ln for ln in out.getvalue().splitlines() if ln.strip()]
    if lines:
        print(lines[-1], flush=True)


if __name__ == "__main__":
    main()
