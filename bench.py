#!/usr/bin/env python
"""bench.py -- benchmark of the B200-native pr-ompl hot path (BASELINE.json: "validated edges/s, kNN
queries/s, plans/s at 1/2/4/8 B200 vs host-CPU OMPL").

Headline line (the default, `--workload knn`): BASELINE.json configs[1], the batched kNN sweep -- one tree
of 10^7 7-DOF nodes, 4096 queries, k=16.  A step is one pass of the hot path over that batch:
  ours       : prgpu_knn_nearest_k_device; with --gpus N the tree is node-sharded over the ranks and a step
               also contains the all-gather of the per-shard top-k candidates and the merge kernel
               (strong scaling).
  reference  : (--impl reference) the CPU nearest-neighbour structure the reference would use -- a k-d tree
               with NearestNeighborsLinear's answers (oracle/kdtree.cpp; OMPL's GNAT is un-vendored), all
               host threads, on a bounded sample of the same queries against the same tree.
The default run ALSO measures the other two metrics BASELINE.json names, each as a complete block under
`extra` (value, e2e, roofline, cpu_baseline at N=1), split over the ranks when --gpus N:
  extra.check_motion : configs[2], 10^6 edges at 0.01 rad, 256 obstacles   (validated edges/s)
  extra.rrtc         : configs[4], 4096 independent RRTConnect queries     (plans/s)
and at N=1 extra.nnf (configs[3], W=200 / M=1024) and extra.dropin_a (one planner instance at a time through
the unmodified reference planner source: the latency-bound path, stated rather than hidden).
`--workload check_motion|rrtc|nnf` makes one of those the headline line instead.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="knn", choices=["knn", "check_motion", "rrtc", "nnf"])
    ap.add_argument("--nodes", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--edges", type=int, default=1_000_000)
    ap.add_argument("--plans", type=int, default=4096)
    ap.add_argument("--plan-iters", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# workload names: ONE place, so that our arm and the reference arm print identical config.workload
# ------------------------------------------------------------------------------------------------
def workload_string(args, wl):
    if wl == "knn":
        return (f"batched kNN sweep: N={args.nodes} nodes, 7-DOF, Q={args.queries}, k={args.k} "
                f"(BASELINE.json configs[1])")
    if wl == "check_motion":
        return (f"batched checkMotion: {args.edges} edges at 0.01 rad, 7-DOF sphere-FK arm (16 link spheres) vs "
                f"256 obstacles (BASELINE.json configs[2]); edge length ~U(0,1] rad")
    if wl == "rrtc":
        return (f"{args.plans} independent RRTConnect queries (ext-con, default range, 0.01 resolution), 7-DOF "
                f"sphere-FK arm vs 256 obstacles, <= {args.plan_iters} samples each (BASELINE.json configs[4], "
                f"query-sharded)")
    return ("NNF Frechet following: W=200 waypoints, 1024 IK samples per layer, k=5, D=1, 7-DOF arm, 32 obstacles "
            "(BASELINE.json configs[3])")


METRICS = {"knn": ("kNN queries/s", "queries/s"), "check_motion": ("validated edges/s", "edges/s"),
           "rrtc": ("RRTConnect plans/s", "plans/s"), "nnf": ("NNF plans/s", "plans/s")}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    """Dense bf16 tensor TFLOP/s (sustained: the kernel is timed inside a long step)."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained, cuBLAS bf16)"
    except Exception:
        return 2250.0, "fallback (B200_PROFILING.md nominal dense bf16)"


def ncu_summary(kernel):
    """Per-launch ncu figures of `kernel` from the committed `ncu --set full` summary (profiles/ncu_summary.json,
    written by scripts/ncu_summary.py): dram bytes, issue-slot utilisation, lanes per instruction.  These come
    from a profiler run of the same command, never from this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def ncu_traffic(kernels):
    total, seen = 0, False
    for k in kernels:
        s = ncu_summary(k)
        if s and s.get("dram_bytes") is not None:
            total += int(s["dram_bytes"])
            seen = True
    return total if seen else None


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def plan_problems(synth, world, n, seed=4):
    """n valid (start, goal) pairs: BASELINE 'independent RRTConnect queries', seed 4.  `world` is any
    object with is_valid_batch (the device world in our arm, the oracle world in the reference arm)."""
    import numpy as np
    q = synth.cloud(seed, 8 * n)
    v = world.is_valid_batch(q)
    v = v[0] if isinstance(v, tuple) else v
    qv = q[v == 1]
    return np.ascontiguousarray(qv[0:2 * n:2]), np.ascontiguousarray(qv[1:2 * n:2])


def nnf_problem(synth, dh, W, M, D, tseed=5, pseed=1):
    """NNF config of BASELINE.json: W waypoints, M IK samples per layer; the IK budget per waypoint is the
    reference's own multinomial (std::default_random_engine draws, NNFrechet.cpp:310-343, restated in
    synth); synthetic IK = trajectory configuration + N(0, 0.15^2) joint noise."""
    R = (W - 1) * (D + 1) + 1
    q, poses = synth.nnf_reference(tseed, 2 * R + 5, dh)
    ids = synth.lin_int_space(0, len(q) - 1, R)
    counts = synth.nnf_layer_counts(pseed, R, W, M)
    ik = synth.nnf_ik_states(2, q[ids], counts)
    return poses, poses[ids], counts, ik


# ------------------------------------------------------------------------------------------------
# CPU baselines (the oracle / compiled reference as the thing timed -- allowed in this leg only)
# ------------------------------------------------------------------------------------------------
class CpuArm:
    """The reference's CPU implementation of one workload on this box's host cores.  prepare() builds the
    state that our arm also keeps outside its timed region (the tree / the world); run(lo, n) processes
    units [lo, lo+n) and returns nothing."""

    def __init__(self, args, wl, synth):
        import numpy as np
        import oracle_lib as O
        O.build()
        self.O, self.np, self.args, self.wl = O, np, args, wl
        self.cores = host_cores()
        self.extra = {}
        if wl == "knn":
            self.pts = synth.cloud(0, args.nodes)
            self.q = synth.cloud(1, args.queries)
            t0 = time.perf_counter()
            self.tree = O.KdTree(self.pts)
            self.extra["kdtree_build_s"] = time.perf_counter() - t0
            self.total = args.queries
            self.kind = "port"
            self.what = ("k-d tree with NearestNeighborsLinear's answers, bit for bit (oracle/kdtree.cpp; the reference "
                         "takes OMPL's GNAT from SelfConfig::getDefaultNearestNeighbors, pr_ompl/src/RRTConnect.cpp:"
                         "131-134, un-vendored); tree built once outside the timed region like the device tree")
        else:
            dh, link, xyzr = synth.default_arm()
            sph, box = synth.obstacles(3)
            self.W = O.World(dh, link, xyzr, sph, box)
            if wl == "rrtc":
                from concurrent.futures import ThreadPoolExecutor
                self.starts, self.goals = plan_problems(synth, self.W, args.plans)
                self.impl = "ref" if O.ref() is not None else "port"
                self.pool = ThreadPoolExecutor(self.cores)
                self.total = args.plans
                self.kind = "reference" if self.impl == "ref" else "port"
                self.what = ("one query per thread, " +
                             ("the reference's own pr_ompl/src/RRTConnect.cpp compiled unmodified over the OMPL-API "
                              "stand-in with NearestNeighborsLinear + DiscreteMotionValidator (oracle/_ref)"
                              if self.impl == "ref" else "oracle port of RRTConnect.cpp"))
            else:
                self.a, self.b = synth.edges(2, args.edges)
                self.total = args.edges
                self.kind = "port"
                self.what = ("DiscreteMotionValidator bisection with early exit over the double-precision synthetic "
                             "checker, threads over edges")

    def run(self, lo, n):
        O, a = self.O, self.args
        if self.wl == "knn":
            self.tree.knn(self.q[lo:lo + n], a.k, nthreads=self.cores)
        elif self.wl == "rrtc":
            np = self.np

            def one(qi):
                return self.W.rrtc_solve(self.starts[qi], self.goals[qi], -np.pi, np.pi, ext=(0, 1), seed=4, query=qi,
                                         max_iters=a.plan_iters, max_nodes=4096, impl=self.impl)["status"]
            list(self.pool.map(one, range(lo, lo + n)))
        else:
            self.W.check_motion_batch(self.a[lo:lo + n], self.b[lo:lo + n], 0.01, want_clearance=False,
                                      nthreads=self.cores)

    def sample_size(self, per_core):
        return min(self.total, per_core * self.cores)

    def linear_scan_rate(self, nq):
        """kNN only: the structure whose SEMANTICS we match (NearestNeighborsLinear), for scale."""
        t0 = time.perf_counter()
        self.O.knn(self.pts, self.q[:nq], self.args.k, nthreads=self.cores)
        return nq / (time.perf_counter() - t0)


PER_CORE = {"knn": 256, "check_motion": 6000, "rrtc": 128}


def cpu_baseline(args, wl, synth):
    """Bounded sample of the same workload on this box's host cores (rank 0, N=1 only)."""
    arm = CpuArm(args, wl, synth)
    ns = arm.sample_size(PER_CORE[wl])
    arm.run(0, min(ns, arm.cores))          # warm-up
    t0 = time.perf_counter()
    arm.run(0, ns)
    dt = time.perf_counter() - t0
    unit = METRICS[wl][1]
    out = {"value": ns / dt, "unit": unit, "cores": arm.cores, "kind": arm.kind,
           "sample": f"first {ns} of the {arm.total} units; {arm.what}; {arm.cores} threads"}
    if wl == "knn":
        nl = min(args.queries, 2 * arm.cores)
        out["structure"] = "kdtree"
        out["linear_scan"] = {"value": arm.linear_scan_rate(nl), "unit": unit,
                              "sample": f"first {nl} queries, NearestNeighborsLinear semantics (oracle port orc_knn)"}
        out.update(arm.extra)
    return out


def run_reference(args, rank):
    """--impl reference: same metric / unit / config.workload as our arm, timed on the host cores."""
    if rank != 0:
        return
    import prgpu_loader
    synth = prgpu_loader.load().synth
    wl = args.workload
    if wl == "nnf":
        return run_reference_nnf(args, synth)
    arm = CpuArm(args, wl, synth)
    per_step = arm.sample_size({"knn": 64, "check_motion": 1500, "rrtc": 16}[wl])
    span = max(1, arm.total - per_step + 1)
    for i in range(args.warmup):
        arm.run((i * per_step) % span, per_step)
    t0 = time.perf_counter()
    for i in range(args.steps):
        arm.run(((args.warmup + i) * per_step) % span, per_step)
    dt = time.perf_counter() - t0
    metric, unit = METRICS[wl]
    value = per_step * args.steps / dt
    cb = {"value": value, "unit": unit, "cores": arm.cores, "kind": arm.kind,
          "sample": f"{per_step} of the {arm.total} units per step; {arm.what}; {arm.cores} threads"}
    cb.update(arm.extra)
    emit_json({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_string(args, wl)},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def run_reference_nnf(args, synth):
    """NNF on the CPU: the literal reference (frechet/src/*.cpp compiled unmodified, oracle/_ref) holds a
    string-keyed Boost graph per tensor-product node and cannot hold the 4.9e8-cell BASELINE problem, so it
    is timed at the largest size it finishes in seconds; the flat-array oracle port is timed at a reduced
    size too.  Both are scaled-down samples of the same workload, stated as such."""
    import oracle_lib as O
    O.build()
    dh, link, xyzr = synth.default_arm()
    sph, box = synth.obstacles(3, n_spheres=16, n_boxes=16)
    ow = O.World(dh, link, xyzr, sph, box)
    cores = host_cores()
    k, D = 5, 1
    out = {}
    if O.ref_nnf() is not None:
        Ws, Ms = 30, 40
        full, ref, counts, ik = nnf_problem(synth, dh, Ws, Ms, D)
        t0 = time.perf_counter()
        r = O.ref_nnf_run(ow, full, Ws, Ms, k, D, ik, seed=1, solve_seconds=120.0)
        dt = time.perf_counter() - t0
        out["literal_reference"] = {"plans_per_s": 1.0 / dt, "setup_s": r["setup_seconds"], "solve_s": r["solve_seconds"],
                                    "tpg_cells": int(r["n_tpg_vertices"]), "solved": int(r["solved"]),
                                    "sample": f"W={Ws}, M={Ms}: the reference's own frechet/ sources (oracle/_ref), 1 thread"}
    Ws, Ms = 50, 256
    full, ref, counts, ik = nnf_problem(synth, dh, Ws, Ms, D)
    t0 = time.perf_counter()
    o = O.nnf_run(ow, ref, counts, ik, k=k, D=D, max_lazy_iters=4000, nthreads=cores)
    dt = time.perf_counter() - t0
    metric, unit = METRICS["nnf"]
    value = 1.0 / dt
    sample = (f"REDUCED problem W={Ws}, M={Ms} ({len(ref) * o['n_vertices']:.3g} TPG cells): oracle port of NNFrechet "
              f"(flat arrays), setup + {o['lazy_iters']} LazySP iterations, {cores} threads for findKnnNodes")
    emit_json({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": 1,
        "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "replicas only",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": workload_string(args, "nnf")},
        "cpu_baseline": dict({"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample}, **out),
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, args, rank, local_rank, world):
        import numpy as np
        import torch
        import prgpu_loader
        self.np, self.torch = np, torch
        self.P = prgpu_loader.load()
        self.synth = self.P.synth
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the prgpu hot path has no CPU fallback")
        torch.cuda.set_device(local_rank)
        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        self.dev = torch.device("cuda", local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist_mod
            self.dist = dist_mod
            # NCCL prints its version banner on stdout; stdout is reserved for the one JSON line
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            self.dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.hbm_peak, self.peak_src = peaks()
        self.K, self.Wm = args.steps, max(args.warmup, 3)
        self.flush = None

    def comm(self):
        """The library's own communicator (prgpu_comm_init over an NCCL id broadcast by torch.distributed)."""
        if self.world == 1:
            return None
        if getattr(self, "_comm", None) is None:
            self._comm = self.P.sharded.make_comm(self.P, self.dist, self.local_rank, self.rank, self.world)
        return self._comm

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def launches(self):
        return int(self.P.lib().prgpu_diag_launch_count())

    def timed(self, step, flush_l2, K=None):
        """W warm-ups, then K steps bracketed by barrier + synchronize, timed with CUDA events on the launching
        stream; max over ranks.  flush_l2: write a 256 MB buffer between timed iterations (inputs smaller than
        L2), otherwise the K steps are timed back to back.  Returns (ms per step, launches per step, clocks)."""
        torch = self.torch
        K = K or self.K
        if flush_l2 and self.flush is None:
            self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)
        for _ in range(self.Wm):
            step()
        self.barrier()
        sampler = ClockSampler(self.local_rank)
        if self.rank == 0:
            sampler.start()
        l0 = self.launches()
        # The timed region is EXACTLY K steps between barrier + synchronize on both sides, CUDA events on the
        # launching stream.  It is measured three times and the MEDIAN region is reported: a sub-millisecond step
        # makes a single 10-step region a 6 ms measurement, within reach of one scheduling hiccup.
        regions = []
        for rep in range(3):
            if rep:
                self.barrier()
            if not flush_l2:
                b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b.record()
                for _ in range(K):
                    step(record=True)
                e.record()
                torch.cuda.synchronize()
                regions.append(b.elapsed_time(e))
            else:
                t_ms = 0.0
                for _ in range(K):
                    self.flush.fill_(1)
                    b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    b.record()
                    step(record=True)
                    e.record()
                    torch.cuda.synchronize()
                    t_ms += b.elapsed_time(e)
                regions.append(t_ms)
        # every rank must pick the SAME region: rank the regions by their max over ranks
        regions = [self.max_over_ranks(r) for r in regions]
        total_ms = sorted(regions)[1]
        l1 = self.launches()
        self.barrier()
        clocks = sampler.stop() if self.rank == 0 else None
        return total_ms / K, (l1 - l0) / (3 * K), clocks

    def timed_host(self, step, K=None, warm=2):
        """End-to-end: host clock around synchronous C-ABI calls with host buffers; max over ranks."""
        K = K or self.K
        for _ in range(warm):
            step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step()
        self.torch.cuda.synchronize()
        s = self.max_over_ranks(time.perf_counter() - t0)
        self.barrier()
        return s / K


def measure_knn(env, args, want_cpu):
    np, torch, P, synth, dist = env.np, env.torch, env.P, env.synth, env.dist
    rank, world, dev, stream = env.rank, env.world, env.dev, env.stream
    N, Q, k = args.nodes, args.queries, args.k
    # One GPU: BASELINE configs[1] as it stands.  N GPUs (headline): every rank holds the WHOLE tree and its own
    # batch of Q queries (weak scaling: the per-GPU work is configs[1]); one all-gather hands every rank all
    # answers.  The node-sharded strong-scaling step SURVEY §8(e) describes is measured below, in extra.
    pts_all = synth.cloud(0, N)
    q_host = torch.from_numpy(synth.cloud(1 if world == 1 else 1000 + rank, Q)).pin_memory()
    b0, b1 = 0, N
    pts_host = torch.from_numpy(pts_all).pin_memory()
    del pts_all
    n_local = N
    tree = P.KnnIndex(7, device=env.local_rank, capacity=n_local)
    pts_dev = pts_host.to(dev, non_blocking=True)
    q_dev = q_host.to(dev, non_blocking=True)
    tree.add_device(pts_dev, n_local, stream)
    torch.cuda.synchronize()
    tree.build_index()   # part of building the structure, like the CPU arm's k-d tree: outside the timed region
    idx = torch.empty((Q, k), dtype=torch.int32, device=dev)
    dd = torch.empty((Q, k), dtype=torch.float64, device=dev)
    bufs = dict(idx=idx, dist=dd)
    comm = env.comm()
    fused = False
    if world > 1:
        # fused search + exchange: the search kernel writes result rows into every rank's window over NVLink
        fused = comm.enable_peer_windows(12 * k * Q * world) and os.environ.get("PRGPU_BENCH_NO_FUSED") != "1"
        gi = torch.empty((world, Q, k), dtype=torch.int32, device=dev)
        gd = torch.empty((world, Q, k), dtype=torch.float64, device=dev)
    tree.set_profiling(True)
    kernel_ms = []

    def local_search(qd, nq_, k_, io, do):
        tree.nearest_k_device(qd, nq_, k_, io, do, stream=stream)

    def merge(gi_, gd_, nparts, nq_, k_, fi_, fd_):
        P.topk_merge_device(env.local_rank, gi_, gd_, nparts, nq_, k_, fi_, fd_, stream)

    phases = {"local_search_ms": [], "gather_merge_ms": []}
    trace = os.environ.get("PRGPU_BENCH_TRACE") == "1"

    def step(record=False):
        # N > 1: the whole step is one C-ABI call (local search of this rank's queries over the whole tree, ONE
        # all-gather of the packed (distance, index) records over the library's own NCCL communicator)
        if world > 1:
            tree.nearest_k_query_sharded_device(comm, q_dev, Q, k, gi, gd, stream=stream)
        else:
            local_search(q_dev, Q, k, idx, dd)
        if record:
            kernel_ms.append(tree.last_kernel_ms())

    def traced_step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        local_search(q_dev, Q, k, idx, dd)
        ev[1].record()
        if world > 1:
            dist.all_gather_into_tensor(gi.view(world * Q, k), idx)
            dist.all_gather_into_tensor(gd.view(world * Q, k), dd)
        ev[2].record()
        torch.cuda.synchronize()
        phases["local_search_ms"].append(ev[0].elapsed_time(ev[1]))
        phases["gather_merge_ms"].append(ev[1].elapsed_time(ev[2]))

    ms_per_step, launches, clocks = env.timed(step, flush_l2=False)
    kms = env.max_over_ranks(sum(kernel_ms) / len(kernel_ms))
    ist = tree.index_stats()          # before anything below refills the tree
    stats_main = tree.search_stats()
    cross = {}
    if world > 1:
        # cross-check of the collective: every rank's slice of the gathered tables is what that rank found locally
        ci, cd = gi.clone(), gd.clone()
        local_search(q_dev, Q, k, idx, dd)
        torch.cuda.synchronize()
        cross = {"own_slice_equals_local_search": bool(torch.equal(ci[rank], idx)) and bool(torch.equal(cd[rank], dd)),
                 "fused_steps": comm.fused_steps(), "exchange": "peer windows (fused)" if fused else "ncclAllGather"}
    # per-phase breakdown (untimed extra steps, events between the phases): what does not shrink with the shard
    for _ in range(5):
        traced_step()
    env.barrier()
    breakdown = {"local_search_ms": env.max_over_ranks(statistics.median(phases["local_search_ms"])),
                 "gather_merge_ms": env.max_over_ranks(statistics.median(phases["gather_merge_ms"])),
                 "dominant_kernel_ms": kms}
    if trace:
        print(f"[rank {rank}] {breakdown}", file=sys.stderr, flush=True)

    e2e_idx = torch.empty((world, Q, k), dtype=torch.int32).pin_memory()
    e2e_dist = torch.empty((world, Q, k), dtype=torch.float64).pin_memory()

    def e2e_step(reupload=False):
        # The reference-facing call: ompl::NearestNeighbors keeps the nodes (add once, query many times), so
        # the step's inputs are the QUERIES: pinned host queries in, host results out; the tree is the
        # structure's state -- exactly what the CPU reference arm does with its tree in host memory.
        if reupload:
            tree.clear()
            P.capi._check(P.lib().prgpu_knn_add(tree.h, P.capi._ptr(pts_host), n_local))
        if world == 1:
            P.capi._check(P.lib().prgpu_knn_nearest_k(tree.h, P.capi._ptr(q_host), Q, k, None,
                                                      P.capi._ptr(e2e_idx), P.capi._ptr(e2e_dist)))
        else:
            P.capi._check(P.lib().prgpu_knn_nearest_k_query_sharded(tree.h, comm.h, P.capi._ptr(q_host), Q, k,
                                                                    P.capi._ptr(e2e_idx), P.capi._ptr(e2e_dist)))

    e2e_s = env.timed_host(e2e_step)
    e2e_s2 = env.timed_host(lambda: e2e_step(reupload=True), K=max(3, env.K // 3))
    tree.build_index()   # the refills left the index stale; everything below searches the final tree again
    h2d, d2h = Q * 56, world * Q * k * 12
    algo_bytes = n_local * 56 + Q * 56 + Q * k * 12
    tpeak, tsrc = tensor_peak()
    mma_peak = None
    if rank == 0:
        v = P.capi.C.c_double()
        if P.lib().prgpu_diag_mma_tf32_peak(env.local_rank, P.capi.C.byref(v)) == 0:
            mma_peak = v.value

    def hbm_block(ms):
        return {"achieved": algo_bytes / (ms * 1e-3) / 1e9, "peak": env.hbm_peak, "unit": "GB/s",
                "frac": algo_bytes / (ms * 1e-3) / 1e9 / env.hbm_peak, "peak_source": env.peak_src,
                "algorithmic_bytes_per_launch": algo_bytes}

    def scan_roofline(ms):
        """The brute-force scan (knn_collect_kernel): tensor-bound, 16 flop per (query, node) pair."""
        pairs = n_local * Q / (ms * 1e-3)
        tf = pairs * 16.0 / 1e12
        summ = ncu_summary("knn_collect_kernel")
        traffic = None if not summ or summ.get("dram_bytes") is None else int(summ["dram_bytes"] * (n_local / float(N)))
        return {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                "traffic": traffic,
                "traffic_note": "ncu dram read+write bytes of one launch over the full 1e7-node tree (profiles/"
                                "ncu_summary.json), scaled to this rank's shard",
                "kernel": "knn_collect_kernel", "kernel_ms": ms, "peak_source": tsrc,
                "flops_per_launch": n_local * Q * 16.0,
                "tf32_mma_sync": {"peak": mma_peak, "frac": None if not mma_peak else tf / mma_peak,
                                  "pairs_per_s": pairs,
                                  "note": "peak = issue rate of mma.sync.m16n8k8 TF32 measured live by "
                                          "prgpu_diag_mma_tf32_peak in this run (16 warps/SM, 8 independent "
                                          "accumulator fragments per warp)"},
                "issue_slots_busy_pct_ncu": None if not summ else summ.get("issue_slots_busy_pct"),
                "hbm": dict(hbm_block(ms), note="SURVEY §8(d)'s fp64 formula 56*N + 56*Q + 12*Q*k; the scan itself "
                                                "streams the 40 B/node TF32 mirror")}

    indexed = ist["indexed_nodes"] > 0
    if indexed:
        summ = ncu_summary("kd_search_team_kernel")
        roof = dict(hbm_block(kms), bound="hbm", kernel="kd_search_team_kernel", kernel_ms=kms,
                    traffic=None if not summ else summ.get("dram_bytes"),
                    issue_slots_busy_pct_ncu=None if not summ else summ.get("issue_slots_busy_pct"),
                    binds="issue slots + latency: a query walks ~250 tree nodes and ~90 leaves (96 slots each, read as "
                          "their fp32 copy; ~7 % of the half-leaves hold a candidate and are read again in fp64), a "
                          "dependent chain per query (the warps of a block take subtrees over from one another); ncu: "
                          "issue slots 64 % busy, SMs active 76 % of the kernel (the slowest queries set its end). The "
                          "algorithmic bytes are SURVEY §8(d)'s brute-force figure (the whole tree once); the search "
                          "touches ~0.07 % of the tree per query but different leaves for different queries: `traffic` "
                          "(ncu, per launch) is 1.7x the algorithmic bytes",
                    traffic_gbs=None if not summ or not summ.get("dram_bytes") else summ["dram_bytes"] / (kms * 1e-3) / 1e9,
                    index={"build_ms": ist["build_ms"], "builds": ist["builds"],
                           "leaves_per_query": ist["leaves_visited"] / max(1, ist["index_calls"] * Q)})
    else:
        roof = scan_roofline(kms)
    extra = {"phase_breakdown": breakdown, "search_path": "k-d index" if indexed else "brute-force scan"}
    if cross:
        extra["query_sharded_c_abi"] = cross
    if indexed and not args.no_extra:
        # the same step with the index switched off: the batched brute-force / tiled scan BASELINE.json's north
        # star names (TF32 tensor-core filter + fp64 re-rank); identical results
        tree.set_search_mode(2)
        kernel_ms.clear()
        bms, _, _ = env.timed(step, flush_l2=False)
        bk = env.max_over_ranks(sum(kernel_ms) / len(kernel_ms))
        extra["brute_force"] = {"queries_per_s": world * Q / (bms * 1e-3), "ms_per_step": bms, "roofline": scan_roofline(bk),
                                "note": "prgpu_knn_set_search_mode(h, 2): no index; this is the path that scales with "
                                        "the shard size when the tree is node-sharded"}
        tree.set_search_mode(0)
        step()
        step()

    # ---- throughput with two batches in flight (one GPU): a 4096-query batch is one query per warp slot of the
    #      machine, so the SMs idle for a quarter of the kernel while the heaviest queries are finished; a second
    #      stream lets the next batch's blocks take the freed slots.  Not the headline (a step there is one batch at
    #      a time); reported because a serving loop would issue batches this way ----
    if indexed and world == 1 and not args.no_extra:
        s2 = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        q2 = [q_dev, torch.from_numpy(synth.cloud(2, Q)).to(dev)]
        o2 = [(torch.empty_like(idx), torch.empty_like(dd)) for _ in range(2)]
        tree.set_profiling(False)
        torch.cuda.synchronize()
        for j in range(4):
            tree.nearest_k_device(q2[j & 1], Q, k, o2[j & 1][0], o2[j & 1][1], stream=s2[j & 1].cuda_stream)
        torch.cuda.synchronize()
        nb = 2 * max(args.steps, 10)
        t0 = time.perf_counter()
        for j in range(nb):
            tree.nearest_k_device(q2[j & 1], Q, k, o2[j & 1][0], o2[j & 1][1], stream=s2[j & 1].cuda_stream)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        local_search(q_dev, Q, k, idx, dd)
        torch.cuda.synchronize()
        extra["two_batches_in_flight"] = {
            "queries_per_s": nb * Q / dt, "ms_per_batch": dt / nb * 1e3, "batches": nb,
            "identical_to_one_at_a_time": bool(torch.equal(o2[0][0], idx)) and bool(torch.equal(o2[0][1], dd)),
            "note": "batches issued alternately on two streams (host clock around the whole run, synchronised on "
                    "both sides); every batch is a complete prgpu_knn_nearest_k_device call"}
        tree.set_profiling(True)

    # ---- strong scaling of ONE configs[1] batch (Q queries in total), both partitionings ----
    if world > 1 and not args.no_extra:
        strong = {}
        q1 = torch.from_numpy(synth.cloud(1, Q)).to(dev)      # the single-GPU batch
        fi = torch.empty((Q, k), dtype=torch.int32, device=dev)
        fd = torch.empty((Q, k), dtype=torch.float64, device=dev)
        # (a) tree nodes sharded: every rank searches its N/G nodes for all Q queries, one
        #     all-gather of the packed per-shard top-k records, merge kernel
        s0, s1 = rank * N // world, (rank + 1) * N // world
        shard = P.KnnIndex(7, device=env.local_rank, capacity=s1 - s0)
        shard.set_index_base(s0)
        shard.add_device(pts_dev[s0:s1].contiguous(), s1 - s0, stream)
        shard.build_index()
        for label, mode in (("indexed", 0), ("brute_force", 2)):
            shard.set_search_mode(mode)
            nms, _, _ = env.timed(lambda record=False: shard.nearest_k_sharded_device(comm, q1, Q, k, fi, fd,
                                                                                       stream=stream), flush_l2=False)
            strong["node_sharded_" + label] = {"queries_per_s": Q / (nms * 1e-3), "ms_per_step": nms,
                                               "path": "k-d index per shard" if mode == 0 else
                                               "brute-force scan (TF32 filter) per shard"}
        ref_i, ref_d = fi.clone(), fd.clone()
        strong["redo_steps"] = comm.redo_steps()
        shard.close()
        # (b) queries split (Q/G per rank), whole tree on every rank, results all-gathered
        if Q % world == 0:
            ql = Q // world
            qs = q1[rank * ql:(rank + 1) * ql].contiguous()
            ai = torch.empty((world, ql, k), dtype=torch.int32, device=dev)
            ad = torch.empty((world, ql, k), dtype=torch.float64, device=dev)
            qms, _, _ = env.timed(lambda record=False: tree.nearest_k_query_sharded_device(comm, qs, ql, k, ai, ad,
                                                                                            stream=stream), flush_l2=False)
            strong["query_sharded"] = {"queries_per_s": Q / (qms * 1e-3), "ms_per_step": qms,
                                       "identical_to_node_sharded": bool(torch.equal(ai.view(Q, k), ref_i)) and
                                       bool(torch.equal(ad.view(Q, k), ref_d))}
        strong["note"] = ("ONE batch of Q queries over all GPUs (total work fixed).  An indexed search is a dependent "
                          "chain per query (~0.2 ms even with 8 warps per query), so a 0.6 ms step cannot be cut 8 ways; "
                          "node sharding is what lets a tree outgrow one GPU, and its brute-force path is the one whose "
                          "time shrinks with the shard")
        extra["strong_scaling_one_batch"] = strong

    # ---- BASELINE configs[4], measured instead of argued: the NN step of a lock-step planner whose TREE is
    #      node-sharded.  One growTree = nearest(q_rand) over the tree + one appended node (RRTConnect.cpp:181,
    #      :209).  Sharded: every rank scans its 1/G of the nodes, ONE all-gather of the per-shard best record,
    #      merge; the new node goes to shard (i mod G).  Beside it: the same step on one GPU holding the whole
    #      tree (what query-sharding does: each GPU plans its own queries on its own trees). ----
    if world > 1 and not args.no_extra:
        lockstep = {}
        for n_tree, q_batch in ((100_000, 1), (10_000_000, 1), (10_000_000, 64)):
            cloud = synth.cloud(21, n_tree + 512)
            mine = torch.from_numpy(np.ascontiguousarray(cloud[:n_tree][rank::world])).to(dev)
            shard = P.KnnIndex(7, device=env.local_rank, capacity=len(mine) + 1024)
            shard.add_device(mine, len(mine), stream)
            shard.set_index_base(0)   # (records carry shard-local indices here; the timing does not depend on them)
            whole = P.KnnIndex(7, device=env.local_rank, capacity=n_tree + 1024)
            whole.add_device(torch.from_numpy(cloud[:n_tree]).to(dev), n_tree, stream)
            newn = torch.from_numpy(cloud[n_tree:n_tree + 512]).to(dev)
            qb = torch.from_numpy(synth.cloud(22, 256 * q_batch)).to(dev)
            oi = torch.empty((q_batch, 1), dtype=torch.int32, device=dev)
            od = torch.empty((q_batch, 1), dtype=torch.float64, device=dev)
            reps = 200

            def grow_sharded():
                for i in range(reps):
                    shard.nearest_k_sharded_device(comm, qb[(i % 256) * q_batch:(i % 256 + 1) * q_batch], q_batch, 1,
                                                   oi, od, stream=stream)
                    if i % world == rank:
                        shard.add_device(newn[i:i + 1], 1, stream)

            def grow_single():
                for i in range(reps):
                    whole.nearest_k_device(qb[(i % 256) * q_batch:(i % 256 + 1) * q_batch], q_batch, 1, oi, od,
                                           stream=stream)
                    whole.add_device(newn[i:i + 1], 1, stream)
            out = {}
            for name, fn in (("node_sharded", grow_sharded), ("single_gpu_whole_tree", grow_single)):
                fn()
                torch.cuda.synchronize()
                env.barrier()
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                out[name + "_us_per_growTree_nn"] = env.max_over_ranks((time.perf_counter() - t0) / reps * 1e6)
            lockstep[f"tree={n_tree},queries_per_step={q_batch}"] = out
            shard.close()
            whole.close()
            del mine, newn, qb
        lockstep["note"] = ("per growTree: nearest over the tree + one appended node; node_sharded = "
                            "prgpu_knn_nearest_k_sharded_device (local scan of N/G nodes + one NCCL all-gather + merge) "
                            "with the new node appended to shard i mod G; single_gpu_whole_tree = prgpu_knn_nearest_k_"
                            "device on one GPU (query-sharded planning: no collective).  Planner trees of the 4096-query "
                            "batch hold tens to hundreds of nodes: far left of the first row")
        extra["configs4_lockstep_node_sharded_nn"] = lockstep

    # ---- the other shapes of the sweep (single GPU) ----
    if world == 1 and not args.no_extra:
        def time_knn(t_, nq, kk, reps=5):
            qd = torch.from_numpy(synth.cloud(7, nq)).to(dev)
            i_ = torch.empty((nq, kk), dtype=torch.int32, device=dev)
            d_ = torch.empty((nq, kk), dtype=torch.float64, device=dev)
            for _ in range(2):
                t_.nearest_k_device(qd, nq, kk, i_, d_, stream=stream)
            b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b.record()
            for _ in range(reps):
                t_.nearest_k_device(qd, nq, kk, i_, d_, stream=stream)
            e.record()
            torch.cuda.synchronize()
            return b.elapsed_time(e) / reps
        tree.clear()
        tree.add_device(pts_dev, n_local, stream)
        torch.cuda.synchronize()
        tree.build_index()
        extra["knn_k1_queries_per_s"] = Q / (time_knn(tree, Q, 1) * 1e-3)
        sweep = {}
        for n_s in (10_000, 100_000, 1_000_000):
            if n_s >= N:
                continue
            ts = P.KnnIndex(7, device=env.local_rank, capacity=n_s)
            ts.add_device(pts_dev[:n_s].contiguous(), n_s, stream)   # (auto mode: these sizes stay on the filter scan)
            for kk in (16, 1):
                sweep[f"N={n_s},k={kk}"] = Q / (time_knn(ts, Q, kk, reps=10) * 1e-3)
            ts.close()
        extra["knn_sweep_queries_per_s"] = sweep
        ms1 = time_knn(tree, 1, 1, reps=20)
        extra["single_query_scan"] = {"ms": ms1, "hbm_gbs": n_local * 56 / (ms1 * 1e-3) / 1e9,
                                      "frac_of_hbm_peak": n_local * 56 / (ms1 * 1e-3) / 1e9 / env.hbm_peak,
                                      "note": "Q=1 (a planner's nearest() on a 1e7-node tree): the HBM-bound shape, "
                                              "exact fp64 kernel"}
    stats = stats_main
    del pts_dev
    tree.close()
    torch.cuda.empty_cache()
    metric, unit = METRICS["knn"]
    line = {
        "metric": metric, "value": world * Q / (ms_per_step * 1e-3), "unit": unit, "ms_per_step": ms_per_step,
        "scaling": "weak",
        "config": {
            "workload": workload_string(args, "knn"),
            "parallelism": "single GPU" if world == 1 else f"{world} GPUs, each holding the whole tree and its own batch "
                           f"of {Q} queries ({world * Q} queries per step in total); one C-ABI call per step "
                           f"(prgpu_knn_nearest_k_query_sharded_device): " + (
                               "the index search kernel writes every result row straight into every rank's window (NVLink "
                               "peer stores from the kernel's epilogue) + one flag-exchange kernel: no collective call"
                               if fused else "local indexed search + ONE NCCL all-gather of the packed (distance, index) "
                               "records") + "; every rank ends with every answer",
            "l2": f"inputs larger than L2 ({n_local * 56 / 1e6:.0f} MB tree shard per GPU, fp64 SoA)",
            "e2e_timing": "host clock around synchronous C-ABI calls, max over ranks; the tree is the "
                          "NearestNeighbors structure's state (as in the CPU reference arm), pinned host queries in / "
                          "host results out every step (one GPU: the caller's result buffers are pinned, so the search kernel "
                          "writes the d2h bytes into them itself as queries finish -- zero copy -- instead of two copies "
                          "after the search); e2e.with_tree_reupload also refills the tree from host memory",
            "arithmetic": "TF32 tensor-core filter -> fp32 re-cut -> fp64 exact distances + proof; results "
                          "bit-identical to the fp64 reference",
            "filter_fallback_queries": stats["fallback_queries"],
        },
        "roofline": roof,
        "e2e": {"value": world * Q / e2e_s, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "with_tree_reupload": {"value": world * Q / e2e_s2, "unit": unit,
                                       "h2d_bytes_per_step": n_local * 56 + Q * 56, "d2h_bytes_per_step": d2h}},
        "gpu_launches_per_step": launches, "clocks": clocks, "extra": extra,
    }
    if want_cpu:
        line["cpu_baseline"] = cpu_baseline(args, "knn", synth)
    return line


def measure_check_motion(env, args, want_cpu):
    np, torch, P, synth = env.np, env.torch, env.P, env.synth
    rank, world, dev, stream = env.rank, env.world, env.dev, env.stream
    E = args.edges
    dh, link, xyzr = synth.default_arm()
    sph, box = synth.obstacles(3)
    world_h = P.World(dh, link, xyzr, sph, box, device=env.local_rank)
    a_all, b_all = synth.edges(2, E)
    e0_, e1_ = rank * E // world, (rank + 1) * E // world
    a_host = torch.from_numpy(a_all[e0_:e1_]).pin_memory()
    b_host = torch.from_numpy(b_all[e0_:e1_]).pin_memory()
    n_local = e1_ - e0_
    a_dev, b_dev = a_host.to(dev), b_host.to(dev)
    valid = torch.empty(n_local, dtype=torch.uint8, device=dev)
    world_h.set_profiling(True)
    kernel_ms = []

    def step(record=False):
        world_h.check_motion_batch_device(a_dev, b_dev, n_local, 0.01, valid, stream=stream)
        if record:
            kernel_ms.append(world_h.last_kernel_ms())

    ms_per_step, launches, clocks = env.timed(step, flush_l2=True)
    kms = env.max_over_ranks(sum(kernel_ms) / len(kernel_ms))
    e2e_valid = torch.empty(n_local, dtype=torch.uint8).pin_memory()

    def e2e_step():
        P.capi._check(P.lib().prgpu_check_motion_batch(world_h.h, P.capi._ptr(a_host), P.capi._ptr(b_host),
                                                       n_local, 0.01, 0, 0, P.capi._ptr(e2e_valid)))
    e2e_s = env.timed_host(e2e_step)
    valid_fraction = float(valid.float().mean().item())
    weak = None
    if world > 1 and not args.no_extra:
        # the same 1e6-edge batch on EVERY GPU (per-GPU work fixed): what splitting independent batches buys
        aw, bw = torch.from_numpy(a_all).to(dev), torch.from_numpy(b_all).to(dev)
        vw = torch.empty(E, dtype=torch.uint8, device=dev)
        wms, _, _ = env.timed(lambda record=False: world_h.check_motion_batch_device(aw, bw, E, 0.01, vw, stream=stream),
                              flush_l2=True)
        weak = {"value": world * E / (wms * 1e-3), "unit": "edges/s", "ms_per_step": wms,
                "note": f"weak scaling: {E} edges per GPU, {world * E} per step"}
        del aw, bw, vw
    algo_bytes = n_local * (2 * 56 + 1)
    kernels = ["cm_sample2_kernel", "cm_lists_kernel<0>", "cm_certify_b_kernel", "cm_lists_kernel<1>"]
    issue = {kname: (ncu_summary(kname) or {}).get("issue_slots_busy_pct") for kname in kernels}
    lanes = {kname: (ncu_summary(kname) or {}).get("threads_per_inst") for kname in kernels}
    roof = {"bound": "hbm", "achieved": algo_bytes / (kms * 1e-3) / 1e9, "peak": env.hbm_peak, "unit": "GB/s",
            "frac": algo_bytes / (kms * 1e-3) / 1e9 / env.hbm_peak, "traffic": ncu_traffic(kernels),
            "kernel": "+".join(kernels), "kernel_ms": kms, "peak_source": env.peak_src,
            "algorithmic_bytes_per_launch": algo_bytes,
            "binds": "issue / latency (fp32 FK chain + dependent L2 loads), not HBM: SURVEY §8(d) puts this kernel "
                     "~1e3x into compute-bound territory, the HBM fraction is reported because the contract asks for it",
            "issue_slots_busy_pct_ncu": issue, "threads_per_inst_ncu": lanes}
    metric, unit = METRICS["check_motion"]
    world_h.close()
    line = {
        "metric": metric, "value": E / (ms_per_step * 1e-3), "unit": unit, "ms_per_step": ms_per_step,
        "scaling": "strong",
        "config": {"workload": workload_string(args, "check_motion"),
                   "parallelism": "single GPU" if world == 1 else f"edges split over {world} GPUs, no collective",
                   "l2": "L2 flushed by a 256 MB write between timed iterations",
                   "e2e_timing": "host clock around synchronous C-ABI calls (prgpu_check_motion_batch, host buffers), "
                                 "max over ranks",
                   "valid_fraction": valid_fraction},
        "roofline": roof,
        "e2e": {"value": E / e2e_s, "unit": unit, "h2d_bytes_per_step": n_local * 112, "d2h_bytes_per_step": n_local},
        "gpu_launches_per_step": launches, "clocks": clocks, "extra": {} if weak is None else {"weak_scaling": weak},
    }
    if want_cpu:
        line["cpu_baseline"] = cpu_baseline(args, "check_motion", synth)
    return line


def measure_rrtc(env, args, want_cpu):
    np, torch, P, synth = env.np, env.torch, env.P, env.synth
    rank, world, dev, stream = env.rank, env.world, env.dev, env.stream
    Bq = args.plans
    dh, link, xyzr = synth.default_arm()
    sph, box = synth.obstacles(3)
    world_h = P.World(dh, link, xyzr, sph, box, device=env.local_rank)
    starts_all, goals_all = plan_problems(synth, world_h, Bq)
    p0, p1 = P.sharded.shard_bounds(Bq, world)[rank]
    n_local = p1 - p0
    s_host = torch.from_numpy(starts_all[p0:p1].copy()).pin_memory()
    g_host = torch.from_numpy(goals_all[p0:p1].copy()).pin_memory()
    s_dev, g_dev = s_host.to(dev), g_host.to(dev)
    rb = P.RrtcBatch(world_h, n_local, -np.pi, np.pi, extension_type="ext-con", seed=4, query_id_base=p0,
                     max_iters=args.plan_iters, max_nodes=2048)
    kernel_ms = []

    def step(record=False):
        if record:
            b_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b_.record()
        rb.run_device(s_dev, g_dev, stream)
        if record:
            e_.record()
            e_.synchronize()
            kernel_ms.append(b_.elapsed_time(e_))

    ms_per_step, launches, clocks = env.timed(step, flush_l2=True)
    kms = env.max_over_ranks(sum(kernel_ms) / len(kernel_ms))

    def e2e_step():
        P.capi._check(P.lib().prgpu_rrtc_batch_run(rb.h, P.capi._ptr(s_host), P.capi._ptr(g_host)))
        rb.summary()
    e2e_s = env.timed_host(e2e_step)
    summ = rb.summary()
    weak = None
    if world > 1 and not args.no_extra:
        # all 4096 queries on EVERY GPU (different sample streams per rank): the batch time is set by the queries
        # that exhaust their budget, so planning more batches side by side is what more GPUs buy
        sw, gw_ = torch.from_numpy(starts_all.copy()).to(dev), torch.from_numpy(goals_all.copy()).to(dev)
        rbw = P.RrtcBatch(world_h, Bq, -np.pi, np.pi, extension_type="ext-con", seed=4 + 1000 * rank, query_id_base=0,
                          max_iters=args.plan_iters, max_nodes=2048)
        wms, _, _ = env.timed(lambda record=False: rbw.run_device(sw, gw_, stream), flush_l2=True)
        weak = {"value": world * Bq / (wms * 1e-3), "unit": "plans/s", "ms_per_step": wms,
                "note": f"weak scaling: {Bq} queries per GPU, {world * Bq} per step"}
        rbw.close()
    nodes = int(summ["n_start"].sum() + summ["n_goal"].sum())
    algo_bytes = n_local * 112 + nodes * 60 + n_local * 16
    kernels = ["rrtc_batch_kernel", "rrtc_team_kernel"]
    roof = {"bound": "hbm", "achieved": algo_bytes / (kms * 1e-3) / 1e9, "peak": env.hbm_peak, "unit": "GB/s",
            "frac": algo_bytes / (kms * 1e-3) / 1e9 / env.hbm_peak, "traffic": ncu_traffic(kernels),
            "kernel": "+".join(kernels), "kernel_ms": kms, "peak_source": env.peak_src,
            "algorithmic_bytes_per_launch": algo_bytes,
            "binds": "latency: a query is a dependent chain sample -> nearest -> validate -> append; the batch time is "
                     "set by the queries that exhaust the sample budget",
            "issue_slots_busy_pct_ncu": {kn: (ncu_summary(kn) or {}).get("issue_slots_busy_pct") for kn in kernels}}
    metric, unit = METRICS["rrtc"]
    line = {
        "metric": metric, "value": Bq / (ms_per_step * 1e-3), "unit": unit, "ms_per_step": ms_per_step,
        "scaling": "strong",
        "config": {"workload": workload_string(args, "rrtc"),
                   "parallelism": "single GPU" if world == 1 else f"queries split over {world} GPUs, no collective",
                   "l2": "L2 flushed by a 256 MB write between timed iterations",
                   "e2e_timing": "host clock around synchronous C-ABI calls (prgpu_rrtc_batch_run with host starts/goals "
                                 "+ prgpu_rrtc_batch_summary), max over ranks",
                   "solved_fraction": float((summ["status"] == 6).mean()),
                   "mean_samples_per_plan": float(summ["iterations"].mean())},
        "roofline": roof,
        "e2e": {"value": Bq / e2e_s, "unit": unit, "h2d_bytes_per_step": n_local * 112, "d2h_bytes_per_step": n_local * 16},
        "gpu_launches_per_step": launches, "clocks": clocks, "extra": {} if weak is None else {"weak_scaling": weak},
    }
    rb.close()
    world_h.close()
    if want_cpu:
        line["cpu_baseline"] = cpu_baseline(args, "rrtc", synth)
    return line


def measure_nnf(env, args, want_cpu):
    """NNF Frechet following (BASELINE.json configs[3]).  Replicas only: every rank plans the same
    problem; rank 0 reports.  value = complete plans (setup + LazySP solve) per second."""
    np, torch, P, synth = env.np, env.torch, env.P, env.synth
    dh, link, xyzr = synth.default_arm()
    sph, box = synth.obstacles(3, n_spheres=16, n_boxes=16)
    gw = P.World(dh, link, xyzr, sph, box, device=env.local_rank)
    W, M, k, D = 200, 1024, 5, 1
    _, ref, counts, ik = nnf_problem(synth, dh, W, M, D)
    # warm-up (untimed): the same problem set up once and searched for a few LazySP iterations
    gwarm = P.Nnf(gw, ref, counts, ik, num_nn=k, discretization=D)
    gwarm.solve(max_lazy_iters=40)
    gwarm.close()
    torch.cuda.synchronize()
    l0 = env.launches()
    t0 = time.perf_counter()
    g = P.Nnf(gw, ref, counts, ik, num_nn=k, discretization=D)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    # the first search sweeps the whole band (dirty level 0): the launch the roofline is quoted on
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    t0 = time.perf_counter()
    r1 = g.solve(max_lazy_iters=1)
    first_search_ms = g.last_dp_ms()
    info = g.band_info()
    r = g.solve(max_lazy_iters=20000) if not r1["solved"] else r1
    solve_s = time.perf_counter() - t0
    info_final = g.band_info()
    nnf_clocks = sampler.stop()
    launches = env.launches() - l0
    iters = int(r1["lazy_iters"] + (0 if r1["solved"] else r["lazy_iters"]))
    nE = g.n_edges
    # bytes one full-band search touches by construction: per cell one fp64 write + (1 + 2 indeg) fp64 reads,
    # mean in-degree E/V
    indeg = nE / float(g.n_vertices)
    algo = info["cells"] * 8.0 * (1.0 + 1.0 + 2.0 * indeg)
    summ = ncu_summary("nnf_flow_dp_kernel")
    metric, unit = METRICS["nnf"]
    line = {
        "metric": metric, "value": 1.0 / (setup_s + solve_s), "unit": unit,
        "ms_per_step": (setup_s + solve_s) * 1e3, "scaling": "replicas only",
        "config": {"workload": workload_string(args, "nnf"), "R": len(ref), "nn_vertices": g.n_vertices,
                   "tpg_cells": len(ref) * g.n_vertices, "setup_s": setup_s, "solve_s": solve_s,
                   "lazy_iterations": iters, "ms_per_lazy_iteration": solve_s / max(1, iters) * 1e3,
                   "solved": int(r["solved"]), "final_frechet_error": r["final_error"],
                   "edges_evaluated": int(r["n_evaluated"]), "edges_in_collision": int(r["n_collisions"]),
                   "band": info, "band_at_the_end": info_final, "l2": "band table 8 B/cell larger than L2 on the first search"},
        "roofline": {"bound": "hbm", "achieved": algo / (first_search_ms * 1e-3) / 1e9, "peak": env.hbm_peak,
                     "unit": "GB/s", "frac": algo / (first_search_ms * 1e-3) / 1e9 / env.hbm_peak,
                     "traffic": None if not summ else summ.get("dram_bytes"), "kernel": "nnf_flow_dp_kernel",
                     "kernel_ms": first_search_ms, "peak_source": env.peak_src,
                     "algorithmic_bytes_per_launch": algo,
                     "note": "the FIRST search of the solve (every band cell recomputed); later searches recompute "
                             "only the levels at or after a knocked-out edge and are latency-bound on the level chain"},
        "e2e": {"value": 1.0 / (setup_s + solve_s), "unit": unit,
                "h2d_bytes_per_step": int(ik.nbytes + ref.nbytes),
                # the kNN table of the NN graph (host-side graph construction) + the path and one record per launch
                "d2h_bytes_per_step": int(len(ik) * k * 4 + len(r.get("path", [])) * 4 + 96 * max(1, launches))},
        "gpu_launches_per_step": launches, "clocks": nnf_clocks, "extra": {},
    }
    ls = g.loop_stats()
    tot = float(sum(ls["phase_cycles"].values())) or 1.0
    line["extra"]["lazysp_loop"] = {
        "kernel_launches": ls["launches"], "columns_repaired": ls["repaired_vertices"],
        "columns_repaired_per_iteration": ls["repaired_vertices"] / max(1, iters),
        "nn_vertices": g.n_vertices, "wide_waves_handed_to_full_sweep": ls["wide_waves_handed_to_full_sweep"],
        "sm_clocks_during_solve": nnf_clocks,
        "phase_share_of_loop_kernel": {kname: c / tot for kname, c in ls["phase_cycles"].items()},
        "note": "the LazySP loop runs in one persistent kernel (hop chase -> per-hop pass -> evaluateEdge -> "
                "bookkeeping -> local repair of the bottleneck table); the host reads one record per launch"}
    g.close()
    gw.close()
    if want_cpu:
        import oracle_lib as O
        O.build()
        ow = O.World(dh, link, xyzr, sph, box)
        Ws, Ms = 50, 256
        _, refs, cs, iks = nnf_problem(synth, dh, Ws, Ms, D)
        cores = host_cores()
        t0 = time.perf_counter()
        o = O.nnf_run(ow, refs, cs, iks, k=k, D=D, max_lazy_iters=4000, nthreads=cores)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": 1.0 / dt, "unit": unit, "cores": cores, "kind": "port",
            "sample": f"REDUCED problem W={Ws}, M={Ms} ({len(refs) * o['n_vertices']:.3g} TPG cells, "
                      f"{len(ref) * line['config']['nn_vertices'] / (len(refs) * o['n_vertices']):.0f}x fewer than the "
                      f"GPU run): oracle port, setup + {o['lazy_iters']} LazySP iterations; the full-size CPU run does "
                      f"not finish in minutes (see --workload nnf --impl reference for the literal reference at toy size)"}
    return line


def measure_dropin_a(env, args, want_cpu):
    """One planner instance at a time through the UNMODIFIED reference planner source over the GPU
    NearestNeighbors / validators (oracle/_ref/libdropin_a.so): a launch + sync per nearest / checkMotion / add.
    SURVEY §7.3 item 3: this path is latency-bound -- one call is a few tens of microseconds whatever the GPU does
    in it (arguments and results travel through a mapped host window, tiny launches read the world's tables in
    place) -- and about breaks even with one CPU thread on the reference's small trees; measured, not hidden."""
    import plugin_lib
    np, synth = env.np, env.synth
    L = plugin_lib.dropin()
    if L is None:
        return {"unavailable": "oracle/_ref/libdropin_a.so not built (needs /root/reference at build time)"}
    dh, link, xyzr = synth.default_arm()
    sph, box = synth.obstacles(3)
    w = plugin_lib.PluginWorld(dh, link, xyzr, sph, box, device=env.local_rank, library=L)
    gw = env.P.World(dh, link, xyzr, sph, box, device=env.local_rank)
    starts, goals = plan_problems(synth, gw, args.plans)
    gw.close()
    n = 64
    w.rrtc_bench(starts[:4], goals[:4], -np.pi, np.pi, max_iters=args.plan_iters, seed=4)
    sec, solved, iters = w.rrtc_bench(starts[:n], goals[:n], -np.pi, np.pi, max_iters=args.plan_iters, seed=4)
    out = {"plans_per_s": n / sec, "us_per_sample": 1e6 * sec / max(1, iters), "solved": int(solved), "queries": n,
           "note": "pr_ompl::RRTConnect exactly as shipped by the reference (compiled unmodified) with "
                   "GpuNearestNeighbors<Motion*> + GpuMotionValidator underneath, one query after the other"}
    import oracle_lib as O
    if want_cpu and O.ref() is not None:
        O.build()
        ow = O.World(dh, link, xyzr, sph, box)
        t0 = time.perf_counter()
        for qi in range(n):
            ow.rrtc_solve(starts[qi], goals[qi], -np.pi, np.pi, ext=(0, 1), seed=4, query=qi,
                          max_iters=args.plan_iters, max_nodes=4096, impl="ref")
        out["cpu_reference_plans_per_s_1_thread"] = n / (time.perf_counter() - t0)
    return out


MEASURE = {"knn": measure_knn, "check_motion": measure_check_motion, "rrtc": measure_rrtc, "nnf": measure_nnf}


def block(line):
    """A workload's line as a block of the headline line's `extra`."""
    keep = ("metric", "value", "unit", "ms_per_step", "scaling", "config", "roofline", "e2e", "gpu_launches_per_step",
            "cpu_baseline")
    out = {k: line[k] for k in keep if k in line}
    if line.get("extra"):
        out["extra"] = line["extra"]
    return out


def run_ours(args, rank, local_rank, world):
    env = Env(args, rank, local_rank, world)
    want_cpu = world == 1 and not args.no_cpu_baseline and rank == 0
    line = MEASURE[args.workload](env, args, want_cpu)
    if args.workload == "knn" and not args.no_extra:
        for wl in ("check_motion", "rrtc"):
            sub = MEASURE[wl](env, args, want_cpu)
            line["extra"][wl] = block(sub)
        if world == 1:
            try:
                line["extra"]["nnf"] = block(measure_nnf(env, args, want_cpu))
            except Exception as e:  # the headline must not die with an optional block
                line["extra"]["nnf"] = {"error": repr(e)}
            try:
                line["extra"]["dropin_a"] = measure_dropin_a(env, args, want_cpu)
            except Exception as e:
                line["extra"]["dropin_a"] = {"error": repr(e)}
    if rank != 0:
        return
    out = {
        "metric": line["metric"], "value": line["value"], "unit": line["unit"], "n_gpus": world, "steps": env.K,
        "warmup": env.Wm, "ms_per_step": line["ms_per_step"], "higher_is_better": True, "scaling": line["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": line["config"],
        "roofline": line["roofline"], "e2e": line["e2e"],
        "gpu_launches": int(round(line["gpu_launches_per_step"] * env.K)), "clocks": line["clocks"],
        "extra": line["extra"],
    }
    out["config"] = dict(out["config"], timed_region=f"exactly {env.K} steps between barrier + synchronize, CUDA events on "
                         "the launching stream, max over ranks; measured three times, the median region is reported")
    if "cpu_baseline" in line:
        out["cpu_baseline"] = line["cpu_baseline"]
    emit_json(out)


_JSON_OUT = None


def reserve_stdout():
    """stdout carries exactly one JSON line: whatever libraries print (the NCCL version banner, for one)
    goes to stderr, the line is written to the saved descriptor."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_json(obj):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    args = parse()
    reserve_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, local_rank, world)
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
