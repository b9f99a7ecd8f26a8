#!/usr/bin/env python
"""bench.py -- converged multi-vehicle OBCA solves/sec (batched) on N B200s of one node.

Workload (BASELINE.json configs[1] batched as configs[3]): the 4-vehicle centralised conflict-resolution NLP
(multi_vehicle_planner.py:343-480) in the parking-lot scenario, one batch of independent instances per GPU that differ
in their initial offsets.  A "step" = one batched joint solve of the per-GPU batch from
the reference's warm start (single-vehicle solutions + pair duals); weak scaling (per-GPU batch fixed).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU restatement (oracle port) on the host cores

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AGENTS = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
METRIC = "converged multi-vehicle OBCA solves/sec (batched)"
# SURVEY.md section 8(d) convention for the 4-vehicle joint problem, per IPM iteration and instance
DRAM_BYTES_PER_ITER_NCU = 21.9e6  # measured, see roofline.traffic_note
BYTES_PER_ITER = 3.456e6
FLOPS_PER_ITER_CONVENTION = 767e6
FP64_PEAK_TFLOPS = 37.0  # B200 data sheet (non-tensor FP64); not measured on this pool


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(self.rows)}


_WORKER = {}


def _cpu_worker_init(prob, guess_list):
    """Per-process setup of the oracle port (sympy block generation + sparse NLP assembly, untimed)."""
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import ipm
    from oracle.nlp import CollocationNLP

    _WORKER["ipm"] = ipm
    _WORKER["cases"] = []
    for b, guess in enumerate(guess_list):
        nlp = CollocationNLP(prob.instance(b))
        _WORKER["cases"].append((nlp, nlp.init_slacks(nlp.pack(guess))))


def cpu_sample(args_tuple):
    """Bounded CPU sample: `iters` interior-point iterations of the oracle port on one joint instance."""
    b, iters = args_tuple
    ipm = _WORKER["ipm"]
    nlp, x0 = _WORKER["cases"][b % len(_WORKER["cases"])]
    t0 = time.perf_counter()
    res = ipm.solve(nlp, x0, ipm.IpmOptions(max_iter=iters))
    return (time.perf_counter() - t0) / max(1, res.iters), res.iters


class CpuBaseline:
    """Oracle port on host cores: seconds per IPM iteration with `n_proc` single-threaded processes running in parallel
    -> solves/sec at a given iteration count per solve (a full CPU solve of the 4-vehicle problem takes minutes)."""

    def __init__(self, plan, n_proc, n_cases=1):
        import multiprocessing as mp

        self.n_proc = n_proc
        B = plan.problem.batch or 1
        guesses = [plan.guess.instance(b % B) for b in range(n_cases)]
        self.pool = mp.get_context("spawn").Pool(n_proc, initializer=_cpu_worker_init, initargs=(plan.problem, guesses))
        self.pool.map(cpu_sample, [(0, 0)] * n_proc)  # make sure every worker finished its setup

    def sample(self, iters_per_solve, sample_iters=4):
        t0 = time.perf_counter()
        out = self.pool.map(cpu_sample, [(b, sample_iters) for b in range(self.n_proc)], chunksize=1)
        wall = time.perf_counter() - t0
        sec_per_iter = float(np.mean([o[0] for o in out]))
        value = self.n_proc / (sec_per_iter * max(1.0, iters_per_solve))
        return {
            "value": value,
            "unit": "solves/s",
            "cores": self.n_proc,
            "kind": "port",
            "sample": "%d IPM iterations of the oracle port (oracle/ipm.py, SuperLU) on %d joint instance(s) in parallel, %.3f s/iteration, "
            "extrapolated to %d iterations per solve; sample wall %.1f s" % (sample_iters, self.n_proc, sec_per_iter, int(iters_per_solve), wall),
        }

    def close(self):
        self.pool.close()
        self.pool.join()


def mpc_latency(rl_file, device, steps):
    """Second half of BASELINE.json's metric: p50 / p99 latency of one distributed-MPC control step (4 vehicles, horizon 30,
    vehicle_follower.py:main), host parameters in -> first input + predictions on the host, all vehicles in one launch."""
    from conflict_rez_b200.control.vehicle_follower import MultiDistributedFollower
    from conflict_rez_b200.pytypes import VehicleState

    heads = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": np.pi / 2}
    mdf = MultiDistributedFollower(rl_file, {a: True for a in AGENTS}, {a: {} for a in AGENTS}, {a: VehicleState() for a in AGENTS}, heads, device=device)
    mdf.setup_multi_vehicles()
    mdf.solve(num_iter=steps)
    t = 1e3 * np.array(mdf.step_time[5:])  # first steps warm the allocator / pinned buffers
    fails = int(sum(v.N - 1 - v.back_up_steps > 0 for v in mdf.vehicles))
    return {"p50_step_ms": float(np.percentile(t, 50)), "p99_step_ms": float(np.percentile(t, 99)), "mean_step_ms": float(t.mean()), "steps": int(len(t)),
            "vehicles": len(AGENTS), "horizon": 30, "vehicles_in_backup_at_end": fails,
            "note": "closed loop, Jacobi exchange of predictions on the host, one batched k_solve launch per control step (4 NLPs)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="instances per GPU")
    ap.add_argument("--tol", type=float, default=1e-2, help="IPOPT tol / constr_viol_tol of the reference (vehicle.py:651-652)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mpc", action="store_true", help="skip the distributed-MPC latency leg (second half of the metric)")
    ap.add_argument("--mpc-steps", type=int, default=150)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import torch
    import torch.distributed as dist

    from conflict_rez_b200.control.batch_planner import prepare_joint_batch, random_init_offsets
    from conflict_rez_b200.control.strategy import write_strategy
    from conflict_rez_b200.solver import ObcaSolver, SolveOptions

    config = {
        "workload": "4-vehicle centralized OBCA conflict resolution (multi_vehicle_planner.py defaults: K=5, N_per_set=5, shrink_tube=0.5, dmin=0.05), "
        "parking-lot scenario, synthetic strategy, randomized initial offsets; batch of independent instances",
        "batch_per_gpu": args.batch,
        "global_batch": args.batch * world,
        "vehicles": 4,
        "obstacles": 6,
        "tol": args.tol,
        "l2": "per-step working set (iterates of the batch) is larger than the 126 MB L2; no explicit flush",
        "parallelism": "instances sharded over %d GPU(s), no data-path collective" % world,
    }

    if args.impl == "reference":
        # CPU arm: the oracle port on the host cores (rank 0 only); each step is a bounded sample of the workload
        if rank != 0:
            return
        fn = os.path.join(tempfile.mkdtemp(), "4v")
        write_strategy(fn)
        from conflict_rez_b200.control.scenario import build_guess, build_problem

        n_proc = max(1, min(os.cpu_count() or 1, 16))
        offs = random_init_offsets(1, 4)
        prob = build_problem(fn, AGENTS, init_offsets=offs)
        guess = build_guess(prob, fn, AGENTS)
        plan = type("Plan", (), {"problem": prob, "guess": guess})
        base = CpuBaseline(plan, n_proc)
        vals = []
        t_all = time.perf_counter()
        for s in range(args.warmup + args.steps):
            cb = base.sample(52.0, sample_iters=3)
            if s >= args.warmup:
                vals.append(cb["value"])
        base.close()
        value = float(np.mean(vals))
        cb["value"] = value
        line = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (time.perf_counter() - t_all) / (args.warmup + args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "impl": "reference", "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "reference arm = CPU restatement (oracle port: numpy/scipy SuperLU interior point from the kinematic warm start, 52 iterations "
            "assumed per solve); CasADi/IPOPT are not installable in this image",
        }
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the OBCA solver has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner to stdout when the communicator is created: keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize(device)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    fn = os.path.join(tempfile.mkdtemp(), "4v")
    write_strategy(fn)
    opts = SolveOptions(tol=args.tol, constr_viol_tol=args.tol, max_iter=600)
    offs_all = random_init_offsets(args.batch * world, 4, seed=0)
    offs = offs_all[rank * args.batch : (rank + 1) * args.batch]
    t_ws = time.perf_counter()
    plan = prepare_joint_batch(fn, AGENTS, offs, opts, device=device)  # warm start: single-vehicle solves + pair duals, on the device
    t_ws = time.perf_counter() - t_ws
    sv = plan.solver

    # ---------------- device-resident timing (value): inputs already in HBM (the warm start never left the device)
    dev_in = plan.dev_guess
    barrier()

    def step_device():
        sv.set_inputs(dev_in)
        sv.run()
        return sv.fetch_stats()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = sv.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ksolve_ms = []
    ev[0].record()
    for _ in range(args.steps):
        sv.set_inputs(dev_in)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        sv.run()
        k1.record()
        st, it, dbl = sv.fetch_stats()
        ksolve_ms.append((k0, k1))
    ev[1].record()
    barrier()
    sampler.stop_flag.set()
    launches = sv.launch_count - launches0
    t_dev = ev[0].elapsed_time(ev[1]) / 1e3
    t_kernel = float(np.mean([a.elapsed_time(b) for a, b in ksolve_ms])) / 1e3
    st_h, it_h = st.cpu().numpy(), it.cpu().numpy()
    converged = int((st_h >= 0).sum())
    sum_iters = float(it_h.sum())

    # ---------------- end-to-end timing (e2e): host buffers in, host results out
    for _ in range(min(1, args.warmup)):
        sv.solve(plan.guess, want_duals=False)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = sv.solve(plan.guess, want_duals=False)
    e1.record()
    barrier()
    t_e2e = e0.elapsed_time(e1) / 1e3
    g = plan.guess
    h2d = sum(int(np.asarray(a).nbytes) for a in (plan.problem.init_pose, g.z, g.lam, g.mu, g.dt, g.pair_lam, g.pair_mu, g.pair_s))
    d2h = sum(int(a.nbytes) for a in (res.z, res.dt, res.pair_lam, res.pair_mu, res.pair_s, res.status, res.iters, res.obj, res.cviol, res.dual_inf, res.compl_inf))
    # fetch_solution always copies lam/mu to the host as well
    d2h += int(np.prod(g.lam.shape)) * 8 * 2
    conv_e2e = int((res.status >= 0).sum())

    # ---------------- reduce over ranks: time = max, counts = sum
    stats = torch.tensor([t_dev, t_e2e, t_kernel], dtype=torch.float64, device=device)
    counts = torch.tensor([converged, conv_e2e, sum_iters, launches, float(len(st_h))], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    t_dev, t_e2e, t_kernel = [float(v) for v in stats.cpu()]
    converged, conv_e2e, sum_iters, launches, total = [float(v) for v in counts.cpu()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = converged * args.steps / t_dev
    e2e_value = conv_e2e * args.steps / t_e2e
    peak, peak_src = load_peaks()
    import ctypes

    fp64_peak = ctypes.c_double(0.0)
    fp64_src = "data sheet"
    if sv.lib.obca_measure_dfma_peak(local_rank, ctypes.byref(fp64_peak)) == 0 and fp64_peak.value > 0:
        fp64_src = "measured (obca_measure_dfma_peak: 8 independent DFMA chains per thread, 8 CTAs of 256 threads per SM)"
    else:
        fp64_peak.value = FP64_PEAK_TFLOPS
    achieved_gbs = sum_iters * BYTES_PER_ITER / t_kernel / 1e9 / world  # per GPU, dominant kernel k_solve
    line = {
        "metric": METRIC,
        "value": value,
        "unit": "solves/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": config,
        "converged_fraction": converged / max(1.0, total),
        "iters_median": float(np.median(it_h)),
        "iters_max": int(it_h.max()),
        "status_hist_rank0": {str(int(k)): int(v) for k, v in zip(*np.unique(st_h, return_counts=True))},
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": {
            "bound": "hbm",
            "kernel": "k_solve (one CTA per instance: NLP evaluation + structured KKT solve + IPM loop)",
            "achieved": achieved_gbs,
            "peak": peak,
            "unit": "GB/s",
            "frac": achieved_gbs / peak,
            "traffic": DRAM_BYTES_PER_ITER_NCU * sum_iters / world,
            "traffic_note": "21.9 MB of DRAM traffic per IPM iteration and instance (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full "
            "capture of k_solve, profiles/r01c_ncu_full_k_solve_summary_final.txt) x the iterations of one launch; 6.3x the algorithmic bytes: "
            "per-CTA work areas (block solves, QR records, T maps) and local-memory arrays stream through L2/HBM every iteration",
            "peak_source": peak_src,
            "note": "algorithmic bytes = 3.456 MB per IPM iteration and instance (one read + one write of the primal-dual iterate, SURVEY.md 8d) "
            "x iterations of all instances / k_solve time; the kernel is FP64-latency bound, not bandwidth bound (DESIGN.md)",
            "fp64_convention": {
                "achieved_tflops": sum_iters * FLOPS_PER_ITER_CONVENTION / t_kernel / 1e12 / world,
                "peak_tflops": fp64_peak.value,
                "frac": sum_iters * FLOPS_PER_ITER_CONVENTION / t_kernel / 1e12 / world / fp64_peak.value,
                "peak_source": fp64_src,
                "note": "SURVEY.md 8d counts 767 MFLOP/iteration for a dense block elimination; the null-space Riccati solve needs far fewer flops, "
                "so this is an equivalent-work figure, not executed flops",
            },
        },
    }
    line["warm_start"] = {
        "wall_s_rank0": t_ws, **plan.timing, "single_vehicle_fail_rank0": int(sum((r.status < 0).sum() for r in plan.singles)),
        "note": "untimed setup of the step (SURVEY.md 8f rank 1): vectorised pose guess on the host, obstacle and pair duals by obca_dual_ws / "
        "obca_joint_dual_ws on the device, 4 batched single-vehicle solves; the joint warm start stays in HBM",
    }
    if world == 1 and not args.no_mpc:
        line["mpc"] = mpc_latency(fn, device, args.mpc_steps)
    if world == 1 and not args.no_cpu_baseline:
        base = CpuBaseline(plan, 1)
        line["cpu_baseline"] = base.sample(float(np.median(it_h)), sample_iters=16)
        base.close()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
