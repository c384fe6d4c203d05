#!/usr/bin/env python
"""bench.py -- Sim3 Levenberg-Marquardt throughput on synthetic pose graphs (BASELINE.json metric).

A "step" is one LM iteration (one OptimizationAlgorithmLevenberg::solve call: 1 linearisation plus
>= 1 damped trial, each trial = preconditioner + PCG solve + retraction + chi2).  The LM iterations
are drawn from repeated solves of the same synthetic sphere graph: a solve runs from the initial
guess until g2o's relative-gain rule (1e-6) or 30 iterations, then the estimates are restored from a
device-side snapshot and the next solve starts.  W warm-up iterations, then exactly K timed ones.

  value   LM iterations/s with the graph resident in HBM (device-timed, max over ranks)
  e2e     the same, driving the LM one iteration at a time through the C ABI with HOST buffers:
          every step uploads the current estimates from pinned host memory (s3o_set_estimates),
          runs one LM iteration (s3o_optimize) and reads the estimates back (s3o_get_vertices)
  roofline  symmetric BSR SpMV (the dominant kernel): algorithmic bytes per launch / sampled
          CUDA-event duration of that kernel inside the timed region, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (g2o-faithful LM: numeric Jacobians h=1e-9, sparse LDLT, 1 thread)
          on a bounded sample of the same generator

--impl reference times the CPU oracle alone (the reference itself cannot be built here:
g2o/Eigen/Sophus/TooN are absent and there is no network).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (laps, poses_per_lap)
    "s1m": (1000, 1000),     # BASELINE configs[3]: 1M poses / 5M edges
    "s100k": (100, 1000),    # BASELINE configs[2]
    "s10k": (10, 1000),
}
CPU_SAMPLE = (10, 1000)      # 10k poses / 50k edges of the same generator
STOP_REL_GAIN = 1e-6
MAX_LM_ITERS = 30


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="s1m", choices=sorted(WORKLOADS))
    ap.add_argument("--pcg-tol", type=float, default=1e-3)
    ap.add_argument("--pcg-max-iter", type=int, default=2000)
    ap.add_argument("--precond", default="auto", choices=["auto", "block-jacobi", "multilevel"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=42)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md 'clocks line')."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(steps, warmup, seed):
    """LM iterations/s of the CPU oracle on the bounded sample (1 thread, g2o-faithful settings)."""
    from oracle import oracle as orc
    from sim3opt_b200 import synth
    laps, per = CPU_SAMPLE
    g = synth.sphere(laps, per, seed=seed)
    orc.set_math_mode(orc.MATH_CORRECTED)
    p = orc.Problem(orc.KIND_SIM3)
    p.set_vertices(g["est"], g["fixed"])
    p.set_edges(g["v0"], g["v1"], g["meas"], g["info"])
    p.set_jacobian_mode(orc.JAC_NUMERIC, 1e-9)
    p.build_structure()
    done = 0
    t_timed = 0.0
    chi2 = float("nan")
    # the oracle has no resume: run one optimize() of warmup+steps iterations and time the tail by
    # differencing two runs would double the cost, so time the whole call and subtract nothing:
    # warm-up here only pages the library in (one chi2 evaluation).
    p.chi2()
    t0 = time.perf_counter()
    n, chi2, lam, hist = p.optimize(steps, 0.0)
    t_timed = time.perf_counter() - t0
    done = n
    sample = f"sphere {laps}x{per} = {laps * per} poses / {len(g['v0'])} edges, {done} LM iterations from the initial guess"
    return done / t_timed, t_timed, done, chi2, sample, len(g["v0"])


def workload_edges(name):
    laps, per = WORKLOADS[name]
    n = laps * per
    return sum(n - o for o in (1, 2, per, per + 1, 2 * per) if o < n)


def cpu_baseline_dict(workload, steps, seed):
    """The oracle timed on the bounded sample, expressed in the metric's unit ON THE BENCH WORKLOAD:
    rate_on_sample * (edges_sample / edges_workload).  Linear-in-edges extrapolation favours the CPU
    (its sparse LDLT grows faster than linearly with the graph)."""
    rate, t, done, chi2, sample, e_sample = cpu_oracle_rate(steps, 0, seed)
    e_full = workload_edges(workload)
    scaled = rate * e_sample / e_full
    return {"value": scaled, "unit": "LM iterations/s", "cores": 1, "kind": "port",
            "sample": sample + f"; measured {rate:.4f} LM iterations/s on the sample, scaled by edges "
                               f"{e_sample}/{e_full} to the {workload} workload",
            "rate_on_sample": rate, "seconds": t, "lm_iterations": done, "final_chi2_sample": chi2}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = cpu_baseline_dict(args.workload, max(args.steps, 1), args.seed)
    rate, done = cpu["value"], cpu["lm_iterations"]
    laps, per = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "Sim3 LM iterations/s", "value": rate, "unit": "LM iterations/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 / rate,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: synthetic Sim3 sphere pose graph, {laps * per} poses / "
                               f"{workload_edges(args.workload)} edges, seed {args.seed}",
                   "sample": cpu["sample"], "jacobians": "numeric h=1e-9 (g2o linearizeOplus)",
                   "linear_solver": "sparse LDLT (up-looking, min-degree), as LinearSolverEigen", "math_mode": "corrected"},
        "cpu_baseline": cpu,
        "e2e": {"value": rate, "unit": "LM iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference (g2o @8564e1e + vio_g2o + Eigen + Sophus + TooN) cannot be built offline; this times "
                "the oracle port, single-threaded like the reference",
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import sim3opt_b200 as s3
    from sim3opt_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner, torchrun notices) must not pollute the one JSON line on stdout
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")

    laps, per = WORKLOADS[args.workload]
    g = synth.sphere(laps, per, seed=args.seed)
    nv, ne = len(g["est"]), len(g["v0"])

    stream = torch.cuda.Stream()
    prob = s3.Problem(s3.KIND_SIM3, device=local_rank, stream=stream.cuda_stream)
    prob.set_math_mode(s3.MATH_CORRECTED)
    prob.set_jacobian_mode(s3.JAC_ANALYTIC)
    prob.set_pcg(args.pcg_tol, args.pcg_max_iter)
    precond = {"auto": s3.PRECOND_AUTO, "block-jacobi": s3.PRECOND_BLOCK_JACOBI, "multilevel": s3.PRECOND_MULTILEVEL}[args.precond]
    prob.set_preconditioner(precond)
    multilevel = precond == s3.PRECOND_MULTILEVEL or (precond == s3.PRECOND_AUTO and nv >= 20000)
    if world > 1:
        # vertex-range partition: rank 0 creates the NCCL id, every rank joins before set_edges
        box = [s3.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        prob.set_comm(rank, world, box[0])
    prob.set_vertices(g["est"], g["fixed"])
    prob.set_edges(g["v0"], g["v1"], g["meas"], g["info"])
    prob.build_structure()
    # global sizes (the per-rank structure holds owned + ghost rows only)
    nf = int((g["fixed"] == 0).sum())
    if world == 1:
        nb = prob.num_blocks
    else:
        from sim3opt_b200 import api as _api
        nb = len(_api.host_structure(nv, g["fixed"], g["v0"], g["v1"])[1])
    prob.snapshot_estimates()
    prob.set_lm_resume(True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Driver:
        """Feeds LM iterations one at a time; restarts from the snapshot when a solve converges."""
        def __init__(self):
            self.in_solve = 0
            self.last_chi = None
            self.solves = []
            self.cur = []
            self.trace = []          # per step: [chi2, lambda, trials, rho, pcg iterations]

        def step(self):
            if self.in_solve == 0:
                prob.restore_estimates()
            n, chi2, lam, hist = prob.optimize(1, 0.0)
            self.in_solve += 1
            self.cur.append(chi2)
            self.trace.append([float(v) for v in np.asarray(hist).reshape(-1)[:5]])
            conv = False
            if self.last_chi is not None and chi2 > 0:
                gain = (self.last_chi - chi2) / chi2
                conv = 0 <= gain < STOP_REL_GAIN
            self.last_chi = chi2
            if conv or self.in_solve >= MAX_LM_ITERS:
                self.solves.append(list(self.cur))
                self.cur, self.in_solve, self.last_chi = [], 0, None
            return chi2

    drv = Driver()
    for _ in range(args.warmup):
        drv.step()
    drv.in_solve, drv.last_chi, drv.cur, drv.trace = 0, None, [], []   # timed region starts a fresh solve

    prob.reset_stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        step_events = []
        for _ in range(args.steps):
            drv.step()
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            step_events.append(e)
        ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    step_ms = [a.elapsed_time(b) for a, b in zip([ev0] + step_events[:-1], step_events)]
    # device time of every completed solve = the sum of its steps (a solve restarts from the snapshot when the
    # previous one has met the relative-gain rule)
    solve_ms, k0 = [], 0
    for sv in drv.solves:
        solve_ms.append(sum(step_ms[k0:k0 + len(sv)]))
        k0 += len(sv)
    st = prob.stats()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = args.steps / (ms * 1e-3)

    # ---- quality of the converged estimate: key-frame centres against the generator's ground truth after a
    # similarity alignment (the reference's evaluation, kitti_surf.cpp:1381-1452), outside every timed region
    quality = None
    if rank == 0 and drv.in_solve == 0 and drv.solves:
        def centres(est):
            q, t, sc = est[:, :4], est[:, 4:7], est[:, 7:8]
            x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
            R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                          2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                          2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1).reshape(-1, 3, 3)
            return -np.einsum("nji,nj->ni", R, t) / sc
        gt_c = centres(g["gt"])
        _, rmse0, _ = s3.align_similarity(centres(g["est"]), gt_c, device=local_rank)
        _, rmse1, max1 = s3.align_similarity(centres(prob.vertices()), gt_c, device=local_rank)
        quality = {"rmse_to_ground_truth_m": {"initial_guess": rmse0, "converged": rmse1, "max_converged": max1},
                   "how": "camera centres, Umeyama-aligned to the generator's ground truth (s3o_align_similarity)"}

    # ---- end-to-end: host estimates in, host estimates out, every step ------------------------
    # Same LM-iteration sequence as the timed region above (solves restart from the initial guess on
    # the 1e-6 gain rule), but the estimates live in pinned HOST memory between steps.
    est_host = torch.empty((nv, 8), dtype=torch.float64).pin_memory()
    est0_host = torch.empty((nv, 8), dtype=torch.float64).pin_memory()
    est_np, est0_np = est_host.numpy(), est0_host.numpy()
    prob.restore_estimates()
    prob.vertices(out=est0_np)
    e2e_steps = args.steps
    prob.set_lm_resume(2)                   # keep lambda/nu across the host round trip of the estimates
    in_solve, last_chi = 0, None
    e2e_trace = []
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        if in_solve == 0:
            prob.restore_estimates()        # drops the LM state: lambda is re-initialised
            prob.set_estimates(est0_np)     # H2D from pinned host memory
        else:
            prob.set_estimates(est_np)
        n_, chi2_, lam_, _h = prob.optimize(1, 0.0)
        prob.vertices(out=est_np)           # D2H of the step's result
        e2e_trace.append([float(v) for v in np.asarray(_h).reshape(-1)[:5]] + [time.perf_counter() - t0])
        in_solve += 1
        conv = last_chi is not None and chi2_ > 0 and 0 <= (last_chi - chi2_) / chi2_ < STOP_REL_GAIN
        last_chi = chi2_
        if conv or in_solve >= MAX_LM_ITERS:
            in_solve, last_chi = 0, None
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = e2e_steps / e2e_s

    # ---- roofline of the dominant kernel (symmetric BSR SpMV) ---------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    n_off = nb - nf
    # B_spmv = 392 N_b + 4 N_b + 4 (N_f + 1) + 2*56 N_f   (each unique block once)
    # (partitioned solve: the launch on one rank covers that rank's rows and blocks only)
    nb_l, nf_l = (nb, nf) if world == 1 else (st["n_blocks"], -(-nf // world))
    bytes_spmv = 392 * nb_l + 4 * nb_l + 4 * (nf_l + 1) + 2 * 56 * nf_l
    roof = {"bound": "hbm", "kernel": "spmv4_kernel<7,128,112,2> (TMA ring, prefetch pipeline)", "achieved": None, "peak": peak, "unit": "GB/s",
            "frac": None, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_spmv}
    if st["n_spmv_sampled"] > 0:
        avg_ms = st["ms_spmv_sampled"] / st["n_spmv_sampled"]
        roof["achieved"] = bytes_spmv / (avg_ms * 1e-3) / 1e9
        roof["frac"] = roof["achieved"] / peak
        roof["avg_launch_ms"] = avg_ms
        roof["launches_sampled"] = st["n_spmv_sampled"]
    traffic_file = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file))
            if tr.get("workload") == args.workload:
                roof["traffic"] = tr.get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline_dict(args.workload, 3, args.seed)

    line = {
        "metric": "Sim3 LM iterations/s", "value": value, "unit": "LM iterations/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: synthetic Sim3 sphere pose graph, {nv} poses / {ne} edges, seed {args.seed}",
                   "free_vertices": nf, "hessian_blocks": nb, "block_dim": 7, "jacobians": "analytic",
                   "linear_solver": ("multilevel (aggregation + block-Jacobi)" if multilevel else "block-Jacobi")
                                    + f" PCG rel_tol={args.pcg_tol} max_iter={args.pcg_max_iter}",
                   "math_mode": "corrected", "l2_policy": "inputs larger than L2 (Hessian blocks %.2f GB)" % (392 * nb / 1e9),
                   "step": "one LM iteration; solves restart from a device snapshot on the 1e-6 gain rule",
                   "partition": "none" if world == 1 else (
                       f"vertex range over {world} ranks; halo: "
                       + ("NVLink peer-to-peer loads inside the SpMV (CUDA IPC)" if st["p2p_halo"] else "NCCL send/recv")
                       + "; NCCL all-reduce / all-gather for the scalars and the level-1 residual; coarse levels of the "
                         "multilevel PCG replicated"),
                   "multilevel_levels": int(st["multilevel_levels"])},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "LM iterations/s", "h2d_bytes_per_step": nv * 64, "d2h_bytes_per_step": nv * 64 + 160,
                "steps": e2e_steps},
        "gpu_launches": int(st["kernel_launches"]),
        "roofline": roof,
        "cpu_baseline": cpu,
        "pcg_iterations": int(st["pcg_iterations"]), "lm_trials": int(st["lm_trials"]),
        "phase_ms": {"linearize": st["ms_linearize"], "solve": st["ms_solve"], "update_chi2": st["ms_update"]},
        "solves_completed": len(drv.solves),
        # time-to-converge (BASELINE metric, second half): device time (CUDA events per step) of one solve from the
        # initial guess to g2o's relative-gain stop (1e-6), averaged over the solves completed inside the timed region
        "time_to_converge_s": (1e-3 * sum(solve_ms) / len(solve_ms)) if solve_ms else None,
        "step_ms": step_ms,
        "lm_iterations_to_converge": (sum(len(sv) for sv in drv.solves) / len(drv.solves)) if drv.solves else None,
        "final_chi2": drv.solves[0][-1] if drv.solves else None,
        "quality": quality,
        "step_trace": drv.trace, "e2e_step_trace": e2e_trace,
        "chi2_history_first_solve": drv.solves[0] if drv.solves else drv.cur,
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
