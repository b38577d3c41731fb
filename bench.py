#!/usr/bin/env python
"""bench.py -- L-BFGS iterations/s (and apply_Hv HBM GB/s) on BASELINE config 2:
paired Rosenbrock, n = 1e7, fp64, m = 10, More-Thuente line search, x0 = 0, on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched under torchrun, one rank per GPU; n is sharded over the ranks, "strong" scaling)

A "step" is one complete LBFGSSolver::minimize() of the problem (22 iterations / 50 objective evaluations) through
the header-only C++ front on top of liblbfgs_b200.so.  `value` = iterations per second with x0 already resident in HBM;
`e2e` = the same with the start point coming from pinned host memory and the solution copied back every step.
The roofline object is for apply_Hv (algorithmic bytes 8*n*(4c+2) per call), timed live with CUDA events
around every call of the timed region.  `cpu_baseline` / `--impl reference` time the CPU restatement of the reference
(oracle/; the reference itself cannot be built: Eigen is absent) on this box's host cores.
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

N_GLOBAL = 10_000_000
M_HIST = 10
WORKLOAD = "C2: paired Rosenbrock n=1e7 fp64, m=10, LineSearchMoreThuente, x0=0 (BASELINE.json configs[1])"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (oracle/ is executed only here, in tests/ and in smoke())
# ---------------------------------------------------------------------------------------------------------------------
def load_oracle(native=True):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    try:
        po.build(native=native)   # native build happens on the box it is timed on
        return po, po.Oracle("orc", native=native)
    except Exception:
        return po, po.Oracle("orc", native=False)


def cpu_solve(po, orc, threads_mode, max_iterations=0):
    import numpy as np
    prm = orc.default_param(m=M_HIST, max_iterations=max_iterations)
    t0 = time.perf_counter()
    r = orc.lbfgs(po.OBJ_ROSENBROCK_PAIRED, np.zeros(N_GLOBAL), po.LS_MORE_THUENTE, prm, sum_mode=threads_mode, trace_cap=1024)
    wall = time.perf_counter() - t0
    return r, wall


def run_reference_arm(args, rank):
    """The reference's CPU implementation of the path on the host cores: the OpenMP build of the restatement (all cores).
    The reference's Eigen code itself is single-threaded (see cpu_baseline in the main arm for the 1-core figure)."""
    if rank != 0:
        return
    po, orc = load_oracle(native=True)
    # all host threads (OpenMP build) unless one thread is faster on this box (memory-bound level-1 code on few cores):
    # calibrate on a 3-iteration solve and keep the faster of the two, so that the baseline is the strongest CPU run we have
    rate = {}
    cpu_solve(po, orc, po.SUM_LANES8, 1)          # untimed: first-touch of the allocator, library load
    for mode in (po.SUM_LANES8_OMP, po.SUM_LANES8):
        rc, _ = cpu_solve(po, orc, mode, 4)
        rate[mode] = rc["niter"] / rc["seconds"]
    mode = max(rate, key=rate.get)
    cores = orc.hw_threads() if mode == po.SUM_LANES8_OMP else 1
    # bound the sample so that warmup+steps finish within a few minutes
    r, wall = cpu_solve(po, orc, mode)
    max_it = 0
    budget = 150.0
    total = args.steps + args.warmup
    if wall * total > budget:
        max_it = max(2, int(r["niter"] * budget / (wall * total)))
    for _ in range(max(0, args.warmup - 1)):
        cpu_solve(po, orc, mode, max_it)
    secs, iters = 0.0, 0
    for _ in range(args.steps):
        r, _ = cpu_solve(po, orc, mode, max_it)
        secs += r["seconds"]
        iters += r["niter"]
    value = iters / secs
    sample = ("%d x minimize() on the full n=1e7 problem" % args.steps) + \
             ("" if max_it == 0 else " truncated at max_iterations=%d (history only partly filled)" % max_it) + \
             "; %d thread(s) (calibrated: %.2f it/s with all %d threads, %.2f it/s with one)" % (
                 cores, rate[po.SUM_LANES8_OMP], orc.hw_threads(), rate[po.SUM_LANES8]) + \
             "; restatement of the reference (oracle/liboracle.so, -O3 -march=native); the unmodified reference headers over the" \
             " minieigen stand-in (oracle/_ref) are the parity checker and ~7x slower, so they are not used as the baseline"
    # for the record: the unmodified reference headers themselves (over the minieigen stand-in), 3 iterations of the same solve
    ref_headers = None
    try:
        import numpy as np
        ref = po.Oracle("ref")
        rr = ref.lbfgs(po.OBJ_ROSENBROCK_PAIRED, np.zeros(N_GLOBAL), po.LS_MORE_THUENTE, ref.default_param(m=M_HIST, max_iterations=3),
                       trace_cap=64)
        ref_headers = {"value": rr["niter"] / rr["seconds"], "unit": "iters/s", "cores": 1,
                       "sample": "3 iterations of the same solve by oracle/_ref (reference headers compiled over oracle/minieigen, -O2)"}
    except Exception:  # noqa: BLE001  (no _ref build on this machine)
        pass
    line = {"impl": "reference", "metric": "lbfgs_iterations_per_sec", "value": value, "unit": "iters/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n": N_GLOBAL, "m": M_HIST, "line_search": "MoreThuente"},
            "cpu_baseline": {"value": value, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample,
                             "reference_headers": ref_headers},
            "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hv", default="auto", choices=["auto", "two_loop", "gram"])
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: in-kernel all-reduce over NVLink peer memory (default) or one ncclAllReduce per reduction")
    ap.add_argument("--solver-loop", default="auto", choices=["auto", "resident", "host"],
                    help="device-resident CUDA graph or host-driven loop (auto = host-driven: at n = 1e7 the two are within noise, see DESIGN.md section 10)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-mode", action="store_true", help="timed region only (for runs under ncu; numbers are not bench values)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import lbfgspp_b200 as lb

    assert torch.cuda.is_available(), "bench.py needs a B200 (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- n-sharding: rank r owns an even-length contiguous block; scalars replicated; every dot is all-reduced ----
    from lbfgspp_b200.sharding import shard_bounds
    lo, hi = shard_bounds(N_GLOBAL, rank, world)
    n_local = hi - lo
    if world > 1 and args.comm == "nccl":
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident = torch.tensor(list(lb.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(ident, src=0)
        lb.comm_init(local_rank, bytes(ident.cpu().numpy().tobytes()), rank, world, index_offset=lo)
    elif world > 1:
        mine = torch.tensor(list(lb.p2p_export(local_rank)), dtype=torch.uint8, device="cuda")
        allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        lb.p2p_attach(local_rank, b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh), rank, world, index_offset=lo)

    hv = {"auto": lb.HV_AUTO, "two_loop": lb.HV_TWO_LOOP, "gram": lb.HV_GRAM}[args.hv]
    prm = lb.LBFGSParam(m=M_HIST)
    resident = (args.solver_loop == "resident") \
        and (world == 1 or args.comm == "p2p") and args.hv != "two_loop"
    sess = lb.Session(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(n_local), prm, "MoreThuente", device=local_rank, hv_algo=hv, resident=resident)
    # per-phase CUDA events exist only on the host-driven path: a second session supplies the phase / roofline numbers
    prof_sess = sess if not resident else lb.Session(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(n_local), prm, "MoreThuente",
                                                     device=local_rank, hv_algo=hv, resident=False)
    ctx = lb.driver_ctx(local_rank)
    abi = lb.abi()

    # ---- warm-up --------------------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        r = sess.solve()
    niter, nfev = r["niter"], r["nfev"]

    # ---- timed region 1: operands resident in HBM ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    import ctypes as C
    barrier()
    abi.lbfgs_b200_timer_start(ctx)
    launches = 0
    iters = 0
    for _ in range(args.steps):
        r = sess.solve()
        launches += r["launches"]
        iters += r["niter"]
    ms = C.c_float(0)
    abi.lbfgs_b200_timer_stop(ctx, C.byref(ms))
    barrier()
    dev_seconds = max_over_ranks(ms.value * 1e-3)
    # ---- phase breakdown / roofline: the same kernels driven from the host with a CUDA-event pair around every call ----
    prof_steps = max(2, min(args.steps, 5)) if not args.profile_mode else 1
    if resident:
        for _ in range(2):
            prof_sess.solve()
    abi.lbfgs_b200_profile_enable(ctx, 1)
    for ph in range(3):
        abi.lbfgs_b200_profile_read(ctx, ph, None, None, 1)
        abi.lbfgs_b200_profile_bytes(ctx, ph, C.byref(C.c_double()), 1)
    barrier()
    for _ in range(prof_steps):
        prof_sess.solve()
    barrier()
    phases = {}
    for ph, name in enumerate(("apply_Hv", "trial", "update")):
        tms, calls, nbytes = C.c_double(0), C.c_uint64(0), C.c_double(0)
        abi.lbfgs_b200_profile_read(ctx, ph, C.byref(tms), C.byref(calls), 1)
        abi.lbfgs_b200_profile_bytes(ctx, ph, C.byref(nbytes), 1)
        phases[name] = {"ms": tms.value, "calls": int(calls.value), "alg_bytes": nbytes.value}
    abi.lbfgs_b200_profile_enable(ctx, 0)

    # ---- timed region 2: end to end (pinned host -> device every step, result back to the host) --------------------
    for _ in range(0 if args.profile_mode else 2):
        sess.solve(from_host=True, to_host=True)
    barrier()
    t0 = time.perf_counter()
    e2e_iters, h2d, d2h = 0, 0, 0
    for _ in range(1 if args.profile_mode else args.steps):
        r2 = sess.solve(from_host=True, to_host=True)
        e2e_iters += r2["niter"]
        h2d, d2h = r2["h2d_bytes"], r2["d2h_bytes"] + 8  # + the fx scalar
    barrier()
    e2e_seconds = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.summary()

    # ---- apply_Hv with a full history (c = m), the steady-state figure --------------------------------------------
    rng = np.random.default_rng(0)
    mctx = lb.Context(local_rank)
    hist = lb.History(mctx, n_local, M_HIST)
    blk = rng.standard_normal(1 << 20)
    def noise(seed):
        return np.resize(np.roll(blk, seed * 7919), n_local)
    if world == 1 and not args.profile_mode:  # the microbenchmark uses a private context without a communicator: N = 1 only
        for k in range(M_HIST):
            s = noise(k)
            hist.add(mctx.array(s), mctx.array(s + 0.1 * noise(100 + k)))
        v, res = mctx.array(noise(999)), mctx.empty(n_local)
        for _ in range(3):
            hist.apply_Hv(v, -1.0, res, hv)
        reps = 30
        mctx.timer_start()
        for _ in range(reps):
            hist.apply_Hv(v, -1.0, res, hv)
        hv_ms = mctx.timer_stop() / reps
        hv_full = {"ms_per_call": hv_ms, "gb_per_s": 8.0 * n_local * (4 * M_HIST + 2) / hv_ms / 1e6, "c": M_HIST}
    else:
        hv_full = None

    peak, peak_src = load_peaks()
    hv_phase = phases["apply_Hv"]
    achieved = hv_phase["alg_bytes"] / (hv_phase["ms"] * 1e-3) / 1e9 if hv_phase["ms"] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("apply_Hv_dram_bytes_per_call_c10_n1e7" if args.hv == "two_loop" else
                                        "update_apply_Hv_dram_bytes_per_call_c10_n1e7")

    value = iters / dev_seconds
    line = {
        "metric": "lbfgs_iterations_per_sec", "value": value, "unit": "iters/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dev_seconds / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n": N_GLOBAL, "n_per_gpu": n_local, "m": M_HIST, "line_search": "MoreThuente",
                   "step": "one full minimize(): %d iterations, %d objective evaluations" % (niter, nfev),
                   "apply_Hv": args.hv, "sharding": "single GPU, no collective" if world == 1 else
                   "n split over %d ranks; reductions all-reduced %s" % (
                       world, "in-kernel over NVLink peer memory" if args.comm == "p2p" else "with ncclAllReduce"),
                   "l2": "inputs larger than L2 (S,Y = %.2f GB per GPU)" % (2 * 8 * n_local * (M_HIST + 1) / 1e9)},
        "clocks": clocks,
        "e2e": {"value": e2e_iters / e2e_seconds, "unit": "iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "pair update + apply_Hv, fused (k_pair_dots + k_gram_combine)" if args.hv != "two_loop" else "apply_Hv (k_hv_stage x 2c+1)",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes": ("8*n*(4c+2) per apply_Hv call" if args.hv == "two_loop" else
                                           "8*n*((4c+2) + 6) per fused call = SURVEY.md 8d's apply_Hv unit + its update unit (the update kernel no longer exists)")
                                          + ", c = pairs in the history at that call",
                     "calls": hv_phase["calls"], "full_history": hv_full},
        "phase_ms_per_step": {k: v["ms"] / prof_steps for k, v in phases.items()},
        "solver_loop": "device-resident (one CUDA graph launch per minimize; conditional WHILE/IF nodes)" if resident else "host-driven",
        "phase_source": "CUDA-event pairs around every call of a host-driven pass of the same kernels (%d solves)" % prof_steps,
        "phase_gb_per_s": {k: (v["alg_bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None) for k, v in phases.items()},
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile_mode:
        po, orc = load_oracle(native=True)
        r_cpu, wall = cpu_solve(po, orc, po.SUM_LANES8)
        line["cpu_baseline"] = {"value": r_cpu["niter"] / r_cpu["seconds"], "unit": "iters/s", "cores": 1, "kind": "port",
                                "sample": "1 x minimize() of the same n=1e7 problem (%d iterations, %d evaluations, %.1f s), "
                                          "single thread like the reference's Eigen level-1 code, 8-lane partial sums"
                                          % (r_cpu["niter"], r_cpu["nfev"], r_cpu["seconds"]),
                                "niter_matches_gpu": bool(r_cpu["niter"] == niter and r_cpu["nfev"] == nfev),
                                "fx_abs_diff": abs(r_cpu["fx"] - r["fx"])}
    if rank == 0:
        print(json.dumps(line), flush=True)
    sess.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
