#!/usr/bin/env python
"""BASELINE config 5: B = 64 independent paired-Rosenbrock problems, n = 1e6 each, fp64, m = 10, More-Thuente,
x0_b ~ U[-1,1] (seed 1000+b), on N B200s.  Two shardings:
  problems : rank r solves problems b = r (mod N), no communication ("replicas only")
  n        : every problem is split over the N ranks along n, reductions all-reduced in-kernel over NVLink
Run:  python bench_batched.py [--B 64] [--n 1000000] [--threads 4]           (1 GPU)
      python -m torch.distributed.run --nproc-per-node N ... bench_batched.py --sharding n|problems
Not the contract bench (that is bench.py / config 2); prints one JSON line per run."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--sharding", default="problems", choices=["problems", "n"])
    args = ap.parse_args()
    import torch
    import lbfgspp_b200 as lb
    from lbfgspp_b200.sharding import shard_bounds
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prm = lb.LBFGSParam(m=10)
    if args.sharding == "n" and world > 1:
        lo, hi = shard_bounds(args.n, rank, world)
        mine = torch.tensor(list(lb.p2p_export(local)), dtype=torch.uint8, device="cuda")
        allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        lb.p2p_attach(local, b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh), rank, world, index_offset=lo)
        X0 = np.stack([np.random.default_rng(1000 + b).uniform(-1, 1, args.n)[lo:hi] for b in range(args.B)])
        sharded, threads = True, 1
    else:
        X0 = np.stack([np.random.default_rng(1000 + b).uniform(-1, 1, args.n) for b in range(rank, args.B, world)])
        sharded, threads = False, args.threads
    lb.solve_batch(lb.OBJ_ROSENBROCK_PAIRED, X0[:2], prm, "MoreThuente", device=local, threads=min(2, threads), sharded=sharded, return_x=False)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res, _, _ = lb.solve_batch(lb.OBJ_ROSENBROCK_PAIRED, X0, prm, "MoreThuente", device=local, threads=threads, sharded=sharded, return_x=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    iters = sum(r["niter"] for r in res)
    ok = sum(r["status"] == "ok" for r in res)
    if dist is not None:
        t = torch.tensor([dt, 0 if sharded else iters, 0 if sharded else ok], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        dt = float(tmax[0])
        if not sharded:
            iters, ok = int(t[1]), int(t[2])
    if rank == 0:
        its = [r["niter"] for r in res]
        print(json.dumps({"config": "C5: B=%d x paired Rosenbrock n=%d fp64 m=10 MoreThuente, x0~U[-1,1] seeds 1000+b" % (args.B, args.n),
                          "n_gpus": world, "sharding": args.sharding, "host_threads_per_gpu": threads, "seconds": dt,
                          "problems_per_s": args.B / dt, "iterations_per_s": iters / dt, "converged": ok,
                          "iterations_min_max_rank0": [min(its), max(its)]}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
