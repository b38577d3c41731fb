"""Multi-GPU parity script (not collected by pytest; tests/test_gpu_multi.py launches it when the box has >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        tests/multi_gpu_check.py p2p|nccl

Every rank owns a contiguous block of every vector (lbfgspp_b200.sharding.shard_bounds).  Checks, against the CPU checker
run on the FULL vectors:
  1. the fused trial / objective kernels of the neighbour-coupled objectives (chained Rosenbrock, tridiagonal quadratic):
     the block of the gradient and the all-reduced {f, g.d, g.g, x.x} -- this is the halo exchange;
  2. whole solves (paired Rosenbrock, chained Rosenbrock, tridiagonal quadratic), host-driven loop and (p2p) the device-resident
     persistent kernel incl. its in-kernel halo exchange: iterations, evaluations, fx, x block;
  3. (p2p) a batch of n-sharded problems in one persistent kernel launch.
Prints one line 'MULTI_GPU_CHECK PASS ...' on rank 0 when every rank passed."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), HERE]
import lbfgspp_b200 as lb  # noqa: E402
import pyoracle as po  # noqa: E402
from lbfgspp_b200.sharding import shard_bounds  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    if mode == "nccl":
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident = torch.tensor(list(lb.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(ident, src=0)
        lb.comm_init(local, bytes(ident.cpu().numpy().tobytes()), rank, world)
    else:
        mine = torch.tensor(list(lb.p2p_export(local)), dtype=torch.uint8, device="cuda")
        allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        lb.p2p_attach(local, b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh), rank, world)
    orc = po.Oracle("orc")
    ctx = lb.Context.borrow(lb.driver_ctx(local))
    failures = []

    def expect(cond, what):
        if not cond:
            failures.append(what)

    # ---- 1. kernels ------------------------------------------------------------------------------------------------
    for kind in (lb.OBJ_ROSENBROCK_CHAINED, lb.OBJ_QUAD_TRIDIAG):
        for n in (8 * world, 4096 * world + 6):
            lo, hi = shard_bounds(n, rank, world)
            lb.set_global_extent(local, lo, n)
            rng = np.random.default_rng(17 * kind + n)           # same stream on every rank
            xp, drt = rng.uniform(-1.5, 1.5, n), rng.standard_normal(n)
            d0 = d1 = None
            if kind == lb.OBJ_QUAD_TRIDIAG:
                d0, d1, _ = po.quad_tridiag_data(n, seed=n)
            step = 0.173
            D0 = ctx.array(d0[lo:hi]) if d0 is not None else None
            D1 = ctx.array(d1[lo:hi]) if d1 is not None else None
            dxp, dd, dx, dg = ctx.array(xp[lo:hi]), ctx.array(drt[lo:hi]), ctx.empty(hi - lo), ctx.empty(hi - lo)
            f, gd, gg, xx = ctx.trial(kind, dxp, dd, step, dx, dg, D0, D1)
            x_full = xp + step * drt
            f_ref, g_ref = orc.objective(kind, x_full, d0, d1)
            scale = np.max(np.abs(g_ref)) + 1.0
            tag = "kind %d n %d rank %d" % (kind, n, rank)
            expect(np.max(np.abs(dx.get() - x_full[lo:hi])) <= 4e-16 * 4.0, "trial x block: " + tag)
            expect(np.max(np.abs(dg.get() - g_ref[lo:hi])) <= 1e-12 * scale, "trial gradient block (halo): " + tag)
            expect(abs(f - f_ref) <= 1e-11 * max(1.0, abs(f_ref)), "trial f: %s %r %r" % (tag, f, f_ref))
            expect(abs(gd - np.dot(g_ref, drt)) <= 1e-10 * (np.sum(np.abs(g_ref * drt)) + 1.0), "trial g.d: " + tag)
            expect(abs(gg - np.dot(g_ref, g_ref)) <= 1e-10 * np.dot(g_ref, g_ref), "trial g.g: " + tag)
            expect(abs(xx - np.dot(x_full, x_full)) <= 1e-10 * np.dot(x_full, x_full), "trial x.x: " + tag)
            dxx = ctx.array(x_full[lo:hi])
            f2, _, gg2, _ = ctx.objective(kind, dxx, dg, D0, D1)
            expect(np.max(np.abs(dg.get() - g_ref[lo:hi])) <= 1e-12 * scale, "objective gradient block (halo): " + tag)
            expect(abs(f2 - f_ref) <= 1e-11 * max(1.0, abs(f_ref)), "objective f: " + tag)

    # ---- 2. solves -------------------------------------------------------------------------------------------------
    def solve(kind, n, x0, prm, ls, d0=None, d1=None, resident=False):
        lo, hi = shard_bounds(n, rank, world)
        lb.set_global_extent(local, lo, n)
        sess = lb.Session(kind, x0[lo:hi], prm, ls, device=local, resident=resident,
                          data0=None if d0 is None else d0[lo:hi], data1=None if d1 is None else d1[lo:hi])
        r = sess.solve(to_host=True)
        r["x"] = sess.result()
        sess.close()
        return r, lo, hi

    def cpu(kind, x0, prm, ls, d0=None, d1=None):
        p = orc.default_param(m=prm.m, max_iterations=prm.max_iterations)
        return orc.lbfgs(kind, x0, po.__dict__["LS_" + ls], p, data0=d0, data1=d1, sum_mode=po.SUM_LANES8)

    for resident in ((False, True) if mode == "p2p" else (False,)):
        tag = "[device-resident] " if resident else "[host-driven] "
        n = 50000 * world
        prm = lb.LBFGSParam(m=10)
        g, lo, hi = solve(lb.OBJ_ROSENBROCK_PAIRED, n, np.zeros(n), prm, "MoreThuente", resident=resident)
        c = cpu(po.OBJ_ROSENBROCK_PAIRED, np.zeros(n), prm, "MORE_THUENTE")
        expect((g["niter"], g["nfev"]) == (c["niter"], c["nfev"]), tag + "paired solve counts %r vs %r" % ((g["niter"], g["nfev"]), (c["niter"], c["nfev"])))
        expect(abs(g["fx"] - c["fx"]) <= 1e-10 * max(1.0, abs(c["fx"])), tag + "paired solve fx")
        expect(np.max(np.abs(g["x"] - c["x"][lo:hi])) <= 1e-7, tag + "paired solve x block")

        n = 10000 * world + 6
        x0 = np.full(n, 3.0)
        g, lo, hi = solve(lb.OBJ_ROSENBROCK_CHAINED, n, x0, prm, "MoreThuente", resident=resident)
        c = cpu(po.OBJ_ROSENBROCK_CHAINED, x0, prm, "MORE_THUENTE")
        expect((g["niter"], g["nfev"]) == (c["niter"], c["nfev"]), tag + "chained solve counts %r vs %r" % ((g["niter"], g["nfev"]), (c["niter"], c["nfev"])))
        expect(abs(g["fx"] - c["fx"]) <= 1e-10 * max(1.0, abs(c["fx"])), tag + "chained solve fx %r %r" % (g["fx"], c["fx"]))
        expect(np.max(np.abs(g["x"] - c["x"][lo:hi])) <= 1e-6, tag + "chained solve x block")

        n = 10000 * world
        d0, d1, xs = po.quad_tridiag_data(n, kappa=1e3, seed=0)
        prm20 = lb.LBFGSParam(m=20)
        g, lo, hi = solve(lb.OBJ_QUAD_TRIDIAG, n, np.zeros(n), prm20, "Bracketing", d0, d1, resident=resident)
        c = cpu(po.OBJ_QUAD_TRIDIAG, np.zeros(n), prm20, "BRACKETING", d0, d1)
        expect(c["status"] == "ok", tag + "cpu tridiag status " + c["status"])
        expect(abs(g["fx"] - c["fx"]) <= 1e-9 * abs(c["fx"]), tag + "tridiag solve fx %r %r" % (g["fx"], c["fx"]))
        expect(abs(g["niter"] - c["niter"]) <= max(3, c["niter"] // 20), tag + "tridiag solve iterations %d vs %d" % (g["niter"], c["niter"]))
        expect(np.max(np.abs(g["x"] - xs[lo:hi])) <= 1e-3, tag + "tridiag solve x block")

    # ---- 3. a batch of problems, every one n-sharded: ONE exchange per round carries all the running problems' sums ----------
    if mode == "p2p":
        n = 20000 * world
        lo, hi = shard_bounds(n, rank, world)
        lb.set_global_extent(local, lo, n)
        starts = [np.zeros(n), np.full(n, 0.5), np.full(n, -0.3), np.zeros(n)]
        bs = lb.BatchSession(lb.OBJ_ROSENBROCK_PAIRED, np.stack([x[lo:hi] for x in starts]), prm, "MoreThuente", device=local)
        res, X, _ = bs.solve()
        bs.close()
        for b, x0 in enumerate(starts):
            c = cpu(po.OBJ_ROSENBROCK_PAIRED, x0, prm, "MORE_THUENTE")
            expect(res[b]["status"] == "ok" and (res[b]["niter"], res[b]["nfev"]) == (c["niter"], c["nfev"]),
                   "sharded batch problem %d counts %r vs %r" % (b, (res[b]["niter"], res[b]["nfev"]), (c["niter"], c["nfev"])))
            expect(abs(res[b]["fx"] - c["fx"]) <= 1e-10 * max(1.0, abs(c["fx"])), "sharded batch problem %d fx" % b)
            expect(np.max(np.abs(X[b] - c["x"][lo:hi])) <= 1e-7, "sharded batch problem %d x block" % b)
        expect(res[0] == res[3] and np.array_equal(X[0], X[3]), "sharded batch: equal problems must give equal bits")

    bad = torch.tensor([len(failures)], dtype=torch.int64, device="cuda")
    dist.all_reduce(bad)
    for msg in failures:
        print("[rank %d] FAIL %s" % (rank, msg), flush=True)
    if rank == 0:
        print("MULTI_GPU_CHECK %s mode=%s world=%d failures=%d" % ("PASS" if bad.item() == 0 else "FAIL", mode, world, bad.item()), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if bad.item() == 0 else 1)


if __name__ == "__main__":
    main()
