"""GPU checks at BASELINE config 2's full size (paired Rosenbrock n = 1e7, fp64, m = 10): the solve against the frozen CPU
run of the unmodified reference headers (tests/golden/c2_full.json, generator tests/golden/make_c2_full.py), and apply_Hv
through properties that need no CPU run of that size: the secant equation H y_new = s_new, linearity, agreement of the two
independent device algorithms (literal two-loop / Gram form) and run-to-run determinism."""
import json
import os

import numpy as np
import pytest

import lbfgspp_b200 as lb

pytestmark = pytest.mark.gpu

N, M = 10_000_000, 10
HERE = os.path.dirname(os.path.abspath(__file__))


def test_config2_full_size_solve_matches_reference_cpu_run():
    with open(os.path.join(HERE, "golden", "c2_full.json")) as fh:
        c = json.load(fh)
    assert (c["n"], c["m"]) == (N, M)
    g = lb.LBFGSSolver(lb.LBFGSParam(m=M), "MoreThuente").minimize(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(N))
    assert g["status"] == "ok"
    assert (g["niter"], g["nfev"]) == (c["niter"], c["nfev"])
    fx_cpu = float.fromhex(c["fx"])
    assert abs(g["fx"] - fx_cpu) <= 1e-10 * max(1.0, abs(fx_cpu))
    x = g["x"]
    # x0 = 0 makes this a replicated 2-D problem: every pair must carry exactly the same two numbers
    assert np.all(x[0::2] == x[0]) and np.all(x[1::2] == x[1])
    assert abs(x[0] - float.fromhex(c["x_even"])) <= 1e-8 and abs(x[1] - float.fromhex(c["x_odd"])) <= 1e-8
    trace_cpu = np.array([float.fromhex(v) for v in c["trace"]])
    k = np.arange(trace_cpu.size)
    tol = np.minimum(1e-10 * 4.0 ** k, 1e-3) * np.maximum(np.abs(trace_cpu), 1e-6 * abs(trace_cpu[0]))
    assert np.all(np.abs(g["trace"] - trace_cpu) <= tol)


def test_apply_Hv_full_size_properties(gpu_ctx):
    rng = np.random.default_rng(0)
    hist = lb.History(gpu_ctx, N, M)
    s = y = None
    for _ in range(M + 2):  # two more than the ring holds: the oldest pairs are overwritten
        s = rng.standard_normal(N)
        y = s + 0.1 * rng.standard_normal(N)
        ds, dy = lb.DeviceArray(gpu_ctx, s), lb.DeviceArray(gpu_ctx, y)
        hist.add(ds, dy)
        del ds, dy
    assert hist.ncorr == M
    dy = lb.DeviceArray(gpu_ctx, y)
    res = gpu_ctx.empty(N)
    s_scale = np.max(np.abs(s))
    for algo in (lb.HV_TWO_LOOP, lb.HV_GRAM):
        hist.apply_Hv(dy, 1.0, res, algo)
        assert np.max(np.abs(res.get() - s)) <= 1e-10 * s_scale      # secant equation of the BFGS inverse update
    v1, v2 = rng.standard_normal(N), rng.standard_normal(N)
    a, b = 2.5, -0.75
    out = {}
    for algo in (lb.HV_TWO_LOOP, lb.HV_GRAM):
        h = []
        for v in (v1, v2, a * v1 + b * v2):
            dv = lb.DeviceArray(gpu_ctx, v)
            hist.apply_Hv(dv, -1.0, res, algo)
            h.append(res.get())
            del dv
        scale = np.max(np.abs(h[2])) + 1.0
        assert np.max(np.abs(h[2] - (a * h[0] + b * h[1]))) <= 1e-11 * scale   # linearity
        out[algo] = h[2]
    scale = np.max(np.abs(out[lb.HV_GRAM])) + 1.0
    assert np.max(np.abs(out[lb.HV_TWO_LOOP] - out[lb.HV_GRAM])) <= 1e-11 * scale  # two device algorithms agree
    dv = lb.DeviceArray(gpu_ctx, a * v1 + b * v2)
    hist.apply_Hv(dv, -1.0, res, lb.HV_GRAM)
    assert np.array_equal(res.get(), out[lb.HV_GRAM])                # bitwise repeatable
