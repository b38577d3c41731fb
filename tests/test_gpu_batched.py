"""BASELINE config 5 (batched independent problems), at sizes a test can afford: every problem of a batch must come out
bit-identical to the same problem solved alone (same kernels, deterministic reductions), whatever the thread count, and
must satisfy the reference examples' convergence criterion."""
import numpy as np
import pytest

import lbfgspp_b200 as lb

pytestmark = pytest.mark.gpu


def test_batch_equals_individual_solves_bit_for_bit():
    B, n = 10, 4096
    X0 = np.stack([np.random.default_rng(1000 + b).uniform(-1, 1, n) for b in range(B)])
    prm = lb.LBFGSParam(m=10)
    res4, X4, _ = lb.solve_batch(lb.OBJ_ROSENBROCK_PAIRED, X0, prm, "MoreThuente", threads=4)
    res1, X1, _ = lb.solve_batch(lb.OBJ_ROSENBROCK_PAIRED, X0, prm, "MoreThuente", threads=1)
    for b in range(B):
        # the batch workers leave the solver loop on automatic: device-resident for built-in objectives up to n = 4e6
        single = lb.LBFGSSolver(prm, "MoreThuente", resident=True).minimize(lb.OBJ_ROSENBROCK_PAIRED, X0[b])
        for r, X in ((res4, X4), (res1, X1)):
            assert r[b]["status"] == "ok"
            assert (r[b]["niter"], r[b]["nfev"]) == (single["niter"], single["nfev"])
            assert r[b]["fx"] == single["fx"] and np.array_equal(X[b], single["x"])
        assert np.max(np.abs(X4[b] - 1.0)) <= 2e-3     # eps_rel * ||x|| loosens with sqrt(n)


@pytest.mark.parametrize("ls", ["Backtracking", "Bracketing", "NocedalWright", "MoreThuente"])
def test_batch_all_line_searches_converge(ls):
    B, n = 8, 24
    X0 = np.stack([np.random.default_rng(50 + b).uniform(-1, 1, n) for b in range(B)])
    res, X, _ = lb.solve_batch(lb.OBJ_ROSENBROCK_PAIRED, X0, lb.LBFGSParam(max_linesearch=256), ls, threads=3)
    assert all(r["status"] == "ok" for r in res)
    assert np.max(np.abs(X - 1.0)) <= 1e-4


def test_batch_reports_per_problem_failures():
    X0 = np.zeros((3, 8))
    res, _, _ = lb.solve_batch(lb.OBJ_ROSENBROCK_PAIRED, X0, lb.LBFGSParam(max_linesearch=1), "Backtracking", threads=2)
    assert [r["status"] for r in res] == ["runtime_error"] * 3
