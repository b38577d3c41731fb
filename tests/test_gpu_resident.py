"""The device-resident solve (whole minimize() = one CUDA graph launch, line search / convergence / gate decided on the device)
must return exactly what the host-driven loop returns: same kernels bodies, same grids, same decision code
(include/LBFGSpp/LineSearchCore.h) -- so the comparison is bit for bit, including the trace of f and the error paths."""
import numpy as np
import pytest

import lbfgspp_b200 as lb
import pyoracle as po
from util import golden_cases, unhex

pytestmark = pytest.mark.gpu
LS_NAMES = ["Backtracking", "Bracketing", "NocedalWright", "MoreThuente"]


def both(objective, x0, prm, ls, dtype=np.float64, **kw):
    # HV_GRAM_UNFUSED: the host-driven loop with the separate update kernel, i.e. the arithmetic the graph runs (the default
    # host-driven loop forms the pair inside the dots pass and takes s'y, y'y from there -- equal up to rounding only)
    host = lb.LBFGSSolver(prm, ls, dtype=dtype, resident=False, hv_algo=lb.HV_GRAM_UNFUSED).minimize(objective, x0, **kw)
    res = lb.LBFGSSolver(prm, ls, dtype=dtype, resident=True).minimize(objective, x0, **kw)
    return host, res


def assert_identical(host, res):
    assert res["status"] == host["status"] and res["msg"] == host["msg"]
    assert (res["niter"], res["nfev"]) == (host["niter"], host["nfev"])
    assert np.array_equal(res["trace"], host["trace"])
    if host["status"] == "ok":
        assert res["fx"] == host["fx"] and res["gnorm"] == host["gnorm"]
        assert np.array_equal(res["x"], host["x"]) and np.array_equal(res["grad"], host["grad"])


@pytest.mark.parametrize("ls", LS_NAMES)
@pytest.mark.parametrize("n", [10, 4098, 100000])
def test_resident_equals_host_driven(ls, n):
    prm = lb.LBFGSParam(m=10 if n > 10 else 6)
    host, res = both(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(n), prm, ls)
    assert host["status"] == "ok"
    assert_identical(host, res)
    assert res["launches"] == 1      # one graph launch


def test_resident_random_start_and_other_objectives():
    rng = np.random.default_rng(3)
    host, res = both(lb.OBJ_ROSENBROCK_PAIRED, rng.uniform(-1, 1, 2000), lb.LBFGSParam(m=7, max_linesearch=64), "MoreThuente")
    assert_identical(host, res)
    n = 5000
    d, b, _ = po.quad_tridiag_data(n, seed=1)
    host, res = both(lb.OBJ_QUAD_TRIDIAG, np.zeros(n), lb.LBFGSParam(m=20), "Bracketing", data0=d, data1=b)
    assert_identical(host, res)
    host, res = both(lb.OBJ_ROSENBROCK_CHAINED, np.full(300, 1.3), lb.LBFGSParam(), "NocedalWright")
    assert_identical(host, res)
    host, res = both(lb.OBJ_QUAD_SHIFT, np.zeros(10), lb.LBFGSParam(), "NocedalWright")
    assert_identical(host, res) and None
    assert res["niter"] == 2


def test_resident_float32():
    host, res = both(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(64), lb.LBFGSParam(), "NocedalWright", dtype=np.float32)
    assert_identical(host, res)


def test_resident_stopping_rules_and_error_paths():
    for prm, ls in ((lb.LBFGSParam(max_iterations=5), "NocedalWright"), (lb.LBFGSParam(past=3, delta=1e-6), "MoreThuente"),
                    (lb.LBFGSParam(max_linesearch=1), "Backtracking"), (lb.LBFGSParam(max_linesearch=1), "Bracketing"),
                    (lb.LBFGSParam(max_linesearch=1, max_iterations=3), "MoreThuente"),
                    (lb.LBFGSParam(max_linesearch=1, max_iterations=3), "NocedalWright"),
                    (lb.LBFGSParam(linesearch=1), "NocedalWright"), (lb.LBFGSParam(linesearch=1, max_linesearch=64), "Backtracking"),
                    (lb.LBFGSParam(linesearch=2, max_linesearch=64), "Bracketing"), (lb.LBFGSParam(m=0), "MoreThuente")):
        host, res = both(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(12), prm, ls)
        assert_identical(host, res)
    # start point already optimal: minimize returns 1 without a line search
    host, res = both(lb.OBJ_ROSENBROCK_PAIRED, np.ones(6), lb.LBFGSParam(), "MoreThuente")
    assert_identical(host, res)
    assert res["niter"] == 1 and res["nfev"] == 1


@pytest.mark.parametrize("case", [c for c in golden_cases("lbfgs") if c["dtype"] == "f64"], ids=lambda c: c["name"])
def test_resident_on_golden_cases(case):
    prm = lb.LBFGSParam(**case["param"])
    d0, d1 = (unhex(case["data"][0]), unhex(case["data"][1])) if case["data"] else (None, None)
    host, res = both(case["objective"], unhex(case["x0"]), prm, case["ls"], data0=d0, data1=d1)
    assert_identical(host, res)
    assert res["status"] == case["status"] and res["msg"] == case["msg"]


@pytest.mark.parametrize("ls", LS_NAMES)
def test_default_host_loop_agrees_with_resident_to_rounding(ls):
    """The default host-driven loop forms the pair inside the dots pass (s'y, y'y summed in a different order than the update
    kernel of the graph): same counts and fx within the parity tolerance on a well-conditioned run."""
    n = 100000
    prm = lb.LBFGSParam(m=10)
    host = lb.LBFGSSolver(prm, ls, resident=False).minimize(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(n))
    res = lb.LBFGSSolver(prm, ls, resident=True).minimize(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(n))
    assert host["status"] == res["status"] == "ok"
    assert (host["niter"], host["nfev"]) == (res["niter"], res["nfev"])
    assert abs(host["fx"] - res["fx"]) <= 1e-10 * max(1.0, abs(res["fx"]))
    assert np.max(np.abs(host["x"] - res["x"])) <= 1e-8
