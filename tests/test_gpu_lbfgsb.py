"""GPU parity tests for the bound-constrained solver (BASELINE config 4) against the reference headers' outputs:
the frozen golden vectors (tests/golden/lbfgs_ref.json, kind "lbfgsb") everywhere, and live runs of oracle/_ref where
that build travelled with the repo.  Parity: same iteration count, evaluation count within +-1 on the short runs (the
Cauchy sweep and the BOXCQP solves re-associate sums), |fx - fx_ref| <= 1e-9 max(1,|fx_ref|), |x - x_ref|_inf <= 1e-6."""
import numpy as np
import pytest

import lbfgspp_b200 as lb
import pyoracle as po
from util import golden_cases, unhex

pytestmark = pytest.mark.gpu


def check(g, c_status, c_niter, c_nfev, c_fx, c_x, iter_slack=0):
    assert g["status"] == c_status, (g["status"], g["msg"])
    assert abs(g["niter"] - c_niter) <= iter_slack, (g["niter"], c_niter)
    assert abs(g["nfev"] - c_nfev) <= iter_slack + 1, (g["nfev"], c_nfev)
    assert abs(g["fx"] - c_fx) <= 1e-9 * max(1.0, abs(c_fx)), (g["fx"], c_fx)
    assert np.max(np.abs(g["x"] - c_x)) <= 1e-6 * max(1.0, np.max(np.abs(c_x)))


@pytest.mark.parametrize("case", golden_cases("lbfgsb"), ids=lambda c: c["name"])
def test_lbfgsb_golden_vectors(case):
    prm = lb.LBFGSBParam(**case["param"])
    g = lb.LBFGSBSolver(prm).minimize(case["objective"], unhex(case["x0"]), unhex(case["lb"]), unhex(case["ub"]))
    check(g, case["status"], case["niter"], case["nfev"], float.fromhex(case["fx"]), unhex(case["x"]),
          iter_slack=0 if case["niter"] < 20 else 2)
    lbv, ubv = unhex(case["lb"]), unhex(case["ub"])
    assert np.all(g["x"] >= lbv) and np.all(g["x"] <= ubv)


@pytest.mark.parametrize("kind,n", [(lb.OBJ_ROSENBROCK_PAIRED, 1000), (lb.OBJ_ROSENBROCK_CHAINED, 1000),
                                    (lb.OBJ_ROSENBROCK_PAIRED, 100000), (lb.OBJ_ROSENBROCK_CHAINED, 100000)])
def test_config4_shape_box_2_4(ref, kind, n):
    """BASELINE config 4 (both readings of 'Rosenbrock-box'): lb = 2, ub = 4, x0 = 3."""
    x0 = np.full(n, 3.0)
    g = lb.LBFGSBSolver(lb.LBFGSBParam()).minimize(kind, x0, 2.0, 4.0)
    c = ref.lbfgsb(kind, x0, 2.0, 4.0, ref.default_param(lbfgsb=True))
    check(g, c["status"], c["niter"], c["nfev"], c["fx"], c["x"])


def test_box_random_interior_and_loose_bounds(ref):
    rng = np.random.default_rng(4)
    n = 5000
    x0 = rng.uniform(-1, 1, n)
    lbv = np.full(n, -0.5)
    ubv = np.full(n, 0.8)
    lbv[::7] = -np.inf
    ubv[::11] = np.inf
    g = lb.LBFGSBSolver(lb.LBFGSBParam()).minimize(lb.OBJ_ROSENBROCK_PAIRED, x0, lbv, ubv)
    c = ref.lbfgsb(po.OBJ_ROSENBROCK_PAIRED, x0, lbv, ubv, ref.default_param(lbfgsb=True))
    assert g["status"] == c["status"] == "ok"
    assert abs(g["fx"] - c["fx"]) <= 1e-6 * max(1.0, abs(c["fx"]))
    assert np.all(g["x"] >= lbv) and np.all(g["x"] <= ubv)


def test_bounds_size_mismatch_is_invalid_argument():
    ctx_err = lb.LBFGSBSolver(lb.LBFGSBParam(m=0)).minimize(lb.OBJ_ROSENBROCK_PAIRED, np.full(4, 3.0), 2.0, 4.0)
    assert ctx_err["status"] == "invalid_argument" and "'m' must be positive" in ctx_err["msg"]
