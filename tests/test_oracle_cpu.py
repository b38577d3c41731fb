"""CPU tests of the checker itself: the restatement must reproduce (a) the unmodified reference headers run over
minieigen, bit for bit, wherever that build exists, and (b) the golden vectors frozen from that build, everywhere."""
import numpy as np
import pytest

import pyoracle as po
from util import LS, golden_cases, same_run, unhex


def run_golden_lbfgs(orc, c):
    dtype = np.float64 if c["dtype"] == "f64" else np.float32
    p = orc.default_param(**c["param"])
    d0, d1 = (unhex(c["data"][0]), unhex(c["data"][1])) if c["data"] else (None, None)
    return orc.lbfgs(c["objective"], unhex(c["x0"]), LS[c["ls"]], p, data0=d0, data1=d1, dtype=dtype)


@pytest.mark.parametrize("case", golden_cases("lbfgs"), ids=lambda c: c["name"])
def test_restatement_matches_golden_lbfgs(orc, case):
    r = run_golden_lbfgs(orc, case)
    assert r["status"] == case["status"] and r["msg"] == case["msg"]
    assert (r["niter"], r["nfev"]) == (case["niter"], case["nfev"])
    assert np.array_equal(r["trace"], unhex(case["trace"]))
    assert np.array_equal(r["x"].astype(np.float64), unhex(case["x"]))
    assert np.array_equal(r["grad"].astype(np.float64), unhex(case["grad"]))
    if case["status"] == "ok":
        assert float(r["fx"]).hex() == case["fx"] and float(r["gnorm"]).hex() == case["gnorm"]


@pytest.mark.parametrize("case", golden_cases("apply_Hv"), ids=lambda c: c["name"])
def test_restatement_matches_golden_apply_Hv(orc, case):
    n, npairs = case["n"], case["npairs"]
    S = unhex(case["S"]).reshape(npairs, n)
    Y = unhex(case["Y"]).reshape(npairs, n)
    res, _, theta = orc.apply_Hv(S, Y, unhex(case["v"]), case["a"], case["m"])
    assert np.array_equal(res, unhex(case["res"]))
    assert float(theta).hex() == case["theta"]


@pytest.mark.parametrize("case", golden_cases("lbfgsb"), ids=lambda c: c["name"])
def test_restatement_matches_golden_lbfgsb(orc, case):
    p = orc.default_param(lbfgsb=True, **case["param"])
    r = orc.lbfgsb(case["objective"], unhex(case["x0"]), unhex(case["lb"]), unhex(case["ub"]), p)
    assert r["status"] == case["status"] and r["msg"] == case["msg"]
    assert (r["niter"], r["nfev"]) == (case["niter"], case["nfev"])
    assert np.array_equal(r["trace"], unhex(case["trace"]))
    assert np.array_equal(r["x"], unhex(case["x"])) and np.array_equal(r["grad"], unhex(case["grad"]))
    assert float(r["fx"]).hex() == case["fx"] and float(r["gnorm"]).hex() == case["gnorm"]


def test_survey_probe_values(orc):
    """BASELINE.md section 4 (independent numpy transliteration): iteration / evaluation counts must agree."""
    expect = {"NocedalWright": (22, 36), "MoreThuente": (21, 28), "Bracketing": (22, 31), "Backtracking": (22, 31)}
    for ls, (it, fev) in expect.items():
        r = orc.lbfgs(po.OBJ_ROSENBROCK_PAIRED, np.zeros(10), LS[ls], orc.default_param())
        assert (r["niter"], r["nfev"]) == (it, fev)
    r = orc.lbfgs(po.OBJ_QUAD_SHIFT, np.zeros(10), LS["NocedalWright"], orc.default_param())
    assert r["niter"] == 2 and np.allclose(r["x"], np.arange(10.0))
    r = orc.lbfgs(po.OBJ_ROSENBROCK_PAIRED, np.zeros(1000), LS["MoreThuente"], orc.default_param(m=10))
    assert (r["niter"], r["nfev"]) == (23, 31)


def test_self_checking_examples_property(orc):
    """example-rosenbrock-comparison.cpp:44-51,83-86: |x - 1|_inf <= 1e-4 from random starts, all line searches."""
    rng = np.random.default_rng(7)
    for n in (2, 8, 24):
        for _ in range(16):
            x0 = rng.uniform(-1, 1, n)
            for ls in range(4):
                r = orc.lbfgs(po.OBJ_ROSENBROCK_PAIRED, x0, ls, orc.default_param(max_linesearch=256))
                assert r["status"] == "ok" and np.max(np.abs(r["x"] - 1.0)) <= 1e-4


def test_lanes_and_omp_modes_agree_with_sequential(orc):
    """Summation-order variants used for the CPU baseline stay within rounding of the faithful one."""
    x0 = np.zeros(2000)
    base = orc.lbfgs(po.OBJ_ROSENBROCK_PAIRED, x0, 3, orc.default_param(m=10))
    for mode in (po.SUM_LANES8, po.SUM_LANES8_OMP):
        r = orc.lbfgs(po.OBJ_ROSENBROCK_PAIRED, x0, 3, orc.default_param(m=10), sum_mode=mode)
        assert (r["niter"], r["nfev"]) == (base["niter"], base["nfev"])
        assert abs(r["fx"] - base["fx"]) <= 1e-10 * max(1.0, abs(base["fx"]))
        assert np.max(np.abs(r["x"] - base["x"])) <= 1e-8


def test_gram_form_matches_two_loop_on_cpu(orc):
    """The vector-free (Gram) two-loop is the same recursion up to rounding: same counts, same fx."""
    for n in (10, 1000, 20000):
        a = orc.lbfgs(po.OBJ_ROSENBROCK_PAIRED, np.zeros(n), 3, orc.default_param(m=10))
        b = orc.lbfgs(po.OBJ_ROSENBROCK_PAIRED, np.zeros(n), 3, orc.default_param(m=10), gram=True)
        assert (a["niter"], a["nfev"]) == (b["niter"], b["nfev"])
        assert abs(a["fx"] - b["fx"]) <= 1e-10 * max(1.0, abs(a["fx"]))


# ---- against the unmodified reference headers (only where oracle/_ref could be built) ------------------------------
def test_pin_against_reference_headers(orc, ref):
    rng = np.random.default_rng(1)
    for n in (2, 6, 24, 100, 2002):
        for t in range(6):
            x0 = rng.uniform(-1, 1, n)
            for ls in range(4):
                p = ref.default_param(max_linesearch=256, m=3 + t % 8)
                assert same_run(ref.lbfgs(po.OBJ_ROSENBROCK_PAIRED, x0, ls, p), orc.lbfgs(po.OBJ_ROSENBROCK_PAIRED, x0, ls, p))
    d, b, _ = po.quad_tridiag_data(3000)
    for ls in range(4):
        p = ref.default_param(m=20)
        assert same_run(ref.lbfgs(po.OBJ_QUAD_TRIDIAG, np.zeros(3000), ls, p, data0=d, data1=b),
                        orc.lbfgs(po.OBJ_QUAD_TRIDIAG, np.zeros(3000), ls, p, data0=d, data1=b))


def test_pin_float32_against_reference_headers(orc, ref):
    for ls in range(4):
        p = ref.default_param()
        a = ref.lbfgs(po.OBJ_ROSENBROCK_PAIRED, np.zeros(10), ls, p, dtype=np.float32)
        b = orc.lbfgs(po.OBJ_ROSENBROCK_PAIRED, np.zeros(10), ls, p, dtype=np.float32)
        assert same_run(a, b)


def test_pin_apply_Hv_against_reference_headers(orc, ref):
    rng = np.random.default_rng(3)
    n, m = 1000, 5
    for npairs in (0, 1, 4, 5, 6, 13):
        S = rng.standard_normal((npairs, n))
        Y = S + 0.1 * rng.standard_normal((npairs, n))
        v = rng.standard_normal(n)
        ra, _, ta = ref.apply_Hv(S, Y, v, -1.0, m)
        rb, _, tb = orc.apply_Hv(S, Y, v, -1.0, m)
        assert np.array_equal(ra, rb) and ta == tb


def test_pin_lbfgsb_against_reference_headers(orc, ref):
    """Cauchy point, subspace minimisation, Bunch-Kaufman LDL' and the L-BFGS-B driver, bit for bit: mixed finite /
    infinite / degenerate (l == u) bounds, history sizes 1..10, 0..10 BOXCQP sweeps, delta-based stopping."""
    rng = np.random.default_rng(3)
    for obj in (po.OBJ_ROSENBROCK_CHAINED, po.OBJ_ROSENBROCK_PAIRED, po.OBJ_QUAD_SHIFT, po.OBJ_QUAD_TRIDIAG):
        for n in (2, 4, 10, 25, 26, 100):
            if obj == po.OBJ_ROSENBROCK_PAIRED and n % 2:
                continue
            for trial in range(6):
                lo = rng.uniform(-2, 1, n)
                hi = lo + rng.uniform(0, 3, n)
                if trial == 0:
                    lo[:], hi[:] = 2.0, 4.0
                elif trial == 1:
                    lo[:], hi[:] = -np.inf, np.inf
                elif trial == 2:
                    hi[::3] = lo[::3]
                elif trial == 3:
                    lo[::2], hi[1::3] = -np.inf, np.inf
                x0 = rng.uniform(-3, 5, n)
                d0, d1 = po.quad_tridiag_data(n, seed=trial)[:2] if obj == po.OBJ_QUAD_TRIDIAG else (None, None)
                kw = dict(m=int(rng.choice([1, 2, 3, 6, 10])), max_iterations=int(rng.choice([0, 5, 50])),
                          max_submin=int(rng.choice([0, 1, 10])), past=int(rng.choice([0, 0, 2])))
                if kw["past"]:
                    kw["delta"] = 1e-8
                a = orc.lbfgsb(obj, x0, lo, hi, orc.default_param(lbfgsb=True, **kw), data0=d0, data1=d1)
                b = ref.lbfgsb(obj, x0, lo, hi, ref.default_param(lbfgsb=True, **kw), data0=d0, data1=d1)
                assert same_run(a, b), (obj, n, trial, kw)


def test_pin_lbfgsb_larger_against_reference_headers(orc, ref):
    for n, obj in ((2000, po.OBJ_ROSENBROCK_CHAINED), (5000, po.OBJ_ROSENBROCK_PAIRED)):
        a = orc.lbfgsb(obj, np.full(n, 3.0), 2.0, 4.0, orc.default_param(lbfgsb=True))
        b = ref.lbfgsb(obj, np.full(n, 3.0), 2.0, 4.0, ref.default_param(lbfgsb=True))
        assert same_run(a, b) and a["status"] == "ok"


def test_lbfgsb_parameter_and_bound_errors(orc):
    """LBFGSB.h:132-133 (size check) and Param.h:351-376 (check_param) surface as invalid_argument."""
    r = orc.lbfgsb(po.OBJ_QUAD_SHIFT, np.zeros(4), 0.0, 1.0, orc.default_param(lbfgsb=True, m=0))
    assert r["status"] == "invalid_argument" and "'m' must be positive" in r["msg"]
    r = orc.lbfgsb(po.OBJ_QUAD_SHIFT, np.zeros(4), 0.0, 1.0, orc.default_param(lbfgsb=True, max_submin=-1))
    assert r["status"] == "invalid_argument" and "max_submin" in r["msg"]


def test_restatement_is_sanitizer_clean():
    """`make -C oracle sanitize`: the restatement under ASan + UBSan over 552 small solves (all line searches, summation modes,
    L-BFGS-B with mixed bounds)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["make", "-C", os.path.join(root, "oracle"), "sanitize"], capture_output=True, text=True, timeout=600)
    if r.returncode != 0 and ("cannot find -lasan" in r.stderr or "cannot find -lubsan" in r.stderr):
        pytest.skip("sanitizer runtimes not installed")
    assert r.returncode == 0 and "selfcheck ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
