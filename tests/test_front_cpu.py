"""CPU tests of the product's host logic: the line-search cores (include/LBFGSpp/LineSearchCore.h, the single source of the
scalar decisions for both the header-only front and the device-resident solve) must behave bit-identically to the restated
reference line searches, including every error path (exception kind + message)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "cpp", "core_harness.so")


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(ROOT, "tests", "cpp", "core_harness.cpp")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-march=x86-64-v3", "-fPIC", "-shared", "-Wall", "-Wl,-Bsymbolic",
                    "-I", os.path.join(ROOT, "oracle"), "-o", SO, src], check=True)
    lib = C.CDLL(SO)
    dp = C.POINTER(C.c_double)
    lib.core_line_search_f64.argtypes = [C.c_int, dp, dp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                         C.c_double, dp, dp, C.c_double, dp, dp, dp, dp, dp, dp, C.c_long, C.POINTER(C.c_long)]
    lib.core_error_message.restype = C.c_char_p
    return lib


def oracle_ls(orc, objective, ls, prm, xp, drt, step, step_max):
    n = xp.size
    dp = C.POINTER(C.c_double)
    P = lambda a: a.ctypes.data_as(dp)
    x, g, trace = np.zeros(n), np.zeros(n), np.zeros(4096)
    st, fx, dg = C.c_double(step), C.c_double(0), C.c_double(0)
    res = po.Result()
    orc.lib.orc_line_search_f64.argtypes = [C.c_int, dp, dp, C.c_long, C.c_int, C.POINTER(po.Param), dp, dp, C.c_double, dp, dp, dp, dp,
                                            dp, dp, C.c_long, C.POINTER(po.Result)]
    orc.lib.orc_line_search_f64(objective, None, None, n, ls, C.byref(prm), P(xp), P(drt), step_max, C.byref(st), C.byref(fx),
                                C.byref(dg), P(x), P(g), P(trace), 4096, C.byref(res))
    return dict(status=res.status, msg=res.msg.decode(), step=st.value, fx=fx.value, dg=dg.value, x=x, g=g,
                trace=trace[:res.trace_len].copy(), nfev=res.nfev)


def core_ls(lib, objective, ls, prm, xp, drt, step, step_max):
    n = xp.size
    dp = C.POINTER(C.c_double)
    P = lambda a: a.ctypes.data_as(dp)
    x, g, trace = np.zeros(n), np.zeros(n), np.zeros(4096)
    st, fx, dg, nfev = C.c_double(step), C.c_double(0), C.c_double(0), C.c_long(0)
    rc = lib.core_line_search_f64(objective, None, None, n, ls, prm.linesearch, prm.max_linesearch, prm.min_step, prm.max_step, prm.ftol,
                                  prm.wolfe, P(xp), P(drt), step_max, C.byref(st), C.byref(fx), C.byref(dg), P(x), P(g), P(trace), 4096,
                                  C.byref(nfev))
    status = 0 if rc == 0 else lib.core_error_kind(rc)
    msg = "" if rc == 0 else lib.core_error_message(rc).decode()
    return dict(status=status, msg=msg, step=st.value, fx=fx.value, dg=dg.value, x=x, g=g, trace=trace[:min(nfev.value, 4096)].copy(),
                nfev=nfev.value)


@pytest.mark.parametrize("ls", [0, 1, 2, 3])
def test_cores_match_restated_line_searches_bit_for_bit(harness, orc, ls):
    rng = np.random.default_rng(40 + ls)
    checked = errors = 0
    for trial in range(300):
        n = int(rng.choice([2, 4, 10, 50]))
        xp = rng.uniform(-1.5, 1.5, n)
        _, g = orc.objective(po.OBJ_ROSENBROCK_PAIRED, xp)
        # mostly descent directions of wildly different scales; sometimes an ascent direction (error path)
        drt = -g * 10.0 ** rng.uniform(-4, 1) + 0.05 * rng.standard_normal(n) * np.linalg.norm(g)
        if trial % 17 == 0:
            drt = g.copy()
        step = float(10.0 ** rng.uniform(-3, 1))
        prm = orc.default_param(linesearch=int(rng.choice([1, 2, 3])) if ls < 2 else 3,
                                max_linesearch=int(rng.choice([1, 2, 5, 20, 64])), min_step=float(rng.choice([1e-20, 1e-3])),
                                max_step=float(rng.choice([1e20, 2.0])))
        step_max = float(rng.choice([1e20, 5.0, 0.5]))
        a = oracle_ls(orc, po.OBJ_ROSENBROCK_PAIRED, ls, prm, xp, drt, step, step_max)
        b = core_ls(harness, po.OBJ_ROSENBROCK_PAIRED, ls, prm, xp, drt, step, step_max)
        assert a["status"] == b["status"], (trial, a["msg"], b["msg"])
        assert a["msg"] == b["msg"]
        assert a["nfev"] == b["nfev"] and np.array_equal(a["trace"], b["trace"])
        if a["status"] == 0:
            assert a["step"] == b["step"] and a["fx"] == b["fx"] and a["dg"] == b["dg"]
            assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["g"], b["g"])
            checked += 1
        else:
            errors += 1
    assert checked > 100 and errors > 5


# ---- host dense algebra of L-BFGS-B (include/LBFGSpp/BKLDLT.h) --------------------------------------------------------------
def bk_solve(lib, a, b, uplo=0, probe=0):
    dp = C.POINTER(C.c_double)
    lib.bkldlt_solve_f64.argtypes = [C.c_int, dp, C.c_int, dp, dp, C.c_int]
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    info = lib.bkldlt_solve_f64(a.shape[0], a.ctypes.data_as(dp), uplo, b.ctypes.data_as(dp), x.ctypes.data_as(dp), probe)
    return info, x


def test_bunch_kaufman_solves_symmetric_indefinite_systems(harness):
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 5, 12, 20, 40):
        for kind in range(5):
            a = rng.standard_normal((n, n))
            a = a + a.T
            if kind == 1 and n > 1:
                np.fill_diagonal(a, 0.0)              # every pivot must be 2x2 or swapped
            elif kind == 2:
                a = a @ a.T + n * np.eye(n)           # SPD: all 1x1, no swaps needed
            elif kind == 3:
                a[np.abs(a) < 0.8] = 0.0              # sparse pattern, many exact zeros
                a += np.diag(rng.choice([-3.0, 3.0], n))
            elif kind == 4 and n % 2 == 0:            # shape of the L-BFGS-B middle matrix [-D L'; L theta*S'S]
                c = n // 2
                s, y = rng.standard_normal((c, 50)), rng.standard_normal((c, 50))
                sy = s @ y.T
                a = np.block([[-np.diag(np.abs(np.diag(sy)) + 0.1), np.tril(sy, -1).T], [np.tril(sy, -1), 1.7 * (s @ s.T)]])
            b = rng.standard_normal(n)
            expect = np.linalg.solve(a, b)
            scale = np.linalg.cond(a) * np.finfo(float).eps * 50
            for uplo in (0, 1):
                tri = np.tril(a) if uplo == 0 else np.triu(a)  # the other triangle must not be read
                info, x = bk_solve(harness, tri + 7.0 * (np.triu(np.ones((n, n)), 1) if uplo == 0 else np.tril(np.ones((n, n)), -1)), b, uplo)
                assert info == 0
                assert np.max(np.abs(x - expect)) <= scale * max(1.0, np.max(np.abs(expect))), (n, kind, uplo)


def test_bunch_kaufman_error_paths(harness):
    """Non-square -> invalid_argument (reference BKLDLT.h:395-396); solve before compute -> logic_error (:446-447);
    a singular pivot block -> info() == NUMERICAL_ISSUE (2)."""
    assert bk_solve(harness, np.eye(2), np.ones(2), probe=-1)[0] == -1
    assert bk_solve(harness, np.eye(2), np.ones(2), probe=-2)[0] == -2
    assert bk_solve(harness, np.zeros((3, 3)), np.ones(3))[0] == 2


# ---- option structs (include/LBFGSpp/Param.h) ------------------------------------------------------------------------------
PARAM_FIELDS = ("m", "epsilon", "epsilon_rel", "past", "delta", "max_iterations", "linesearch", "max_submin", "max_linesearch",
                "min_step", "max_step", "ftol", "wolfe")
BAD_PARAMS = [dict(m=0), dict(m=-3), dict(epsilon=-1e-9), dict(epsilon_rel=-1.0), dict(past=-1), dict(delta=-1e-3),
              dict(max_iterations=-1), dict(linesearch=0), dict(linesearch=4), dict(max_linesearch=0), dict(min_step=-1.0),
              dict(max_step=1e-30), dict(ftol=0.0), dict(ftol=0.5), dict(wolfe=1e-5), dict(wolfe=1.0), dict(max_submin=-1),
              dict(m=0, epsilon=-1.0), dict(ftol=0.7, wolfe=0.6)]


@pytest.mark.parametrize("lbfgsb", [False, True])
def test_check_param_messages_equal_the_reference(harness, orc, lbfgsb):
    """LBFGSParam / LBFGSBParam::check_param (reference Param.h:191-218, 350-376): same rule order, same messages; defaults valid."""
    harness.param_check_message.restype = C.c_char_p
    harness.param_check_message.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
    for bad in [dict()] + BAD_PARAMS:
        if lbfgsb and "linesearch" in bad or not lbfgsb and "max_submin" in bad:
            continue
        p = orc.default_param(lbfgsb=lbfgsb, **bad)
        got = harness.param_check_message(int(lbfgsb), *[getattr(p, f) for f in PARAM_FIELDS]).decode()
        if lbfgsb:
            r = orc.lbfgsb(po.OBJ_QUAD_SHIFT, np.zeros(4), -1.0, 1.0, p)
        else:
            r = orc.lbfgs(po.OBJ_QUAD_SHIFT, np.zeros(4), po.LS_BACKTRACKING, p)
        expect = r["msg"] if r["status"] == "invalid_argument" else ""
        assert got == expect, (bad, got, expect)
        assert (got == "") == (not bad)


def test_product_sources_never_reach_into_the_oracle():
    """The checker is test infrastructure: nothing under include/ or lbfgspp_b200/ may include, link or load oracle/ files."""
    offenders = []
    for top in ("include", "lbfgspp_b200"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".h", ".hpp", ".cu", ".cuh", ".cpp", ".py")):
                    for line in open(os.path.join(dirpath, f), errors="replace"):
                        uses = any(tok in line for tok in ("#include", "import ", "CDLL", "dlopen", "sys.path", "-L", "-l:"))
                        if uses and any(needle in line for needle in ("oracle", "libref_lbfgspp", "minieigen")):
                            offenders.append((os.path.join(dirpath, f), line.strip()))
    assert not offenders, offenders
