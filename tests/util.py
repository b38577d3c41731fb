import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LS = {"Backtracking": 0, "Bracketing": 1, "NocedalWright": 2, "MoreThuente": 3}


def unhex(xs):
    return np.array([float.fromhex(v) for v in xs], dtype=np.float64)


def golden_cases(kind=None):
    with open(os.path.join(HERE, "golden", "lbfgs_ref.json")) as fh:
        cases = json.load(fh)["cases"]
    return [c for c in cases if kind is None or c["kind"] == kind]


def same_run(a, b):
    """Bit-for-bit equality of two solver result dicts (status, counts, trace, x, grad)."""
    return (a["status"] == b["status"] and a["msg"] == b["msg"] and a["niter"] == b["niter"] and a["nfev"] == b["nfev"]
            and np.array_equal(a["trace"], b["trace"]) and np.array_equal(a["x"], b["x"])
            and np.array_equal(a["grad"], b["grad"]))
