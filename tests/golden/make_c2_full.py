"""Generate tests/golden/c2_full.json: BASELINE config 2 at its full size (paired Rosenbrock, n = 1e7, fp64, m = 10,
More-Thuente, x0 = 0) run on the CPU by the unmodified reference headers (oracle/_ref) -- about a minute of one core.
Run from the repo root in the build container:  python tests/golden/make_c2_full.py
The run also asserts that the restatement (oracle/liboracle.so, sequential sums) reproduces it bit for bit."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as po  # noqa: E402

N, M = 10_000_000, 10


def main():
    ref, orc = po.Oracle("ref"), po.Oracle("orc")
    r = ref.lbfgs(po.OBJ_ROSENBROCK_PAIRED, np.zeros(N), po.LS_MORE_THUENTE, ref.default_param(m=M))
    o = orc.lbfgs(po.OBJ_ROSENBROCK_PAIRED, np.zeros(N), po.LS_MORE_THUENTE, orc.default_param(m=M))
    assert r["status"] == "ok"
    assert (r["niter"], r["nfev"], r["fx"]) == (o["niter"], o["nfev"], o["fx"]) and np.array_equal(r["x"], o["x"])
    assert np.all(r["x"][0::2] == r["x"][0]) and np.all(r["x"][1::2] == r["x"][1])
    out = dict(n=N, m=M, linesearch="MoreThuente", niter=r["niter"], nfev=r["nfev"], fx=float(r["fx"]).hex(),
               gnorm=float(r["gnorm"]).hex(), x_even=float(r["x"][0]).hex(), x_odd=float(r["x"][1]).hex(),
               trace=[float(v).hex() for v in r["trace"]], seconds_ref=r["seconds"], seconds_restatement=o["seconds"])
    with open(os.path.join(ROOT, "tests", "golden", "c2_full.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(out["niter"], out["nfev"], r["fx"], r["seconds"], o["seconds"])


if __name__ == "__main__":
    main()
