"""Small whole solves for tests/gpu_sanitize.sh (compute-sanitizer target; not collected by pytest): both solver loops, all four line
searches, the neighbour-coupled objectives (staged tiles with margins), fp32, a batch, the literal two-loop recursion, L-BFGS-B."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import lbfgspp_b200 as lb  # noqa: E402
import pyoracle as po  # noqa: E402  (only the deterministic data generator)

n = 6002
for resident in (False, True):
    for ls in ("Backtracking", "Bracketing", "NocedalWright", "MoreThuente"):
        g = lb.LBFGSSolver(lb.LBFGSParam(m=7), ls, resident=resident).minimize(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(n))
        print("lbfgs", "resident" if resident else "host", ls, g["status"], g["niter"], g["nfev"], g["fx"])
d, b, _ = po.quad_tridiag_data(5003, seed=2)
for resident in (False, True):
    g = lb.LBFGSSolver(lb.LBFGSParam(m=12, max_iterations=25), "Bracketing", resident=resident).minimize(lb.OBJ_QUAD_TRIDIAG, np.zeros(5003), data0=d, data1=b)
    print("tridiag", "resident" if resident else "host", g["status"], g["niter"], g["nfev"], g["fx"])
    g = lb.LBFGSSolver(lb.LBFGSParam(m=5), "NocedalWright", resident=resident).minimize(lb.OBJ_ROSENBROCK_CHAINED, np.full(3000, 1.3))
    print("chained", "resident" if resident else "host", g["status"], g["niter"], g["nfev"], g["fx"])
g = lb.LBFGSSolver(lb.LBFGSParam(), "MoreThuente", dtype=np.float32, resident=True).minimize(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(4096))
print("lbfgs f32 resident", g["status"], g["niter"], g["fx"])
bs = lb.BatchSession(lb.OBJ_ROSENBROCK_PAIRED, np.stack([np.zeros(3000), np.full(3000, 0.5), np.ones(3000)]), lb.LBFGSParam(m=6), "MoreThuente")
res, _, _ = bs.solve()
bs.close()
print("batch", [(r["status"], r["niter"]) for r in res])
g = lb.LBFGSSolver(lb.LBFGSParam(m=5), "NocedalWright", hv_algo=lb.HV_TWO_LOOP).minimize(lb.OBJ_ROSENBROCK_CHAINED, np.full(3000, 1.3))
print("lbfgs two-loop chained", g["status"], g["niter"], g["fx"])
g = lb.LBFGSBSolver(lb.LBFGSBParam()).minimize(lb.OBJ_ROSENBROCK_CHAINED, np.full(5000, 3.0), 2.0, 4.0)
print("lbfgsb", g["status"], g["niter"], g["nfev"], g["fx"])
