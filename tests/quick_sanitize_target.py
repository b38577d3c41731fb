"""Small whole solves for tests/gpu_sanitize.sh (compute-sanitizer target; not collected by pytest)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import lbfgspp_b200 as lb  # noqa: E402

n = 6002
for resident in (False, True):
    for ls in ("Backtracking", "Bracketing", "NocedalWright", "MoreThuente"):
        g = lb.LBFGSSolver(lb.LBFGSParam(m=7), ls, resident=resident).minimize(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(n))
        print("lbfgs", "resident" if resident else "host", ls, g["status"], g["niter"], g["nfev"], g["fx"])
g = lb.LBFGSSolver(lb.LBFGSParam(), "MoreThuente", dtype=np.float32).minimize(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(4096))
print("lbfgs f32", g["status"], g["niter"], g["fx"])
g = lb.LBFGSSolver(lb.LBFGSParam(m=5), "NocedalWright", hv_algo=lb.HV_TWO_LOOP).minimize(lb.OBJ_ROSENBROCK_CHAINED, np.full(3000, 1.3))
print("lbfgs two-loop chained", g["status"], g["niter"], g["fx"])
g = lb.LBFGSBSolver(lb.LBFGSBParam()).minimize(lb.OBJ_ROSENBROCK_CHAINED, np.full(5000, 3.0), 2.0, 4.0)
print("lbfgsb", g["status"], g["niter"], g["nfev"], g["fx"])
