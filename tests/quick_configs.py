"""Timing of BASELINE configs 3 and 4 at full size (not part of pytest; numbers quoted in DESIGN.md)."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import lbfgspp_b200 as lb, pyoracle as po
orc = po.Oracle('orc')
# config 3: quadratic n=1e6, m=20, Bracketing
n = 1_000_000
d, b, xs = po.quad_tridiag_data(n, kappa=1e3, seed=0)
prm = lb.LBFGSParam(m=20)
s = lb.LBFGSSolver(prm, "Bracketing")
for rep in range(2):
    g = s.minimize(lb.OBJ_QUAD_TRIDIAG, np.zeros(n), data0=d, data1=b, want_grad=False)
t0 = time.time()
c = orc.lbfgs(po.OBJ_QUAD_TRIDIAG, np.zeros(n), po.LS_BRACKETING, orc.default_param(m=20), data0=d, data1=b, sum_mode=po.SUM_LANES8)
cpu_s = time.time() - t0
print("C3 gpu", g['status'], g['msg'][:50], g['niter'], g['nfev'], repr(g['fx']), "trace_last", repr(g['trace'][-1]), "solve_s", g['seconds'],
      "| cpu", c['status'], c['niter'], c['nfev'], "trace_last", repr(c['trace'][-1]), "wall_s", cpu_s,
      "| max|x-x*| gpu", np.max(np.abs(g['x'] - xs)), "cpu", np.max(np.abs(c['x'] - xs)))
# config 4: box [2,4], x0 = 3, n = 1e6, both objectives, default LBFGSBParam
for kind, name in ((lb.OBJ_ROSENBROCK_PAIRED, "paired"), (lb.OBJ_ROSENBROCK_CHAINED, "chained")):
    for eps_rel in (1e-5, 0.0):
        prm = lb.LBFGSBParam(epsilon_rel=eps_rel)
        for rep in range(2):
            g = lb.LBFGSBSolver(prm).minimize(kind, np.full(n, 3.0), 2.0, 4.0)
        print("C4", name, "eps_rel", eps_rel, g['status'], g['msg'][:60], g['niter'], g['nfev'], repr(g['fx']), "solve_s", g['seconds'],
              "launches", g['launches'])
ref = None
try:
    ref = po.Oracle('ref')
except Exception as e:
    print("no _ref:", e)
if ref is not None:
    n2 = 100_000
    for kind, name in ((po.OBJ_ROSENBROCK_PAIRED, "paired"), (po.OBJ_ROSENBROCK_CHAINED, "chained")):
        c = ref.lbfgsb(kind, np.full(n2, 3.0), 2.0, 4.0, ref.default_param(lbfgsb=True, epsilon_rel=0.0))
        g = lb.LBFGSBSolver(lb.LBFGSBParam(epsilon_rel=0.0)).minimize(kind, np.full(n2, 3.0), 2.0, 4.0)
        print("C4 n=1e5 eps_rel=0", name, "gpu", g['niter'], g['nfev'], repr(g['fx']), g['seconds'], "| ref(minieigen)", c['niter'], c['nfev'], repr(c['fx']), c['seconds'])
