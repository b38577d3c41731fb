import sys, time, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'oracle')
import lbfgspp_b200 as lb, pyoracle as po
orc=po.Oracle('orc')
for n in (10, 1000000, 10000000):
    prm=lb.LBFGSParam(m=10 if n>10 else 6)
    s=lb.LBFGSSolver(prm,'MoreThuente')
    for rep in range(2):
        g=s.minimize(lb.OBJ_ROSENBROCK_PAIRED,np.zeros(n),want_grad=False)
    print(n, g['status'], g['msg'], g['niter'], g['nfev'], g['fx'], 'solve s', g['seconds'], 'e2e', g['seconds_e2e'], 'launches', g['launches'], 'it/s', g['niter']/g['seconds'])
