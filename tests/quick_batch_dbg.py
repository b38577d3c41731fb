import sys, numpy as np
sys.path.insert(0,'.')
import lbfgspp_b200 as lb
B,n=6,4096
X0=np.stack([np.random.default_rng(1000+b).uniform(-1,1,n) for b in range(B)])
prm=lb.LBFGSParam(m=10)
for th in (1,2,4):
    res,X,_=lb.solve_batch(lb.OBJ_ROSENBROCK_PAIRED,X0,prm,"MoreThuente",threads=th)
    print(th,[ (r['status'],r['niter']) for r in res])
import time
# n = 1e6 single solves: resident vs host-driven
n=1_000_000
for resident in (False, True):
    s=lb.Session(lb.OBJ_ROSENBROCK_PAIRED, np.zeros(n), prm, "MoreThuente", resident=resident)
    for _ in range(3): r=s.solve()
    t0=time.perf_counter()
    for _ in range(20): r=s.solve()
    dt=(time.perf_counter()-t0)/20
    print('n=1e6 resident',resident,r['niter'],r['nfev'],'ms/solve',dt*1e3,'it/s',r['niter']/dt)
    s.close()
