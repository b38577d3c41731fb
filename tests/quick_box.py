import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'oracle')
import lbfgspp_b200 as lb, pyoracle as po
ref=po.Oracle('ref')
for kind,n in ((0,10),(2,25),(0,200),(2,200),(0,100000)):
    x0=np.full(n,3.0)
    g=lb.LBFGSBSolver(lb.LBFGSBParam()).minimize(kind,x0,2.0,4.0)
    c=ref.lbfgsb(kind,x0,2.0,4.0,ref.default_param(lbfgsb=True))
    print(kind,n,'gpu',g['status'],g['msg'][:80],g['niter'],g['nfev'],g['fx'],'ref',c['niter'],c['nfev'],c['fx'], 'dx',np.max(np.abs(g['x']-c['x'])), 'sec', g['seconds'])
