import os, sys
import numpy as np
ROOT="/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"tests"))
import __graft_entry__ as ge
pkg=ge.package()
t=pkg.RayMarcherExample(0)
t.InitGrid(64); t.fill_synthetic(20.0,0); t.zeroGrad()
t.kernel2D_TVGrad(); t.sync()
a=t.get_grad()[:,0].copy()
t.set_option("tv_sweep",0); t.zeroGrad(); t.kernel2D_TVGrad(); t.sync()
b=t.get_grad()[:,0]
print("maxdiff", np.abs(a-b).max(), np.abs(b).max())
