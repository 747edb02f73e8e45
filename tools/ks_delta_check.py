import importlib, os, sys
import numpy as np
from scipy import stats
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("metropolis-hastings-gpgpu_b200"); S = pkg.synth
from oracle_lib import Oracle
k, o = pkg.Kernel(), Oracle()
room = S.make_config(3)
pd, pf = [], []
for t in range(8):
    _, co = o.run(room, 4096, 260, seed=9000 + t)
    _, cd = k.wrapper_ex(room, 4096, 260, seed=100 + t, eval_mode=1)
    _, cf = k.wrapper_ex(room, 4096, 260, seed=100 + t, eval_mode=0)
    pd.append(stats.ks_2samp(cd["totalCosts"], co["totalCosts"]).pvalue)
    pf.append(stats.ks_2samp(cf["totalCosts"], co["totalCosts"]).pvalue)
    print(t, "delta %.3f full %.3f  mean delta %.3f full %.3f oracle %.3f" % (pd[-1], pf[-1], cd["totalCosts"].mean(), cf["totalCosts"].mean(), co["totalCosts"].mean()), flush=True)
print("delta p-values", np.round(pd, 3), "full", np.round(pf, 3))
