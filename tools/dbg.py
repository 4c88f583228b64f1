import sys; sys.path.insert(0,"/root/repo")
import numpy as np, visionworkbench_b200 as v, oracle
from visionworkbench_b200.synth import make_rasters
for (W,H),s,k in [((420,300),(33,17),(9,9)), ((420,300),(32,17),(9,9)), ((300,70),(33,8),(21,21)), ((300,70),(9,8),(7,7))]:
    l,r = make_rasters(W,H,s,k,seed=5)
    got = v.calc_disparity(0,l,r,s,k); ref = oracle.calc_disparity(0,l,r,s,k)
    bad = (got!=ref).any(-1)
    print(W,H,s,k,v.last_k1_stats()["path"],"bad",bad.sum(), "of", bad.size, "first bad", np.argwhere(bad)[:3].tolist(), got[bad][:2].tolist(), ref[bad][:2].tolist())
