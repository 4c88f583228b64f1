import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import visionworkbench_b200 as v
import oracle
from visionworkbench_b200.synth import make_rasters
cases = [((300, 70), (16, 8), (21, 21), 12), ((300, 70), (16, 8), (21, 21), 8), ((300, 70), (8, 8), (21, 21), 12), ((300, 70), (16, 8), (7, 7), 12),
         ((100, 30), (16, 8), (21, 21), 12), ((100, 30), (8, 8), (7, 7), 12), ((100, 30), (8, 8), (7, 7), 10), ((64, 64), (8, 8), (3, 5), 12)]
import itertools
for ((W, H), s, k, bits), cost in itertools.product(cases, (2, 1)):
    l, r = make_rasters(W, H, s, k, seed=11 + W, bits=bits)
    got = np.asarray(v.calc_disparity(cost, l, r, s, k)); st = v.last_k1_stats()
    ref = np.asarray(oracle.calc_disparity(cost, l, r, s, k))
    bad = (got[..., 0] != ref[..., 0]) | (got[..., 1] != ref[..., 1]) | (got[..., 2] != ref[..., 2])
    print("cost", cost, (W, H), s, k, bits, st["path"], "bad", int(bad.sum()), "of", bad.size, "valid got/ref", int((got[..., 2] != 0).sum()), int((ref[..., 2] != 0).sum()))
    if bad.any():
        ys, xs = np.nonzero(bad)
        print("   rows", np.bincount(ys // 8)[:12], "cols", np.bincount(xs // 32)[:12])
        for y, x in list(zip(ys, xs))[:4]:
            print("   ", y, x, "got", got[y, x], "ref", ref[y, x])
