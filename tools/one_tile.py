import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import visionworkbench_b200 as v
from visionworkbench_b200.synth import make_pair
search = (-64, -64, 64, 64)
left, right, lm, rm, _ = make_pair(2048, 2048, search, 103)
view = v.pyramid_correlate(left, right, lm, rm, 0, 0.0, search, (15, 15), 1, 0, 0.0, 2.0, 0, 5, 5)
for i in range(2):
    t0 = time.perf_counter(); o = view.rasterize(None, (0, 0, 1024, 1024)); print("tile ms", (time.perf_counter() - t0) * 1e3)
