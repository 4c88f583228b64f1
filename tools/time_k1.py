"""Quick device-resident timing of calc_disparity (CUDA events).  Usage: python tools/time_k1.py W H sx sy k [cost]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import visionworkbench_b200 as v
from visionworkbench_b200.synth import make_rasters
W, H, sx, sy, k = [int(a) for a in sys.argv[1:6]]
cost = int(sys.argv[6]) if len(sys.argv) > 6 else 0
left, right = make_rasters(W, H, (sx, sy), (k, k), seed=106)
dl, dr = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = v.calc_disparity(cost, dl, dr, (sx, sy), (k, k))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{W}x{H} search {sx}x{sy} k{k} cost{cost}: {ms:.2f} ms  {W*H/ms/1e3:.2f} Mpix/s  {W*H*sx*sy/ms/1e9:.3f} Teval/s  path={v.last_k1_stats()}")
