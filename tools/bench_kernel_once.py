"""One device-resident calc_disparity at the bench workload (for ncu captures of the dominant kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import visionworkbench_b200 as v
from visionworkbench_b200.synth import make_rasters
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
left, right = make_rasters(S, S, (128, 128), (21, 21), seed=106)
dl, dr = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
for _ in range(2):
    out = v.calc_disparity(0, dl, dr, (128, 128), (21, 21))
torch.cuda.synchronize()
print(v.last_k1_stats())
