"""Device-resident timing of calc_disparity_sgm.  Usage: python tools/time_sgm.py SIZE SEARCH KERNEL  (search box [0, SEARCH]^2)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import visionworkbench_b200 as v
S, s, k = [int(a) for a in sys.argv[1:4]]
rng = np.random.default_rng(104)
base = np.floor(rng.random((S + s + 8, S + s + 8)) * 256)
base = np.floor((base + np.roll(base, 1, 0) + np.roll(base, 1, 1)) / 3).astype(np.float32)
left = np.ascontiguousarray(base[4:4 + S, 4:4 + S]); right = np.ascontiguousarray(base[4 - s // 2:4 - s // 2 + S + s, 4 - s // 2:4 - s // 2 + S + s])
dl, dr = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
for it in range(3):
    n0 = v.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = v.calc_disparity_sgm(dl, dr, (s, s), k); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nd = (s + 1) ** 2
    ok = float(((out[..., 0] == s // 2) & (out[..., 1] == s // 2)).float().mean())
    print(f"sgm {S}x{S} search [0,{s}]^2 ({nd} disparities) census {k}: {ms:.1f} ms  {out.shape[0]*out.shape[1]/ms/1e3:.1f} Mpix/s  "
          f"{out.shape[0]*out.shape[1]*nd*41/ms/1e6:.0f} GB/s of the reference-equivalent 41 B/(px*d)  launches {v.kernel_launches()-n0}  correct {ok:.3f}")
