"""Device-resident timing of BASELINE config 4: SemiGlobalMatcher, census 5, 8 paths, SIZE^2 pair, global window
[0, SEARCH]^2, per-pixel boxes from a synthetic half-resolution prior with search_buffer (2, 2).
Usage: python tools/time_sgm4.py [SIZE=4096] [SEARCH=128] [mgm]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import visionworkbench_b200 as v
from visionworkbench_b200 import api
from visionworkbench_b200.synth import make_sgm_case

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
s = int(sys.argv[2]) if len(sys.argv) > 2 else 128
mgm = "mgm" in sys.argv
left, right, prev, true = make_sgm_case(S, s, 5, seed=104)
dl, dr = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
dp = torch.from_numpy(prev).cuda()
sp = api._sgm_params((s, s), 5, api.CENSUS_TRANSFORM, mgm, api.SUBPIXEL_LC_BLEND, (2, 2), 1e9)
ow, oh = C.c_int(0), C.c_int(0)
L = v.lib()
head = (C.byref(sp), dl.data_ptr(), dl.shape[1], dl.shape[0], dl.stride(0), dr.data_ptr(), dr.shape[1], dr.shape[0], dr.stride(0))
api._check(L.vwb200_sgm_calc_disparity_ex(*head, None, 0, None, 0, 0, 0, None, 0, 0, 0, None, None, 0, None, 0, None, C.byref(ow), C.byref(oh), 1, None))
out = torch.empty((oh.value, ow.value, 3), dtype=torch.int32, device="cuda")
sub = torch.empty((oh.value, ow.value, 3), dtype=torch.float32, device="cuda")
for it in range(4):
    n0 = v.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    api._check(L.vwb200_sgm_calc_disparity_ex(*head, None, 0, None, 0, 0, 0, dp.data_ptr(), dp.shape[1], dp.shape[0], dp.stride(0) // 3, None,
                                              out.data_ptr(), ow.value, sub.data_ptr(), ow.value, None, C.byref(ow), C.byref(oh), 1,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    o = out.cpu().numpy()
    hk = 2
    t = true[hk:hk + oh.value, hk:hk + ow.value]
    ok = float(((o[..., 0] == t[..., 0]) & (o[..., 1] == t[..., 1]) & (o[..., 2] == 1)).mean())
    npx = oh.value * ow.value
    print(f"cfg4 {'MGM' if mgm else 'SGM'} {S}^2 window [0,{s}]^2 boxes 5x5: {ms:.2f} ms  {npx / ms / 1e3:.0f} Mpix/s  "
          f"{npx * 25 * 41 / ms / 1e6:.0f} GB/s of the reference-equivalent 1025 B/px  launches {v.kernel_launches() - n0}  correct {ok:.4f}")
