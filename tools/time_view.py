"""Time PyramidCorrelationView.rasterize over all tiles of a synthetic pair (BASELINE config 3 shape).
Usage: [THREADS=n] python tools/time_view.py SIZE TILE LEVELS COST KERNEL SEARCH [check]
THREADS > 1 rasterises tiles from a pool of host threads, as VW's block writer does (Image/ImageIO.h:228-235)."""
import sys, os, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import visionworkbench_b200 as v
from visionworkbench_b200.synth import make_pair
S, T, LV, cost, k, s = [int(a) for a in sys.argv[1:7]]
check = len(sys.argv) > 7
NT = int(os.environ.get("THREADS", "1"))
search = (-s // 2, -s // 2, s // 2, s // 2)
t0 = time.time()
left, right, lm, rm, _ = make_pair(S, S, search, 103, dropout=0.0 if os.environ.get("NODROP") else 0.03)
print(f"generated in {time.time()-t0:.1f}s")
dl, dr, dlm, drm = [torch.from_numpy(a).cuda() for a in (left, right, lm, rm)]
view = v.pyramid_correlate(dl, dr, dlm, drm, 0, 0.0, search, (k, k), cost, 0, 0.0, 2.0, 0, 5, LV)
out = torch.empty((S, S, 3), dtype=torch.float32, device="cuda")
streams = [torch.cuda.Stream() for _ in range(NT)]      # one stream per host thread, as bench.py's config 3 does
for it in range(3):
    n0 = v.kernel_launches()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tiles = [(x, y) for y in range(0, S, T) for x in range(0, S, T)]
    def work(i):
        with torch.cuda.stream(streams[i]):
            for x, y in tiles[i::NT]:
                view.rasterize(out[y:y+T, x:x+T], (x, y, min(S, x+T), min(S, y+T)))
    if NT > 1:
        with ThreadPoolExecutor(NT) as ex:
            list(ex.map(work, range(NT)))
    else:
        work(0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"view {S}x{S} tiles {T} levels {LV} cost {cost} k{k} search {s} threads {NT}: {dt*1e3:.1f} ms  {S*S/dt/1e6:.2f} Mpix/s  launches {v.kernel_launches()-n0}  valid {float((out[...,2]>0).float().mean()):.3f}")
if check:
    import oracle
    p = oracle.make_params(search, (k, k), cost=cost, consistency_threshold=2.0, filter_half_kernel=5, max_pyramid_levels=LV)
    bb = (T, T, min(S, 2*T), min(S, 2*T))
    t0 = time.time(); ref = oracle.pyramid_correlate(p, left, right, lm, rm, bbox=bb); dt = time.time() - t0
    got = out[bb[1]:bb[3], bb[0]:bb[2]].cpu().numpy()
    print(f"oracle tile {bb}: {dt:.1f}s ({(bb[2]-bb[0])*(bb[3]-bb[1])/dt/1e6:.3f} Mpix/s/thread); equal: {np.array_equal(got, ref)}; mismatches {(got!=ref).any(-1).sum()}")
