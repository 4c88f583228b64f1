#!/usr/bin/env python
"""bench.py -- disparity Mpix/s of the dense block-matching hot path (BASELINE.json metric).

Workload (default): vw::stereo::calc_disparity / best_of_search_convolution over a synthetic
8192x8192 stereo pair, 128x128 search window, 21x21 kernel, AbsoluteCost (--cost sq|ncc for the
others), integer-valued 12-bit imagery.  A step = one full-image pass.

  python bench.py --gpus N --steps K --warmup W          our arm (torchrun for N > 1)
  python bench.py --impl reference ...                   the reference algorithm on the host cores

N > 1: the image is sharded into contiguous output-row bands, one per rank (strong scaling).  Inputs
are row-sharded too: every step each rank receives the halo rows it needs (kernel-1 rows of the left
raster, kernel-1 + search-1 rows of the right raster) from the next rank with NCCL send/recv over
NVLink, inside the timed region.  No other collective exists on this path.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

COSTS = {"abs": 0, "sq": 1, "ncc": 2}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cost", default="abs", choices=list(COSTS))
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--search", type=int, default=128)
    ap.add_argument("--kernel", type=int, default=21)
    ap.add_argument("--seed", type=int, default=106)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"calc_disparity single level {a.size}x{a.size}, search {a.search}x{a.search}, kernel {a.kernel}x{a.kernel}, "
            f"cost {a.cost.upper()}, synthetic 12-bit integer-valued pair (seed {a.seed})")


def gen_rasters(a):
    from visionworkbench_b200.synth import make_rasters
    return make_rasters(a.size, a.size, (a.search, a.search), (a.kernel, a.kernel), seed=a.seed)


# ------------------------------------------------------------------------------------------------
# CPU side: the reference algorithm (oracle restatement) on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_sample(a, left, right, nthreads):
    """One 128x128-output tile per thread with the full search window (the reference parallelises over
    independent tiles, Image/ImageIO.h:289-311).  Returns (pixels, seconds)."""
    import oracle
    k, s, t = a.kernel, a.search, 128
    per_row = max(1, min(nthreads, a.size // t))
    rows = (nthreads + per_row - 1) // per_row
    Wc, Hc = per_row * t, rows * t
    l = np.ascontiguousarray(left[:Hc + k - 1, :Wc + k - 1])
    r = np.ascontiguousarray(right[:Hc + k - 1 + s - 1, :Wc + k - 1 + s - 1])
    t0 = time.perf_counter()
    oracle.calc_disparity_tiled(COSTS[a.cost], l, r, (s, s), (k, k), tile=t, nthreads=nthreads)
    return Wc * Hc, time.perf_counter() - t0


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    nthreads = oracle.max_threads()
    left, right = gen_rasters(a)
    for _ in range(min(a.warmup, 1)):
        cpu_sample(a, left, right, nthreads)
    pix = 0
    t = 0.0
    for _ in range(a.steps):
        p, dt = cpu_sample(a, left, right, nthreads)
        pix += p
        t += dt
    v = pix / t / 1e6
    line = {
        "impl": "reference", "metric": "disparity Mpix/s", "value": v, "unit": "Mpix/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 cost / f64 sums", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "reference algorithm (CPU restatement of best_of_search_convolution; the "
                   "reference itself cannot be compiled here: no Boost/GDAL headers), tile-parallel like block_write_image"},
        "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": nthreads, "kind": "port",
                         "sample": f"per step: {nthreads} tiles of 128x128 output pixels, full {a.search}x{a.search} window, one tile per thread"},
        "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.split(",") for ln in open(self.f.name).read().strip().splitlines() if ln.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][2])
        out["power_w_max"] = max(float(r[3]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, n in enumerate(names):
            if any(r[5 + i].strip().lower().startswith("active") for r in rows):
                out["reasons"].append(n)
        out["samples"] = len(rows)
        return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    import visionworkbench_b200 as v

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert v.device_count() > 0
    cost = COSTS[a.cost]
    k, s, S = a.kernel, a.search, a.size
    left, right = gen_rasters(a)                        # same seed on every rank
    # ---- shard output rows into bands; each rank OWNS the input rows of its band (visionworkbench_b200/sharding.py) ----
    from visionworkbench_b200 import sharding
    p = sharding.plan(rank, world, S, k, s, left.shape[0], right.shape[0])
    y0, y1, H = p.y0, p.y1, p.y1 - p.y0
    lh_need, rh_need = p.left_rows, p.right_rows          # rows this rank's kernel launch reads
    dl = torch.empty((lh_need, left.shape[1]), dtype=torch.float32, device="cuda")
    dr = torch.empty((rh_need, right.shape[1]), dtype=torch.float32, device="cuda")
    dl[:p.own_left].copy_(torch.from_numpy(left[y0:y0 + p.own_left]))
    dr[:p.own_right].copy_(torch.from_numpy(right[y0:y0 + p.own_right]))

    def halo_exchange():
        return sharding.exchange_halos(p, dl, dr)

    def step_device():
        halo_exchange()
        return v.calc_disparity(cost, dl, dr, (s, s), (k, k))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- host-resident copies for the e2e leg (pinned) ----
    hl = torch.empty((lh_need, left.shape[1]), dtype=torch.float32, pin_memory=True)
    hr = torch.empty((rh_need, right.shape[1]), dtype=torch.float32, pin_memory=True)
    hl.copy_(torch.from_numpy(left[y0:y0 + lh_need]))
    hr.copy_(torch.from_numpy(right[y0:y0 + rh_need]))
    hout = torch.empty((H, S, 3), dtype=torch.int32, pin_memory=True)
    hl_np, hr_np, hout_np = hl.numpy(), hr.numpy(), hout.numpy()
    L = v.lib()

    def step_e2e():
        rc = L.vwb200_calc_disparity(cost, hl_np.ctypes.data, hl_np.shape[1], hl_np.shape[0], hl_np.shape[1],
                                     hr_np.ctypes.data, hr_np.shape[1], hr_np.shape[0], hr_np.shape[1],
                                     s, s, k, k, hout_np.ctypes.data, S, 0, None)
        if rc:
            raise RuntimeError(L.vwb200_last_error().decode())

    # ---- warm-up ----
    for _ in range(a.warmup):
        out = step_device()
    path = v.last_k1_stats()["path"]
    # ---- timed: device-resident ----
    clocks = Clocks(local)
    barrier()
    if rank == 0:
        clocks.start()
    n0 = v.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    e0.record()
    for _ in range(a.steps):
        out = step_device()
        kernel_ms.append(v.last_k1_stats()["kernel_ms"])
    e1.record()
    barrier()
    launches = v.kernel_launches() - n0
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    nl = torch.tensor([float(launches)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    value = S * S * a.steps / (ms * 1e-3) / 1e6
    # ---- timed: end to end through the C ABI with host buffers ----
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    torch.cuda.synchronize()
    te = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([te], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    te = float(t.item())
    e2e = S * S * a.steps / (te * 1e-3) / 1e6
    # e2e result must equal the device-resident result
    same = bool(np.array_equal(hout_np, out.cpu().numpy()))
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, which = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
        kms = float(np.mean(kernel_ms))
        alg_bytes = H * S * (4 + 4 + 12)                 # left + right + 12-byte disparity pixel
        achieved = alg_bytes / (kms * 1e-3) / 1e9
        evals = H * S * s * s
        kname = "k1_generic_kernel" if path != "exact-int" else ("k1_fast_abs_kernel" if a.cost == "abs" else "k1_screen_kernel")
        traffic = None
        try:          # measured once under ncu for the default workload; null for any other
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json"))).get(f"{kname}|{S}|{s}|{a.kernel}|{a.cost}")
            if tr and world == 1:
                traffic = tr["dram_read_bytes"] + tr["dram_write_bytes"]
        except Exception:
            pass
        line = {
            "metric": "disparity Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32 on u16 (exact)" if path == "exact-int" else "f32 cost / f64 sums", "data": "synthetic",
            "config": {"workload": workload_name(a), "kernel_path": path, "l2": "inputs larger than L2 (2 x 270 MB rasters, 805 MB output)",
                       "parallelism": f"{world} output-row band(s), NCCL send/recv halo rows" if world > 1 else "1 GPU, persistent CTAs",
                       "e2e_equals_device_result": same},
            "e2e": {"value": e2e, "unit": "Mpix/s", "h2d_bytes_per_step": int(hl_np.nbytes + hr_np.nbytes) * world if world == 1 else int(hl_np.nbytes + hr_np.nbytes),
                    "d2h_bytes_per_step": int(hout_np.nbytes), "ms_per_step": te / a.steps, "bytes_are": "per rank"},
            "gpu_launches": int(nl.item()),
            "clocks": clk,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu, profiles/traffic_r01.json)",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "peak_is": which, "kernel": kname,
                         "kernel_ms": kms, "kernel_share_of_step": kms * a.steps / ms if world == 1 else None,
                         "note": "ALU/issue-bound by construction (SURVEY 8d): ~10 issue slots per pixel*disparity vs 20 B per pixel",
                         "alu": {"achieved_Teval_s": evals / (kms * 1e-3) / 1e12, "evals_per_launch": evals}},
        }
        if world == 1 and not a.no_cpu_baseline:
            import oracle
            oracle.build()
            nthreads = oracle.max_threads()
            pix, dt = 0, 0.0
            while dt < 10.0:
                p, d = cpu_sample(a, left, right, nthreads)
                pix += p
                dt += d
            line["cpu_baseline"] = {"value": pix / dt / 1e6, "unit": "Mpix/s", "cores": nthreads, "kind": "port",
                                    "sample": f"{pix} output pixels in {nthreads}-tile batches of 128x128, full {s}x{s} window, {dt:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
