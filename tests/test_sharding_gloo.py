"""world_size-2 (and 4) gloo tests of the band/halo logic bench.py uses for N > 1 GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visionworkbench_b200 import sharding  # noqa: E402


def test_plan_covers_every_output_row_once():
    for world in (1, 2, 4, 8):
        rows = []
        for r in range(world):
            p = sharding.plan(r, world, 8192, 21, 128)
            rows += list(range(p.y0, p.y1))
            assert p.left_rows == p.own_left + p.recv_left and p.right_rows == p.own_right + p.recv_right
            if r == world - 1:
                assert p.recv_left == 0 and p.recv_right == 0
            else:
                assert p.recv_left == 20 and p.recv_right == 20 + 127
            assert p.send_left == (20 if r else 0)
        assert rows == list(range(8192))
    with pytest.raises(ValueError):
        sharding.plan(0, 8, 512, 21, 128)       # 64-row bands cannot hold a 147-row halo


def _worker(rank, world, port, out_rows, ky, sy, width, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    left = rng.random((out_rows + ky - 1, width + ky - 1)).astype(np.float32)
    right = rng.random((out_rows + ky - 1 + sy - 1, width + ky - 1 + sy - 1)).astype(np.float32)
    p = sharding.plan(rank, world, out_rows, ky, sy)
    lb = torch.zeros((p.left_rows, left.shape[1]))
    rb = torch.zeros((p.right_rows, right.shape[1]))
    lb[:p.own_left] = torch.from_numpy(left[p.y0:p.y0 + p.own_left])
    rb[:p.own_right] = torch.from_numpy(right[p.y0:p.y0 + p.own_right])
    nbytes = sharding.exchange_halos(p, lb, rb)
    ok = np.array_equal(lb.numpy(), left[p.y0:p.y0 + p.left_rows]) and np.array_equal(rb.numpy(), right[p.y0:p.y0 + p.right_rows])
    q.put((rank, ok, nbytes))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_reassembles_the_rasters(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000 + world
    out_rows, ky, sy, width = 96 * world, 7, 9, 40
    procs = [ctx.Process(target=_worker, args=(r, world, port, out_rows, ky, sy, width, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(ok for _, ok, _ in res), res
    by_rank = {r: n for r, _, n in res}
    assert by_rank[world - 1] == 0
    assert by_rank[0] == ((ky - 1) * (width + ky - 1) + (ky - 1 + sy - 1) * (width + ky - 1 + sy - 1)) * 4


def test_c_abi_plan_and_single_rank_comm():
    """vwb200_shard_plan (the arithmetic behind sharding.plan) against the closed form, and the world-size-1 communicator:
    no NCCL, no device, exchange is a no-op (the N > 1 NCCL path is exercised by bench.py on GPUs)"""
    import ctypes as C
    for world in (1, 2, 3, 8):
        for out_rows, ky, sy in ((8192, 21, 128), (16384, 15, 256), (1000, 7, 9)):
            covered = []
            for r in range(world):
                p = sharding.plan(r, world, out_rows, ky, sy)
                band = (out_rows + world - 1) // world
                assert (p.y0, p.y1) == (min(out_rows, r * band), min(out_rows, (r + 1) * band))
                h = p.y1 - p.y0
                assert p.left_rows == h + ky - 1 and p.right_rows == h + ky - 1 + sy - 1
                last = r == world - 1 or p.y1 >= out_rows
                assert p.recv_left == (0 if last else ky - 1) and p.recv_right == (0 if last else ky - 1 + sy - 1)
                assert p.send_left == (ky - 1 if r and h else 0) and p.send_right == (ky - 1 + sy - 1 if r and h else 0)
                covered += list(range(p.y0, p.y1))
            assert covered == list(range(out_rows))
    comm = sharding.ShardComm(0, 1)
    p = sharding.plan(0, 1, 64, 5, 7)
    from visionworkbench_b200 import api
    cp = sharding._cplan(p)
    buf = (C.c_float * 16)()
    assert api.lib().vwb200_shard_exchange_halos(comm._h, C.byref(cp), buf, 4, 4, buf, 4, 4, None) == 0
    comm.close()


def test_view_tile_rows_cover_the_raster_and_margins_reach_far_enough():
    """config 3 over N GPUs (bench.py cfg3_sharded): tile rows split over the ranks with no collective"""
    from visionworkbench_b200 import sharding
    for rows, tile, world in [(8192, 1024, 1), (8192, 1024, 2), (8192, 1024, 8), (8192, 1024, 3), (5000, 1024, 4), (1024, 1024, 4)]:
        cover = []
        for r in range(world):
            y0, y1, top, bot = sharding.view_rows(r, world, rows, tile, levels=5, kernel=15, search_rows=64)
            assert 0 <= y0 <= y1 <= rows
            assert y0 % tile == 0 and (y1 % tile == 0 or y1 == rows)
            # the rows a tile of this rank reads: bbox grown by half_kernel * 2^levels (pyramid padding) and the search rows
            assert top == min(7 * 32 + 64, y0) and bot == min(7 * 32 + 64, rows - y1)
            cover.append((y0, y1))
        assert cover[0][0] == 0 and cover[-1][1] == rows
        assert all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
        sizes = [(b - a + tile - 1) // tile for a, b in cover]
        assert max(sizes) - min(sizes) <= 1                   # balanced to one tile row
