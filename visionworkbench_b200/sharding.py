"""Output-row band sharding of a correlation job over ranks.

The unit of the path is the output tile; tiles are independent, so a raster shards into contiguous
output-row bands, one per rank, with no reduction.  When the INPUT rasters are row-sharded the same way,
rank r must receive from rank r+1 the rows its last kernel windows (and search rows) reach into:
    left  raster: ky - 1          rows
    right raster: ky - 1 + sy - 1 rows
`plan()` returns the arithmetic, `exchange_halos()` performs the send/recv with torch.distributed
(NCCL on GPUs, gloo in the CPU tests).
"""
from dataclasses import dataclass


@dataclass
class BandPlan:
    rank: int
    world: int
    y0: int            # first output row of this rank
    y1: int            # one past the last output row
    left_rows: int     # rows of the left raster this rank's launch reads  (H + ky - 1)
    right_rows: int    # rows of the right raster                          (H + ky - 1 + sy - 1)
    own_left: int      # rows resident on this rank before the exchange
    own_right: int
    recv_left: int     # halo rows to receive from rank + 1
    recv_right: int
    send_left: int     # rows to send to rank - 1 (its halo)
    send_right: int


def plan(rank, world, out_rows, ky, sy, left_total_rows=None, right_total_rows=None):
    """Band plan for `out_rows` output rows.  The padded rasters have out_rows + ky - 1 (left) and
    out_rows + ky - 1 + sy - 1 (right) rows unless given."""
    if left_total_rows is None:
        left_total_rows = out_rows + ky - 1
    if right_total_rows is None:
        right_total_rows = out_rows + ky - 1 + sy - 1
    band = (out_rows + world - 1) // world
    y0, y1 = min(out_rows, rank * band), min(out_rows, (rank + 1) * band)
    h = y1 - y0
    left_rows, right_rows = h + ky - 1, h + ky - 1 + sy - 1
    last = rank == world - 1 or y1 >= out_rows
    own_left = min(left_rows, (left_total_rows - y0) if last else h)
    own_right = min(right_rows, (right_total_rows - y0) if last else h)
    recv_left, recv_right = left_rows - own_left, right_rows - own_right
    send_left = (ky - 1) if rank > 0 and h > 0 else 0
    send_right = (ky - 1 + sy - 1) if rank > 0 and h > 0 else 0
    if band < ky - 1 + sy - 1 and world > 1:
        raise ValueError("band height smaller than the halo: a rank would need rows from beyond its neighbour")
    return BandPlan(rank, world, y0, y1, left_rows, right_rows, own_left, own_right, recv_left, recv_right, send_left, send_right)


def exchange_halos(p, left_band, right_band):
    """left_band / right_band: 2-D tensors with p.left_rows / p.right_rows rows whose first own_* rows are
    valid.  Fills the halo rows from rank + 1 and serves rank - 1.  Returns bytes received."""
    import torch.distributed as dist
    if p.world == 1:
        return 0
    ops = []
    if p.recv_left:
        ops.append(dist.P2POp(dist.irecv, left_band[p.own_left:p.own_left + p.recv_left], p.rank + 1))
    if p.recv_right:
        ops.append(dist.P2POp(dist.irecv, right_band[p.own_right:p.own_right + p.recv_right], p.rank + 1))
    if p.send_left:
        ops.append(dist.P2POp(dist.isend, left_band[:p.send_left], p.rank - 1))
    if p.send_right:
        ops.append(dist.P2POp(dist.isend, right_band[:p.send_right], p.rank - 1))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return (p.recv_left * left_band.shape[1] + p.recv_right * right_band.shape[1]) * left_band.element_size()
