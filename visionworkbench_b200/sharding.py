"""Output-row band sharding of a correlation job over ranks.

The unit of the path is the output tile; tiles are independent, so a raster shards into contiguous
output-row bands, one per rank, with no reduction.  When the INPUT rasters are row-sharded the same way,
rank r must receive from rank r+1 the rows its last kernel windows (and search rows) reach into:
    left  raster: ky - 1          rows
    right raster: ky - 1 + sy - 1 rows
`plan()` returns the arithmetic (the C ABI's vwb200_shard_plan), `ShardComm` moves the halo rows on GPUs through the C
ABI (vwb200_shard_exchange_halos: ncclSend / ncclRecv in one NCCL group over NVLink), and `exchange_halos()` is the same
exchange over torch.distributed for host tensors (the gloo CPU tests of the band logic).
"""
import ctypes as C
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


class _CPlan(C.Structure):          # vwb200_band_plan (include/vwb200.h)
    _fields_ = [(n, C.c_int32) for n in ("rank", "world", "y0", "y1", "left_rows", "right_rows", "own_left", "own_right",
                                         "recv_left", "recv_right", "send_left", "send_right")]


def _cplan(p):
    return _CPlan(p.rank, p.world, p.y0, p.y1, p.left_rows, p.right_rows, p.own_left, p.own_right, p.recv_left, p.recv_right,
                  p.send_left, p.send_right)


def plan(rank, world, out_rows, ky, sy, left_total_rows=None, right_total_rows=None):
    """Band plan for `out_rows` output rows (vwb200_shard_plan).  The padded rasters have out_rows + ky - 1 (left) and
    out_rows + ky - 1 + sy - 1 (right) rows unless given."""
    from .api import lib
    c = _CPlan()
    f = lib().vwb200_shard_plan
    f.argtypes = [C.c_int] * 7 + [C.POINTER(_CPlan)]
    rc = f(rank, world, out_rows, ky, sy, left_total_rows or 0, right_total_rows or 0, C.byref(c))
    if rc:
        raise ValueError(lib().vwb200_last_error().decode())
    return BandPlan(*[getattr(c, n) for n, _ in _CPlan._fields_])


class ShardComm:
    """The NCCL communicator of the C ABI (vwb200_shard_create): rank 0 draws the unique id, `broadcast` hands the 128
    bytes to every rank (here: torch.distributed), every rank joins on its current CUDA device."""

    def __init__(self, rank, world, broadcast=None):
        from .api import lib, _check
        self._lib, self._check, self._h = lib(), _check, C.c_void_p()
        self.rank, self.world = rank, world
        ident = (C.c_ubyte * 128)()
        if world > 1:
            if rank == 0:
                _check(self._lib.vwb200_shard_unique_id(ident))
            data = bytes(ident)
            data = broadcast(data) if broadcast else data
            ident = (C.c_ubyte * 128).from_buffer_copy(data)
        self._lib.vwb200_shard_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        _check(self._lib.vwb200_shard_create(ident, rank, world, C.byref(self._h)))
        self._lib.vwb200_shard_exchange_halos.argtypes = [C.c_void_p, C.POINTER(_CPlan), C.c_void_p, C.c_int, C.c_ssize_t, C.c_void_p, C.c_int,
                                                          C.c_ssize_t, C.c_void_p]
        self._lib.vwb200_shard_destroy.argtypes = [C.c_void_p]
        self._lib.vwb200_shard_destroy.restype = None

    def exchange_halos(self, p, left_band, right_band, stream=None):
        """left_band / right_band: float32 CUDA tensors with p.left_rows / p.right_rows rows whose first own_* rows are
        valid; the halo rows arrive from rank + 1 (asynchronously, on the current torch stream)."""
        import torch
        assert left_band.dtype == torch.float32 and right_band.dtype == torch.float32 and left_band.is_cuda and right_band.is_cuda
        assert left_band.stride(1) == 1 and right_band.stride(1) == 1
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        cp = _cplan(p)
        self._check(self._lib.vwb200_shard_exchange_halos(self._h, C.byref(cp), left_band.data_ptr(), left_band.shape[1], left_band.stride(0),
                                                          right_band.data_ptr(), right_band.shape[1], right_band.stride(0), C.c_void_p(st)))
        return (p.recv_left * left_band.shape[1] + p.recv_right * right_band.shape[1]) * 4

    def close(self):
        if self._h:
            self._lib.vwb200_shard_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def torch_broadcast(data):
    """hand rank 0's bytes to every rank over the default torch.distributed group"""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(data), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


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


def view_rows(rank, world, rows, tile, levels, kernel, search_rows):
    """Tile-row split of a PyramidCorrelationView job (SURVEY 8e, config 3 over N GPUs): the tiles are the units, a rank takes
    a contiguous run of tile rows and needs NO data from its neighbours at run time if it holds, besides the rows of its tiles,
    the margin a tile reaches above and below: the kernel padding at the coarsest level, (kernel // 2) * 2^levels rows
    (CorrelationView.cc:89-97), plus the search rows.  Returns (y0, y1, top, bottom): the output rows [y0, y1) of this rank and
    the margins clipped to the raster."""
    ntr = (rows + tile - 1) // tile                       # tile rows of the job
    per, extra = divmod(ntr, world)
    t0 = rank * per + min(rank, extra)
    t1 = t0 + per + (1 if rank < extra else 0)
    y0, y1 = min(t0 * tile, rows), min(t1 * tile, rows)
    reach = (kernel // 2) * (1 << levels) + search_rows
    return y0, y1, min(reach, y0), min(reach, rows - y1)
