"""Multi-GPU row bands: the image shards into contiguous row bands, one per rank.

Pixels are independent and the only shared input is the (read-only) Gaussian array, so the data path needs exactly
two exchanges per scene / frame and no halo:
    * one broadcast of the scene (n x 40 B) from rank 0 when it changes          -- torch.distributed.broadcast
    * one gather of the finished u32 bands to rank 0 per frame                   -- grouped send/recv (bands differ in size)
One process per GPU; NCCL over NVLink on the GPU box, gloo in the CPU tests.  Band boundaries come from the per-row
cost K1 reports (sum of pixels * 5 * n^2 per cell row), split by vrt_host_row_bands so the slowest rank's share is minimal.
"""
import ctypes

import numpy as np

from . import _ffi


def split_rows(row_cost, row_px, n_parts, height, align):
    """Work-balanced band boundaries in PIXEL rows.

    row_cost[i] is the cost of pixel rows [i*row_px, (i+1)*row_px); boundaries are multiples of `align` pixels (a
    multiple of row_px: tile rows keep every cell inside one band).  Returns n_parts+1 ascending pixel rows."""
    row_cost = np.asarray(row_cost, np.float64)
    if align % row_px:
        raise ValueError("align must be a multiple of the cost row height")
    group = align // row_px
    n_groups = (len(row_cost) + group - 1) // group
    cost = np.zeros(n_groups, np.float64)
    for g in range(n_groups):
        cost[g] = row_cost[g * group : (g + 1) * group].sum()
    # a small per-row constant keeps empty regions from collapsing into zero-height bands
    cost = cost + max(cost.sum(), 1.0) * 1e-6
    out = np.zeros(n_parts + 1, np.uint32)
    rc = _ffi.host_lib().vrt_host_row_bands(cost.ctypes.data_as(ctypes.c_void_p), n_groups, n_parts, out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("vrt_host_row_bands failed")
    bounds = [min(int(b) * align, height) for b in out]
    bounds[-1] = height
    return bounds


def balanced_bands(renderer, n_parts, height, align):
    """Bands from the row costs of the renderer's last full-frame vrt_cuda_tile()."""
    rows, row_px = renderer.row_costs()
    return split_rows(rows, row_px, n_parts, height, align)


def rebalance(row_cost, row_px, bounds, measured_ms, height, align):
    """Feedback step of the band balance: the n^2 cost model is only proportional to time up to a factor that varies over the
    image (list lengths change the share of per-occluder setup, K1 is not in the model), so after a frame every rank reports
    its device time and each band's rows are re-weighted by that band's measured time per unit of modelled cost; the bands are
    then split again.  An orbiting camera changes the picture slowly, so the previous frame is a good predictor
    (one or two steps bring the slowest rank within a few percent of the mean).  Returns the new n_parts+1 pixel rows."""
    row_cost = np.asarray(row_cost, np.float64)
    weighted = row_cost.copy()
    floor = max(row_cost.sum(), 1.0) * 1e-6  # empty rows still cost their band something
    for r, ms in enumerate(measured_ms):
        a, b = bounds[r] // row_px, (bounds[r + 1] + row_px - 1) // row_px
        if b <= a:
            continue
        modelled = row_cost[a:b].sum() + floor * (b - a)
        weighted[a:b] = (row_cost[a:b] + floor) * (float(ms) / modelled)
    return split_rows(weighted, row_px, len(measured_ms), height, align)


def gather_bands(image, bounds, rank, world, dist):
    """Rank r owns rows [bounds[r], bounds[r+1]) of `image` (a [H, W] tensor on every rank); after the call rank 0
    holds every band.  One grouped batch of point-to-point transfers (ncclSend/ncclRecv under NCCL)."""
    ops = []
    if rank == 0:
        for src in range(1, world):
            if bounds[src + 1] > bounds[src]:
                ops.append(dist.P2POp(dist.irecv, image[bounds[src] : bounds[src + 1]], src))
    elif bounds[rank + 1] > bounds[rank]:
        ops.append(dist.P2POp(dist.isend, image[bounds[rank] : bounds[rank + 1]], 0))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class PeerImage:
    """The frame buffer of a multi-GPU frame: ONE image on rank 0's GPU that every rank's render kernel stores its row band
    into directly (vrt_cuda_peer_image_create / _open: CUDA IPC, peer stores over NVLink / NVSwitch), so the exchange step of a
    frame shrinks from a gather of the bands to one barrier (`complete`).  `ok` is False on every rank when any rank could not
    map the image (no peer access between the GPUs, or a CPU test backend): the caller then keeps the gather path."""

    def __init__(self, renderer, height, width, rank, world, dist, torch):
        self.renderer, self.rank, self.dist, self.torch = renderer, rank, dist, torch
        self.ptr, self.ok, self.why = 0, True, ""
        handle = [None]
        if rank == 0:
            try:
                self.ptr, handle[0] = renderer.peer_image_create(height * width * 4)
            except RuntimeError as exc:
                self.ok, self.why = False, str(exc)
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            try:
                if handle[0] is None:
                    raise RuntimeError("the owner could not create the image")
                self.ptr = renderer.peer_image_open(handle[0])
            except RuntimeError as exc:
                self.ok, self.why = False, str(exc)
        dev = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        flag = torch.tensor([1.0 if self.ok else 0.0], dtype=torch.float32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        everyone = bool(flag.item() > 0.5)
        if not everyone and self.ptr:
            renderer.peer_image_close(self.ptr)
            self.ptr = 0
        self.ok = everyone
        self._token = torch.zeros(1, dtype=torch.float32, device=dev)
        self.tensor = None
        if self.ok and rank == 0 and dev.type == "cuda":
            class _Raw:  # rank 0's view of its own allocation as a [H, W] int32 tensor
                __cuda_array_interface__ = {"shape": (height, width), "typestr": "<i4", "data": (self.ptr, False), "version": 2}

            self.tensor = torch.as_tensor(_Raw(), device=dev)

    def complete(self):
        """Enqueue the frame's barrier on the current stream: once it has passed on rank 0, every rank's render kernel has
        finished and its stores have landed in rank 0's memory."""
        self.dist.all_reduce(self._token)

    def close(self):
        self.tensor = None
        if self.ptr:
            self.renderer.peer_image_close(self.ptr)
            self.ptr = 0
