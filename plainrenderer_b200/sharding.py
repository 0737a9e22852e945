"""Row-sharded frames across GPUs (DESIGN.md "Multi-GPU").

One process per GPU. Each rank owns a band of screen rows (multiples of 32) and drives its frontend segment by segment
(`Frontend.begin_frame` / `run_segment`); between segments the ranks exchange what the next passes read across band
boundaries, over `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests):

  * ALLREDUCE_SUM_U32  the 128 luminance-histogram counters
  * ALLGATHER_ROWS     images read at arbitrary rows by a later pass (HiZ level 3, half-res depth, the GI buffers around
                       the world-space spatial filter, TAA history, bloom mip 1)
  * HALO_ROWS          a few rows on each side of the band boundary

Stencils with a small fixed footprint (shading -> TAA -> bloom mip 1, upscale, bloom upsample mip 0) use overlapped
computation instead of an exchange: the frontend extends the row window of the producing pass.
The exchanges write straight into the images' device memory (tensors aliasing the pointers of the C-ABI).

Two transports:
  * `DistComm(..., peer=False)`: every exchange is a batch of NCCL send/recv issued from Python between two segments
  * `DistComm(..., peer=True)` (default on GPUs): **peer exchange over NVLink** (include/plain_b200.h). The ranks map each other's
    images with CUDA IPC (handles shipped once with torch.distributed); from then on `run_segment` performs the exchanges on
    the device - a kernel stores the rank's rows into the peers' copies of the image, a flag barrier in peer memory orders
    it - and a whole frame is enqueued by one call without a host round trip. Images are mapped lazily: the first time an
    exchange names an unmapped image it is returned to Python, which maps it on all ranks and performs that one exchange
    over NCCL.
"""
import ctypes as C

import numpy as np

from . import ffi


def shard_band(api, full_height, count, rank, divisor, rows):
    a, b = ffi.u32(), ffi.u32()
    api.f["shard_band"](ffi.u32(full_height), ffi.u32(count), ffi.u32(rank), ffi.u32(divisor), ffi.u32(rows), C.byref(a), C.byref(b))
    return a.value, b.value


def full_res_band(api, full_height, count, rank):
    return shard_band(api, full_height, count, rank, 1, full_height)


class _CudaMem:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def alias_tensor(ptr, nbytes, device):
    """uint8 torch tensor over `nbytes` at `ptr` (device memory of the CUDA backend, or host memory of the CPU oracle)."""
    import torch
    if device is None or str(device) == "cpu":
        return torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr)))
    return torch.as_tensor(_CudaMem(ptr, nbytes), device=device)


def exchange_views(x, device, cache=None):
    """[(tensor(rows, slices, pitch) uint8, rows, divisor)] for the images of a plain_exchange. Rows come first so that
    t[r0:r1] is a row range in every slice of a 3-D level (contiguous for 2-D images, strided for 3-D ones)."""
    out = []
    for i in range(x.n_images):
        slices = max(int(x.depth[i]), 1)
        key = (x.device_ptr[i], x.rows[i], x.row_pitch_bytes[i], slices)
        t = cache.get(key) if cache is not None else None
        if t is None:
            t = alias_tensor(x.device_ptr[i], slices * x.rows[i] * x.row_pitch_bytes[i], device).view(slices, x.rows[i], x.row_pitch_bytes[i]).permute(1, 0, 2)
            if cache is not None:
                cache[key] = t
        out.append((t, x.rows[i], x.row_divisor[i]))
    return out


def plan_row_exchange(kind, halo, bands, me):
    """Pure planning (unit-tested on CPU): what rank `me` sends and receives for one image.
    bands: [(a, b)] per rank. Returns (sends [(peer, r0, r1)], recvs [(peer, r0, r1)]) in rows of the image."""
    sends, recvs = [], []
    a, b = bands[me]
    n = len(bands)
    if kind == ffi.EXCHANGE_ALLGATHER_ROWS:
        for p in range(n):
            if p == me:
                continue
            if b > a:
                sends.append((p, a, b))
            pa, pb = bands[p]
            if pb > pa:
                recvs.append((p, pa, pb))
    elif kind == ffi.EXCHANGE_HALO_ROWS:
        for p in (me - 1, me + 1):
            if p < 0 or p >= n:
                continue
            pa, pb = bands[p]
            if p < me:   # neighbour above: it needs my first rows, I need its last rows
                s0, s1 = a, min(a + halo, b)
                r0, r1 = max(pb - halo, pa), pb
            else:        # neighbour below
                s0, s1 = max(b - halo, a), b
                r0, r1 = pa, min(pa + halo, pb)
            if s1 > s0:
                sends.append((p, s0, s1))
            if r1 > r0:
                recvs.append((p, r0, r1))
    return sends, recvs


class DistComm:
    """Exchanges over torch.distributed (one rank per process)."""

    def __init__(self, api, full_height, device=None, stream=None, frontend=None, peer=False):
        import torch.distributed as dist
        self.api, self.H, self.device, self.stream = api, full_height, device, stream
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.cache = {}
        self.bytes_sent = 0
        self.fe, self.peer = frontend, False
        self.python_exchanges = 0  # exchanges that went through Python (all of them without peer exchange)
        if peer and self.world > 1:
            self._peer_setup()

    # ---- peer exchange over NVLink: CUDA IPC set-up (collective) ----
    def _all_gather_handles(self, handle_bytes):
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, bytes(handle_bytes))
        return out

    def _peer_setup(self):
        be = self.fe.backend
        be._check(self.api.b["peer_init"](be.ctx, ffi.u32(self.rank), ffi.u32(self.world)), "peer_init")
        h = (C.c_uint8 * 64)()
        be._check(self.api.b["peer_get_sync_handle"](be.ctx, h), "peer_get_sync_handle")
        for r, hb in enumerate(self._all_gather_handles(h)):
            if r != self.rank:
                buf = (C.c_uint8 * 64).from_buffer_copy(hb)
                be._check(self.api.b["peer_open_sync"](be.ctx, ffi.u32(r), buf), "peer_open_sync")
        self.fe._check(self.api.f["set_peer_exchange"](self.fe.fe, C.c_int32(1)), "set_peer_exchange")
        self.peer = True

    def _peer_map_images(self, x):
        """Maps the images of exchange x on all ranks (every rank reaches the same exchange: the calls below are collective)."""
        be = self.fe.backend
        for i in range(x.n_images):
            img = x.image[i]
            h = (C.c_uint8 * 64)()
            be._check(self.api.b["peer_get_image_handle"](be.ctx, img, h), "peer_get_image_handle")
            for r, hb in enumerate(self._all_gather_handles(h)):
                if r != self.rank:
                    buf = (C.c_uint8 * 64).from_buffer_copy(hb)
                    be._check(self.api.b["peer_open_image"](be.ctx, img, ffi.u32(r), buf), "peer_open_image")

    def check_peer_error(self, blocking=True):
        """Raises when a peer barrier gave up waiting (the frame consumed incomplete exchanges). blocking=False is the per-frame poll:
        no synchronisation, the value is at most one frame old (the error word is sticky)."""
        if not self.peer:
            return
        e = ffi.u32()
        name = "peer_error" if blocking else "peer_error_poll"
        self.fe.backend._check(self.api.b[name](self.fe.backend.ctx, C.byref(e)), name)
        if e.value:
            raise RuntimeError("peer exchange: a barrier timed out waiting for another rank - the frames since the previous check are invalid")

    def bands(self, rows, divisor):
        return [shard_band(self.api, self.H, self.world, r, divisor, rows) for r in range(self.world)]

    def exchange(self, x):
        import torch
        import torch.distributed as dist
        self.python_exchanges += 1
        if self.peer and x.kind != ffi.EXCHANGE_ALLREDUCE_SUM_U32:
            self._peer_map_images(x)  # from the next frame on run_segment performs this exchange on the device
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _Null()
        with ctx:
            if x.kind == ffi.EXCHANGE_ALLREDUCE_SUM_U32:
                t = alias_tensor(x.device_ptr[0], x.element_count * 4, self.device).view(torch.int32)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                self.bytes_sent += x.element_count * 4
                return
            ops, staged = [], []
            for t, rows, div in exchange_views(x, self.device, self.cache):
                sends, recvs = plan_row_exchange(x.kind, x.halo_rows, self.bands(rows, div), self.rank)
                for p, r0, r1 in sends:
                    ops.append(dist.P2POp(dist.isend, t[r0:r1].contiguous(), p))  # a copy only for 3-D levels (strided row ranges)
                    self.bytes_sent += (r1 - r0) * t.shape[1] * t.shape[2]
                for p, r0, r1 in recvs:
                    dst = t[r0:r1]
                    if dst.is_contiguous():
                        ops.append(dist.P2POp(dist.irecv, dst, p))
                    else:
                        tmp = torch.empty(dst.shape, dtype=dst.dtype, device=dst.device)
                        staged.append((dst, tmp))
                        ops.append(dist.P2POp(dist.irecv, tmp, p))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            for dst, tmp in staged:
                dst.copy_(tmp)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class LocalComm:
    """All ranks in ONE process (R frontends on one device): emulation of the sharded frame for tests. The exchange copies
    between the ranks' images; it is given the descriptors of all ranks for the same exchange point."""

    def __init__(self, api, full_height, world, device):
        self.api, self.H, self.world, self.device = api, full_height, world, device

    def exchange_all(self, xs):
        import torch
        if self.device is not None and str(self.device) != "cpu":
            torch.cuda.synchronize()
        x0 = xs[0]
        if x0.kind == ffi.EXCHANGE_ALLREDUCE_SUM_U32:
            ts = [alias_tensor(x.device_ptr[0], x.element_count * 4, self.device).view(torch.int32) for x in xs]
            total = torch.stack(ts).sum(dim=0, dtype=torch.int32)
            for t in ts:
                t.copy_(total)
        else:
            views = [exchange_views(x, self.device) for x in xs]
            for i in range(x0.n_images):
                rows, div = views[0][i][1], views[0][i][2]
                bands = [shard_band(self.api, self.H, self.world, r, div, rows) for r in range(self.world)]
                for me in range(self.world):
                    _, recvs = plan_row_exchange(x0.kind, x0.halo_rows, bands, me)
                    for p, r0, r1 in recvs:
                        views[me][i][0][r0:r1].copy_(views[p][i][0][r0:r1])
        if self.device is not None and str(self.device) != "cpu":
            torch.cuda.synchronize()


def copy_exchange(x):
    """run_segment reuses one descriptor per frontend: a deferred exchange is kept by value."""
    y = type(x)()
    C.memmove(C.byref(y), C.byref(x), C.sizeof(x))
    return y


def run_frame(fe, comm, cam, time, delta_time, inputs=None, async_upload=False, upload_rows=None):
    """One frame of one rank (DistComm)."""
    i = inputs or {}
    fe.begin_frame(cam, time, delta_time, i.get("depth"), i.get("motion"), i.get("normal"), i.get("gbuffer"), i.get("shadow_maps"), async_upload=async_upload, rows=upload_rows)
    n = 0
    while True:
        x = fe.run_segment()
        if x is None:
            if hasattr(comm, "check_peer_error"):
                comm.check_peer_error(blocking=False)  # every frame, without a synchronisation
            return n
        comm.exchange(x)
        n += 1


def run_frame_local(fes, comm, cam, time, delta_time, inputs=None, upload_rows=None):
    """One frame of all ranks in lockstep (LocalComm emulation)."""
    i = inputs or {}
    for r, fe in enumerate(fes):
        fe.begin_frame(cam, time, delta_time, i.get("depth"), i.get("motion"), i.get("normal"), i.get("gbuffer"), i.get("shadow_maps"), rows=upload_rows[r] if upload_rows else None)
    n = 0
    late = []  # deferred exchanges (next-frame data) are performed AFTER the frame's last pass, as the peer exchange may: a pass of this
    #            frame that needed their rows would read stale data here and fail the parity tests
    while True:
        xs = [fe.run_segment() for fe in fes]
        if xs[0] is None:
            assert all(x is None for x in xs)
            for held in late:
                comm.exchange_all(held)
            return n
        if xs[0].deferred:
            late.append([copy_exchange(x) for x in xs])
        else:
            comm.exchange_all(xs)
        n += 1
