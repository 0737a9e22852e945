"""Host-side logic of the row-sharded frame (N > 1): band partition, exchange planning, and the exchanges themselves over
torch.distributed with the gloo backend, world size 2, on CPU tensors."""
import ctypes as C
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


@pytest.mark.parametrize("H,count", [(2160, 8), (2160, 4), (2160, 2), (1080, 8), (192, 3), (4320, 8), (64, 1)])
def test_bands_partition_every_level(cuda, H, count):
    from plainrenderer_b200 import sharding
    for divisor, rows in [(1, H), (2, H // 2), (32, (H + 31) // 32), (16, H // 16), (2, (H // 2))]:
        bands = [sharding.shard_band(cuda, H, count, r, divisor, rows) for r in range(count)]
        assert bands[0][0] == 0 and bands[-1][1] == rows
        for (a0, b0), (a1, b1) in zip(bands, bands[1:]):
            assert b0 == a1 and b0 >= a0  # contiguous, ordered, no overlap
    full = [sharding.full_res_band(cuda, H, count, r) for r in range(count)]
    assert all(a % 32 == 0 for a, _ in full)  # histogram tiles, 16-row trace blocks, froxel rows, 4 fused HiZ levels never straddle ranks
    sizes = [b - a for a, b in full]
    assert max(sizes) - min(sizes) <= 32 + (32 - H % 32) % 32


def test_exchange_plans_are_consistent(cuda):
    from plainrenderer_b200 import ffi, sharding
    H, count = 2160, 8
    for kind, halo in [(ffi.EXCHANGE_ALLGATHER_ROWS, 0), (ffi.EXCHANGE_HALO_ROWS, 2), (ffi.EXCHANGE_HALO_ROWS, 8)]:
        for divisor, rows in [(1, H), (2, H // 2), (16, H // 16)]:
            bands = [sharding.shard_band(cuda, H, count, r, divisor, rows) for r in range(count)]
            plans = [sharding.plan_row_exchange(kind, halo, bands, r) for r in range(count)]
            for me, (sends, recvs) in enumerate(plans):
                for p, r0, r1 in sends:  # what I send to p is exactly what p expects from me
                    assert (me, r0, r1) in plans[p][1]
                for p, r0, r1 in recvs:
                    assert (me, r0, r1) in plans[p][0]
                    assert bands[p][0] <= r0 < r1 <= bands[p][1]  # only rows the sender owns
            if kind == ffi.EXCHANGE_ALLGATHER_ROWS:
                for me, (_, recvs) in enumerate(plans):
                    covered = sorted([(r0, r1) for _, r0, r1 in recvs] + [bands[me]])
                    assert covered[0][0] == 0 and covered[-1][1] == rows and all(x[1] == y[0] for x, y in zip(covered, covered[1:]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        import torch.distributed as dist
        import plainrenderer_b200 as pr
        from plainrenderer_b200 import ffi, sharding
        dist.init_process_group("gloo", rank=rank, world_size=world)
        api = pr.load()
        comm = sharding.DistComm(api, H, device="cpu")
        results = {}
        # two images exchanged together (like Y_SH + CoCg): full-res rows with pitch 24, half-res rows with pitch 8
        for kind, halo in [(ffi.EXCHANGE_ALLGATHER_ROWS, 0), (ffi.EXCHANGE_HALO_ROWS, 3)]:
            imgs = []
            x = ffi.Exchange()
            x.kind, x.halo_rows, x.n_images = kind, halo, 2
            for i, (rows, div, pitch) in enumerate([(H, 1, 24), (H // 2, 2, 8)]):
                a, b = sharding.shard_band(api, H, world, rank, div, rows)
                img = np.zeros((rows, pitch), np.uint8)
                img[a:b] = (np.arange(a, b)[:, None] * 7 + np.arange(pitch)[None, :] + 100 * i) % 251 + 1  # value depends on (row, column, image), not on the rank
                imgs.append(img)
                x.device_ptr[i], x.rows[i], x.row_pitch_bytes[i], x.row_divisor[i] = img.ctypes.data, rows, pitch, div
            comm.exchange(x)
            results[(kind, halo)] = [m.copy() for m in imgs]
        hist = np.arange(128, dtype=np.uint32) * (rank + 1)
        x = ffi.Exchange()
        x.kind, x.n_images, x.element_count = ffi.EXCHANGE_ALLREDUCE_SUM_U32, 1, 128
        x.device_ptr[0] = hist.ctypes.data
        comm.exchange(x)
        results["hist"] = hist.copy()
        q.put((rank, results))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "ERROR: " + traceback.format_exc()))


def test_exchanges_over_gloo_world_size_2(cuda):
    import torch.multiprocessing as mp
    from plainrenderer_b200 import ffi, sharding
    H, world = 256, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    for r in range(world):
        assert not isinstance(out[r], str), out[r]
    expect = lambda rows, pitch, i: ((np.arange(rows)[:, None] * 7 + np.arange(pitch)[None, :] + 100 * i) % 251 + 1).astype(np.uint8)
    for r in range(world):
        full, half = out[r][(ffi.EXCHANGE_ALLGATHER_ROWS, 0)]
        assert np.array_equal(full, expect(H, 24, 0)) and np.array_equal(half, expect(H // 2, 8, 1))  # every rank holds every row
        full, half = out[r][(ffi.EXCHANGE_HALO_ROWS, 3)]
        for img, rows, div, i in ((full, H, 1, 0), (half, H // 2, 2, 1)):
            a, b = sharding.shard_band(cuda, H, world, r, div, rows)
            lo, hi = max(a - 3, 0), min(b + 3, rows)
            assert np.array_equal(img[lo:hi], expect(rows, img.shape[1], i)[lo:hi])  # own band + 3 halo rows
            assert not img[:lo].any() and not img[hi:].any()                         # nothing else was touched
        assert np.array_equal(out[r]["hist"], np.arange(128, dtype=np.uint32) * 3)


def test_exchange_schedule_of_a_sharded_frame(ffi, oracle):
    """the exchanges a row-sharded frontend asks for, in order - with uploaded inputs (8) and with rasterised inputs (the motion
    vectors become a 9th, right after the depth prepass). Driven against the CPU backend: the schedule is host-side logic."""
    from conftest import PlainSceneSequence
    from plainrenderer_b200 import assets

    deferred = {}

    def schedule(fe, cam):
        names = []
        fe.begin_frame(cam, 1 / 60.0, 1 / 60.0)
        while True:
            x = fe.run_segment()
            if x is None:
                return names
            names.append((x.name.decode(), x.kind, x.n_images))
            deferred[x.name.decode()] = int(x.deferred)
    lib = assets.Assets(ROOT / "oracle" / "_build" / "liboracle.so", "oracle_asset_")
    seq = PlainSceneSequence(ffi, oracle, lib, 64, 64, shard_rank=0, shard_count=2)
    raster = schedule(seq.fe, seq.camera_at(0, False))
    seq.close()
    s = ffi.default_settings(oracle, 64, 64, shard_rank=1, shard_count=2)
    fe = ffi.Frontend(oracle, s)
    scene = ffi.SyntheticScene(oracle, n_instances=4)
    scene.attach(fe)
    uploaded = schedule(fe, ffi.camera((-13.0, -1.7, 0.5), (1, 0, 0), (0, 0, 1), (0, -1, 0)))
    scene.close()
    fe.close()
    # round 2: the half-res depth travels with the pyramid level (one barrier), the 2-row halo of the first spatial filter became overlapped
    # computation, and the three exchanges of next-frame data (giSpatial1, froxelHistory, taaHistory) are marked deferred - over peer
    # exchange they run behind the frame; through Python (this test, NCCL, LocalComm) they are ordinary all-gathers at the same place
    # round 2: the half-res depth travels with the pyramid level, the 2-row halo of the first spatial filter became overlapped
    # computation, the three exchanges of next-frame data (froxelHistory, giSpatial1, taaHistory) are marked deferred - over peer exchange
    # they run behind the frame; through Python (this test, NCCL, LocalComm) they are ordinary all-gathers - and with uploaded inputs the
    # rank-local work is recorded first so that the row all-gather and the histogram all-reduce are back to back (one barrier on the device)
    tail = ["froxelHistory", "giTrace", "giTemporal", "giSpatial1", "taaHistory", "bloomMip1"]  # the froxel chain is recorded beside the trace
    assert [n for n, _, _ in uploaded] == ["hiz+depthHalf", "histogram"] + tail
    assert dict((n, c) for n, _, c in uploaded)["hiz+depthHalf"] == 2
    assert [n for n, _, _ in raster] == ["histogram", "motion", "hiz+depthHalf"] + tail
    assert dict((n, k) for n, k, _ in raster)["motion"] == ffi.EXCHANGE_ALLGATHER_ROWS
    assert uploaded[1][1] == ffi.EXCHANGE_ALLREDUCE_SUM_U32
    # exactly the three all-gathers of next-frame data carry the deferred mark (include/plain_frontend.h plain_exchange.deferred)
    assert sorted(n for n, d in deferred.items() if d) == ["froxelHistory", "giSpatial1", "taaHistory"]
