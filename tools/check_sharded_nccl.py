#!/usr/bin/env python
"""Parity of the row-sharded frame over NCCL: run under torchrun with N ranks (one GPU each).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/check_sharded_nccl.py [W H]

Every rank renders the same 4-frame sequence twice on its own GPU: unsharded (whole frame, no exchange) and as rank r of
N with the exchanges of plainrenderer_b200.sharding.DistComm over NCCL. The rank's band of the tonemapped frame, the
all-reduced histogram and the all-gathered TAA history must equal the unsharded result bit for bit.
--raster: the frames are rendered from `.plain` meshes (raster_inputs = 1): nothing is uploaded, every rank
rasterises its band + halo, the motion vectors travel as an 11th exchange.
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

CAMERA = ((-13.0, -1.7, 0.5), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, -1.0, 0.0))


def main():
    import torch
    import torch.distributed as dist
    import plainrenderer_b200 as pr
    from plainrenderer_b200 import ffi, sharding

    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    peer = "--nccl" not in sys.argv  # default: peer pushes over NVLink (CUDA IPC); --nccl: send/recv batches from Python
    W, H = (int(args[0]), int(args[1])) if len(args) > 1 else (512, 512)
    frames = int(args[2]) if len(args) > 2 else 6
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api = pr.load()

    def make(r, n):
        s = ffi.default_settings(api, W, H, sun_direction_deg=(40.0, 35.0), shard_rank=r, shard_count=n)
        fe = ffi.Frontend(api, s, device=local)
        scene = ffi.SyntheticScene(api, n_instances=14)
        scene.attach(fe)
        fe.set_exposure(2e-5)
        return s, fe, scene

    raster = "--raster" in sys.argv
    if raster:
        sys.path.insert(0, str(ROOT / "tests"))
        from conftest import PlainSceneSequence  # the scene of tests/test_raster_gpu.py: the golden .plain assets, instanced
        from plainrenderer_b200 import assets
        seq0 = PlainSceneSequence(ffi, api, assets.Assets(), W, H, device=local)
        seq1 = PlainSceneSequence(ffi, api, assets.Assets(), W, H, device=local, shard_rank=rank, shard_count=world)
        s0, ref, scene0, s1, fe, scene1 = seq0.s, seq0.fe, None, seq1.s, seq1.fe, None
    else:
        s0, ref, scene0 = make(0, 0)
        s1, fe, scene1 = make(rank, world)
    stream_ptr = C.c_void_p()
    api.b["get_stream"](fe.backend.ctx, C.byref(stream_ptr))
    stream = torch.cuda.ExternalStream(stream_ptr.value, device=torch.device("cuda", local))
    comm = sharding.DistComm(api, H, device=torch.device("cuda", local), stream=stream, frontend=fe, peer=peer)
    band = sharding.full_res_band(api, H, world, rank)
    upload = (max(band[0] - 16, 0), min(band[1] + 16, H))
    ok = True
    prev = None
    skew = [a for a in sys.argv[1:] if a.startswith("--skew=")]  # --skew=RANK:MS: that rank is late on the host by MS before every frame and mid-frame
    skew_rank, skew_ms = (int(skew[0][7:].split(":")[0]), float(skew[0][7:].split(":")[1])) if skew else (-1, 0.0)
    if rank == skew_rank:
        import time
        real_exchange = comm.exchange
        orig_run_segment = fe.run_segment

        def slow_run_segment():
            time.sleep(skew_ms * 1e-3)
            return orig_run_segment()
        fe.run_segment = slow_run_segment
    for f in range(frames):
        p, fw, r, u = CAMERA
        cam = ffi.camera((p[0] + 0.02 * f, p[1], p[2] + 0.01 * f), fw, r, u)
        if raster:
            ref.render_frame(cam, (f + 1) / 60.0, 1 / 60.0)
            n = sharding.run_frame(fe, comm, cam, (f + 1) / 60.0, 1 / 60.0)
        else:
            inputs = scene0.render_inputs(s0, cam, f + 1, prev_cam=prev, shadows=True)
            prev = cam
            ref.render_frame(cam, (f + 1) / 60.0, 1 / 60.0, inputs["depth"], inputs["motion"], inputs["normal"], inputs["gbuffer"], inputs["shadow_maps"])
            n = sharding.run_frame(fe, comm, cam, (f + 1) / 60.0, 1 / 60.0, inputs, upload_rows=upload)
        torch.cuda.synchronize()
        want = ref.read_output().reshape(H, W * 4)
        got = np.zeros((H, W * 4), np.uint8)
        fe.read_output_rows(got, band)
        a, b = band
        same_frame = np.array_equal(got[a:b], want[a:b])
        hist_w = ref.backend.read_storage_buffer(ref.storage_buffer("histogram"), 512)
        hist_g = fe.backend.read_storage_buffer(fe.storage_buffer("histogram"), 512)
        hname = "taaHist%d" % (f % 2)
        same_hist = np.array_equal(hist_w, hist_g)
        taa_w, taa_g = ref.backend.read_image(ref.image(hname)), fe.backend.read_image(fe.image(hname))
        same_taa = np.array_equal(taa_w, taa_g)
        if not same_taa:  # which rows: the rank's own band, its overlapped rows or rows a peer pushed
            bad = np.nonzero((taa_w.reshape(H, -1) != taa_g.reshape(H, -1)).any(axis=1))[0]
            print("rank %d frame %d: TAA history differs in %d rows, first %d last %d (band [%d,%d))" % (rank, f, len(bad), bad[0], bad[-1], band[0], band[1]), flush=True)
        if raster:  # the all-gathered motion vectors of this frame
            mname = "motion%d" % ((f + 1) % 3)
            same_taa = same_taa and np.array_equal(ref.backend.read_image(ref.image(mname)), fe.backend.read_image(fe.image(mname)))
        print("rank %d/%d frame %d: %d exchanges through Python, band [%d,%d) frame %s, histogram %s, TAA history %s" %
              (rank, world, f, n, a, b, "equal" if same_frame else "DIFFERS", "equal" if same_hist else "DIFFERS", "equal" if same_taa else "DIFFERS"), flush=True)
        ok = ok and same_frame and same_hist and same_taa
    comm.check_peer_error()
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("SHARDED_%s_PARITY %s (%d ranks, %dx%d, %d frames%s)" % ("PEER" if peer else "NCCL", "OK" if int(t.item()) == 0 else "FAILED", world, W, H, frames, ", rasterised inputs" if raster else ""), flush=True)
    for sc, f_ in ((scene0, ref), (scene1, fe)):
        if sc is not None:
            sc.close()
        f_.close()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
