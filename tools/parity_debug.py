"""Debug tool (GPU box): renders a short sequence with the CUDA backend and the CPU oracle and prints, per frame and per
named resource, the number of differing bytes. Usage: python tools/parity_debug.py [W H frames instances]"""
import sys, time
import numpy as np
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
import plainrenderer_b200 as pr
from plainrenderer_b200 import ffi

W, H, FRAMES, INST = (int(a) for a in (sys.argv[1:5] + ["256", "144", "3", "24"][len(sys.argv) - 1:]))
IMAGES = ["skyTransmission", "skyMultiscatter", "skyLut", "hiz", "depthHalf", "giY0", "giC0", "giY1", "giC1", "giHistY0", "giHistC0", "giHistY1", "giHistC1", "giFullY", "giFullC",
          "froxelMaterial", "froxelScatter", "froxelHist0", "froxelHist1", "froxelIntegration", "color0", "color1", "taaHist0", "taaHist1", "post1", "bloomDown", "bloomUp", "brdfLut", "output"]
BUFFERS = [("histogramPerTile", None), ("histogram", 512), ("light", 20), ("sunShadowInfo", 304), ("sdfCulled", None), ("sdfTiles", None)]

cuda = pr.load()
orc = ffi.Api(str(pr.PKG_DIR.parent / "oracle" / "_build" / "liboracle.so"), "oracle_", "oracle_frontend_")
cam = ffi.camera((-13.0, -1.7, 0.5), (1, 0, 0), (0, 0, 1), (0, -1, 0))
sides = []
for api in (cuda, orc):
    s = ffi.default_settings(api, W, H, sun_direction_deg=(40.0, 35.0))
    fe = ffi.Frontend(api, s)
    sc = ffi.SyntheticScene(api, n_instances=INST)
    sc.attach(fe)
    fe.set_exposure(2e-5)
    sides.append((api, s, fe, sc))
mipcount = lambda fe, h: 1 + int(np.floor(np.log2(max(fe.backend.image_description(h).width, fe.backend.image_description(h).height)))) if fe.backend.image_description(h).mip_count == 1 else (fe.backend.image_description(h).manual_mip_count if fe.backend.image_description(h).mip_count == 2 else 1)
inp = None
for f in range(FRAMES):
    outs = []
    for api, s, fe, sc in sides:
        if inp is None or api is cuda:
            inp = sc.render_inputs(s, cam, f + 1, shadows=True)
        t0 = time.time()
        fe.render_frame(cam, (f + 1) / 60.0, 1 / 60.0, inp["depth"], inp["motion"], inp["normal"], inp["gbuffer"], inp["shadow_maps"])
        fe.backend._check(api.b["wait_for_gpu_idle"](fe.backend.ctx), "idle")
        outs.append(time.time() - t0)
    print("frame %d: cuda %.3fs oracle %.3fs" % (f, outs[0], outs[1]))
    (ca, cs, cfe, _), (oa, os_, ofe, _) = sides
    for name in IMAGES:
        hc, ho = cfe.image(name), ofe.image(name)
        for mip in range(mipcount(cfe, hc)):
            a, b = cfe.backend.read_image(hc, mip), ofe.backend.read_image(ho, mip)
            n = int((a != b).sum())
            if n:
                idx = np.nonzero(a != b)[0]
                print("   %-18s mip %d: %d / %d bytes differ (first at byte %d: cuda %d oracle %d)" % (name, mip, n, a.size, idx[0], a[idx[0]], b[idx[0]]))
    for name, size in BUFFERS:
        hc, ho = cfe.storage_buffer(name), ofe.storage_buffer(name)
        if size is None:
            p, sz = ffi.C.c_void_p(), ffi.C.c_size_t()
            ca.b["get_storage_buffer_device_pointer"](cfe.backend.ctx, ffi.u32(hc), ffi.C.byref(p), ffi.C.byref(sz))
            size = sz.value
        a, b = cfe.backend.read_storage_buffer(hc, size), ofe.backend.read_storage_buffer(ho, size)
        n = int((a != b).sum())
        if n:
            idx = np.nonzero(a != b)[0]
            print("   buffer %-12s: %d / %d bytes differ (first at byte %d)" % (name, n, a.size, idx[0]))
    print("   launches/frame:", cfe.backend.last_frame_launch_count(), " output mean:", cfe.read_output().reshape(H, W, 4)[..., :3].mean(axis=(0, 1)))
print("done")
