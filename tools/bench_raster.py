#!/usr/bin/env python
"""Times the rasterisation passes (depth prepass, G-buffer fill, sun shadow cascades; csrc/passes_raster.cu) inside
whole frames rendered end to end from meshes (raster_inputs = 1) at 3840x2160, per pass with CUDA events (the backend's timing mode).

    python tools/bench_raster.py [--tess 1,8,32] [--boxes 100] [--frames 10]

Scene: `--boxes` boxes of the size range of the bench's synthetic scene, each face tessellated into tess x tess quads, so the same
picture is drawn with 12 ... 12 * tess^2 triangles per box: tess = 1 is the few-huge-triangles case (walls covering a quarter of the
frame: the persistent big-triangle path), tess = 32 is 1.2 M small triangles (the warp-per-triangle path). Prints one JSON line per
tessellation: triangles, ms per pass, Mtriangles/s and covered Mpixels/s of the prepass. Not the headline bench (bench.py)."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
CAMERA = ((-13.0, -1.7, 0.5), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, -1.0, 0.0))


def box_mesh(ffi, tess, outward_is_cross):
    """unit box [-1, 1]^3, every face tess x tess quads; triangle winding chosen so that cross(b - a, c - a) points outward
    (or inward) as in the reference asset pipeline's own cube (tests/golden/sdf/cube.plain)"""
    pos, nrm, tan, bit, uv, idx = [], [], [], [], [], []
    g = np.linspace(-1.0, 1.0, tess + 1)
    for axis in range(3):
        for sign in (-1.0, 1.0):
            u_axis, v_axis = (axis + 1) % 3, (axis + 2) % 3
            n = np.zeros(3); n[axis] = sign
            u = np.zeros(3); u[u_axis] = 1.0
            v = np.zeros(3); v[v_axis] = 1.0
            base = len(pos)
            for b in g:
                for a in g:
                    p = n + a * u + b * v
                    pos.append(p); nrm.append(n); tan.append(u); bit.append(v); uv.append(((a + 1) * 2, (b + 1) * 2))
            flip = (np.dot(np.cross(u, v), n) > 0) != outward_is_cross
            for j in range(tess):
                for i in range(tess):
                    q = [base + j * (tess + 1) + i, base + j * (tess + 1) + i + 1, base + (j + 1) * (tess + 1) + i + 1, base + (j + 1) * (tess + 1) + i]
                    tris = [(q[0], q[1], q[2]), (q[0], q[2], q[3])]
                    for t in tris:
                        idx.extend(t[::-1] if flip else t)
    return np.array(idx, np.uint32), ffi.pack_vertices(np.array(pos, np.float32), uvs=np.array(uv), normals=np.array(nrm), tangents=np.array(tan), bitangents=np.array(bit))


def box_sdf(res=16):
    """analytic SDF brick of the unit box padded like the asset pipeline pads (half floats, (d, h, w))"""
    c = (np.arange(res) + 0.5) / res * 3.0 - 1.5  # the brick spans the padded box: padSDFBoundingBox adds max(7.5 %, 0.5 m) per side
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    q = np.stack([np.abs(x) - 1, np.abs(y) - 1, np.abs(z) - 1], -1)
    d = np.linalg.norm(np.maximum(q, 0), axis=-1) + np.minimum(q.max(-1), 0)
    return d.astype(np.float16).view(np.uint16)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tess", default="1,8,32")
    ap.add_argument("--boxes", type=int, default=100)
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    args = ap.parse_args()
    import plainrenderer_b200 as pr
    from plainrenderer_b200 import assets, ffi
    api = pr.load()
    cube = assets.Assets().load_scene(ROOT / "tests" / "golden" / "sdf" / "cube.plain").meshes[0]
    a, b, c = (cube.positions[cube.indices[k]] for k in range(3))
    packed = int(np.frombuffer(cube.vertices[cube.indices[0], 16:20].tobytes(), np.uint32)[0])
    comp = lambda s: ((packed >> s) & 1023) - (1024 if ((packed >> s) & 1023) >= 512 else 0)
    vertex_normal = np.array([comp(20), comp(10), comp(0)], np.float64)
    outward_is_cross = bool(np.dot(np.cross(b - a, c - a), vertex_normal) > 0)
    rng = np.random.default_rng(0x504c4149)
    placement = []
    for k in range(args.boxes):
        half = rng.uniform(0.4, 2.5, 3) * (3.0 if k < 6 else 1.0)
        centre = np.array([rng.uniform(-12, 18), rng.uniform(-6, 2.5), rng.uniform(-9, 9)])
        placement.append((half, centre))
    placement.append((np.array([20.0, 0.3, 12.0]), np.array([3.0, 3.0, 0.0])))  # the floor
    for tess in [int(t) for t in args.tess.split(",")]:
        s = ffi.default_settings(api, args.width, args.height, raster_inputs=1, sun_direction_deg=(40.0, 35.0))
        fe = ffi.Frontend(api, s)
        mesh = fe.register_sdf_mesh(box_sdf(), (-1, -1, -1), (1, 1, 1), (0.7, 0.65, 0.6))
        idx, vtx = box_mesh(ffi, tess, outward_is_cross)
        fe.set_mesh_geometry(mesh, idx, vtx)
        objects = []
        for half, centre in placement:
            M = np.diag([half[0], half[1], half[2], 1.0]).astype(np.float32)
            M[:3, 3] = centre
            objects.append((mesh, M.T.ravel(), centre - half, centre + half))
        fe.set_scene(objects)
        fe.set_exposure(2e-5)
        cam = ffi.camera(*CAMERA)
        be = fe.backend
        for f in range(3):
            fe.render_frame(cam, (f + 1) / 60.0, 1 / 60.0)
        be.set_timing_enabled(True)
        acc = {}
        for f in range(args.frames):
            fe.render_frame(cam, (f + 4) / 60.0, 1 / 60.0)
            for name, ms in be.pass_timings():
                acc[name] = acc.get(name, 0.0) + ms / args.frames
        be.set_timing_enabled(False)
        depth = be.read_image(fe.image("depth0"), 0, np.float32)
        tris = len(idx) // 3 * len(objects)
        raster = {k: round(v, 4) for k, v in acc.items() if k.startswith(("Depth prepass", "G-buffer fill", "Shadow map cascade"))}
        pre = acc.get("Depth prepass", 0.0)
        print(json.dumps({"tool": "bench_raster", "resolution": [args.width, args.height], "boxes": len(objects), "tess": tess, "triangles_per_pass": tris,
                          "coverage": round(float((depth > 0).mean()), 3), "raster_passes_ms": raster, "raster_total_ms": round(sum(raster.values()), 4),
                          "frame_total_ms": round(sum(acc.values()), 4),
                          "prepass_mtris_per_s": round(tris / pre / 1e3, 1) if pre else None,
                          "prepass_covered_mpix_per_s": round(float((depth > 0).sum()) / pre / 1e3, 1) if pre else None}), flush=True)
        fe.close()


if __name__ == "__main__":
    main()
