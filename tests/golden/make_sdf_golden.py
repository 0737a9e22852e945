"""Generates tests/golden/sdf/* with the REFERENCE's own asset pipeline (oracle/_ref/PlainAssetPipeline, built by
oracle/build_ref.sh from /root/reference: glTF import + CPU SDF bake, SceneSDF.cpp:296-514).

For each case a glTF derived from the reference's vendored tinygltf Cube model (positions transformed in the .bin) is
converted; the fixture is the reference's own output: <case>.plain (mesh exactly as the bake saw it: positions, indices,
bounding box) and <case>.dds (R16F 3-D brick). Run from the repo root in the build container:
    python tests/golden/make_sdf_golden.py
The binary hangs at exit (detached job-system workers, JobSystem.cpp:34-42): it is stopped once the brick is written."""
import json
import shutil
import struct
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF_BIN = ROOT / "oracle" / "_ref" / "PlainAssetPipeline"
CUBE = Path("/root/reference/Plain/vendor/tinygltf/models/Cube")
OUT = ROOT / "tests" / "golden" / "sdf"

# name -> 3x3 matrix applied to the cube's positions (rows) and a translation
CASES = {
    "cube": (np.eye(3), np.zeros(3)),                                                        # 2 m cube -> 16^3
    "slab": (np.array([[3.0, 0.4, 0.0], [0.0, 1.0, 0.3], [0.2, 0.0, 1.7]]), np.array([0.5, -0.25, 1.0])),  # sheared, off-centre -> 32 x 16 x 16
    "tall": (np.array([[0.8, 0.0, 0.1], [0.3, 4.5, 0.0], [0.0, 0.2, 1.1]]), np.array([-1.0, 2.0, 0.0])),   # -> 16 x 64 x 16
    # BASELINE configs[0]: the cube scaled x8 (16 m edge) -> 64^3, the reference's maximum (SceneSDF.cpp:120-131). 25 s on one core.
    # Only its sha256 is committed (cube64.dds.sha256; the brick itself is 512 KB) together with the mesh the bake saw (cube64.plain).
    "cube64": (np.eye(3) * 8.0, np.zeros(3)),
}
HASH_ONLY = {"cube64"}


def run_case(name, matrix, offset):
    work = Path(tempfile.mkdtemp(prefix="sdfgold_"))
    model_dir = work / "resources" / "models" / name
    model_dir.mkdir(parents=True)
    for f in CUBE.iterdir():
        shutil.copy(f, model_dir / f.name)
    gltf = json.loads((model_dir / "Cube.gltf").read_text())
    acc = gltf["accessors"][gltf["meshes"][0]["primitives"][0]["attributes"]["POSITION"]]
    view = gltf["bufferViews"][acc["bufferView"]]
    raw = bytearray((model_dir / "Cube.bin").read_bytes())
    start = view.get("byteOffset", 0) + acc.get("byteOffset", 0)
    pos = np.frombuffer(bytes(raw[start:start + acc["count"] * 12]), np.float32).reshape(-1, 3).astype(np.float64)
    pos = (pos @ matrix.T + offset).astype(np.float32)
    raw[start:start + acc["count"] * 12] = pos.tobytes()
    (model_dir / "Cube.bin").write_bytes(bytes(raw))
    acc["min"], acc["max"] = pos.min(0).tolist(), pos.max(0).tolist()
    (model_dir / "Cube.gltf").write_text(json.dumps(gltf))
    dds = model_dir / "sdfTextures" / "Cube.dds"
    logf = open(work / "log.txt", "w")
    p = subprocess.Popen([str(REF_BIN), "models/%s/Cube.gltf" % name], cwd=work, stdout=logf, stderr=subprocess.STDOUT)
    t0, last = time.time(), -1
    while time.time() - t0 < 300 and p.poll() is None:  # stdout is block-buffered and the process never exits: watch the output file instead
        time.sleep(0.5)
        size = dds.stat().st_size if dds.exists() else -1
        if size > 148 and size == last:
            break
        last = size
    p.kill()
    p.wait()
    logf.close()
    log = (work / "log.txt").read_text().splitlines()
    if not dds.exists():
        raise SystemExit("reference asset pipeline produced no brick for %s:\n%s" % (name, "\n".join(log)))
    OUT.mkdir(parents=True, exist_ok=True)
    shutil.copy(model_dir / "Cube.plain", OUT / (name + ".plain"))
    if name in HASH_ONLY:
        import hashlib
        raw = dds.read_bytes()
        (OUT / (name + ".dds.sha256")).write_text("%s  file (%d bytes)\n%s  texels (the %d bytes after the 148-byte DDS + DX10 header)\n"
                                                   % (hashlib.sha256(raw).hexdigest(), len(raw), hashlib.sha256(raw[148:]).hexdigest(), len(raw) - 148))
    else:
        shutil.copy(dds, OUT / (name + ".dds"))
    hdr = dds.read_bytes()[:148]
    h, w, d = struct.unpack_from("<I", hdr, 12)[0], struct.unpack_from("<I", hdr, 16)[0], struct.unpack_from("<I", hdr, 24)[0]
    print("%s: brick %dx%dx%d, %s" % (name, w, h, d, [l for l in log if "SDF computation time" in l]))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    if not REF_BIN.exists():
        raise SystemExit("build the reference asset pipeline first: bash oracle/build_ref.sh")
    for name, (m, o) in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        run_case(name, m, o)
