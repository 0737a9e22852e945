"""Generates tests/golden/frame_sequence.json from the CPU oracle (the reference ships no golden vectors; its GLSL
cannot execute here - SURVEY.md 8c). Run from the repo root: python tests/golden/make_golden.py

A 3-frame 96x54 sequence with a moving camera; per frame the sha256 of every named resource, plus the raw luminance
histogram, LightBuffer and a 16x9 patch of the tonemapped frame for readable diffs."""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parents[1]))
import conftest  # noqa: E402
from plainrenderer_b200 import ffi  # noqa: E402

W, H, FRAMES, INSTANCES = 96, 54, 3, 10


def run(api):
    seq = conftest.Sequence(ffi, api, W, H, instances=INSTANCES)
    frames = []
    for f in range(FRAMES):
        seq.step(moving=True)
        snap = seq.snapshot()
        frames.append({"sha256": {k: hashlib.sha256(v.tobytes()).hexdigest() for k, v in snap.items()},
                       "histogram": snap["buf:histogram"].view(np.uint32).tolist(),
                       "light": snap["buf:light"].view(np.float32).tolist(),
                       "output_patch": snap["output/0"].reshape(H, W, 4)[20:29, 40:56].tolist()})
    seq.close()
    return {"width": W, "height": H, "instances": INSTANCES, "frames": frames}


N4_W, N4_H, N4_FRAMES, N4_INSTANCES = 64, 36, 2, 8
N4_IMAGES = ["post0", "post1", "taaLum0", "taaLum1", "taaHist0", "taaHist1", "output"]


def run_n4(api):
    """SURVEY.md 8f N4 variants (supersampling, debug modes 1, 2 and 4 of conftest.N4_VARIANTS; the GPU parity tests run all of
    them against the oracle): sha256 of the images the extra passes write, 2 frames at 64x36."""
    out = []
    for settings in [conftest.N4_VARIANTS[i] for i in (0, 2, 3, 5)]:
        seq = conftest.Sequence(ffi, api, N4_W, N4_H, instances=N4_INSTANCES, **settings)
        frames = []
        for f in range(N4_FRAMES):
            seq.step(moving=True)
            snap = seq.snapshot(N4_IMAGES, [("histogram", 512)])
            frames.append({k: hashlib.sha256(v.tobytes()).hexdigest() for k, v in snap.items()})
        seq.close()
        out.append({"settings": settings, "frames": frames})
    return out


if __name__ == "__main__":
    api = ffi.Api(str(HERE.parents[1] / "oracle" / "_build" / "liboracle.so"), "oracle_", "oracle_frontend_")
    (HERE / "frame_sequence.json").write_text(json.dumps(run(api)))
    print("wrote", HERE / "frame_sequence.json")
    (HERE / "n4_variants.json").write_text(json.dumps(run_n4(api)))
    print("wrote", HERE / "n4_variants.json")
