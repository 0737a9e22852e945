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


if __name__ == "__main__":
    api = ffi.Api(str(HERE.parents[1] / "oracle" / "_build" / "liboracle.so"), "oracle_", "oracle_frontend_")
    (HERE / "frame_sequence.json").write_text(json.dumps(run(api)))
    print("wrote", HERE / "frame_sequence.json")
