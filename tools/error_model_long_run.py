"""Fast-contract error model (liboracle_sfu.so) against the exact oracle: 16 frames, 100 instances, 8 frames of motion then 8 at rest. CPU only."""
import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')  # run from the repository root
import numpy as np
from plainrenderer_b200 import ffi
import tolerance
from conftest import Sequence
o=ffi.Api('oracle/_build/liboracle.so','oracle_','oracle_frontend_'); s=ffi.Api('oracle/_build/liboracle_sfu.so','oracle_','oracle_frontend_')
a,b=Sequence(ffi,s,320,180,100),Sequence(ffi,o,320,180,100)
for f in range(16):
    i=a.step(moving=(f<8)); b.step(moving=(f<8),inputs=i)
    if f in (0,3,7,11,15):
        bad,log=tolerance.compare_snapshots(a.snapshot(),b.snapshot())
        print("frame",f,"FAIL" if bad else "ok", bad)
        print("   "+log[0]); print("   "+[l for l in log if l.startswith("color0")][0]); print("   "+[l for l in log if l.startswith("giFullY")][0]); print("   "+log[-1])
