"""Fast-contract error model (liboracle_sfu.so) against the exact oracle over the non-default configurations. CPU only."""
import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')  # run from the repository root
import numpy as np
from plainrenderer_b200 import ffi
import tolerance
from conftest import Sequence
o=ffi.Api('oracle/_build/liboracle.so','oracle_','oracle_frontend_'); s=ffi.Api('oracle/_build/liboracle_sfu.so','oracle_','oracle_frontend_')
variants=[dict(indirect_lighting_tech=1), dict(diffuse_brdf=0,direct_multiscatter=1), dict(diffuse_brdf=1,direct_multiscatter=2), dict(diffuse_brdf=3,direct_multiscatter=3),
          dict(half_res_trace=0), dict(sun_shadow_cascade_count=4), dict(taa_history_sampling_tech=1), dict(taa_history_sampling_tech=2), dict(taa_history_sampling_tech=0,taa_use_clipping=0),
          dict(taa_use_separate_supersampling=1), dict(use_geometry_aa=0), dict(strict_influence_radius_cutoff=1)]
for v in variants:
    a,b=Sequence(ffi,s,160,90,16,**v),Sequence(ffi,o,160,90,16,**v)
    try:
        for f in range(4):
            i=a.step(moving=True); b.step(moving=True,inputs=i)
        bad,log=tolerance.compare_snapshots(a.snapshot(),b.snapshot())
        print(v, "OK" if not bad else "FAIL %s"%bad, "|", log[0])
    except Exception as e:
        print(v,"EXC",e)
    finally:
        a.close(); b.close()
