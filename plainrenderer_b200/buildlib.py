"""Builds plainrenderer_b200/libplain_b200.so in-tree: the CUDA backend (csrc/*.cu, sm_100a) + the host-side frontend
mirror (host/*.cpp) behind the C-ABI of include/plain_b200.h and include/plain_frontend.h.

    python -m plainrenderer_b200.buildlib [--force] [--jobs N]

nvcc cross-compiles without a GPU. Flags that are part of the numeric contract: -fmad=false (no
contraction), default -prec-div/-prec-sqrt/-ftz=false; host side -ffp-contract=off.

A second library, libplain_b200_fast.so, is the same sources with the floating-point passes (FAST_SOURCES) compiled under
the "fast" contract (-DPLAIN_FAST_CONTRACT -fmad=true -ftz=true -prec-sqrt=false -> SFU approximations + contraction); the
integer / LUT / rasterisation / bake kernels and the host side are the very same objects as in the exact library.
"""
import os
import platform
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
OUT = PKG / "_build"
LIB = PKG / "libplain_b200.so"
LIB_FAST = PKG / "libplain_b200_fast.so"
FAST_SOURCES = ("passes_gi", "passes_post", "passes_shading", "passes_volumetrics")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")
INCLUDES = ["-I%s" % (ROOT / "include"), "-I%s" % (PKG / "csrc"), "-I%s" % (PKG / "host")]
# the contract's explicit fmaf (pvec.h / detmath.h) must be the hardware instruction on the host too: -mfma on x86-64
HOST_FMA = ["-mfma"] if platform.machine() in ("x86_64", "AMD64") else []
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false", "--extended-lambda", "--expt-relaxed-constexpr",
              "-Xcompiler", ",".join(["-fPIC", "-fvisibility=hidden", "-ffp-contract=off"] + HOST_FMA), "-Xptxas", "-v", "-diag-suppress", "177,550"]
# fast contract: contraction, flush-to-zero, approximate sqrt, SFU transcendentals (detmath.h) - but IEEE division and correctly rounded
# reciprocals stay (texel selection at uv = iUV / size sits exactly on texel borders)
NVCC_FLAGS_FAST = [f for f in NVCC_FLAGS if f != "-fmad=false"] + ["-DPLAIN_FAST_CONTRACT", "-fmad=true", "-ftz=true", "-prec-sqrt=false", "-prec-div=true"]
CXX_FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wno-unused-function"] + HOST_FMA


def _newer(src, deps, obj):
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in [src] + deps)


def build(force=False, jobs=None, verbose=False):
    OUT.mkdir(exist_ok=True)
    headers = [str(p) for p in list((ROOT / "include").glob("*.h")) + list((PKG / "csrc").glob("*.h")) + list((PKG / "csrc").glob("*.cuh")) + list((PKG / "host").glob("*.h"))]
    tasks = []
    for src in sorted((PKG / "csrc").glob("*.cu")):
        obj = OUT / (src.stem + ".o")
        tasks.append((src, obj, [NVCC] + NVCC_FLAGS + INCLUDES + ["-c", str(src), "-o", str(obj)]))
        if src.stem in FAST_SOURCES:
            obj = OUT / (src.stem + "_fast.o")
            tasks.append((src, obj, [NVCC] + NVCC_FLAGS_FAST + INCLUDES + ["-c", str(src), "-o", str(obj)]))
    for src in sorted((PKG / "host").glob("*.cpp")):
        obj = OUT / ("host_" + src.stem + ".o")
        tasks.append((src, obj, [CXX] + CXX_FLAGS + INCLUDES + ["-c", str(src), "-o", str(obj)]))
    todo = [t for t in tasks if force or _newer(t[0], headers, t[1])]

    def run(t):
        r = subprocess.run(t[2], capture_output=True, text=True)
        log = OUT / (t[1].stem + ".log")
        log.write_text(r.stdout + r.stderr)  # ptxas -v resource usage per kernel lands here
        if r.returncode != 0:
            raise RuntimeError("compile failed: %s\n%s" % (" ".join(t[2]), (r.stdout + r.stderr)[-6000:]))
        if verbose:
            print("built", t[1].name)

    with ThreadPoolExecutor(max_workers=jobs or min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(run, todo))
    exact = [str(t[1]) for t in tasks if not t[1].stem.endswith("_fast")]
    fast = [str(t[1]) for t in tasks if t[1].stem.endswith("_fast") or t[1].stem not in FAST_SOURCES]
    for lib, objs in ((LIB, exact), (LIB_FAST, fast)):
        if todo or not lib.exists():
            cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib)] + objs + ["-Xcompiler", "-fPIC", "-lpthread"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose=True)
    print(lib)
