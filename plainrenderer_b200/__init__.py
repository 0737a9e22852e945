"""plainrenderer_b200 - B200-native frame path of Gaukler/PlainRenderer behind the reference's RenderBackend interface.

The product is `libplain_b200.so` (CUDA kernels for sm_100a + the C-ABI of include/plain_b200.h + the host-side
RenderFrontend mirror). This package only builds it in-tree and binds it with ctypes. There is no CPU path: `load()`
raises if the library is missing, and `plain_backend_create` fails without a CUDA device.
"""
from pathlib import Path

from . import ffi

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libplain_b200.so"
LIB_FAST_PATH = PKG_DIR / "libplain_b200_fast.so"  # same C-ABI, floating-point passes under the "fast" contract
_api = {}


def build(force=False):
    import importlib
    return importlib.import_module(__name__ + ".buildlib").build(force=force)


def load(contract="exact"):
    """Binds the product library (symbols plain_* / plain_frontend_*). Raises if it has not been built.
    contract: "exact" (default; bit-exact against the oracle) or "fast" (SFU approximations + contraction in the floating-point passes)."""
    path = {"exact": LIB_PATH, "fast": LIB_FAST_PATH}[contract]
    if contract not in _api:
        if not path.exists():
            raise RuntimeError("%s is missing - run `python -m plainrenderer_b200.buildlib` (there is no CPU fallback)" % path)
        _api[contract] = ffi.Api(path, "plain_", "plain_frontend_")
    return _api[contract]
