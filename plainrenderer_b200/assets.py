"""ctypes binding of include/plain_assets.h: .plain scenes, R16F .dds bricks (N1) and the SDF bake (N2).

    lib = assets.Assets()                      # product library (CUDA bake); raises if libplain_b200.so is missing
    scene = lib.load_scene("model.plain")      # -> Scene(objects, meshes)
    texels = lib.bake(mesh)                    # uint16 half floats, shape (d, h, w)
The CPU oracle is bound the same way by the tests: Assets(path_to_liboracle, "oracle_asset_")."""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np


class MeshInfo(C.Structure):
    _fields_ = [("index_count", C.c_uint32), ("vertex_count", C.c_uint32), ("bb_min", C.c_float * 3), ("bb_max", C.c_float * 3), ("mean_albedo", C.c_float * 3),
                ("albedo_path", C.c_char * 256), ("normal_path", C.c_char * 256), ("specular_path", C.c_char * 256), ("sdf_path", C.c_char * 256)]


class SceneObject(C.Structure):
    _fields_ = [("model_matrix", C.c_float * 16), ("mesh_index", C.c_uint64)]


@dataclass
class Mesh:
    positions: np.ndarray   # (n, 3) float32
    indices: np.ndarray     # (m,) uint32
    bb_min: np.ndarray
    bb_max: np.ndarray
    mean_albedo: np.ndarray
    paths: dict = field(default_factory=dict)
    vertices: np.ndarray = None  # (n, 28) uint8: the vertex buffer as stored


@dataclass
class Scene:
    objects: list           # [(model_matrix (4, 4) column-major as float32[16], mesh_index)]
    meshes: list


class AssetError(RuntimeError):
    pass


SYMBOLS = ["last_error", "scene_load", "scene_destroy", "scene_counts", "scene_object", "scene_mesh_info", "scene_mesh_geometry", "scene_mesh_vertices", "dds_r16f_info", "dds_r16f_load",
           "dds_r16f_save", "sdf_resolution", "sdf_bake"]


class Assets:
    def __init__(self, lib_path=None, prefix="plain_asset_"):
        if lib_path is None:
            from . import LIB_PATH
            if not LIB_PATH.exists():
                raise AssetError("libplain_b200.so has not been built (python -m plainrenderer_b200.buildlib)")
            lib_path = LIB_PATH
        self.lib = C.CDLL(str(lib_path))
        self.f = {n: getattr(self.lib, prefix + n) for n in SYMBOLS}
        self.f["last_error"].restype = C.c_char_p
        self.f["scene_destroy"].restype = None
        self.f["sdf_resolution"].restype = None

    def _check(self, rc, what):
        if rc:
            raise AssetError("%s: %s" % (what, (self.f["last_error"]() or b"").decode() or "failed"))

    def load_scene(self, path):
        p = C.c_void_p()
        self._check(self.f["scene_load"](str(path).encode(), C.byref(p)), "scene_load")
        try:
            no, nm = C.c_uint64(), C.c_uint64()
            self.f["scene_counts"](p, C.byref(no), C.byref(nm))
            objects, meshes = [], []
            for i in range(no.value):
                o = SceneObject()
                self._check(self.f["scene_object"](p, C.c_uint64(i), C.byref(o)), "scene_object")
                objects.append((np.array(o.model_matrix, np.float32), int(o.mesh_index)))
            for m in range(nm.value):
                info = MeshInfo()
                self._check(self.f["scene_mesh_info"](p, C.c_uint64(m), C.byref(info)), "scene_mesh_info")
                pos = np.zeros((info.vertex_count, 3), np.float32)
                idx = np.zeros(info.index_count, np.uint32)
                self._check(self.f["scene_mesh_geometry"](p, C.c_uint64(m), pos.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p)), "scene_mesh_geometry")
                vtx = np.zeros((info.vertex_count, 28), np.uint8)
                self._check(self.f["scene_mesh_vertices"](p, C.c_uint64(m), vtx.ctypes.data_as(C.c_void_p)), "scene_mesh_vertices")
                meshes.append(Mesh(pos, idx, np.array(info.bb_min, np.float32), np.array(info.bb_max, np.float32), np.array(info.mean_albedo, np.float32),
                                   {k: getattr(info, k + "_path").decode() for k in ("albedo", "normal", "specular", "sdf")}, vtx))
            return Scene(objects, meshes)
        finally:
            self.f["scene_destroy"](p)

    def load_brick(self, path):
        e = (C.c_uint32 * 3)()
        self._check(self.f["dds_r16f_info"](str(path).encode(), e), "dds_r16f_info")
        out = np.zeros((e[2], e[1], e[0]), np.uint16)
        self._check(self.f["dds_r16f_load"](str(path).encode(), out.ctypes.data_as(C.c_void_p), C.c_size_t(out.size)), "dds_r16f_load")
        return out

    def save_brick(self, path, texels):
        texels = np.ascontiguousarray(texels, np.uint16)
        d, h, w = texels.shape
        self._check(self.f["dds_r16f_save"](str(path).encode(), (C.c_uint32 * 3)(w, h, d), texels.ctypes.data_as(C.c_void_p)), "dds_r16f_save")

    def resolution(self, bb_min, bb_max):
        e = (C.c_uint32 * 3)()
        self.f["sdf_resolution"]((C.c_float * 3)(*bb_min), (C.c_float * 3)(*bb_max), e)
        return e[0], e[1], e[2]

    def bake(self, mesh, extent=None, device=0):
        """SDF brick of a mesh: uint16 half floats (d, h, w) and the device time of the kernel in ms."""
        w, h, d = extent or self.resolution(mesh.bb_min, mesh.bb_max)
        pos = np.ascontiguousarray(mesh.positions, np.float32)
        idx = np.ascontiguousarray(mesh.indices, np.uint32)
        out = np.zeros((d, h, w), np.uint16)
        ms = C.c_float(0)
        rc = self.f["sdf_bake"](C.c_int(device), pos.ctypes.data_as(C.c_void_p), C.c_uint32(len(pos)), idx.ctypes.data_as(C.c_void_p), C.c_uint32(len(idx)),
                                (C.c_float * 3)(*mesh.bb_min), (C.c_float * 3)(*mesh.bb_max), (C.c_uint32 * 3)(w, h, d), out.ctypes.data_as(C.c_void_p), C.byref(ms))
        if rc:
            raise AssetError("sdf_bake failed (no CUDA device? the bake has no CPU fallback)")
        return out, ms.value
