"""ctypes binding of oracle/_ref/libptap_ref.so (the reference's own Renderer.cpp/Scene.cpp
compiled for the host, see oracle/build_ref.sh).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libptap_ref.so")
REFERENCE_ROOT = os.environ.get("PTAP_REFERENCE", "/root/reference")

# numpy views of the reference's PODs (Primitive.h:23-178); sizes checked against the library.
MATERIAL = np.dtype([("type", "<i4"), ("refractive_index", "<f4"), ("reflectivity", "<f4"), ("color", "<f4", 3)])
MODEL = np.dtype([("grid_index", "<i4"), ("mesh_index", "<i4"), ("model_to_world", "<f4", 16),
                  ("world_to_model", "<f4", 16), ("mat", MATERIAL)])
MESH = np.dtype([("v_start", "<i4"), ("v_end", "<i4"), ("t_start", "<i4"), ("t_end", "<i4"),
                 ("bb_min", "<f4", 3), ("bb_max", "<f4", 3)])
VERTEX = np.dtype([("position", "<f4", 3), ("normal", "<f4", 3), ("uv", "<f4", 2)])
TRIANGLE = np.dtype([("v", "<i4", 3)])
GRID = np.dtype([("v_start", "<i4"), ("v_end", "<i4"), ("width", "<f4", 3), ("entity_type", "<i4"), ("entity_index", "<i4")])
VOXEL = np.dtype([("start", "<i4"), ("end", "<i4"), ("entity_type", "<i4")])
RAY = np.dtype([("orig", "<f4", 3), ("dir", "<f4", 3), ("t_orig", "<f4", 3), ("t_dir", "<f4", 3), ("inv_dir", "<f4", 3),
                ("ipixel", "<i4"), ("remaining_bounces", "<i4"), ("color", "<f4", 3)])
HIT = np.dtype([("dist", "<f4"), ("normal", "<f4", 3), ("mat", MATERIAL), ("ipixel", "<i4")])
REFHIT = np.dtype([("model", "<i4"), ("tri", "<i4"), ("t_model", "<f4"), ("dist", "<f4"), ("u", "<f4"), ("v", "<f4"),
                   ("normal", "<f4", 3), ("mat_type", "<i4")])
PROBE = np.dtype([("model", "<i4"), ("tri", "<i4"), ("t_model", "<f4"), ("u", "<f4"), ("v", "<f4")])
SCENE_DTYPES = [MODEL, MESH, VERTEX, TRIANGLE, GRID, VOXEL, np.dtype("<i4")]
SCENE_NAMES = ["models", "meshes", "vertices", "triangles", "grids", "voxels", "refs"]

FLOAT_MAX = np.float32(9999999.0)

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run oracle/build_ref.sh where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.ref_scene_builtin.restype = vp; L.ref_scene_builtin.argtypes = [C.c_char_p]
        L.ref_scene_from_arrays.restype = vp
        L.ref_scene_from_arrays.argtypes = [vp, ci, vp, ci, vp, ci, vp, ci]
        L.ref_scene_free.argtypes = [vp]
        L.ref_scene_counts.argtypes = [vp, vp]
        L.ref_scene_get.argtypes = [vp, ci, vp]
        L.ref_scene_set_models.argtypes = [vp, vp, ci]
        L.ref_renderer_create.restype = vp; L.ref_renderer_create.argtypes = [vp, ci, ci, ci]
        L.ref_renderer_free.argtypes = [vp]
        for name in ("ref_init_image", "ref_generate", "ref_gather"):
            getattr(L, name).argtypes = [vp]
        L.ref_trace_step.argtypes = [vp, ci]
        L.ref_trace_step_r1.argtypes = [vp]
        L.ref_shade_step.argtypes = [vp, ci]
        L.ref_compact_step.argtypes = [vp]; L.ref_compact_step.restype = ci
        L.ref_nrays.argtypes = [vp]; L.ref_nrays.restype = ci
        for name in ("ref_get_rays", "ref_set_rays", "ref_get_hits", "ref_set_hits", "ref_get_probe"):
            getattr(L, name).argtypes = [vp, vp, ci]
        L.ref_get_image.argtypes = [vp, vp]
        L.ref_render_loop.argtypes = [vp, ci]; L.ref_render_loop.restype = C.c_double
        L.ref_write_bmp.argtypes = [vp, C.c_char_p, ci]; L.ref_write_bmp.restype = ci
        L.ref_trace.argtypes = [vp, vp, ci, ci, vp]; L.ref_trace.restype = ci
        L.ref_util_hash.argtypes = [C.c_uint]; L.ref_util_hash.restype = C.c_uint
        L.ref_rng_u01.argtypes = [ci, ci, ci, ci, vp]
        L.ref_scatter.argtypes = [ci, vp, vp, ci, ci, ci, vp]
        L.ref_set_grid.argtypes = [ci, ci, ci]
        L.ref_set_threads.argtypes = [ci]
        L.ref_max_threads.restype = ci
        for i, dt in enumerate(SCENE_DTYPES + [RAY, HIT]):
            assert L.ref_sizeof(i) == dt.itemsize, (i, L.ref_sizeof(i), dt.itemsize)
        assert L.ref_sizeof(11) == REFHIT.itemsize
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class RefScene:
    """The reference's `Scene` (Scene.h:21-39)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference scene construction failed")
        self.h = handle

    @classmethod
    def builtin(cls, root: str | None = None) -> "RefScene":
        root = root or os.path.join(REFERENCE_ROOT, "PathTracerAP")
        return cls(lib().ref_scene_builtin(root.encode()))

    @classmethod
    def from_arrays(cls, models, meshes, vertices, triangles) -> "RefScene":
        models = np.ascontiguousarray(models, MODEL); meshes = np.ascontiguousarray(meshes, MESH)
        vertices = np.ascontiguousarray(vertices, VERTEX); triangles = np.ascontiguousarray(triangles, TRIANGLE)
        return cls(lib().ref_scene_from_arrays(_ptr(models), len(models), _ptr(meshes), len(meshes),
                                               _ptr(vertices), len(vertices), _ptr(triangles), len(triangles)))

    def arrays(self) -> dict:
        counts = np.zeros(7, np.int32)
        lib().ref_scene_counts(self.h, _ptr(counts))
        out = {}
        for i, (name, dt) in enumerate(zip(SCENE_NAMES, SCENE_DTYPES)):
            a = np.zeros(int(counts[i]), dt)
            if len(a):
                lib().ref_scene_get(self.h, i, _ptr(a))
            out[name] = a
        return out

    def set_models(self, models):
        models = np.ascontiguousarray(models, MODEL)
        lib().ref_scene_set_models(self.h, _ptr(models), len(models))

    def close(self):
        if self.h:
            lib().ref_scene_free(self.h)
            self.h = None


class RefRenderer:
    """The reference's `Renderer` (Renderer.h:46-55) with its render loop opened up launch by launch."""

    def __init__(self, scene: RefScene, W: int, H: int, depth: int = 5):
        assert (W * H) % 32 == 0, "the reference launches ceil(N/32) with integer division (Renderer.cpp:573)"
        self.W, self.H, self.depth, self.N = W, H, depth, W * H
        self.h = lib().ref_renderer_create(scene.h, W, H, depth)

    def init_image(self): lib().ref_init_image(self.h)
    def generate(self): lib().ref_generate(self.h)
    def trace(self, probe: bool = False): lib().ref_trace_step(self.h, int(probe))
    def trace_r1(self): lib().ref_trace_step_r1(self.h)      # tier R1: every triangle, the reference's predicate (probe always recorded)
    def shade(self, it: int): lib().ref_shade_step(self.h, it)
    def compact(self) -> int: return lib().ref_compact_step(self.h)
    def gather(self): lib().ref_gather(self.h)
    @property
    def nrays(self) -> int: return lib().ref_nrays(self.h)

    def rays(self, n=None) -> np.ndarray:
        n = self.N if n is None else n
        a = np.zeros(n, RAY); lib().ref_get_rays(self.h, _ptr(a), n); return a

    def hits(self, n=None) -> np.ndarray:
        n = self.N if n is None else n
        a = np.zeros(n, HIT); lib().ref_get_hits(self.h, _ptr(a), n); return a

    def probe(self, n=None) -> np.ndarray:
        n = self.N if n is None else n
        a = np.zeros(n, PROBE); lib().ref_get_probe(self.h, _ptr(a), n); return a

    def image(self) -> np.ndarray:
        a = np.zeros((self.H, self.W, 3), np.float32); lib().ref_get_image(self.h, _ptr(a)); return a

    def render_loop(self, iters: int) -> float:
        """The reference's own Renderer::renderLoop, untouched; returns wall seconds."""
        return lib().ref_render_loop(self.h, iters)

    def write_bmp(self, directory: str, iters: int):
        rc = lib().ref_write_bmp(self.h, directory.encode(), iters)
        if rc != 0:
            raise RuntimeError(f"ref_write_bmp failed: {rc}")

    def trace_rays(self, rays_od: np.ndarray, mode: int = 0) -> np.ndarray:
        """Closest hit of the reference for an arbitrary ray set; mode 0 = R0 grid walk, 1 = R1 brute force."""
        rays_od = np.ascontiguousarray(rays_od, np.float32).reshape(-1, 6)
        out = np.zeros(len(rays_od), REFHIT)
        rc = lib().ref_trace(self.h, _ptr(rays_od), len(rays_od), mode, _ptr(out))
        if rc != 0:
            raise RuntimeError(f"ref_trace failed: {rc}")
        return out

    def run_iteration(self, it: int, probe: bool = False, on_bounce=None, mode: int = 0):
        """One iteration of the loop at Renderer.cpp:582-640 without the first-hit cache; returns active rays per bounce.
        mode 1 swaps the closest-hit launch for the R1 tier (brute force with the reference's predicate)."""
        self.generate()
        counts = []
        b = 0
        while self.nrays > 0:
            counts.append(self.nrays)
            if mode == 1:
                self.trace_r1()
            else:
                self.trace(probe)
            if on_bounce is not None:
                on_bounce(b, self)
            self.shade(it)
            self.compact()
            b += 1
        self.gather()
        return counts

    def close(self):
        if self.h:
            lib().ref_renderer_free(self.h)
            self.h = None


def util_hash(a: int) -> int:
    return lib().ref_util_hash(C.c_uint(a & 0xFFFFFFFF))


def rng_u01(it: int, index: int, depth: int, n: int) -> np.ndarray:
    out = np.zeros(n, np.float32); lib().ref_rng_u01(it, index, depth, n, _ptr(out)); return out


def scatter(kind: int, normal, direction, it: int, index: int, depth: int) -> np.ndarray:
    n = np.ascontiguousarray(normal, np.float32); d = np.ascontiguousarray(direction, np.float32)
    out = np.zeros(3, np.float32)
    lib().ref_scatter(kind, _ptr(n), _ptr(d), it, index, depth, _ptr(out))
    return out
