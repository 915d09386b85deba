"""ctypes binding of oracle/libptap_oracle.so, the plain-C restatement of the reference's hot path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .ref import (GRID, HIT, MESH, MODEL, RAY, REFHIT, TRIANGLE, VERTEX, VOXEL, FLOAT_MAX)  # noqa: F401  (dtypes only)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libptap_oracle.so")
OHIT = REFHIT  # identical layout (model, tri, t_model, dist, u, v, normal[3], mat_type)


class _OScene(C.Structure):
    _fields_ = [("models", C.c_void_p), ("nmodels", C.c_int), ("meshes", C.c_void_p), ("nmeshes", C.c_int),
                ("vertices", C.c_void_p), ("nvertices", C.c_int), ("triangles", C.c_void_p), ("ntriangles", C.c_int),
                ("grids", C.c_void_p), ("ngrids", C.c_int), ("voxels", C.c_void_p), ("nvoxels", C.c_int),
                ("refs", C.c_void_p), ("nrefs", C.c_int), ("grid_dim", C.c_int * 3)]


_lib = None


def build(force: bool = False) -> str:
    src = [os.path.join(HERE, f) for f in ("ptap_oracle.c", "synth.c", "ptap_oracle.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-s", "-C", HERE, "libptap_oracle.so"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, ci = C.c_void_p, C.c_int
        L.oracle_hash.argtypes = [C.c_uint]; L.oracle_hash.restype = C.c_uint
        L.oracle_rng_seed.argtypes = [ci, ci, ci]; L.oracle_rng_seed.restype = C.c_uint
        L.oracle_rng_next.argtypes = [vp]; L.oracle_rng_next.restype = C.c_float
        L.oracle_scatter.argtypes = [ci, vp, vp, ci, ci, ci, vp]
        L.oracle_normal_matrix.argtypes = [vp, vp]
        L.oracle_build_grids.argtypes = [vp, ci, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp]; L.oracle_build_grids.restype = ci
        L.oracle_trace.argtypes = [vp, vp, ci, ci, vp]
        L.oracle_model_hits.argtypes = [vp, vp, ci, ci, vp, vp]; L.oracle_model_hits.restype = ci
        L.oracle_grid_path.argtypes = [vp, vp, ci, ci, vp]; L.oracle_grid_path.restype = ci
        L.oracle_hit_distance.argtypes = [vp, vp, ci, C.c_float]; L.oracle_hit_distance.restype = C.c_float
        L.oracle_wavefront_create.argtypes = [vp, ci, ci, ci]; L.oracle_wavefront_create.restype = vp
        L.oracle_wavefront_free.argtypes = [vp]
        L.oracle_wavefront_set_mode.argtypes = [vp, ci]
        for n in ("oracle_init_image", "oracle_generate", "oracle_trace_step", "oracle_gather"):
            getattr(L, n).argtypes = [vp]
        L.oracle_shade_step.argtypes = [vp, ci]
        L.oracle_compact_step.argtypes = [vp]; L.oracle_compact_step.restype = ci
        L.oracle_nrays.argtypes = [vp]; L.oracle_nrays.restype = ci
        for n in ("oracle_rays", "oracle_hits", "oracle_probe", "oracle_image"):
            getattr(L, n).argtypes = [vp]; getattr(L, n).restype = vp
        L.oracle_render.argtypes = [vp, ci, ci, ci, vp]
        L.oracle_write_bmp.argtypes = [vp, ci, ci, ci, C.c_char_p]; L.oracle_write_bmp.restype = ci
        L.oracle_set_threads.argtypes = [ci]
        L.oracle_compose_trs.argtypes = [vp, C.c_float, vp, vp, vp]
        L.oracle_icosphere_triangles.argtypes = [ci]; L.oracle_icosphere_triangles.restype = ci
        L.oracle_icosphere.argtypes = [ci, C.c_float, C.c_float, C.c_uint, vp, vp, vp]; L.oracle_icosphere.restype = ci
        L.oracle_max_threads.restype = ci
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def build_grids(models, meshes, vertices, triangles, grid_dim=(25, 25, 25)):
    """Scene::addMeshesToGrid restated (Scene.cpp:318-396). Returns (models', grids, voxels, refs)."""
    models = np.array(models, MODEL, copy=True); meshes = np.ascontiguousarray(meshes, MESH)
    vertices = np.ascontiguousarray(vertices, VERTEX); triangles = np.ascontiguousarray(triangles, TRIANGLE)
    gd = (C.c_int * 3)(*grid_dim)
    ng, nv, nr = C.c_int(), C.c_int(), C.c_int()
    L = lib()
    L.oracle_build_grids(_ptr(models), len(models), _ptr(meshes), len(meshes), _ptr(vertices), _ptr(triangles), gd,
                         None, C.byref(ng), None, C.byref(nv), None, C.byref(nr))
    grids = np.zeros(ng.value, GRID); voxels = np.zeros(nv.value, VOXEL); refs = np.zeros(max(nr.value, 1), np.int32)
    L.oracle_build_grids(_ptr(models), len(models), _ptr(meshes), len(meshes), _ptr(vertices), _ptr(triangles), gd,
                         _ptr(grids), C.byref(ng), _ptr(voxels), C.byref(nv), _ptr(refs), C.byref(nr))
    return models, grids, voxels, refs[:nr.value]


class OracleScene:
    """Holds the seven scene arrays (reference layouts) alive and exposes them as an OScene."""

    def __init__(self, arrays: dict, grid_dim=(25, 25, 25)):
        self.a = {
            "models": np.ascontiguousarray(arrays["models"], MODEL), "meshes": np.ascontiguousarray(arrays["meshes"], MESH),
            "vertices": np.ascontiguousarray(arrays["vertices"], VERTEX), "triangles": np.ascontiguousarray(arrays["triangles"], TRIANGLE),
        }
        if arrays.get("voxels") is not None and arrays.get("refs") is not None and arrays.get("grids") is not None:
            self.a["grids"] = np.ascontiguousarray(arrays["grids"], GRID)
            self.a["voxels"] = np.ascontiguousarray(arrays["voxels"], VOXEL)
            self.a["refs"] = np.ascontiguousarray(arrays["refs"], np.int32)
        else:
            m, g, v, r = build_grids(self.a["models"], self.a["meshes"], self.a["vertices"], self.a["triangles"], grid_dim)
            self.a.update(models=m, grids=g, voxels=v, refs=r)
        a = self.a
        self.c = _OScene(_ptr(a["models"]), len(a["models"]), _ptr(a["meshes"]), len(a["meshes"]),
                         _ptr(a["vertices"]), len(a["vertices"]), _ptr(a["triangles"]), len(a["triangles"]),
                         _ptr(a["grids"]), len(a["grids"]), _ptr(a["voxels"]), len(a["voxels"]),
                         _ptr(a["refs"]), len(a["refs"]), (C.c_int * 3)(*grid_dim))

    def arrays(self):
        return self.a

    # -- probes of single steps (tests/test_emulation_model.py) --------------------------------------------------------------------
    def model_hits(self, ray_od, imodel, cap=64):
        """(global triangle ids, model-space t) of every triangle of the model that the reference's predicate accepts for the ray."""
        ray = np.ascontiguousarray(ray_od, np.float32).reshape(6)
        tri = np.zeros(cap, np.int32); t = np.zeros(cap, np.float32)
        n = lib().oracle_model_hits(C.byref(self.c), _ptr(ray), imodel, cap, _ptr(tri), _ptr(t))
        assert n <= cap
        return tri[:n].copy(), t[:n].copy()

    def grid_path(self, ray_od, imodel):
        """The (n, 3) voxel indices the model's grid walk visits when nothing is hit; empty when the walk is not entered."""
        ray = np.ascontiguousarray(ray_od, np.float32).reshape(6)
        cap = int(sum(self.c.grid_dim)) + 3
        out = np.zeros((cap, 3), np.int32)
        n = lib().oracle_grid_path(C.byref(self.c), _ptr(ray), imodel, cap, _ptr(out))
        assert n <= cap
        return out[:n].copy()

    def hit_distance(self, ray_od, imodel, t) -> np.float32:
        ray = np.ascontiguousarray(ray_od, np.float32).reshape(6)
        return np.float32(lib().oracle_hit_distance(C.byref(self.c), _ptr(ray), imodel, C.c_float(float(t))))

    def trace(self, rays_od, mode=0) -> np.ndarray:
        rays_od = np.ascontiguousarray(rays_od, np.float32).reshape(-1, 6)
        out = np.zeros(len(rays_od), OHIT)
        lib().oracle_trace(C.byref(self.c), _ptr(rays_od), len(rays_od), mode, _ptr(out))
        return out


class OracleWavefront:
    """Renderer::renderLoop restated launch by launch (Renderer.cpp:567-648)."""

    def __init__(self, scene: OracleScene, W: int, H: int, depth: int = 5, mode: int = 0):
        """mode 0: the reference's grid walk (R0); mode 1: brute force over every triangle with the same predicate (R1)."""
        self.scene, self.W, self.H, self.depth, self.N = scene, W, H, depth, W * H
        self.h = lib().oracle_wavefront_create(C.byref(scene.c), W, H, depth)
        lib().oracle_wavefront_set_mode(self.h, mode)

    def init_image(self): lib().oracle_init_image(self.h)
    def generate(self): lib().oracle_generate(self.h)
    def trace(self): lib().oracle_trace_step(self.h)
    def shade(self, it): lib().oracle_shade_step(self.h, it)
    def compact(self): return lib().oracle_compact_step(self.h)
    def gather(self): lib().oracle_gather(self.h)
    @property
    def nrays(self): return lib().oracle_nrays(self.h)

    def _view(self, fn, dtype, n):
        p = getattr(lib(), fn)(self.h)
        buf = (C.c_char * (n * dtype.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype, n)

    def rays(self, n=None): return self._view("oracle_rays", RAY, self.N if n is None else n).copy()
    def hits(self, n=None): return self._view("oracle_hits", HIT, self.N if n is None else n).copy()
    def probe(self, n=None): return self._view("oracle_probe", OHIT, self.N if n is None else n).copy()
    def image(self): return self._view("oracle_image", np.dtype("<f4"), self.N * 3).reshape(self.H, self.W, 3).copy()

    def render(self, it0, it1, first_hit_cache=False) -> int:
        traced = C.c_longlong(0)
        lib().oracle_render(self.h, it0, it1, int(first_hit_cache), C.byref(traced))
        return traced.value

    def run_iteration(self, it, on_bounce=None):
        self.generate()
        counts = []
        b = 0
        while self.nrays > 0:
            counts.append(self.nrays)
            self.trace()
            if on_bounce is not None:
                on_bounce(b, self)
            self.shade(it)
            self.compact()
            b += 1
        self.gather()
        return counts

    def close(self):
        if self.h:
            lib().oracle_wavefront_free(self.h)
            self.h = None


def compose_trs(translate, rotate_y_degrees, scale):
    """(model_to_world, world_to_model) as Scene.cpp:34-39 composes them: glm::translate * rotate(Y) * scale, glm::inverse."""
    t = np.ascontiguousarray(translate, np.float32); s = np.ascontiguousarray(scale, np.float32)
    m2w = np.zeros(16, np.float32); w2m = np.zeros(16, np.float32)
    lib().oracle_compose_trs(_ptr(t), float(rotate_y_degrees), _ptr(s), _ptr(m2w), _ptr(w2m))
    return m2w, w2m


def icosphere(level, radius=1000.0, displacement=0.05, seed=1):
    """(vertices, triangles, bb_min, bb_max) of the synthetic displaced icosphere: 20 * 4^level triangles, un-joined vertices."""
    n = lib().oracle_icosphere_triangles(level)
    v = np.zeros(3 * n, VERTEX); lo = np.zeros(3, np.float32); hi = np.zeros(3, np.float32)
    got = lib().oracle_icosphere(level, radius, displacement, seed, _ptr(v), _ptr(lo), _ptr(hi))
    assert got == n
    t = np.zeros(n, TRIANGLE); t["v"] = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    return v, t, lo, hi


def scene_with_mesh(base, keep_models, mesh_vertices, mesh_triangles, bb_min, bb_max, model):
    """The four reference-layout arrays of `base` restricted to `keep_models`, plus one more mesh and one model using it."""
    models = np.concatenate([np.ascontiguousarray(base["models"], MODEL)[keep_models], np.zeros(1, MODEL)])
    meshes = np.concatenate([np.ascontiguousarray(base["meshes"], MESH), np.zeros(1, MESH)])
    nv, nt = len(base["vertices"]), len(base["triangles"])
    m = meshes[-1]
    m["v_start"], m["v_end"], m["t_start"], m["t_end"], m["bb_min"], m["bb_max"] = nv, nv + len(mesh_vertices), nt, nt + len(mesh_triangles), bb_min, bb_max
    tri = mesh_triangles.copy(); tri["v"] += nv
    mm = models[-1]
    mm["mesh_index"] = len(meshes) - 1
    mm["model_to_world"], mm["world_to_model"] = model["m2w"], model["w2m"]
    mm["mat"]["type"] = model["type"]; mm["mat"]["color"] = model["color"]
    return {"models": models, "meshes": meshes, "vertices": np.concatenate([np.ascontiguousarray(base["vertices"], VERTEX), mesh_vertices]),
            "triangles": np.concatenate([np.ascontiguousarray(base["triangles"], TRIANGLE), tri])}


def write_bmp(image_sum, iters, path):
    img = np.ascontiguousarray(image_sum, np.float32)
    H, W, _ = img.shape
    rc = lib().oracle_write_bmp(_ptr(img), W, H, iters, str(path).encode())
    if rc != 0:
        raise OSError(f"oracle_write_bmp({path}) failed")


def util_hash(a): return lib().oracle_hash(C.c_uint(a & 0xFFFFFFFF))


def rng_u01(it, index, depth, n):
    st = C.c_uint(lib().oracle_rng_seed(it, index, depth))
    return np.array([lib().oracle_rng_next(C.byref(st)) for _ in range(n)], np.float32)


def scatter(kind, normal, direction, it, index, depth):
    n = np.ascontiguousarray(normal, np.float32); d = np.ascontiguousarray(direction, np.float32)
    out = np.zeros(3, np.float32)
    lib().oracle_scatter(kind, _ptr(n), _ptr(d), it, index, depth, _ptr(out))
    return out
