"""Closest-hit parity on scenes that stress the two-level BVH: many instances with arbitrary (non-uniform, rotated) transforms,
model matrices that are NOT inverses of each other (the reference applies them literally), empty and single-triangle meshes."""
import numpy as np
import pytest

from conftest import have_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="no CUDA device")]


def _rot(axis, deg):
    a = np.asarray(axis, np.float64); a /= np.linalg.norm(a)
    t = np.deg2rad(deg); c, s = np.cos(t), np.sin(t)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) * c + s * K + (1 - c) * np.outer(a, a)


def _trs(translate, R, scale):
    M = np.eye(4)
    M[:3, :3] = R @ np.diag(scale)
    M[:3, 3] = translate
    return M


def _rays(n, seed, lo, hi):
    rs = np.random.RandomState(seed)
    o = rs.uniform(lo, hi, (n, 3))
    tgt = rs.uniform(lo * 0.6, hi * 0.6, (n, 3))
    d = tgt - o
    d *= rs.uniform(0.2, 5.0, (n, 1)) / np.linalg.norm(d, axis=1, keepdims=True)
    d[: n // 40, 0] = 0.0
    d[n // 40: n // 20, 2] = 0.0
    return np.concatenate([o, d], 1).astype(np.float32)


def _assert_equal(got, want, what):
    assert np.array_equal(got["model"], want["model"]), f"{what}: model ids differ on {(got['model'] != want['model']).sum()} rays"
    assert np.array_equal(got["tri"], want["tri"]), f"{what}: triangle ids differ"
    hit = want["model"] >= 0
    for f in ("t_model", "dist", "u", "v"):
        assert np.array_equal(got[f][hit], want[f][hit]), f"{what}: {f} not bit-equal"
    assert np.array_equal(got["normal"][hit], want["normal"][hit])
    return hit


def _scene_and_oracle(port, models, meshes, vertices, triangles):
    from pathtracerap_b200 import ACCEL_BVH, Renderer, Scene
    s = Scene.from_arrays(models, meshes, vertices, triangles)
    s.build_bvh()
    r = Renderer(width=32, height=32, depth=5, accel=ACCEL_BVH)
    r.allocateOnGPU(s)
    o = port.OracleScene({"models": models, "meshes": meshes, "vertices": vertices, "triangles": triangles})
    r._scene = s
    return r, o


def _check_grid_tiers(r, o, rays, what):
    """The same scene through the reference's grids (built on the device): walked and emulated through the BVH, both against the oracle's R0."""
    from pathtracerap_b200 import ACCEL_GRID_EMULATED
    want = o.trace(rays, 0)
    r.build_grids_device(r._scene, 25, 25, 25)
    _assert_equal(r.trace(rays), want, what + ": grid walk vs oracle R0")
    r.set_accel(ACCEL_GRID_EMULATED)
    _assert_equal(r.trace(rays), want, what + ": emulated walk vs oracle R0")


def test_many_instances_arbitrary_transforms(libptap, port, golden_scene):
    """96 instances of the three bundled meshes, rotated about random axes, non-uniformly scaled, overlapping: exercises a TLAS several
    levels deep, the per-instance pruning bound under anisotropic scale, and cross-instance ties."""
    from pathtracerap_b200 import MODEL
    g = golden_scene
    rs = np.random.RandomState(11)
    models = np.zeros(96, MODEL)
    for k in range(96):
        mesh = [2, 1, 0][k % 3] if k % 7 else 1
        sc = rs.uniform(0.02, 0.12, 3) * (0.3 if mesh == 0 else 1.0)
        if k % 5 == 0:
            sc[:] = sc[0]
        M = _trs(rs.uniform(-600, 600, 3), _rot(rs.randn(3), rs.uniform(0, 360)), sc)
        models[k]["mesh_index"] = mesh
        models[k]["model_to_world"] = M.T.astype(np.float32).reshape(16)              # column-major
        models[k]["world_to_model"] = np.linalg.inv(M.astype(np.float32).astype(np.float64)).T.astype(np.float32).reshape(16)
        models[k]["mat"]["type"] = k % 7
        models[k]["mat"]["color"] = rs.uniform(0.1, 0.99, 3)
    # two exactly coincident instances: equal distances, the lower model index must win
    models[95] = models[3]
    r, o = _scene_and_oracle(port, models, g["meshes"], g["vertices"], g["triangles"])
    rays = _rays(150_000, 5, -900.0, 900.0)
    got, want = r.trace(rays), o.trace(rays, 1)
    hit = _assert_equal(got, want, "96 instances")
    assert 0.3 < hit.mean() < 0.999
    assert (want["model"] == 3).sum() > 100 and (want["model"] == 95).sum() == 0
    assert len(np.unique(want["model"][hit])) > 60
    _check_grid_tiers(r, o, rays[:60_000], "96 instances")
    r.free()


def test_matrices_that_are_not_inverses(libptap, port, golden_scene):
    """The reference maps the ray with world_to_model and the hit back with model_to_world, whatever their relation (Renderer.cpp:381-391).
    With inconsistent pairs the distance-based pruning is invalid; the library must detect that and still match brute force."""
    g = golden_scene
    models = g["models"].copy()
    rs = np.random.RandomState(2)
    for k in (0, 4, 7):
        M = models[k]["model_to_world"].reshape(4, 4).T.astype(np.float64)
        M[:3, :3] *= rs.uniform(0.7, 1.4)                   # model_to_world rescaled: world distances no longer follow t
        M[:3, 3] += rs.uniform(-30, 30, 3)
        models[k]["model_to_world"] = M.T.astype(np.float32).reshape(16)
    r, o = _scene_and_oracle(port, models, g["meshes"], g["vertices"], g["triangles"])
    rays = _rays(100_000, 6, -450.0, 850.0)
    _assert_equal(r.trace(rays), o.trace(rays, 1), "inconsistent matrices")
    _check_grid_tiers(r, o, rays, "inconsistent matrices")
    r.free()


def test_empty_and_tiny_meshes(libptap, port, golden_scene):
    from pathtracerap_b200 import MESH, MODEL, TRIANGLE, VERTEX
    g = golden_scene
    # mesh 3: one triangle; mesh 4: empty (a failed load leaves such a mesh behind, Scene.cpp:231-235)
    nv, nt = len(g["vertices"]), len(g["triangles"])
    v = np.zeros(3, VERTEX)
    v["position"] = [[-800, -800, 0], [800, -800, 0], [0, 900, 0]]
    v["normal"] = [[0, 0, 1000.0]] * 3
    vertices = np.concatenate([g["vertices"], v])
    t = np.zeros(1, TRIANGLE); t["v"] = [[nv, nv + 1, nv + 2]]
    triangles = np.concatenate([g["triangles"], t])
    meshes = np.zeros(5, MESH)
    meshes[:3] = g["meshes"]
    meshes[3] = (nv, nv + 3, nt, nt + 1, (-800, -800, 0), (800, 900, 0))
    meshes[4] = (nv + 3, nv + 3, nt + 1, nt + 1, (9999999.0,) * 3, (-9999990.0,) * 3)
    models = np.zeros(4, MODEL)
    models[0] = g["models"][3]
    for k, (mesh, tr) in enumerate([(3, (0, 300, -100)), (4, (0, 0, 0)), (3, (50, 320, 150))], start=1):
        M = _trs(tr, _rot((0, 1, 0), 20.0 * k), (0.3, 0.3, 0.3))
        models[k]["mesh_index"] = mesh
        models[k]["model_to_world"] = M.T.astype(np.float32).reshape(16)
        models[k]["world_to_model"] = np.linalg.inv(M.astype(np.float32).astype(np.float64)).T.astype(np.float32).reshape(16)
        models[k]["mat"]["color"] = (0.5, 0.5, 0.5)
    r, o = _scene_and_oracle(port, models, meshes, vertices, triangles)
    rays = _rays(50_000, 8, -450.0, 850.0)
    want = o.trace(rays, 1)
    _assert_equal(r.trace(rays), want, "tiny meshes")
    assert (want["model"] == 1).any() and (want["model"] == 3).any() and not (want["model"] == 2).any()
    _check_grid_tiers(r, o, rays, "tiny meshes")
    r.free()


CONFIG_SCENE = """
RESOLUTION
[96, 64]

ITER
3

DEPTH
4

DIFFUSE
wall
[0.8, 0.7, 0.6]

EMISSIVE
lamp
[5.0, 5.0, 5.0]

BOX
room
[2.0, 2.0, 2.0]
[-2.0, -2.0, -2.0]
translate:[0, 4, 890]
rotateX:[0, 0, 0]
scale:[0.03, 0.03, 0.03]
material: wall

SPHERE
ball
1.0
[0, 0, 0]
translate:[2, 2, 890]
scale:[0.004, 0.004, 0.004]
material: lamp

MESH
plate
plate.obj
translate:[-3, 3, 892]
rotateX:[40, 0, 0]
scale:[0.005, 0.005, 0.005]
material: wall
"""


def test_config_txt_scene_end_to_end(libptap, port, tmp_path):
    """Config.txt -> Scene (parser, OBJ reader, box / sphere generators, grid builder) -> Renderer, against the oracle rendering the SAME
    arrays: closest hits of the frame's own primary and bounce rays bit-equal (grid walk vs R0, BVH vs R1), film within the frame bound.
    (The reference reads no Config.txt, SURVEY 8f row 1: what is pinned here is that whatever the parser builds is rendered exactly as
    the reference's kernels would render it.)"""
    from pathtracerap_b200 import ACCEL_BVH, ACCEL_GRID_COMPAT, Renderer, Scene
    (tmp_path / "plate.obj").write_text("v -1 0 -1\nv 1 0 -1\nv 1 0 1\nv -1 0 1\nvn 0 1 0\nf 1//1 2//1 3//1\nf 1//1 3//1 4//1\n")
    cfg = tmp_path / "Config.txt"
    cfg.write_text(CONFIG_SCENE)
    s = Scene(str(cfg))
    p = s.config_params()
    W, H, iters, depth = p["W"], p["H"], p["iters"], p["depth"]
    assert (W, H, iters, depth) == (96, 64, 3, 4)
    a = s.arrays()
    assert len(a["models"]) == 3 and len(a["grids"]) == 3
    osc = port.OracleScene({k: a[k] for k in ("models", "meshes", "vertices", "triangles")})
    assert osc.arrays()["voxels"].tobytes() == a["voxels"].tobytes() and osc.arrays()["refs"].tobytes() == a["refs"].tobytes()
    # the oracle's own wavefront of iteration 0: every ray it traces, bounce by bounce
    w = port.OracleWavefront(osc, W, H, depth)
    w.init_image()
    sets = []
    w.run_iteration(0, on_bounce=lambda b, ww: sets.append(np.concatenate([ww.rays(ww.nrays)["orig"], ww.rays(ww.nrays)["dir"]], 1)))
    rays = np.concatenate(sets).astype(np.float32)
    assert len(sets) >= 2 and len(rays) > W * H
    r = Renderer(width=W, height=H, depth=depth, first_hit_cache=True)
    r.allocateOnGPU(s)
    for accel, mode in ((ACCEL_GRID_COMPAT, 0), (ACCEL_BVH, 1)):
        r.set_accel(accel)
        hit = _assert_equal(r.trace(rays), osc.trace(rays, mode), f"config scene, accel {accel}")
        assert hit.mean() > 0.3
    # the frame itself, through the reference's grid walk
    w.init_image(); w.render(0, iters, True)
    want = w.image()
    r.set_accel(ACCEL_GRID_COMPAT)
    r.set_params(W, H, depth, first_hit_cache=True)
    r.render(0, iters)
    film = r.film()
    assert want.mean() > 0.05
    assert float(np.sqrt(np.mean((film - want) ** 2))) <= 0.005 * float(want.mean())
    assert float(np.mean(film == want)) > 0.97
    w.close(); r.free()
