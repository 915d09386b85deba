"""Closest-hit parity on the GPU, through the C ABI (ptap_trace): BASELINE.json's contract is bit-exact primitive ids
on a fixed seeded ray set and t / barycentrics within 1e-5 relative (we expect, and assert, bit equality)."""
import numpy as np
import pytest

from conftest import have_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="no CUDA device")]

FLOAT_MAX = np.float32(9999999.0)
REL_TOL = 1e-5      # BASELINE.json north_star: "hit t and barycentrics must agree to 1e-5 relative"


def assert_hits_equal(got, want, what):
    assert np.array_equal(got["model"], want["model"]), f"{what}: model ids differ on {(got['model'] != want['model']).sum()} rays"
    assert np.array_equal(got["tri"], want["tri"]), f"{what}: triangle ids differ on {(got['tri'] != want['tri']).sum()} rays"
    hit = want["model"] >= 0
    for f in ("t_model", "dist", "u", "v"):
        a, b = got[f][hit].astype(np.float64), want[f][hit].astype(np.float64)
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
        assert rel.max(initial=0.0) <= REL_TOL, f"{what}: {f} differs by {rel.max()}"
        assert np.array_equal(got[f][hit], want[f][hit]), f"{what}: {f} not bit-equal"
    assert np.array_equal(got["normal"][hit], want["normal"][hit]), f"{what}: normals not bit-equal"
    assert np.array_equal(got["mat_type"], want["mat_type"])
    assert (got["dist"][~hit] >= FLOAT_MAX).all()


@pytest.fixture(scope="module")
def renderer(gpu_scene):
    from pathtracerap_b200 import Renderer
    r = Renderer(width=64, height=32, depth=5)
    r.allocateOnGPU(gpu_scene)
    yield r
    r.free()


def test_grid_compat_bit_exact_vs_reference_golden(renderer, golden_trace):
    from pathtracerap_b200 import ACCEL_GRID_COMPAT
    renderer.set_accel(ACCEL_GRID_COMPAT)
    got = renderer.trace(golden_trace["rays"])
    assert_hits_equal(got, golden_trace["r0"], "grid-compat vs reference R0 (golden)")


def test_bvh_bit_exact_vs_brute_force_golden(renderer, golden_trace):
    from pathtracerap_b200 import ACCEL_BVH
    renderer.set_accel(ACCEL_BVH)
    got = renderer.trace(golden_trace["rays"])
    assert_hits_equal(got, golden_trace["r1"], "BVH vs reference brute force R1 (golden)")


def _random_rays(n, seed):
    rs = np.random.RandomState(seed)
    # origins inside the box interior, directions uniform; plus axis-parallel and zero-component directions (slab special cases)
    o = np.stack([rs.uniform(-450, 500, n), rs.uniform(-100, 850, n), rs.uniform(-450, 900, n)], 1)
    d = rs.randn(n, 3)
    d[: n // 50, 0] = 0.0
    d[n // 50: n // 25, 1] = 0.0
    d[n // 25: 3 * n // 50] = np.eye(3)[rs.randint(0, 3, 3 * n // 50 - n // 25)] * rs.choice([-1.0, 1.0], (3 * n // 50 - n // 25, 1))
    d *= rs.uniform(0.1, 30.0, (n, 1))            # Ray::base.dir is not normalised (Renderer.cpp:548)
    return np.concatenate([o, d], 1).astype(np.float32)


@pytest.mark.parametrize("seed", [1, 2])
def test_random_rays_vs_oracle(renderer, oracle_scene, seed):
    from pathtracerap_b200 import ACCEL_BVH, ACCEL_GRID_COMPAT
    rays = _random_rays(200_000, seed)
    renderer.set_accel(ACCEL_GRID_COMPAT)
    assert_hits_equal(renderer.trace(rays), oracle_scene.trace(rays, 0), "grid-compat vs oracle R0 (random rays)")
    renderer.set_accel(ACCEL_BVH)
    assert_hits_equal(renderer.trace(rays), oracle_scene.trace(rays, 1), "BVH vs oracle R1 (random rays)")


def test_edge_cases(renderer, oracle_scene):
    from pathtracerap_b200 import ACCEL_BVH, ACCEL_GRID_COMPAT
    # empty ray set; a single ray; rays starting outside everything and pointing away (all-miss); rays inside a bbox
    for accel, mode in ((ACCEL_GRID_COMPAT, 0), (ACCEL_BVH, 1)):
        renderer.set_accel(accel)
        assert len(renderer.trace(np.zeros((0, 6), np.float32))) == 0
        away = np.array([[0, 0, 5000, 0, 0, 1], [0, 5000, 0, 0, 1, 0], [9000, 0, 0, 1, 0, 0]], np.float32)
        got = renderer.trace(away)
        assert (got["model"] == -1).all() and (got["tri"] == -1).all()
        assert_hits_equal(got, oracle_scene.trace(away, mode), "all-miss")
        one = np.array([[0, 0, 920, 0.5, 0.25, -20]], np.float32)
        assert_hits_equal(renderer.trace(one), oracle_scene.trace(one, mode), "single ray")
        n = 33  # ragged: not a multiple of the warp or block size
        rays = _random_rays(n, 5)
        assert_hits_equal(renderer.trace(rays), oracle_scene.trace(rays, mode), "ragged count")


def test_traversal_counts(renderer, golden_trace):
    from pathtracerap_b200 import ACCEL_BVH, ACCEL_GRID_COMPAT
    rays = golden_trace["rays"]
    renderer.set_accel(ACCEL_GRID_COMPAT)
    hits, cnt = renderer.trace(rays, counts=True)
    assert np.array_equal(hits["tri"], golden_trace["r0"]["tri"])
    cells, refs, tris = cnt[:, 0].mean(), cnt[:, 1].mean(), cnt[:, 2].mean()
    assert refs == tris and 10 < cells < 100 and 5 < tris < 100          # ~40 cells, ~32 tests per ray
    renderer.set_accel(ACCEL_BVH)
    hits, cnt = renderer.trace(rays, counts=True)
    assert np.array_equal(hits["tri"], golden_trace["r1"]["tri"])
    assert cnt[:, 2].mean() < tris                                        # the BVH tests fewer triangles than the 25^3 grid
