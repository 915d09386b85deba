"""Closest-hit parity beyond the bundled scene: the synthetic 82k-, 1.3M- and 5.2M-triangle meshes of bench.py (BASELINE.json configs[1], configs[3]).

The oracle's brute force (tier R1) finishes a few thousand rays on these sizes in seconds; at the full ray counts the checks are
size-independent properties: a permutation of the ray set permutes the hits (no cross-ray state in the persistent, work-stealing
kernel), tracing twice is idempotent, and every reported hit re-traces to itself from a ray restarted just in front of it."""
import numpy as np
import pytest

from conftest import have_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="no CUDA device")]


def _camera_rays(W, H):
    i = np.arange(W * H)
    x, y = (i % W).astype(np.float32), (i // W).astype(np.float32)
    wx = (-10.0 + (x * np.float32(20.0 / W)).astype(np.float64)).astype(np.float32)      # Renderer.cpp:538-543
    wy = (-4.0 + (y * np.float32(16.0 / H)).astype(np.float64)).astype(np.float32)
    r = np.zeros((W * H, 6), np.float32)
    r[:, 2] = 920.0
    r[:, 3], r[:, 4], r[:, 5] = wx, wy, np.float32(900.0) - np.float32(920.0)
    return r


def _bounce_rays(n, seed):
    rs = np.random.RandomState(seed)
    o = np.stack([rs.uniform(-450, 500, n), rs.uniform(-100, 850, n), rs.uniform(-450, 900, n)], 1)
    tgt = np.array([25.0, 230.0, -50.0]) + rs.randn(n, 3) * 200.0                          # aimed around the displaced sphere
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], 1).astype(np.float32)


@pytest.fixture(scope="module", params=["mesh100k", "mesh1m", "mesh5m"])
def big(request, libptap):
    import bench
    from pathtracerap_b200 import ACCEL_BVH, Renderer
    scene, arrays = bench.build_scene(request.param)
    scene.build_bvh()
    r = Renderer(width=64, height=32, depth=5, accel=ACCEL_BVH)
    r.allocateOnGPU(scene)
    yield request.param, r, arrays
    r.free()


def test_bvh_vs_brute_force_oracle(big, port):
    name, r, arrays = big
    n = {"mesh100k": 20000, "mesh1m": 20000, "mesh5m": 2000}[name]     # brute force: n x triangles predicate evaluations on the host
    cam = _camera_rays(1920, 1080)
    rays = np.concatenate([cam[np.random.RandomState(3).choice(len(cam), n // 2, replace=False)], _bounce_rays(n - n // 2, 4)])
    oscene = port.OracleScene(arrays)
    want = oscene.trace(rays, 1)
    got = r.trace(rays)
    assert np.array_equal(got["model"], want["model"]) and np.array_equal(got["tri"], want["tri"])
    hit = want["model"] >= 0
    assert hit.mean() > 0.9
    for f in ("t_model", "dist", "u", "v"):
        assert np.array_equal(got[f][hit], want[f][hit]), f
    assert (want["model"][hit] == len(arrays["models"]) - 1).mean() > 0.2      # a good share of the rays ends on the big mesh


def test_production_frame_rays_vs_brute_force_oracle(big, port):
    """The benchmarked configuration itself: a real 1920x1080 iteration through the PRODUCTION kernel instantiation (ptap_render_probe),
    every round; a seeded subset of each round's rays is re-traced by the brute-force oracle (R1): ids, model t and the world distance
    the shade kernel derives must be bit-equal."""
    name, r, arrays = big
    per_round = {"mesh100k": 4000, "mesh1m": 4000, "mesh5m": 500}[name]
    W, H, depth = 1920, 1080, 5
    r.set_params(W, H, depth, first_hit_cache=True)
    oscene = port.OracleScene(arrays)
    rs = np.random.RandomState(21)
    for rnd in range(depth):
        rays, pix, hits = r.render_probe(0, rnd)
        assert len(rays) > per_round
        sel = np.sort(rs.choice(len(rays), per_round, replace=False))
        want = oscene.trace(rays[sel], 1)
        got = hits[sel]
        assert np.array_equal(got["model"], want["model"]) and np.array_equal(got["tri"], want["tri"]), f"round {rnd}"
        hit = want["model"] >= 0
        for f in ("t_model", "dist", "normal"):
            assert np.array_equal(got[f][hit], want[f][hit]), f"round {rnd}: {f}"
    r.frame_begin()
    r.set_params(64, 32, 5)


def test_full_size_properties(big):
    name, r, arrays = big
    cam = _camera_rays(1920, 1080)                               # the full 2,073,600-ray primary wavefront of the bench config
    rays = np.concatenate([cam, _bounce_rays(1 << 20, 7)])
    a = r.trace(rays)
    assert np.array_equal(r.trace(rays).tobytes(), a.tobytes())                          # idempotent / deterministic
    perm = np.random.RandomState(9).permutation(len(rays))
    b = r.trace(rays[perm])
    assert np.array_equal(b.tobytes(), a[perm].tobytes())                                # no cross-ray state
    hit = a["model"] >= 0
    assert hit.mean() > 0.95 and (a["tri"][hit] < len(arrays["triangles"])).all() and (a["dist"][hit] > 0).all()
    # restart every ray 1 % in front of its hit: the same primitive must win again, at the remaining distance
    sel = np.nonzero(hit)[0][:: 7]
    o, d = rays[sel, :3].astype(np.float64), rays[sel, 3:].astype(np.float64)
    dn = d / np.linalg.norm(d, axis=1, keepdims=True)
    dist = a["dist"][sel].astype(np.float64)
    r2 = np.concatenate([o + dn * (dist * 0.99)[:, None], d], 1).astype(np.float32)
    c = r.trace(r2)
    same = (c["model"] == a["model"][sel]) & (c["tri"] == a["tri"][sel])
    assert same.mean() > 0.999                                    # the rest: equal-distance neighbours across a shared edge
    assert np.abs(c["dist"][same] - dist[same] * 0.01).max() <= 1e-3 * np.maximum(dist[same], 1.0).max()


def test_full_size_frame_invariants(big):
    """Whole frames at the bench resolution (1920x1080, depth 5): size-independent properties of the wavefront.
    Every pixel owns exactly one path per iteration and a path deposits sqrt(throughput) <= 1 once, so after k iterations every film
    value lies in (0, k]; the active count never grows; rays traced = sum of the active counts; two renders are bit-identical."""
    name, r, arrays = big
    from pathtracerap_b200 import ACCEL_BVH
    W, H, depth, iters = 1920, 1080, 5, 3
    r.set_params(W, H, depth, first_hit_cache=False)
    r.render(0, iters)
    film = r.film()
    st = r.stats()
    assert np.isfinite(film).all() and film.min() > 0.0 and film.max() <= iters * (1 + 1e-6)
    act = st["active_per_round"][:depth]
    assert act[0] == W * H and all(a >= b for a, b in zip(act, act[1:])) and act[-1] > 0
    assert st["paths"] == iters * W * H
    assert iters * act[0] <= st["rays_traced"] <= iters * sum(act) * 1.01 and st["rays_traced"] >= sum(act)
    r.set_params(W, H, depth, first_hit_cache=False)
    r.render(0, iters)
    assert np.array_equal(r.film(), film)                                   # deterministic: no atomics on the film, stable compaction
    # the first-hit cache changes the launch schedule, not the image (camera rays are identical every iteration, Renderer.cpp:594-613)
    r.set_params(W, H, depth, first_hit_cache=True)
    r.render(0, iters)
    assert np.array_equal(r.film(), film)
    assert r.stats()["rays_traced"] == st["rays_traced"] - (iters - 1) * W * H
    r.set_params(64, 32, 5)
