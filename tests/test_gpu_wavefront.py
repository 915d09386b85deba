"""Shading, compaction and whole-frame parity on the GPU through the C ABI."""
import numpy as np
import pytest

from conftest import have_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="no CUDA device")]

FLOAT_MAX = np.float32(9999999.0)
# New directions go through libdevice sinf/cosf/powf instead of glibc's: allow a few ulp, relative to unit-length vectors.
DIR_TOL = 2e-6
ORIG_TOL = 1e-5      # relative to the scene scale (~1e3): positions inherit the direction error times the hit distance


@pytest.fixture(scope="module")
def renderer(gpu_scene):
    from pathtracerap_b200 import Renderer
    r = Renderer(width=64, height=48, depth=5, first_hit_cache=False)
    r.allocateOnGPU(gpu_scene)
    yield r
    r.free()


def test_shade_and_compaction_vs_reference_states(renderer, golden_wavefront):
    """Feeds the reference's own pre-shade wavefront (rays + closest-hit ids) to ptap_shade and compares the post-shade
    state slot by slot, the survivor order, and the film contribution of terminated paths."""
    from pathtracerap_b200 import PATH_IN
    g = golden_wavefront
    for k in range(int(g["nsteps"])):
        n, it = int(g["n"][k]), int(g["iter"][k])
        p = np.zeros(n, PATH_IN)
        p["orig"], p["dir"], p["color"], p["ipixel"] = g[f"pre_orig_{k}"], g[f"pre_dir_{k}"], g[f"pre_color_{k}"], g[f"pre_ipixel_{k}"]
        miss = g[f"hit_dist_{k}"] >= FLOAT_MAX
        p["model"] = np.where(miss, -1, g[f"hit_model_{k}"]); p["tri"] = np.where(miss, -1, g[f"hit_tri_{k}"]); p["dist"] = g[f"hit_dist_{k}"]
        remaining = int(g[f"pre_bounces_{k}"][0])
        assert (g[f"pre_bounces_{k}"] == remaining).all()          # every live path of a round has the same count
        out, order = renderer.shade(p, it, remaining)
        want_alive = g[f"post_bounces_{k}"] > 0
        assert np.array_equal(out["alive"].astype(bool), want_alive), f"step {k}: survivor set differs"
        # stable compaction: survivors keep their relative order (thrust::stable_partition, Renderer.cpp:628)
        assert np.array_equal(order, np.nonzero(want_alive)[0])
        assert np.array_equal(out["ipixel"], g[f"pre_ipixel_{k}"])
        a = want_alive
        assert np.array_equal(out["color"][a], g[f"post_color_{k}"][a])                       # throughput: exact arithmetic only
        assert np.abs(out["dir"][a] - g[f"post_dir_{k}"][a]).max(initial=0) <= DIR_TOL * max(1.0, np.abs(g[f"post_dir_{k}"][a]).max(initial=1))
        assert np.abs(out["orig"][a] - g[f"post_orig_{k}"][a]).max(initial=0) <= ORIG_TOL * 1e3
        # terminated paths: the film receives sqrt(throughput) (gatherImageDataKernel, Renderer.cpp:489-495)
        assert np.array_equal(out["color"][~a], np.sqrt(g[f"post_color_{k}"][~a]))


def test_frame_vs_reference_film(renderer, golden_films, golden_scene):
    """Whole frames at matched spp and seed.  Trace and throughput arithmetic are exact; only the sampled directions differ
    in the last ulp (libdevice vs glibc), which very rarely flips a later hit.  Bound: RMSE <= 0.5 % of the mean film value
    and PSNR >= 40 dB, against Monte-Carlo noise at 4 spp of roughly 30 %."""
    from pathtracerap_b200 import ACCEL_BVH, ACCEL_GRID_COMPAT
    f = golden_films
    W, H, depth, iters = (int(x) for x in f["bundled_params"])
    want = f["bundled_film"]
    for cache in (False, True):
        renderer.set_accel(ACCEL_GRID_COMPAT)
        renderer.set_params(W, H, depth, first_hit_cache=cache)
        renderer.render(0, iters)
        film = renderer.film()
        st = renderer.stats()
        rmse = float(np.sqrt(np.mean((film - want) ** 2)))
        peak = float(want.max())
        psnr = 20 * np.log10(peak / max(rmse, 1e-12))
        exact = float(np.mean(film == want))
        print(f"cache={cache}: rmse={rmse:.3e} mean={want.mean():.3f} psnr={psnr:.1f} dB bit-equal pixels={exact:.4f}")
        assert rmse <= 0.005 * want.mean() and psnr >= 40.0
        assert exact > 0.97
        traced = int(np.sum(f["bundled_counts"])) - ((iters - 1) * W * H if cache else 0)
        assert abs(st["rays_traced"] - traced) <= 0.002 * traced
        assert st["paths"] == iters * W * H
    # last iteration's active rays per round against the reference's counts (tiny drift allowed: ulp-level direction differences)
    got = np.array(st["active_per_round"][:depth]); ref_counts = f["bundled_counts"][-1]
    assert got[0] == ref_counts[0] and got[1] == ref_counts[1]
    assert np.abs(got - ref_counts).max() <= 0.01 * ref_counts[0]
    # the BVH frame: same scene, R1 semantics (differs from the grid walk on ~0.4 % of rays) - statistical bound only
    renderer.set_accel(ACCEL_BVH)
    renderer.set_params(W, H, depth, first_hit_cache=True)
    renderer.render(0, iters)
    bf = renderer.film()
    rel = abs(bf.mean() - want.mean()) / want.mean()
    assert rel < 0.01


def test_committed_render_bmp(gpu_scene, golden_render_bmp, tmp_path):
    """The reference's only golden artefact, PathTracerAP/Render.bmp (the author's GPU render of the coded scene at 1000x800, iteration
    count unrecorded), against the file this library writes for the same scene at 256 iterations: per-channel byte means within 0.1/255
    and PSNR >= 46 dB after an 8x8 box filter (the committed image still carries its own Monte-Carlo noise).  The grid walk is the
    reference's algorithm (measured: means within 0.003, 54.3 dB).  The BVH answers the exact closest-hit query, i.e. it also finds the
    hits the reference's grid walk loses (0.3-0.6 % of rays per bounce), and is therefore measurably FARTHER from the
    reference's own image: means +0.42 / +0.11 / +0.35, 42.0 dB.  Its bound only guards against gross errors."""
    from conftest import box8_of_film
    from pathtracerap_b200 import ACCEL_BVH, ACCEL_GRID_COMPAT, Renderer
    g = golden_render_bmp
    H, W = (int(x) for x in g["shape"])
    iters = 256
    r = Renderer(width=W, height=H, depth=5, first_hit_cache=True)
    r.allocateOnGPU(gpu_scene)
    for accel, mean_tol, min_psnr in ((ACCEL_GRID_COMPAT, 0.1, 46.0), (ACCEL_BVH, 0.8, 39.0)):
        r.set_accel(accel)
        r.set_params(W, H, 5, first_hit_cache=True)
        r.render(0, iters)
        b8, means = box8_of_film(r.film(), iters)
        rmse = float(np.sqrt(np.mean((b8 - g["box8"]) ** 2)))
        psnr = 20 * np.log10(255.0 / rmse)
        print(f"accel {accel}: channel means {means} vs {g['channel_means']}, box-filtered rmse {rmse:.2f}/255, psnr {psnr:.1f} dB")
        assert np.abs(means - g["channel_means"]).max() <= mean_tol
        assert psnr >= min_psnr
    # the file itself: the bytes ptap_write_bmp stores are the ones compared above
    r.set_accel(ACCEL_GRID_COMPAT); r.set_params(W, H, 5, first_hit_cache=True); r.render(0, iters)
    r.renderImage(str(tmp_path / "Render.bmp"))
    raw = np.frombuffer((tmp_path / "Render.bmp").read_bytes(), np.uint8, W * H * 3, 54).reshape(H, W, 3)
    assert np.abs(raw.reshape(-1, 3).astype(np.float64).mean(0) - g["channel_means"]).max() <= 0.1
    r.free()


def test_cornell_frame(libptap, golden_scene, golden_films):
    from pathtracerap_b200 import Renderer, Scene
    f = golden_films
    s = Scene.from_arrays(f["cornell_models"], golden_scene["meshes"], golden_scene["vertices"], golden_scene["triangles"])
    s.build_grids()
    assert s.arrays()["grids"].tobytes() == f["cornell_grids"].tobytes()
    W, H, depth, iters = (int(x) for x in f["cornell_params"])
    r = Renderer(width=W, height=H, depth=depth, first_hit_cache=True)
    r.allocateOnGPU(s)
    r.render(0, iters)
    film = r.film()
    want = f["cornell_film"]
    rmse = float(np.sqrt(np.mean((film - want) ** 2)))
    assert rmse <= 0.005 * want.mean()
    assert float(np.mean(film == want)) > 0.97
    r.free()


def test_iteration_ranges_add_up(renderer, golden_films):
    """sample partitioning. Rendering [0,2) and [2,4) on separate 'virtual ranks' and summing the films equals
    one run over [0,4) up to float-sum reassociation."""
    f = golden_films
    W, H, depth, iters = (int(x) for x in f["bundled_params"])
    from pathtracerap_b200 import ACCEL_GRID_COMPAT
    renderer.set_accel(ACCEL_GRID_COMPAT)
    renderer.set_params(W, H, depth, first_hit_cache=True)
    renderer.render(0, iters)
    whole = renderer.film()
    renderer.set_params(W, H, depth, first_hit_cache=True)
    renderer.render(0, 2)
    a = renderer.film()
    renderer.set_params(W, H, depth, first_hit_cache=True)
    renderer.render(2, 4)
    b = renderer.film()
    assert np.allclose(a + b, whole, rtol=1e-6, atol=1e-6)
    # film_add: the receive side of the reduce
    renderer.film_add(a)
    assert np.allclose(renderer.film(), whole, rtol=1e-6, atol=1e-6)


def test_bmp_matches_oracle_writer(renderer, port, tmp_path, golden_films):
    f = golden_films
    W, H, depth, iters = (int(x) for x in f["bundled_params"])
    renderer.set_params(W, H, depth, first_hit_cache=True)
    renderer.render(0, iters)
    film = renderer.film()
    renderer.renderImage(str(tmp_path / "gpu.bmp"))
    port.write_bmp(film, iters, tmp_path / "cpu.bmp")
    assert (tmp_path / "gpu.bmp").read_bytes() == (tmp_path / "cpu.bmp").read_bytes()


def test_lanes_do_not_change_a_bit(gpu_scene, golden_films, monkeypatch):
    """Multi-lane rendering pipelines iterations over several streams but keeps every film add in iteration order:
    the film of a 7-iteration frame is bit-identical for 1, 2, 3, 4 and 8 lanes, with and without the first-hit cache, in one call
    or split over two calls, and the traced-ray count is the same."""
    from pathtracerap_b200 import ACCEL_BVH, Renderer
    W, H, depth, _ = (int(x) for x in golden_films["bundled_params"])
    iters = 7
    ref_film, ref_rays = {}, {}
    for lanes in (1, 2, 3, 4, 8):
        monkeypatch.setenv("PTAP_LANES", str(lanes))
        r = Renderer(width=W, height=H, depth=depth, accel=ACCEL_BVH, first_hit_cache=True)
        r.allocateOnGPU(gpu_scene)
        assert r.stats()["lanes"] == lanes
        for cache in (True, False):
            r.set_params(W, H, depth, first_hit_cache=cache)
            r.render(0, iters)
            film, rays = r.film(), r.stats()["rays_traced"]
            if lanes == 1:
                ref_film[cache], ref_rays[cache] = film, rays
            assert np.array_equal(film, ref_film[cache]), f"{lanes} lanes, cache={cache}: film differs from the one-lane film"
            assert rays == ref_rays[cache]
            r.set_params(W, H, depth, first_hit_cache=cache)
            r.render(0, 3); r.render(3, iters)
            assert np.array_equal(r.film(), ref_film[cache]), f"{lanes} lanes, cache={cache}: split render differs"
        r.free()
    monkeypatch.delenv("PTAP_LANES")
    r = Renderer(width=W, height=H, depth=depth)
    r.allocateOnGPU(gpu_scene)
    assert r.stats()["lanes"] == 8                      # frames of at most 2^20 pixels default to 8 lanes, larger ones to 4
    r.set_params(1280, 1024, depth)
    assert r.stats()["lanes"] == 4
    r.free()


def test_material_class_regrouping_does_not_change_a_bit(gpu_scene, golden_films, monkeypatch):
    """PTAP_SHADE_SORT=1 (k_scan's counting-sort permutation, k_shade<true>) only changes which thread shades a slot: the
    film of the five-material scene and the traced-ray count are bit-identical to the default schedule."""
    from pathtracerap_b200 import ACCEL_BVH, Renderer
    W, H, depth, _ = (int(x) for x in golden_films["bundled_params"])
    out = {}
    for sort in ("0", "1"):
        monkeypatch.setenv("PTAP_SHADE_SORT", sort)
        r = Renderer(width=W, height=H, depth=depth, accel=ACCEL_BVH, first_hit_cache=True)
        r.allocateOnGPU(gpu_scene)
        r.render(0, 5)
        out[sort] = (r.film(), r.stats()["rays_traced"])
        r.free()
    assert np.array_equal(out["0"][0], out["1"][0]) and out["0"][1] == out["1"][1]
    assert out["0"][0].mean() > 0.1


def test_supersampled_resolve(renderer, port, tmp_path, golden_films):
    """SAMPLESX x SAMPLESY.  The lattice render is the ordinary path at (W*SX, H*SY) (generateRaysKernel, Renderer.cpp:527-542);
    the resolve is pixel = sum over its samples (row-major) of avg * sample, avg = 1.0f / (SX*SY) - checked bit for bit against the same
    fp32 arithmetic in numpy, and the BMP against the oracle's writer on the resolved film.  (The reference's own gather leaves the image
    black for SAMPLES > 1, Renderer.cpp:493: there is no reference output to match.)"""
    from pathtracerap_b200 import PtapError
    f = golden_films
    W, H, depth, iters = (int(x) for x in f["bundled_params"])
    renderer.set_params(W, H, depth, first_hit_cache=True)
    renderer.render(0, iters)
    hi = renderer.film()
    for sx, sy in ((2, 2), (4, 1), (1, 3), (1, 1)):
        got = renderer.film_resolved((sx, sy))
        avg = np.float32(1.0) / np.float32(sx * sy)
        want = np.zeros((H // sy, W // sx, 3), np.float32)
        for j in range(sy):
            for i in range(sx):
                want = want + avg * hi[j::sy, i::sx]
        assert got.shape == want.shape and np.array_equal(got, want), (sx, sy)
        renderer.renderImage(str(tmp_path / "gpu.bmp"), samples=(sx, sy))
        port.write_bmp(want, iters, tmp_path / "cpu.bmp")
        assert (tmp_path / "gpu.bmp").read_bytes() == (tmp_path / "cpu.bmp").read_bytes()
    with pytest.raises(PtapError):
        renderer.film_resolved((5, 1))        # 128 is not divisible by 5
    with pytest.raises(PtapError):
        renderer.film_resolved((0, 1))


def test_errors_are_loud(libptap):
    from pathtracerap_b200 import PtapError, Renderer
    r = Renderer(width=32, height=32, depth=5)
    with pytest.raises(PtapError):
        r.render(0, 1)                      # no scene uploaded
    with pytest.raises(PtapError):
        r.set_params(32, 32, 99)            # depth beyond the reserved rounds
    r.free()


def test_camera_parameters_and_jitter(gpu_scene):
    """generateRaysKernel's camera as parameters (ptap_set_camera).  Parity unpinned beyond the default (the reference
    hard-codes its camera, Renderer.cpp:538-545, and has no jitter): the rays of a custom camera and of a jittered one are compared bit for
    bit with a numpy restatement of the documented arithmetic (include/ptap.h); the default camera still gives the reference's rays; with
    jitter the first-hit cache is not used and two renders are bit-identical."""
    from pathtracerap_b200 import ACCEL_BVH, Renderer
    W, H, depth = 96, 64, 5
    r = Renderer(width=W, height=H, depth=depth, accel=ACCEL_BVH, first_hit_cache=True)
    r.allocateOnGPU(gpu_scene)
    f32 = np.float32

    def expect(origin, pmin, span, jit=None):
        i = np.arange(W * H)
        fx, fy = (i % W).astype(f32), (i // W).astype(f32)
        if jit is not None:
            fx, fy = fx + jit[0], fy + jit[1]
        sx, sy = f32(np.float64(f32(span[0])) / W), f32(np.float64(f32(span[1])) / H)
        wx = (np.float64(f32(pmin[0])) + (fx * sx).astype(np.float64)).astype(f32)
        wy = (np.float64(f32(pmin[1])) + (fy * sy).astype(np.float64)).astype(f32)
        out = np.zeros((W * H, 6), f32)
        out[:, :3] = f32(origin)
        out[:, 3], out[:, 4], out[:, 5] = wx - f32(origin[0]), wy - f32(origin[1]), f32(pmin[2]) - f32(origin[2])
        return out

    rays, pix, _ = r.render_probe(0, 0)
    assert np.array_equal(rays, expect((0, 0, 920), (-10, -4, 900), (20, 16)))              # the reference's camera (Renderer.cpp:538-548)
    cam = dict(origin=(35.5, 120.25, 880.0), plane_min=(-7.0, -2.5, 861.0), span=(14.0, 9.75))
    r.set_camera(**cam)
    rays, pix, _ = r.render_probe(0, 0)
    assert np.array_equal(rays, expect(cam["origin"], cam["plane_min"], cam["span"])) and np.array_equal(pix, np.arange(W * H))

    def hash32(a):                                                                          # utility.h:43-53
        a = a.astype(np.uint32)
        a = (a + np.uint32(0x7ed55d16)) + (a << np.uint32(12)); a = (a ^ np.uint32(0xc761c23c)) ^ (a >> np.uint32(19))
        a = (a + np.uint32(0x165667b1)) + (a << np.uint32(5)); a = (a + np.uint32(0xd3a2646c)) ^ (a << np.uint32(9))
        a = (a + np.uint32(0xfd7046c5)) + (a << np.uint32(3)); a = (a ^ np.uint32(0xb55a4f09)) ^ (a >> np.uint32(16))
        return a

    seed = 1234
    r.set_camera(jitter=True, jitter_seed=seed)
    seen = []
    with np.errstate(over="ignore"):
        for it in (0, 5):
            i = np.arange(W * H, dtype=np.uint32)
            h = hash32(np.uint32(seed) ^ hash32(np.array([it], np.uint32) + np.uint32(0x9e3779b9))) ^ hash32(i)
            jx = (hash32(h) >> np.uint32(8)).astype(f32) * f32(2.0 ** -24)
            jy = (hash32(h ^ np.uint32(0x85ebca6b)) >> np.uint32(8)).astype(f32) * f32(2.0 ** -24)
            assert jx.min() >= 0 and jx.max() < 1 and jy.max() < 1
            rays, _, _ = r.render_probe(it, 0)
            assert np.array_equal(rays, expect((0, 0, 920), (-10, -4, 900), (20, 16), (jx, jy))), f"jittered rays of iteration {it}"
            seen.append(rays)
    assert not np.array_equal(seen[0], seen[1])
    r.frame_begin(); r.render(0, 4); a = r.film(); st = r.stats()
    assert st["rays_traced"] >= 4 * W * H                                                   # every iteration traces its own camera rays: no first-hit cache
    r.frame_begin(); r.render(0, 4)
    assert np.array_equal(r.film(), a)
    r.set_camera()                                                                          # back to the reference's camera: cache in use again
    r.frame_begin(); r.render(0, 4)
    assert r.stats()["rays_traced"] < st["rays_traced"]
    r.free()
