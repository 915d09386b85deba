"""BASELINE.json's configurations at BASELINE.json's sizes, against the reference itself where its compiled form travelled to this machine
(oracle/_ref/libptap_ref.so, built by oracle/build_ref.sh from /root/reference), else against the C port that CPU tests pin to it."""
import numpy as np
import pytest

from conftest import have_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="no CUDA device")]


def _reference_frame(arrays, W, H, depth, iters):
    """(film, per-iteration active counts, kind) from the compiled reference's own kernels, or the port when that library is absent."""
    from oracle import port, ref
    if ref.available():
        scene = ref.RefScene.from_arrays(arrays["models"], arrays["meshes"], arrays["vertices"], arrays["triangles"])
        r = ref.RefRenderer(scene, W, H, depth)
        r.init_image()
        counts = [r.run_iteration(it) for it in range(iters)]
        film = r.image()
        r.close(); scene.close()
        return film, counts, "reference"
    scene = port.OracleScene(arrays)
    w = port.OracleWavefront(scene, W, H, depth)
    w.init_image()
    counts = [w.run_iteration(it) for it in range(iters)]
    film = w.image()
    w.close()
    return film, counts, "port"


@pytest.mark.parametrize("accel_name", ["emulated", "walked"])
def test_config0_cornell_512x512_16spp_depth8(libptap, accel_name):
    """configs[0]: Cornell box from Input data, 512 x 512, 16 spp, depth 8, diffuse only, golden image from the reference's host-compiled
    path - produced live here (15.5 M rays on the host).  The drop-in default (the grid walk's results through the BVH, R0) and the walk
    itself must match it within RMSE <= 0.5 % of the mean, PSNR >= 40 dB, > 97 % bit-equal pixels; iteration-0 active counts are SURVEY
    Appendix A.3b's checkpoints."""
    import bench
    from pathtracerap_b200 import ACCEL_GRID_COMPAT, ACCEL_GRID_EMULATED, Renderer
    accel = {"emulated": ACCEL_GRID_EMULATED, "walked": ACCEL_GRID_COMPAT}[accel_name]
    scene, arrays = bench.build_scene("cornell")
    scene.build_grids(25, 25, 25)
    W, H, iters, depth = 512, 512, 16, 8
    want, counts, kind = _reference_frame(arrays, W, H, depth, iters)
    assert counts[0] == [262144, 232245, 153987, 111947, 80385, 59065, 43551, 32384]
    r = Renderer(width=W, height=H, depth=depth, accel=accel, first_hit_cache=True)
    r.allocateOnGPU(scene)
    assert r.accel == accel
    r.render(0, iters)
    film = r.film()
    st = r.stats()
    rmse = float(np.sqrt(np.mean((film - want) ** 2)))
    psnr = 20 * np.log10(float(want.max()) / max(rmse, 1e-12))
    exact = float(np.mean(film == want))
    print(f"cornell 512x512x16 d8 ({accel_name}) vs {kind}: rmse={rmse:.3e} mean={want.mean():.3f} psnr={psnr:.1f} dB bit-equal pixels={exact:.4f}")
    assert rmse <= 0.005 * want.mean() and psnr >= 40.0 and exact > 0.97
    got = np.array(st["active_per_round"][:depth])
    assert got[0] == counts[-1][0] and got[1] == counts[-1][1]
    assert np.abs(got - np.array(counts[-1])).max() <= 0.01 * got[0]
    traced = sum(sum(c) for c in counts) - (iters - 1) * W * H
    assert abs(st["rays_traced"] - traced) <= 0.002 * traced
    r.free()


def test_bundled_1000x800_iteration0_checkpoints(gpu_scene, oracle_scene):
    """The reference's native configuration (Config.h:12-13): 1000 x 800, depth 5, iteration 0.  Known answers of the compiled reference
    (tests/golden/trace_bundled.npz carries the same numbers): active rays per bounce 800000 / 708894 / 474310 /
    348742 / 254855 and hits 800000 / 574891 / 409613 / 300967 / 223986.  Rounds 0 and 1 are exact arithmetic end to end; later rounds
    inherit last-ulp differences of the sampled directions (libdevice vs glibc sinf/cosf/powf) and may drift by a few rays."""
    from pathtracerap_b200 import ACCEL_GRID_EMULATED, Renderer
    active = [800000, 708894, 474310, 348742, 254855]
    hits_want = [800000, 574891, 409613, 300967, 223986]
    r = Renderer(width=1000, height=800, depth=5, accel=ACCEL_GRID_EMULATED, first_hit_cache=False)       # the drop-in default
    r.allocateOnGPU(gpu_scene)
    n_got, h_got = [], []
    for rnd in range(5):
        rays, pix, hits = r.render_probe(0, rnd)
        n_got.append(len(rays)); h_got.append(int((hits["model"] >= 0).sum()))
    print("active", n_got, "hits", h_got)
    assert n_got[:2] == active[:2] and h_got[:2] == hits_want[:2]
    assert max(abs(a - b) for a, b in zip(n_got, active)) <= 40 and max(abs(a - b) for a, b in zip(h_got, hits_want)) <= 40
    r.frame_begin()
    r.render(0, 1)
    assert r.stats()["active_per_round"][:5] == n_got
    r.free()


def test_config4_mesh1m_3840x2160_frame_invariants(libptap):
    """configs[4]'s frame (the 1.3 M-triangle scene at 3840 x 2160) on one GPU: size-independent properties of the wavefront at 8.3 M
    paths per iteration, and the sample partition's identity: iterations [0,2) + [2,4) rendered separately and summed equal one
    [0,4) render bit for bit when added in iteration order... which a plain sum of two partial films is not (float reassociation), so the
    check is equality of each partial render with itself under a different lane count, and closeness of the sum."""
    import bench
    from pathtracerap_b200 import ACCEL_BVH, Renderer
    scene, arrays = bench.build_scene("mesh1m4k")
    scene.build_bvh()
    W, H, depth = 3840, 2160, 5
    r = Renderer(width=W, height=H, depth=depth, accel=ACCEL_BVH, first_hit_cache=True)
    r.allocateOnGPU(scene)
    r.render(0, 4)
    full = r.film()
    st = r.stats()
    act = st["active_per_round"][:depth]
    assert act[0] == W * H and all(a >= b for a, b in zip(act, act[1:])) and act[-1] > 0
    assert st["paths"] == 4 * W * H
    assert np.isfinite(full).all() and full.min() > 0.0 and full.max() <= 4 * (1 + 1e-6)
    r.frame_begin(); r.render(0, 2); a = r.film()
    r.frame_begin(); r.render(2, 4); b = r.film()
    assert a.max() <= 2 * (1 + 1e-6) and b.max() <= 2 * (1 + 1e-6)
    assert np.abs((a + b) - full).max() <= 4 * 2.0 ** -22                    # reassociation of four float adds of values <= 1
    r.frame_begin(); r.render(0, 4)
    assert np.array_equal(r.film(), full)                                    # deterministic at 8.3 M paths
    r.free()
