"""Parity of the path that the benchmark and Renderer::renderLoop's replacement actually run.

ptap_trace (tests/test_gpu_trace.py) launches the barycentric-recording instantiations of the closest-hit kernels; ptap_render launches
k_trace_bvh<false,false> / k_trace_grid<false,false> - different code objects - and, for BVH hits, leaves the exact world distance to the
shade kernel.  These tests go through ptap_render_probe, which enqueues an iteration exactly as ptap_render does and hands back the
wavefront of one round, and through whole frames at oracle tier R1 (brute force with the reference's predicate inside the reference's loop),
so that the BVH path carries the same bounds as the bit-compatible grid path."""
import numpy as np
import pytest

from conftest import have_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="no CUDA device")]

FLOAT_MAX = np.float32(9999999.0)
EMISSIVE = 4
DIR_TOL = 2e-6


def _f32(x):
    return np.asarray(x, np.float32)


def _assert_production_hits(got, want, what):
    assert np.array_equal(got["model"], want["model"]), f"{what}: model ids differ on {(got['model'] != want['model']).sum()} rays"
    assert np.array_equal(got["tri"], want["tri"]), f"{what}: triangle ids differ on {(got['tri'] != want['tri']).sum()} rays"
    hit = want["model"] >= 0
    for f in ("t_model", "dist"):
        assert np.array_equal(got[f][hit], want[f][hit]), f"{what}: {f} not bit-equal"
    assert np.array_equal(got["normal"][hit], want["normal"][hit]), f"{what}: normals not bit-equal"
    assert np.array_equal(got["mat_type"], want["mat_type"])
    assert (got["dist"][~hit] >= FLOAT_MAX).all()


def _expected_next_origins(rays, hits, remaining):
    """k_shade / shadeRayKernel (Renderer.cpp:426-478) in numpy binary32, un-contracted: survivors of the round in slot order and their
    new origins hit + 0.1 * n, where hit = orig + normalize(dir) * dist uses the distance the shade kernel had to evaluate itself."""
    alive = (hits["model"] >= 0) & (hits["mat_type"] != EMISSIVE) & (remaining > 1)
    o, d = _f32(rays[:, :3]), _f32(rays[:, 3:])
    s = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    inv = np.float32(1.0) / np.sqrt(s)
    dn = d * inv[:, None]
    pt = o + dn * hits["dist"][:, None]
    no = pt + hits["normal"] * np.float32(0.1)
    return alive, no[alive]


@pytest.fixture(scope="module")
def renderer(gpu_scene):
    from pathtracerap_b200 import Renderer
    r = Renderer(width=256, height=192, depth=5, first_hit_cache=False)
    r.allocateOnGPU(gpu_scene)
    yield r
    r.free()


@pytest.mark.parametrize("accel_name,mode", [("grid", 0), ("emu", 0), ("bvh", 1), ("lbvh", 1)])
def test_production_kernels_round_by_round(renderer, oracle_scene, accel_name, mode):
    """Every round of a real iteration: the rays the production kernel read, re-traced by the oracle (R0 for the grid walk and for its
    emulation through the BVH, R1 for the BVHs), must give bit-equal ids, model t, world distance (for the BVH: the value deferred to the consumer), normal and material;
    and the next round's wavefront must be exactly the survivors, in stable order, restarted at hit + 0.1 n."""
    from pathtracerap_b200 import ACCEL_BVH, ACCEL_BVH_DEVICE, ACCEL_GRID_COMPAT, ACCEL_GRID_EMULATED
    accel = {"grid": ACCEL_GRID_COMPAT, "emu": ACCEL_GRID_EMULATED, "bvh": ACCEL_BVH, "lbvh": ACCEL_BVH_DEVICE}[accel_name]
    W, H, depth = 256, 192, 5
    renderer.set_accel(accel)
    for cache in (False, True):
        renderer.set_params(W, H, depth, first_hit_cache=cache)
        for it in (0, 3):
            prev = None
            for rnd in range(depth):
                rays, pix, hits = renderer.render_probe(it, rnd)
                assert len(rays) > 0
                if rnd == 0:
                    assert len(rays) == W * H and np.array_equal(pix, np.arange(W * H))
                want = oracle_scene.trace(rays, mode)
                _assert_production_hits(hits, want, f"{accel_name} iter {it} round {rnd}")
                if prev is not None:
                    alive, origins = prev
                    assert len(rays) == int(alive.sum()), f"round {rnd}: survivor count"
                    assert np.array_equal(pix, prev_pix[alive]), f"round {rnd}: compaction is not stable"
                    assert np.array_equal(rays[:, :3], origins), f"round {rnd}: origins (hit distance consumed by k_shade) not bit-equal"
                    assert np.isfinite(rays).all()
                prev = _expected_next_origins(rays, hits, depth - rnd)
                prev_pix = pix
    renderer.frame_begin()


def test_probe_equals_parity_entry(renderer):
    """The two instantiations must agree with each other too: ptap_trace (UV build) on the probe's rays."""
    from pathtracerap_b200 import ACCEL_BVH
    renderer.set_accel(ACCEL_BVH)
    renderer.set_params(256, 192, 5, first_hit_cache=False)
    rays, pix, hits = renderer.render_probe(1, 2)
    again = renderer.trace(rays)
    for f in ("model", "tri", "t_model", "dist", "normal", "mat_type"):
        assert np.array_equal(again[f], hits[f]), f
    renderer.frame_begin()


@pytest.mark.parametrize("accel_name", ["bvh", "lbvh"])
def test_bvh_frame_vs_reference_film_r1(renderer, golden_films, accel_name):
    """BVH frames against the film the compiled reference produces when only its closest-hit launch is swapped for brute force
    (tests/golden/films.npz: bundled_film_r1, made by tools/make_golden.py from oracle/_ref).  Same bound as the grid path carries against
    the unmodified reference: RMSE <= 0.5 % of the mean film value, PSNR >= 40 dB, > 97 % of the pixels bit-equal, active counts within 1 %."""
    from pathtracerap_b200 import ACCEL_BVH, ACCEL_BVH_DEVICE
    f = golden_films
    W, H, depth, iters = (int(x) for x in f["bundled_params"])
    want = f["bundled_film_r1"]
    renderer.set_accel(ACCEL_BVH if accel_name == "bvh" else ACCEL_BVH_DEVICE)
    for cache in (False, True):
        renderer.set_params(W, H, depth, first_hit_cache=cache)
        renderer.render(0, iters)
        film = renderer.film()
        st = renderer.stats()
        rmse = float(np.sqrt(np.mean((film - want) ** 2)))
        psnr = 20 * np.log10(float(want.max()) / max(rmse, 1e-12))
        exact = float(np.mean(film == want))
        print(f"{accel_name} cache={cache}: rmse={rmse:.3e} mean={want.mean():.3f} psnr={psnr:.1f} dB bit-equal pixels={exact:.4f}")
        assert rmse <= 0.005 * want.mean() and psnr >= 40.0
        assert exact > 0.97
        traced = int(np.sum(f["bundled_counts_r1"])) - ((iters - 1) * W * H if cache else 0)
        assert abs(st["rays_traced"] - traced) <= 0.002 * traced
        got = np.array(st["active_per_round"][:depth]); ref_counts = f["bundled_counts_r1"][-1]
        assert got[0] == ref_counts[0] and got[1] == ref_counts[1]
        assert np.abs(got - ref_counts).max() <= 0.01 * ref_counts[0]


def test_bvh_frame_mesh100k_vs_oracle_r1(libptap, port):
    """The same bound on BASELINE configs[1]'s scene (81,920-triangle displaced icosphere in the box) at a resolution the brute-force
    oracle finishes in seconds: 96 x 64, 2 iterations, depth 5 (about 3e9 predicate evaluations on the host)."""
    import bench
    from pathtracerap_b200 import ACCEL_BVH, Renderer
    scene, arrays = bench.build_scene("mesh100k")
    scene.build_bvh()
    W, H, depth, iters = 96, 64, 5, 2
    r = Renderer(width=W, height=H, depth=depth, accel=ACCEL_BVH, first_hit_cache=True)
    r.allocateOnGPU(scene)
    r.render(0, iters)
    film = r.film()
    got = np.array(r.stats()["active_per_round"][:depth])
    r.free()
    oscene = port.OracleScene(arrays)
    w = port.OracleWavefront(oscene, W, H, depth, mode=1)
    w.init_image()
    counts = [w.run_iteration(it) for it in range(iters)]
    want = w.image()
    w.close()
    rmse = float(np.sqrt(np.mean((film - want) ** 2)))
    exact = float(np.mean(film == want))
    print(f"mesh100k R1 frame: rmse={rmse:.3e} mean={want.mean():.3f} bit-equal pixels={exact:.4f} counts {counts[-1]} vs {got.tolist()}")
    assert rmse <= 0.005 * want.mean()
    assert exact > 0.97
    assert got[0] == counts[-1][0] and got[1] == counts[-1][1]
    assert np.abs(got[:len(counts[-1])] - np.array(counts[-1])).max() <= 0.01 * got[0]


def _sphere_scene(scales, seed):
    """Unit-radius icospheres (model space) instanced at world scales >= 500, overlapping and nested, so that the predicate's t >= -EPSILON
    band is several world units wide and instances compete for rays that start a hair off a surface."""
    from pathtracerap_b200 import COAT, DIFFUSE, METAL, Scene
    rs = np.random.RandomState(seed)
    s = Scene.empty()
    mi = s.add_icosphere(3, radius=1.0, displacement=0.02, seed=5)
    centres = []
    for k, sc in enumerate(scales):
        c = rs.uniform(-400, 400, 3) if k else np.zeros(3)
        s.add_model(mi, translate=tuple(float(x) for x in c), rotate_y_degrees=float(rs.uniform(0, 360)),
                    scale=(sc, sc * float(rs.uniform(0.9, 1.1)), sc), material=[DIFFUSE, METAL, COAT][k % 3], color=(0.8, 0.7, 0.6))
        centres.append((c, sc))
    return s, centres


def test_large_instance_scale_grazing_rays(libptap, port):
    """Instances scaled x500 ... x900 of a unit mesh.  Rays start 0.1 world units off (and inside) the surfaces at every angle including
    grazing: self-hits with model t in [-EPSILON, 0) lie up to 4.5 world units BEHIND the origin, the reference ranks them by
    length() >= 0 (Renderer.cpp:391-393), and a nearer instance must not be pruned or mis-ranked.  Bit-equal to brute force."""
    from pathtracerap_b200 import ACCEL_BVH, ACCEL_BVH_DEVICE, Renderer
    scales = [500.0, 650.0, 900.0, 520.0, 700.0]
    s, centres = _sphere_scene(scales, 11)
    a = s.arrays()
    s.build_bvh()
    rs = np.random.RandomState(12)
    rays = []
    for c, sc in centres:
        n = 6000
        p = rs.randn(n, 3); p /= np.linalg.norm(p, axis=1, keepdims=True)
        surf = c + p * sc                                            # roughly on the instance's surface (displacement 2 %)
        off = rs.choice([-20.0, -3.0, -0.1, 0.1, 1.0, 3.0, 20.0], n)[:, None]
        o = surf + p * off
        t = rs.randn(n, 3); t -= p * np.sum(t * p, axis=1, keepdims=True); t /= np.linalg.norm(t, axis=1, keepdims=True)
        ang = rs.choice([0.0, 1e-3, 1e-2, 0.1, 0.5, 1.5], n)[:, None] * rs.choice([-1.0, 1.0], n)[:, None]
        d = t * np.cos(ang) + p * np.sin(ang)
        d *= rs.uniform(0.2, 40.0, (n, 1))                           # Ray::base.dir is not normalised
        rays.append(np.concatenate([o, d], 1))
    rays = np.concatenate(rays).astype(np.float32)
    oscene = port.OracleScene({k: a[k] for k in ("models", "meshes", "vertices", "triangles")})
    want = oscene.trace(rays, 1)
    hit = want["model"] >= 0
    assert hit.mean() > 0.5
    neg = hit & (want["t_model"] < 0)
    assert neg.sum() > 100, "the test must exercise winners behind the origin"
    r = Renderer(width=64, height=32, depth=5, accel=ACCEL_BVH)
    r.allocateOnGPU(s)
    for accel in (ACCEL_BVH, ACCEL_BVH_DEVICE):
        r.set_accel(accel)
        got = r.trace(rays)
        assert np.array_equal(got["model"], want["model"]), f"model ids differ on {(got['model'] != want['model']).sum()} rays"
        assert np.array_equal(got["tri"], want["tri"])
        for f in ("t_model", "dist", "u", "v"):
            assert np.array_equal(got[f][hit], want[f][hit]), f
    r.free()
