"""PTAP_ACCEL_BVH_DEVICE: the per-mesh trees are built on the GPU (LBVH, csrc/bvh_device.cu).  A different tree than the host's SAH one,
the same hits: bit-exact against the brute-force oracle (tier R1) and against the host-built BVH at full size."""
import numpy as np
import pytest

from conftest import have_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="no CUDA device")]


def _assert_equal(got, want, what):
    assert np.array_equal(got["model"], want["model"]), f"{what}: model ids differ on {(got['model'] != want['model']).sum()} rays"
    assert np.array_equal(got["tri"], want["tri"]), f"{what}: triangle ids differ on {(got['tri'] != want['tri']).sum()} rays"
    hit = want["model"] >= 0
    for f in ("t_model", "dist", "u", "v"):
        assert np.array_equal(got[f][hit], want[f][hit]), f"{what}: {f} not bit-equal"


def test_device_bvh_bundled_vs_golden(gpu_scene, golden_trace):
    from pathtracerap_b200 import ACCEL_BVH, ACCEL_BVH_DEVICE, Renderer
    r = Renderer(width=64, height=32, depth=5, accel=ACCEL_BVH_DEVICE)
    r.allocateOnGPU(gpu_scene)
    _assert_equal(r.trace(golden_trace["rays"]), golden_trace["r1"], "device-built BVH vs reference brute force R1 (golden)")
    bs = r.build_stats()
    assert bs["bvh_nodes"] > 0 and 1 <= bs["bvh_depth"] < 50 and bs["ms_build"] > 0
    # a frame through the device-built tree equals the frame through the host-built one bit for bit (same hits => same paths)
    r.set_params(160, 120, 5)
    r.render(0, 3)
    a = r.film()
    r.set_accel(ACCEL_BVH_DEVICE)
    r.free()
    r2 = Renderer(width=160, height=120, depth=5, accel=ACCEL_BVH)
    r2.allocateOnGPU(gpu_scene)
    r2.render(0, 3)
    assert np.array_equal(r2.film(), a)
    r2.free()


@pytest.mark.parametrize("workload", ["mesh100k", "mesh1m"])
def test_device_bvh_large(workload, libptap, port):
    import bench
    from test_gpu_large import _bounce_rays, _camera_rays
    from pathtracerap_b200 import ACCEL_BVH, ACCEL_BVH_DEVICE, Renderer
    scene, arrays = bench.build_scene(workload)
    r = Renderer(width=64, height=32, depth=5, accel=ACCEL_BVH_DEVICE)
    r.allocateOnGPU(scene)                                   # no host BVH in the scene: built on the device at this call
    bs = r.build_stats()
    print(workload, "device build:", bs)
    cam = _camera_rays(1920, 1080)
    n = 20000 if workload == "mesh100k" else 2000
    rays = np.concatenate([cam[np.random.RandomState(3).choice(len(cam), n // 2, replace=False)], _bounce_rays(n - n // 2, 4)])
    _assert_equal(r.trace(rays), port.OracleScene(arrays).trace(rays, 1), "device-built BVH vs brute force")
    # full size: the host-built (SAH) tree must give the same hits on the whole primary wavefront + 1M incoherent rays
    big = np.concatenate([cam, _bounce_rays(1 << 20, 7)])
    a = r.trace(big)
    r.free()
    scene.build_bvh()
    r2 = Renderer(width=64, height=32, depth=5, accel=ACCEL_BVH)
    r2.allocateOnGPU(scene)
    b = r2.trace(big)
    r2.free()
    assert np.array_equal(a.tobytes(), b.tobytes())
