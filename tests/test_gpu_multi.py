"""Multi-GPU sample partitioning on real devices (needs >= 2 GPUs; `gpurun --gpus 2`): two ranks over NCCL render disjoint
iteration ranges of the same frame and ONE reduce assembles it; the result equals the one-GPU frame up to float-sum reassociation."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, have_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="no CUDA device")]

WORKER = """
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from pathtracerap_b200 import ACCEL_BVH, Renderer, Scene, multi_gpu
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = np.load({scene!r})
s = Scene.from_arrays(g["models"], g["meshes"], g["vertices"], g["triangles"])
s.build_bvh()
r = Renderer(device=local, width=160, height=120, depth=5, accel=ACCEL_BVH)
r.allocateOnGPU(s)
t = multi_gpu.render_partitioned(r, 8, rank, world)
r.sync(); torch.cuda.synchronize()
if rank == 0:
    via_torch = t.cpu().numpy().reshape(120, 160, 3).copy()
    np.save({out!r}, via_torch)
    st = r.stats()
    print("RANK0_DONE", st["rays_traced"])
dist.barrier()
# the same frame with the reduce issued through the library's C ABI (ptap_nccl_init + ptap_reduce: its own communicator, its own stream)
how = multi_gpu.nccl_join(r, rank, world)
assert how is not None, "ptap_nccl_init failed: libnccl not loadable"
b, e = multi_gpu.iteration_range(rank, world, 8)
r.frame_begin(); r.render(b, e); r.reduce(0); r.sync()
if rank == 0:
    assert np.array_equal(r.film(), via_torch), "ptap_reduce differs from torch.distributed.reduce"
    print("C_ABI_REDUCE_OK")
dist.barrier()
r.nccl_finalize()
r.free()
dist.destroy_process_group()
"""


def test_two_gpu_frame_equals_one_gpu_frame(libptap, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from pathtracerap_b200 import ACCEL_BVH, Renderer, Scene
    scene_path = os.path.join(GOLDEN, "bundled_scene.npz")
    out = str(tmp_path / "film2.npy")
    w = tmp_path / "worker.py"
    w.write_text(textwrap.dedent(WORKER.format(root=ROOT, scene=scene_path, out=out)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29544", str(w)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    assert "RANK0_DONE" in p.stdout and "C_ABI_REDUCE_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
    film2 = np.load(out)
    g = np.load(scene_path)
    s = Scene.from_arrays(g["models"], g["meshes"], g["vertices"], g["triangles"])
    s.build_bvh()
    r = Renderer(width=160, height=120, depth=5, accel=ACCEL_BVH)
    r.allocateOnGPU(s)
    r.render(0, 8)
    film1 = r.film()
    r.free()
    assert np.allclose(film2, film1, rtol=1e-6, atol=1e-6)
    assert float(np.mean(film2 == film1)) > 0.5


@pytest.fixture(scope="module")
def scene_bvh(libptap):
    from pathtracerap_b200 import Scene
    g = np.load(os.path.join(GOLDEN, "bundled_scene.npz"))
    s = Scene.from_arrays(g["models"], g["meshes"], g["vertices"], g["triangles"])
    s.build_bvh()
    return s


def test_peer_reduce_of_two_contexts(scene_bvh):
    """ptap_reduce_peer: two contexts of one process (here both on device 0; on a multi-GPU box the copy is a peer copy over NVLink)
    render disjoint iteration ranges; film(A) += film(B) must equal the float32 sum of the two films bit for bit (a FIXED order of
    additions, unlike a tree reduction), and the one-context frame up to reassociation."""
    from pathtracerap_b200 import ACCEL_BVH, Renderer
    W, H = 160, 120
    a = Renderer(width=W, height=H, depth=5, accel=ACCEL_BVH); a.allocateOnGPU(scene_bvh)
    b = Renderer(width=W, height=H, depth=5, accel=ACCEL_BVH); b.allocateOnGPU(scene_bvh)
    a.render(0, 4); b.render(4, 8)
    fa, fb = a.film(), b.film()
    a.frame_begin(); b.frame_begin()
    a.render(0, 4); b.render(4, 8)              # asynchronous on both contexts; the reduce is ordered after both without a host sync
    a.reduce_peer(b)
    got = a.film()
    assert np.array_equal(got, fa + fb)
    assert np.array_equal(b.film(), fb)         # the source film is untouched
    a.frame_begin(); a.render(0, 8)
    one = a.film()
    assert np.allclose(got, one, rtol=1e-6, atol=1e-6) and float(np.mean(got == one)) > 0.5
    a.free(); b.free()


def test_nccl_reduce_through_the_c_abi_single_rank(scene_bvh):
    """ptap_nccl_unique_id / ptap_nccl_init / ptap_reduce with one rank: proves that libnccl is found and bound at run time (dlopen) and
    that the collective is ordered on the library's stream; with one rank the film must come back unchanged."""
    from pathtracerap_b200 import ACCEL_BVH, Renderer, multi_gpu
    r = Renderer(width=160, height=120, depth=5, accel=ACCEL_BVH); r.allocateOnGPU(scene_bvh)
    r.render(0, 3)
    before = r.film()
    r.nccl_init(multi_gpu.nccl_unique_id(), 1, 0)
    r.frame_begin(); r.render(0, 3)
    r.reduce(0)
    assert np.array_equal(r.film(), before)
    r.nccl_finalize()
    r.free()


def test_stamps_and_iteration_times(scene_bvh):
    """PTAP_FLAG_STAMP / PTAP_FLAG_ITER_TIMES: the closest-hit launches' residency measured on the device clock inside the multi-lane
    schedule (union <= the call's duration <= sum over overlapping lanes is possible), and one completion time per iteration, in order."""
    from pathtracerap_b200 import ACCEL_BVH, Renderer
    r = Renderer(width=640, height=480, depth=5, accel=ACCEL_BVH); r.allocateOnGPU(scene_bvh)
    r.set_params(640, 480, 5, first_hit_cache=True, stamp=True, iter_times=True)
    r.render(0, 12); r.sync()
    plain = r.film()
    st = r.stats()
    assert st["trace_launches"] == 12 * 5 - 11
    assert 0.0 < st["ms_trace_inflight"] <= st["ms_render"] * 1.02 + 0.05
    assert st["ms_trace_sum"] >= st["ms_trace_inflight"] * 0.999
    t = r.iteration_times()
    assert len(t) == 12 and (np.diff(t) >= 0).all() and t[0] > 0 and t[-1] <= st["ms_render"] * 1.02 + 0.05
    r.set_params(640, 480, 5, first_hit_cache=True)
    r.render(0, 12)
    assert np.array_equal(r.film(), plain)      # the instrumentation changes no pixel
    r.free()
