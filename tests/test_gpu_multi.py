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
    np.save({out!r}, t.cpu().numpy().reshape(120, 160, 3))
    st = r.stats()
    print("RANK0_DONE", st["rays_traced"])
dist.barrier()
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
    assert "RANK0_DONE" in p.stdout
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
