# PLOC with the upper levels built by the host's binned SAH over the last clusters: parity, then trace rate and build time by cluster count
python -m pytest tests/test_gpu_device_bvh.py tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -2
show() { python -c "
import json,sys
d=json.load(open('$1')); r=d['roofline']; print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'nodes', r['avg_nodes_per_ray'], 'tris', r['avg_tris_per_ray'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --accel lbvh"
for t in 0 1024 4096 16384 65536; do PTAP_PLOC_TOP=$t $B --workload mesh1m > gpurun_out/r43_mesh1m_top$t.json 2>>gpurun_out/r43.err; show gpurun_out/r43_mesh1m_top$t.json; done
for t in 0 16384; do PTAP_PLOC_TOP=$t $B --workload mesh5m > gpurun_out/r43_mesh5m_top$t.json 2>>gpurun_out/r43.err; show gpurun_out/r43_mesh5m_top$t.json; done
for t in 0 4096 16384 65536; do PTAP_PLOC_TOP=$t python -c "
import sys; sys.path.insert(0,'.')
import bench
from pathtracerap_b200 import ACCEL_BVH_DEVICE, Renderer
s,a = bench.build_scene('mesh1m')
r = Renderer(width=64,height=32,depth=5,accel=ACCEL_BVH_DEVICE); r.allocateOnGPU(s); r.upload(s); r.upload(s); print('top $t device build', r.build_stats()); r.free()
" 2>&1 | tail -1; done
tail -2 gpurun_out/r43.err
