# device builder: leaves re-laid out by the collapse, subtrees of <= PTAP_PLOC_LEAF triangles as one leaf
for l in 1 4; do PTAP_PLOC_LEAF=$l python -m pytest tests/test_gpu_device_bvh.py -m gpu -x -q 2>&1 | tail -1; done
PTAP_PLOC_LEAF=2 PTAP_DEVICE_BUILDER=lbvh python -m pytest tests/test_gpu_device_bvh.py -m gpu -x -q 2>&1 | tail -1
show() { python -c "
import json,sys
d=json.load(open('$1')); r=d['roofline']; print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'nodes', r['avg_nodes_per_ray'], 'tris', r['avg_tris_per_ray'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --accel lbvh --workload mesh1m"
for l in 1 2 3 4; do PTAP_PLOC_LEAF=$l $B > gpurun_out/r46_leaf$l.json 2>>gpurun_out/r46.err; show gpurun_out/r46_leaf$l.json; done
tail -2 gpurun_out/r46.err
