show() { python -c "
import json,sys
d=json.load(open('$1')); r=d['roofline']; print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'nodes', r['avg_nodes_per_ray'], 'tris', r['avg_tris_per_ray'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --accel lbvh --workload mesh1m"
PTAP_PLOC_ADAPT=0 $B > gpurun_out/r45_a0.json 2>>gpurun_out/r45.err; show gpurun_out/r45_a0.json
PTAP_PLOC_ADAPT=1 $B > gpurun_out/r45_a1.json 2>>gpurun_out/r45.err; show gpurun_out/r45_a1.json
PTAP_PLOC_ADAPT=1 PTAP_PLOC_TOP=1024 $B > gpurun_out/r45_a1_t1024.json 2>>gpurun_out/r45.err; show gpurun_out/r45_a1_t1024.json
PTAP_PLOC_ADAPT=1 python -m pytest tests/test_gpu_device_bvh.py -m gpu -x -q 2>&1 | tail -1
tail -2 gpurun_out/r45.err
