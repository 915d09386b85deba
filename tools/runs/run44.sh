python -m pytest tests/test_gpu_device_bvh.py tests/test_gpu_scenes.py -m gpu -x -q 2>&1 | tail -2
show() { python -c "
import json,sys
d=json.load(open('$1')); r=d['roofline']; print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])" 2>&1 | tail -1; }
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --accel lbvh"
for w in mesh1m mesh5m mesh100k; do $B --workload $w > gpurun_out/r44_${w}_lbvh.json 2>>gpurun_out/r44.err; show gpurun_out/r44_${w}_lbvh.json; done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r44_default.json 2>>gpurun_out/r44.err; python -c "
import json; d=json.load(open('gpurun_out/r44_default.json')); print(d['value'], d['e2e']['value'], json.dumps(d['extras']['e2e_build_included']))"
tail -3 gpurun_out/r44.err
