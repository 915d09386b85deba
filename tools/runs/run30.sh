# cold ray state in shared memory (k_trace_bvh): parity of the refactored default, then A/B on mesh1m / mesh5m / bundled
python -m pytest tests/test_gpu_trace.py tests/test_gpu_production.py tests/test_gpu_scenes.py -m gpu -x -q 2>&1 | tail -2
show() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', d['value'], d['ms_per_step'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
for w in mesh1m mesh5m bundled; do
$B --workload $w > gpurun_out/r30_${w}_default.json 2>>gpurun_out/r30.err; show gpurun_out/r30_${w}_default.json
for v in cs8 cs9 cs10; do PTAP_LIB=$PWD/pathtracerap_b200/variants/libptap_$v.so $B --workload $w > gpurun_out/r30_${w}_$v.json 2>>gpurun_out/r30.err; show gpurun_out/r30_${w}_$v.json; done
done
PTAP_LIB=$PWD/pathtracerap_b200/variants/libptap_cs9.so python -m pytest tests/test_gpu_trace.py tests/test_gpu_production.py -m gpu -x -q 2>&1 | tail -2
tail -3 gpurun_out/r30.err
