python -m pytest tests/test_gpu_emulated.py -m gpu -x -q 2>&1 | tail -2
show() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', d['value'], d['ms_per_step'])" 2>&1 | tail -1; }
for w in "bundled" "mesh1m --spp 16"; do n=${w%% *}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --accel emu --workload $w"
$B > gpurun_out/r41_${n}_fb8.json 2>>gpurun_out/r41.err; show gpurun_out/r41_${n}_fb8.json
for v in fb32 fb4; do PTAP_LIB=$PWD/pathtracerap_b200/variants/libptap_$v.so $B > gpurun_out/r41_${n}_$v.json 2>>gpurun_out/r41.err; show gpurun_out/r41_${n}_$v.json; done
done
