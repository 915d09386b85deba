show() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', d['value'], d['ms_per_step'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --accel emu --workload bundled"
$B > gpurun_out/r42_base.json 2>>gpurun_out/r42.err; show gpurun_out/r42_base.json
for v in sb0 sb0s1 sb2 s3; do PTAP_LIB=$PWD/pathtracerap_b200/variants/libptap_$v.so $B > gpurun_out/r42_$v.json 2>>gpurun_out/r42.err; show gpurun_out/r42_$v.json; done
