show() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', d['value'], d['ms_per_step'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --workload bundled --accel emu"
$B > gpurun_out/r37_e7.json 2>gpurun_out/r37.err; show gpurun_out/r37_e7.json
for v in e6 e8; do PTAP_LIB=$PWD/pathtracerap_b200/variants/libptap_$v.so $B > gpurun_out/r37_$v.json 2>>gpurun_out/r37.err; show gpurun_out/r37_$v.json; done
tail -2 gpurun_out/r37.err
