python -m pytest tests/test_gpu_emulated.py -m gpu -x -q 2>&1 | tail -2
show() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', d['value'], d['ms_per_step'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --workload bundled --accel emu"
for r in 8 16 24 28 32; do PTAP_EMU_REFILL=$r $B > gpurun_out/r21_refill$r.json 2>>gpurun_out/r21.err; show gpurun_out/r21_refill$r.json; done
PTAP_EMU_REFILL=24 PTAP_EMU_REPLAY_CTAS=8 $B > gpurun_out/r21_refill24_c8.json 2>>gpurun_out/r21.err; show gpurun_out/r21_refill24_c8.json
tail -3 gpurun_out/r21.err
