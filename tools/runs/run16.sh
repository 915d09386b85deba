# knobs of the emulated walk on the bundled scene (after a parity pass of the changed kernels)
python -m pytest tests/test_gpu_emulated.py -m gpu -x -q 2>&1 | tail -2
show() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', d['value'], d['ms_per_step'])" 2>&1 | tail -1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --workload bundled --accel emu"
$B > gpurun_out/r16_base.json 2>gpurun_out/r16.err; show gpurun_out/r16_base.json
PTAP_LANES=8 $B > gpurun_out/r16_lanes8.json 2>>gpurun_out/r16.err; show gpurun_out/r16_lanes8.json
PTAP_LANES=6 $B > gpurun_out/r16_lanes6.json 2>>gpurun_out/r16.err; show gpurun_out/r16_lanes6.json
PTAP_EMU_WALK_CTAS=7 $B > gpurun_out/r16_walk7.json 2>>gpurun_out/r16.err; show gpurun_out/r16_walk7.json
PTAP_EMU_WALK_CTAS=1 $B > gpurun_out/r16_walk1.json 2>>gpurun_out/r16.err; show gpurun_out/r16_walk1.json
PTAP_EMU_REPLAY_CTAS=4 $B > gpurun_out/r16_replay4.json 2>>gpurun_out/r16.err; show gpurun_out/r16_replay4.json
PTAP_EMU_REPLAY_CTAS=16 $B > gpurun_out/r16_replay16.json 2>>gpurun_out/r16.err; show gpurun_out/r16_replay16.json
PTAP_VOTE_TRI=4 $B > gpurun_out/r16_vt4.json 2>>gpurun_out/r16.err; show gpurun_out/r16_vt4.json
PTAP_VOTE_TRI=12 $B > gpurun_out/r16_vt12.json 2>>gpurun_out/r16.err; show gpurun_out/r16_vt12.json
tail -3 gpurun_out/r16.err
