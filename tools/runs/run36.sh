# final bench lines of the other workloads (CUDA events, no profiler), for profiles/r02
O=gpurun_out/r36; mkdir -p $O
for w in mesh100k mesh5m bundled cornell; do python bench.py --workload $w --steps 3 --warmup 3 --no-extras > $O/bench_$w.json 2> $O/bench_$w.err; python -c "
import json; d=json.load(open('$O/bench_$w.json')); print('$w', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])"; done
python bench.py --workload bundled --accel emu --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_bundled_emu.json 2> $O/err.log; python -c "
import json; d=json.load(open('$O/bench_bundled_emu.json')); print('bundled emu', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
python bench.py --workload mesh1m --accel lbvh --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_mesh1m_ploc.json 2>> $O/err.log; python -c "
import json; d=json.load(open('$O/bench_mesh1m_ploc.json')); print('mesh1m ploc', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>> $O/err.log; head -c 300 $O/bench_reference.json; echo
