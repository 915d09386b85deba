#!/bin/bash
# Regenerates the measurements under profiles/rNN on a GPU box:  gpurun --timeout 1500 -- 'bash tools/profile_round.sh'
# Bench lines first (CUDA events, no profiler attached), then the ncu launch list and the --set full captures of the same commands.
# Outputs land in gpurun_out/round/; summarise the .ncu-rep files with tools/ncu_summary.py and copy what is judged into profiles/.
O=gpurun_out/round; mkdir -p $O
for w in mesh1m mesh100k bundled cornell mesh5m; do
  python bench.py --workload $w --steps 3 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err
done
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --workload mesh1m --accel lbvh --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_mesh1m_lbvh.json 2> $O/err.log
python bench.py --workload bundled --accel grid --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_bundled_grid.json 2>> $O/err.log
python bench.py --workload mesh1m --accel grid --grid-dim 160 --spp 4 --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_mesh1m_grid160.json 2>> $O/err.log
python bench.py --workload mesh1m --accel grid --grid-dim 25 --spp 1 --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_mesh1m_grid25.json 2>> $O/err.log
# launch list of the default workload (2 spp keeps the list short; shares are per-launch-class, not absolute)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_mesh1m.csv python bench.py --spp 2 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
# --set full: primary, bounce-1 and bounce-2 launches of the closest-hit kernel; scan + shade of bounces 1-2; the grid walk on the bundled scene
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_trace_bvhILb0ELb0 -c 3 -o $O/prof_trace_bvh_mesh1m python bench.py --spp 2 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_scan' -s 2 -c 4 -o $O/prof_scan_shade_mesh1m python bench.py --spp 2 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_shade.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_scan' -s 2 -c 2 -o $O/prof_scan_shade_bundled python bench.py --workload bundled --spp 2 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_shade_b.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_trace_gridILb0ELb0 -s 1 -c 1 -o $O/prof_trace_grid_bundled python bench.py --workload bundled --accel grid --spp 2 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_grid.log 2>&1
ls -la $O > $O/ls.txt
