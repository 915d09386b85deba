# ncu --set full captures of the current closest-hit kernels (after the same commands ran without ncu in run17 / run15)
O=gpurun_out/r19; mkdir -p $O
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_trace_bvhILb0ELb0 -c 3 -o $O/prof_trace_bvh_mesh1m python bench.py --spp 2 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_trace_emu|k_emu_replay' -s 2 -c 4 -o $O/prof_emu_bundled python bench.py --workload bundled --accel emu --spp 2 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_emu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_bundled_emu.csv python bench.py --workload bundled --accel emu --spp 2 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_launches.log 2>&1
ls -la $O
