# ncu --set full of the final emulated-walk kernels, bounce-1 launches (the same command ran without ncu in run47's bench extras)
O=gpurun_out/r48; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:'k_trace_emu|k_emu_setup|k_emu_tail|k_emu_full' -s 4 -c 4 -o $O/prof_emu_final_bundled python bench.py --workload bundled --accel emu --spp 2 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu.log 2>&1
ls -la $O
