# runs tools/trace_micro.py for the default library and every variant build; one JSON line each
python tools/trace_micro.py ${1:-mesh1m} default 2>>gpurun_out/micro.err | tail -1
for v in pathtracerap_b200/variants/*.so; do PTAP_LIB=$PWD/$v python tools/trace_micro.py ${1:-mesh1m} 2>>gpurun_out/micro.err | tail -1; done
tail -2 gpurun_out/micro.err
