# 2-GPU sanity of the final tree: multi-GPU tests, then bench.py --gpus 2 (weak) under torchrun, as the driver launches it
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r31_n2.json 2> gpurun_out/r31_n2.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r31_n2.json')); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['scaling'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r31_ref_n2.json 2>> gpurun_out/r31_n2.err; echo "ref rc=$?"; head -c 400 gpurun_out/r31_ref_n2.json; echo
tail -3 gpurun_out/r31_n2.err
