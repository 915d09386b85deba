# strong scaling of BASELINE configs[4] (3840x2160, 1024 spp of the 1.3 M-triangle scene split over the GPUs) and weak scaling of the
# default workload, on one multi-GPU box:  gpurun --gpus 8 -- 'bash tools/scale_run.sh "8 4 2"'
O=gpurun_out/scale; mkdir -p $O
for n in ${1:-8 4 2}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n bench.py --gpus $n --steps 2 --warmup 3 --no-cpu-baseline --workload mesh1m4k --scaling strong > $O/strong_n$n.json 2> $O/strong_n$n.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 296$n bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > $O/weak_n$n.json 2> $O/weak_n$n.err
done
for f in $O/*.json; do python -c "
import json
d=json.load(open('$f')); print('$f', d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['parallelism'][:60])" 2>&1 | tail -1; done
