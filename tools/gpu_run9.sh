set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_block_${N}gpu.json 2> gpurun_out/r2_bench_block_${N}gpu.err
tail -5 gpurun_out/r2_bench_block_${N}gpu.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_block_${N}gpu.json').read().strip().splitlines()[-1])
print('N',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'gather',d['config']['seal_gather_ms'],'verified',d['config']['receipt_verified_segments'],'launches',d['gpu_launches'])
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 0 | cut -c1-300
