set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_block_${N}gpu.json 2> gpurun_out/r2_bench_block_${N}gpu.err
tail -3 gpurun_out/r2_bench_block_${N}gpu.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_block_${N}gpu.json').read().strip().splitlines()[-1])
print('N',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'gather',d['config']['seal_gather_ms'],'verified',d['config']['receipt_verified_segments'],'launches',d['gpu_launches'])
P
if [ "$N" = "2" ]; then
timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_hal.py -x -q -k "seal_cache or session or evaluate_any or poly_divide or fri_fold or seal_parity_device_api or golden_seal_digests_on_gpu or mux" 2>&1 | tail -4
fi
