# N-GPU block bench on the final build (cross-rank retry protocol on NCCL) + latency distribution diagnostics
set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_final_bench_block_${N}gpu.json 2> gpurun_out/r2_final_bench_block_${N}gpu.err
tail -3 gpurun_out/r2_final_bench_block_${N}gpu.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_final_bench_block_${N}gpu.json').read().strip().splitlines()[-1])
c=d['config']
print('N',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'gather',c['seal_gather_ms'],'launches',d['gpu_launches'])
print('rank_prove_ms',c['rank_prove_ms'],'lat',c['median_segment_latency_ms'],c['segment_latency_ms'],'reassigned',c['segments_reassigned_across_ranks'])
P
