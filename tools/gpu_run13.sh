# final single-GPU validation of the round-2 build: default bench, spin-vs-blocking host waits A/B, smoke, full GPU parity suite
set -x
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_block.json 2> gpurun_out/r2_final_bench_block.err
tail -2 gpurun_out/r2_final_bench_block.err
B2Z_SPIN_SYNC=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_bench_block_spin.json 2> gpurun_out/r2_final_bench_block_spin.err
python - <<P
import json
for f in ('r2_final_bench_block','r2_final_bench_block_spin'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f,'value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'launches',d['gpu_launches'],'lat',d['config']['median_segment_latency_ms'],'cpu',(d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
P
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
