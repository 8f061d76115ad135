set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_gpu7_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# the driver's command line
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_block_po2_20.json 2> gpurun_out/r2_bench_block.err
tail -3 gpurun_out/r2_bench_block.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_block_po2_20.json').read().strip().splitlines()[-1])
print('BLOCK value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'gather',d['config']['seal_gather_ms'],'clocks',d['clocks'])
print(d['config']['phase_ms_per_segment']); print(d['cpu_baseline']); print([(k['kernel'][:30],round(k['ms'],2)) for k in d['kernels']], d['roofline']['ms_per_launch'], d['roofline']['frac'])
P
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/r2_bench_reference.json | cut -c1-400
timeout 900 python bench.py --in-flight 1 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_block_f1.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_block_f1.json').read().strip().splitlines()[-1]); print('F1 value',d['value'],'e2e',d['e2e']['value'])"
timeout 900 python bench.py --circuit planning --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_block_planning.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_block_planning.json').read().strip().splitlines()[-1]); print('PLANNING block value',d['value'],'e2e',d['e2e']['value'], d['config']['phase_ms_per_segment'])"
timeout 900 python bench.py --workload segment --circuit planning --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_segment_planning_po2_22.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_segment_planning_po2_22.json').read().strip().splitlines()[-1]); print('SEGMENT planning po2 22 value',d['value'],'e2e',d['e2e']['value'], d['phase_ms_per_step'])"
timeout 900 python bench.py --workload segment --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_segment_rv32_po2_22.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_segment_rv32_po2_22.json').read().strip().splitlines()[-1]); print('SEGMENT rv32 po2 22 value',d['value'],'e2e',d['e2e']['value'], d['phase_ms_per_step'])"
# launch list of the default bench command (shares) and full captures of the top kernels (CSV only: small)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/r2_launches_bench.log 2>&1
ls -la gpurun_out/r2_launches.csv
timeout 900 ncu --set full --clock-control none -k regex:hash_rows_mixed -c 1 --csv --page raw --log-file gpurun_out/r2_prof_hash_rows_mixed_po2_20.csv python tools/prof_ops.py --po2 20 --cols 192 --ops hash_rows --reps 1 > /dev/null 2>&1
ls -la gpurun_out/*.csv; du -sh gpurun_out
