set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_prover.py -x -q -k "mux or session or pinned or unsatisfied or golden_seal_digests_on_gpu" 2>&1 | tail -8 | tee gpurun_out/r2_gpu5_pytest.log
timeout 900 python tools/prof_evalcheck.py --circuit rv32im_scale --po2 20 --reps 3 --prove 2>&1 | tail -1 | tee gpurun_out/r2_gpu5_evalcheck.json
B2Z_EVALCHECK_DOT=1 timeout 900 ncu --set full --clock-control none -k regex:b2z_eval_check -c 1 --csv --page raw --log-file gpurun_out/r2_prof_evalcheck_rv32.csv python tools/prof_evalcheck.py --circuit rv32im_scale --po2 18 --reps 1 > gpurun_out/r2_gpu5_ncu.log 2>&1
tail -2 gpurun_out/r2_gpu5_ncu.log; ls -la gpurun_out/r2_prof_evalcheck_rv32.csv
timeout 1200 python bench.py --steps 5 --warmup 2 > gpurun_out/r2_gpu5_bench_block.json 2> gpurun_out/r2_gpu5_bench_block.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_gpu5_bench_block.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'gather',d['config']['seal_gather_ms'],'lat',d['config']['median_segment_latency_ms'])
print(d['config']['phase_ms_per_segment']); print(d['cpu_baseline'])
"
tail -3 gpurun_out/r2_gpu5_bench_block.err
du -sh gpurun_out
