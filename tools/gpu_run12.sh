# tiled eval_check: parity first, then A/B timing (eval_check alone, whole segment, block bench)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_hal.py -x -q -k "eval_check or mux or baseline_size or golden or unsatisfied or session or parity" 2>&1 | tail -4
for T in 1 0; do
  B2Z_EVALCHECK_TILED=$T timeout 300 python tools/prof_evalcheck.py --circuit rv32im_scale --po2 20 --reps 5 --prove 2>&1 | tail -1 | cut -c1-400
done
B2Z_EVALCHECK_TILED=1 timeout 300 python tools/prof_evalcheck.py --circuit planning --po2 22 --reps 5 2>&1 | tail -1 | cut -c1-300
B2Z_EVALCHECK_TILED=0 timeout 300 python tools/prof_evalcheck.py --circuit planning --po2 22 --reps 5 2>&1 | tail -1 | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_block_tiled.json 2> gpurun_out/r2_bench_block_tiled.err
tail -2 gpurun_out/r2_bench_block_tiled.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_block_tiled.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'launches',d['gpu_launches'])
print(json.dumps(d.get('kernel_rooflines', d.get('roofline')))[:1500])
P
