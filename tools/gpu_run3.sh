set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_hal.py tests/test_gpu_prover.py -x -q -k "ntt or two_pass or config5 or mux or roundtrip or full_size or hash_rows_mixed_default" 2>&1 | tail -12 > gpurun_out/r2_gpu3_pytest.log
cat gpurun_out/r2_gpu3_pytest.log
{
for cfg in "1 16"; do
  set -- $cfg
  echo "== NTT_BIG=$1 LANES=$2 (lazy/shoup butterflies)"
  B2Z_NTT_BIG=$1 B2Z_NTT_BIG_LANES=$2 timeout 600 python tools/prof_ops.py --po2 22 --cols 192 --ops expand,interp --reps 5 2>&1 | tail -2
  B2Z_NTT_BIG=$1 B2Z_NTT_BIG_LANES=$2 timeout 600 python tools/prof_ops.py --po2 20 --cols 192 --ops expand,interp --reps 5 2>&1 | tail -2
done
echo "== NTT_BIG=0 (multi-pass, lazy/shoup butterflies)"
B2Z_NTT_BIG=0 timeout 600 python tools/prof_ops.py --po2 22 --cols 192 --ops expand,interp --reps 5 2>&1 | tail -2
echo "== hash: mixed baseline"
timeout 600 python tools/prof_ops.py --po2 22 --cols 192 --ops hash_rows --reps 3 2>&1 | tail -1
for cfg in "4 4" "5 3" "4 3" "3 4" "6 3" "5 4" "6 2" "8 2" "10 0" "0 6"; do
  set -- $cfg
  echo "== hash split INT=$1 F64=$2"
  B2Z_HASH_SPLIT=1 B2Z_HASH_SPLIT_INT=$1 B2Z_HASH_SPLIT_F64=$2 timeout 600 python tools/prof_ops.py --po2 22 --cols 192 --ops hash_rows --reps 3 2>&1 | tail -1
done
} > gpurun_out/r2_gpu3_micro.log 2>&1
cat gpurun_out/r2_gpu3_micro.log
timeout 900 python tools/prof_evalcheck.py --circuit rv32im_scale --po2 20 --reps 3 --prove > gpurun_out/r2_gpu3_evalcheck_128.json 2> gpurun_out/r2_gpu3_evalcheck_128.err
cat gpurun_out/r2_gpu3_evalcheck_128.json; tail -3 gpurun_out/r2_gpu3_evalcheck_128.err
timeout 900 python tools/prof_evalcheck.py --circuit planning --po2 20 --reps 3 --prove > gpurun_out/r2_gpu3_evalcheck_planning.json 2>&1
cat gpurun_out/r2_gpu3_evalcheck_planning.json
timeout 1200 python bench.py --steps 3 --warmup 1 > gpurun_out/r2_gpu3_bench_block.json 2> gpurun_out/r2_gpu3_bench_block.err
tail -c 3000 gpurun_out/r2_gpu3_bench_block.json; tail -5 gpurun_out/r2_gpu3_bench_block.err
timeout 1200 python bench.py --steps 3 --warmup 1 --in-flight 1 --no-cpu-baseline > gpurun_out/r2_gpu3_bench_block_f1.json 2> gpurun_out/r2_gpu3_bench_block_f1.err
head -c 600 gpurun_out/r2_gpu3_bench_block_f1.json; tail -3 gpurun_out/r2_gpu3_bench_block_f1.err
