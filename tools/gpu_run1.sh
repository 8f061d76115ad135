set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
nproc
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_gpu1_pytest.log
cat gpurun_out/r2_gpu1_pytest.log
for cfg in "0 16" "1 16" "1 32"; do
  set -- $cfg
  echo "== NTT_BIG=$1 LANES=$2"
  B2Z_NTT_BIG=$1 B2Z_NTT_BIG_LANES=$2 timeout 600 python tools/prof_ops.py --po2 22 --cols 192 --ops expand,interp --reps 5 2>&1 | tail -3
  B2Z_NTT_BIG=$1 B2Z_NTT_BIG_LANES=$2 timeout 600 python tools/prof_ops.py --po2 20 --cols 192 --ops expand,interp --reps 5 2>&1 | tail -3
done > gpurun_out/r2_gpu1_ntt.log 2>&1
cat gpurun_out/r2_gpu1_ntt.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_gpu1_bench.json 2> gpurun_out/r2_gpu1_bench.err
tail -c 2500 gpurun_out/r2_gpu1_bench.json
tail -5 gpurun_out/r2_gpu1_bench.err
