set -x
mkdir -p gpurun_out
for T in 128 256 512 1024; do
  B2Z_EVALCHECK_THREADS=$T timeout 900 python tools/prof_evalcheck.py --circuit rv32im_scale --po2 20 --reps 3 > gpurun_out/r2_gpu4_evalcheck_$T.json 2> gpurun_out/r2_gpu4_evalcheck_$T.err
  cat gpurun_out/r2_gpu4_evalcheck_$T.json; tail -2 gpurun_out/r2_gpu4_evalcheck_$T.err
done
# ncu of the 128-thread and the best-guess 512-thread variants (short: po2 16)
for T in 128 512; do
B2Z_EVALCHECK_THREADS=$T timeout 900 ncu --set full --clock-control none -k regex:b2z_eval_check -c 1 -o gpurun_out/r2_prof_evalcheck_rv32_$T python tools/prof_evalcheck.py --circuit rv32im_scale --po2 18 --reps 1 > gpurun_out/r2_gpu4_ncu_$T.log 2>&1
tail -2 gpurun_out/r2_gpu4_ncu_$T.log
done
