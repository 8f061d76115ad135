set -x
mkdir -p gpurun_out
for K in 1 2 3 4; do
  B2Z_EVALCHECK_TILES=$K timeout 900 python tools/prof_evalcheck.py --circuit rv32im_scale --po2 20 --reps 3 2>&1 | tail -1 | tee gpurun_out/r2_gpu6_evalcheck_K$K.json
done
B2Z_EVALCHECK_TILES=2 timeout 900 python -m pytest tests/test_gpu_prover.py -x -q -k "mux or golden_seal_digests_on_gpu or parity_device_api_planning" 2>&1 | tail -4
B2Z_EVALCHECK_TILES=2 timeout 900 ncu --set full --clock-control none -k regex:b2z_eval_check -c 1 --csv --page raw --log-file gpurun_out/r2_prof_evalcheck_rv32_K2.csv python tools/prof_evalcheck.py --circuit rv32im_scale --po2 18 --reps 1 > gpurun_out/r2_gpu6_ncu.log 2>&1
tail -1 gpurun_out/r2_gpu6_ncu.log
