set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_prover.py -x -q -k "mux or parity_host_api or parity_device_api or golden_seal_digests_on_gpu or witgen_accumulate or unsatisfied" 2>&1 | tail -6 | tee gpurun_out/r2_gpu8_pytest.log
for G in 8 6 4; do
  B2Z_EVALCHECK_GROUPS=$G timeout 900 python tools/prof_evalcheck.py --circuit rv32im_scale --po2 20 --reps 3 2>&1 | tail -1 | tee gpurun_out/r2_gpu8_evalcheck_staged_G$G.json
done
B2Z_EVALCHECK_STAGED=0 timeout 900 python tools/prof_evalcheck.py --circuit rv32im_scale --po2 20 --reps 3 2>&1 | tail -1 | tee gpurun_out/r2_gpu8_evalcheck_unstaged.json
timeout 900 python tools/prof_evalcheck.py --circuit rv32im_scale --po2 20 --reps 3 --prove 2>&1 | tail -1 | tee gpurun_out/r2_gpu8_segment_staged.json
timeout 900 python tools/prof_evalcheck.py --circuit planning --po2 20 --reps 3 2>&1 | tail -1
B2Z_EVALCHECK_STAGED=0 timeout 900 python tools/prof_evalcheck.py --circuit planning --po2 20 --reps 3 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none -k regex:b2z_eval_check -c 1 --csv --page raw --log-file gpurun_out/r2_prof_evalcheck_rv32_staged.csv python tools/prof_evalcheck.py --circuit rv32im_scale --po2 18 --reps 1 > gpurun_out/r2_gpu8_ncu.log 2>&1
tail -1 gpurun_out/r2_gpu8_ncu.log
