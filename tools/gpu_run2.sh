set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_prover.py -x -q 2>&1 | tail -15 > gpurun_out/r2_gpu2_pytest.log
cat gpurun_out/r2_gpu2_pytest.log
# ncu: new two-pass NTT kernels at po2 22 (LDE 2^24), 32 columns
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_big -c 6 -o gpurun_out/r2_prof_ntt_big python tools/prof_ops.py --po2 22 --cols 32 --ops expand,interp --reps 1 > gpurun_out/r2_gpu2_ncu.log 2>&1
tail -3 gpurun_out/r2_gpu2_ncu.log
ls -la gpurun_out/*.ncu-rep
