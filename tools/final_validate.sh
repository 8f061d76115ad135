# Round-end validation on one B200: GPU parity suite, smoke, benches, ncu evidence (outputs under gpurun_out/)
set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_po2_22.json 2> gpurun_out/final_bench_po2_22.err
timeout 300 python bench.py --po2 20 --steps 5 --warmup 3 > gpurun_out/final_bench_po2_20.json 2> gpurun_out/final_bench_po2_20.err
timeout 300 python bench.py --po2 20 --steps 500 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_config3_500x_po2_20.json 2> gpurun_out/final_bench_config3.err
timeout 300 python tools/bench_micro.py > gpurun_out/final_micro_config5.json 2> gpurun_out/final_micro.err
timeout 300 ncu --set full --clock-control none -k regex:hash_rows -c 1 -f -o gpurun_out/final_hash_rows_po2_22 python tools/prof_ops.py --po2 22 --cols 192 --ops expand,hash_rows --reps 1 > gpurun_out/final_ncu_hash.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/final_launches.csv python bench.py --po2 20 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_launch_bench.log 2>&1
tail -c 600 gpurun_out/final_bench_po2_22.json
