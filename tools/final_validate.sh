# Round-2 validation on one B200 (outputs under gpurun_out/): GPU parity suite, smoke, the driver's bench command, the
# reference arm, the config-5 microbench, a launch list and one full ncu capture of the dominant kernel.
#   gpurun --timeout 1500 -- 'bash tools/final_validate.sh'
set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_block.json 2> gpurun_out/bench_block.err
timeout 400 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python bench.py --workload segment --circuit planning --po2 22 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_segment_planning_po2_22.json 2> gpurun_out/bench_segment.err
timeout 300 python tools/bench_micro.py > gpurun_out/micro_config5.json 2> gpurun_out/micro.err
# launch list (cold-cache, serialised: compare SHARES) and one full capture; text/CSV pages only -- a .ncu-rep embeds the 32 MB cubin
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --in-flight 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:hash_rows_mixed -c 1 --csv --page raw --log-file gpurun_out/prof_hash_rows_mixed.csv python tools/prof_ops.py --po2 20 --cols 192 --ops expand,hash_rows --reps 1 > gpurun_out/prof_hash.log 2>&1
tail -c 700 gpurun_out/bench_block.json
