# ncu captures of the mixed INT/FP64 hash_rows kernel: args = "period n_int ctas" triples
set -x
for cfg in "$@"; do
  set -- $cfg
  B2Z_HASH_ROLE_PERIOD=$1 B2Z_HASH_ROLE_INT=$2 B2Z_HASH_CTAS=$3 timeout 150 ncu --set full --clock-control none -k regex:hash_rows -c 1 -f -o gpurun_out/mixed_p$1_i$2_c$3 python tools/prof_ops.py --po2 20 --cols 64 --ops expand,hash_rows --reps 1 > gpurun_out/ncu_mixed_p$1_i$2_c$3.log 2>&1
done
