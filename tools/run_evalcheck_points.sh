# A/B of B2Z_EVALCHECK_POINTS (coset points per thread of the generated eval_check): parity tests, then timings
B2Z_EVALCHECK_POINTS=4 timeout 100 python -m pytest tests/test_gpu_prover.py -x -q -m gpu -k "eval_check or seal_parity" 2>&1 | tail -1
for K in 1 4; do B2Z_EVALCHECK_POINTS=$K timeout 60 python tools/prof_evalcheck.py 20 2>&1 | tail -1; done
B2Z_EVALCHECK_POINTS=2 timeout 100 python -m pytest tests/test_gpu_prover.py -x -q -m gpu -k "eval_check or seal_parity" 2>&1 | tail -1
B2Z_EVALCHECK_POINTS=2 timeout 60 python tools/prof_evalcheck.py 20 2>&1 | tail -1
for K in 1 4 2; do B2Z_EVALCHECK_POINTS=$K timeout 60 python tools/prof_evalcheck.py 22 2>&1 | tail -1; done
