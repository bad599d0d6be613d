python -m pytest tests/test_gpu_batched.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -2
for rho in 2.0 2.8; do
python scripts/groupwise_c4.py --frames 16 --iters 3 --lockstep 1 --rho $rho 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rho', $rho, d['lockstep_lbfgs'], 'M', d['support_points'], [round(x) for x in d['reg_opt_ms']], d['FE'])"
done
