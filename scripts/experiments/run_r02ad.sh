python -m pytest tests/test_gpu_batched.py -x -q > gpurun_out/r02ad_pytest.log 2>&1; tail -3 gpurun_out/r02ad_pytest.log
for rho in 1.4142 2.0 2.8 4.0; do
python scripts/groupwise_c4.py --frames 16 --iters 3 --lockstep 1 --rho $rho 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rho', $rho, d['lockstep_lbfgs'], 'M', d['support_points'], [round(x) for x in d['reg_opt_ms']], d['FE'])"
done 2>&1 | tee gpurun_out/r02ad_gw.txt
