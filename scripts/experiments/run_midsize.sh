for rho in 2.0 2.8 4.0; do for ls in 1 0; do
python scripts/groupwise_c4.py --frames 16 --iters 3 --lockstep $ls --rho $rho 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rho', $rho, 'lockstep', $ls, d['lockstep_lbfgs'], 'M', d['support_points'], [round(x) for x in d['reg_opt_ms']], d['FE'])"
done; done
