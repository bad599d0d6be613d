for fr in 8 16 32; do
python scripts/groupwise_c4.py --frames $fr --iters 3 --lockstep 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('frames', $fr, d['lockstep_lbfgs'], 'M', d['support_points'], [round(x) for x in d['reg_opt_ms']], d['FE'])"
done 2>&1 | tee gpurun_out/r02al_gw.txt
