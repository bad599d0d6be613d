python -m pytest tests/test_gpu_batched.py -x -q 2>&1 | tail -1
for fr in 8 16; do for R in 4 2; do
DICP_SMALL_MID_R=$R python scripts/groupwise_c4.py --frames $fr --iters 3 --lockstep 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('frames', $fr, 'R', $R, d['lockstep_lbfgs'], 'M', d['support_points'], [round(x) for x in d['reg_opt_ms']], d['FE'])"
done; done 2>&1 | tee gpurun_out/r02ah_gw.txt
python scripts/groupwise_c4.py --frames 8 --iters 3 --lockstep 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('frames 8 per-frame path', d['lockstep_lbfgs'], 'M', d['support_points'], [round(x) for x in d['reg_opt_ms']], d['FE'])" | tee -a gpurun_out/r02ah_gw.txt
