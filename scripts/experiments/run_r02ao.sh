for fr in 8 16; do for G in 1 2; do
python scripts/groupwise_c4.py --frames $fr --iters 3 --lockstep 1 --groups $G --group-min 2 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('frames', $fr, 'groups', $G, 'M', d['support_points'], [round(x) for x in d['reg_opt_ms']], d['FE'])"
done; done 2>&1 | tee gpurun_out/r02ao_gw.txt
python scripts/groupwise_c4.py --frames 8 --iters 3 --lockstep 1 --groups 4 --group-min 2 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('frames 8 groups 4', 'M', d['support_points'], [round(x) for x in d['reg_opt_ms']], d['FE'])" | tee -a gpurun_out/r02ao_gw.txt
