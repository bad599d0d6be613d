for cfg in "64 4" "64 2" "32 4" "32 1" "4 2" "4 1"; do set -- $cfg
python scripts/groupwise_c4.py --frames $1 --iters 3 --lockstep 1 --groups $2 --group-min 2 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('frames', $1, 'groups', $2, 'M', d['support_points'], [round(x) for x in d['reg_opt_ms']], d['FE'])"
done 2>&1 | tee gpurun_out/r02ap_gw.txt
