python -m pytest tests -m gpu -x -q > gpurun_out/r02w_pytest.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/r02w_pytest.log
rm -f gpurun_out/r02w_gw.txt
for g in 1 2; do python scripts/groupwise_iteration.py --iters 6 --groups $g 2>/dev/null | tail -1 >> gpurun_out/r02w_gw.txt; done
python scripts/groupwise_iteration.py --iters 6 --frames 8 2>/dev/null | tail -1 >> gpurun_out/r02w_gw.txt
python - <<EOP
import json
for l in open("gpurun_out/r02w_gw.txt"):
    d=json.loads(l); print(d["frames"], d["lockstep_groups"], [round(x,2) for x in d["gmm_opt_ms"]], [round(x,2) for x in d["reg_opt_ms"]], d["FE"])
EOP
