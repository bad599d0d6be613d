python -m pytest tests/test_gpu_batched.py tests/test_gpu_em_psr.py -x -q > gpurun_out/r02n_pytest.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/r02n_pytest.log
for K in 64 8; do
for sh in "" "128,8" "64,16" "128,16" "64,8"; do
  echo "K=$K shape=$sh" >> gpurun_out/r02n_closure.txt
  DICP_CC_SHAPE=$sh python scripts/bench_batched_closure.py --K $K >> gpurun_out/r02n_closure.txt 2>&1
done; done
python scripts/bench_batched_closure.py --K 32 >> gpurun_out/r02n_closure.txt 2>&1
python scripts/bench_batched_closure.py --K 16 >> gpurun_out/r02n_closure.txt 2>&1
python scripts/bench_batched_closure.py --K 64 --D 3 --M 40 --N 12000 >> gpurun_out/r02n_closure.txt 2>&1
grep -o "shape=.*\|\"K\": [0-9]*\|closure_ms_median\": [0-9.]*" gpurun_out/r02n_closure.txt | paste - - - 
python scripts/groupwise_iteration.py --iters 5 2>/dev/null | tail -1 > gpurun_out/r02n_gw.txt; python scripts/groupwise_iteration.py --iters 5 --frames 8 2>/dev/null | tail -1 >> gpurun_out/r02n_gw.txt
python - <<EOP
import json
for l in open("gpurun_out/r02n_gw.txt"):
    d=json.loads(l); print(d["frames"], d["lockstep_groups"], [round(x,2) for x in d["gmm_opt_ms"]], [round(x,2) for x in d["reg_opt_ms"]], d["FE"])
EOP
