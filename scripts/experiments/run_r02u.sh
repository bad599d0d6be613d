python -m pytest tests/test_gpu_em_psr.py tests/test_gpu_api_parity.py -x -q 2>&1 | tail -3
python - <<EOP
import sys, json, torch
sys.argv=["bench.py"]
import bench
from diff_icp_b200 import ops
dev=torch.device("cuda:0")
def timeit(fn, n=20):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e-3)
    ts.sort(); return ts[len(ts)//2]
peaks={"ffma":3.535e13,"mufu_ex2":4.63e12}
r=bench.em_roofline(dev,timeit,peaks)
for k,v in r.items():
    if isinstance(v, dict):
        print(k, v["workload"], "em_step_us", round(v["em_step_s"]*1e6,1), v["binding"])
        for n in ("row_lite","col_stats","row_full","first_sweep_fused"):
            print("   ", n, round(v[n]["s_per_call"]*1e6,1), "us fp32", round(v[n]["frac_fp32"],3), "hbm", round(v[n]["frac_hbm"],3))
EOP
python scripts/groupwise_iteration.py --iters 6 2>/dev/null | tail -1 > gpurun_out/r02u_gw.txt
python - <<EOP
import json
for l in open("gpurun_out/r02u_gw.txt"):
    d=json.loads(l); print(d["frames"], d["lockstep_groups"], [round(x,2) for x in d["gmm_opt_ms"]], [round(x,2) for x in d["reg_opt_ms"]], d["FE"])
EOP
