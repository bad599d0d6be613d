"""Pipe-throughput probe: measures FFMA / FFMA2 / MUFU.EX2 rates on the device (roofline denominators)."""
import json, sys, torch
sys.path.insert(0, ".")
from diff_icp_b200 import ops
dev = torch.device("cuda:0")
sms = ops.load().dicp_sm_count()
res = {"sms": sms}
for name, which, per in [("ffma", 0, 1), ("ffma2", 1, 2), ("mufu_ex2", 2, 1)]:
    blocks, iters = sms * 8, 20000
    out = torch.empty(blocks * 256, device=dev)
    ops.pipe_probe(which, blocks, 200, out)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.pipe_probe(which, blocks, iters, out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    n_instr = blocks * 256 * iters * (8 if which != 1 else 4)
    res[name] = {"ops_per_s": n_instr * per / best, "per_sm_per_clk_at_1965MHz": n_instr * per / best / sms / 1.965e9}
print(json.dumps(res))
