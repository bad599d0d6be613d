"""Experiment driver: packed-f32x2 classic RHS kernel vs the scalar product kernel (needs libdicp_b200_F2.so)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DICP_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diff_icp_b200", "libdicp_b200_F2.so")
from diff_icp_b200 import ops, shooting, _lib
from diff_icp_b200.core.LDDMM import LDDMMModel
lib = _lib.load()
M, D, dev = 20480, 3, torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
q = torch.rand(M, D, generator=g).to(dev); p = (1e-3 * torch.randn(M, D, generator=g)).to(dev)
LM = LDDMMModel(sigma=0.2, D=3, lambd=500.0, spec={"device": dev, "dtype": torch.float32}, version="classic", scheme="Euler", nt=10)
sp = LM._spec_for(M, 0, dev)
ws = ops.alloc_workspace(M, M, dev)
state = torch.cat([q.reshape(-1), p.reshape(-1), torch.zeros(1, device=dev)]); F = torch.zeros(sp.S + 3, device=dev)
def timeit(fn, n=20):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
t_ref = timeit(lambda: shooting._rhs(sp, state, F, ws))
vq_ref, dp_ref = F[:M * 3].view(M, 3).clone(), F[M * 3:2 * M * 3].view(M, 3).clone()
fn = lib.dicp_exp_f2
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4
vq, dp = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev)
ws2 = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
print("scalar classic fwd ms", t_ref)
for R in (2, 4):
    for ns in (8, 14, 15, 29, 30, 44, 59):
        call = lambda: fn(q.data_ptr(), p.data_ptr(), M, 0.2, R, ns, vq.data_ptr(), dp.data_ptr(), ws2.data_ptr(), torch.cuda.current_stream().cuda_stream)
        rc = call(); torch.cuda.synchronize()
        err = max(float((vq - vq_ref).abs().max() / vq_ref.abs().max()), float((dp - dp_ref).abs().max() / dp_ref.abs().max()))
        print("f2 R", R, "nsplit", ns, "ms", round(timeit(call), 4), "rc", rc, "relerr vs scalar", err)
