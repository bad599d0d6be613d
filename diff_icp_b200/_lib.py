"""ctypes binding of libdicp_b200.so (the C ABI declared in include/dicp_b200.h).

There is NO fallback: if the library is missing or was not built for this device the import of any
compute entry point raises.  torch is used for device memory, streams and autograd plumbing only.
"""

from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DICP_B200_LIB", os.path.join(_HERE, "libdicp_b200.so"))   # override: tuning sweeps only

_lib = None
_lock = threading.Lock()

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_f = ctypes.c_float
_int = ctypes.c_int
_u = ctypes.c_uint
_sz = ctypes.c_size_t

_dbl = ctypes.c_double


class LbfgsDev(ctypes.Structure):
    """`dicp_lbfgs_dev` of include/dicp_b200.h (device-resident lock-step L-BFGS state)."""
    _fields_ = [("K", _int), ("stride", ctypes.c_longlong), ("history", _int), ("max_iter", _int), ("max_eval", _int),
                ("max_ls", _int), ("tol_grad", _dbl), ("tol_change", _dbl), ("lr", _dbl), ("c1", _dbl), ("c2", _dbl),
                ("ints", _vp), ("dbl", _vp), ("vec", _vp), ("best_x", _vp), ("dirs", _vp), ("stps", _vp), ("ro", _vp),
                ("al", _vp), ("counters", _vp)]


_ldp = ctypes.POINTER(LbfgsDev)
# arguments of dicp_batch_closure_cluster up to `nscal` (shared by the dicp_lbfgs_dev_round / _loop_create entry points)
_CLOSURE_ARGS = [_int, _int, _f, _f, _int, _vp, _vp, _i64, _i64, _i64, _int, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _f, _vp, _i64,
                 _int]

# name -> (restype, argtypes); mirrors include/dicp_b200.h one to one
SIGNATURES = {
    "dicp_version": (_int, []),
    "dicp_sm_count": (_int, []),
    "dicp_sym_mode": (_int, [_int]),
    "dicp_launch_count": (ctypes.c_ulonglong, []),
    "dicp_pair_workspace_bytes": (_sz, [_i64, _i64]),
    "dicp_ksum": (_int, [_int, _u, _f, _vp, _i64, _vp, _i64, _vp, _vp, _vp] + [_vp] * 11 + [_vp, _sz, _vp]),
    "dicp_rhs_forward": (_int, [_int, _int, _f, _f, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _int]),
    "dicp_rhs_adjoint": (_int, [_int, _int, _f, _f, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _vp, _sz, _vp, _int]),
    "dicp_em_rowpass": (_int, [_int, _int, _f, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dicp_em_colstats": (_int, [_int, _f, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "dicp_em_lse_colstats": (_int, [_int, _f, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dicp_em_state_workspace_bytes": (_sz, [_i64, _i64]),
    "dicp_em_state_step": (_int, [_int, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int,
                                  _int, _int, _vp, _sz, _vp]),
    "dicp_em_loop_create": (_vp, [_int, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int,
                                  _int, _int, _vp, _sz, _vp]),
    "dicp_em_loop_launch": (_int, [_vp, _vp]),
    "dicp_em_loop_destroy": (None, [_vp]),
    "dicp_em_reduce_pack": (_int, [_int, _vp, _vp, _i64, _vp, _int, _vp, _vp]),
    "dicp_em_mstep_merged": (_int, [_int, _vp, _vp, _vp, _vp, _i64, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dicp_em_mstep": (_int, [_int, _vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "dicp_log_resp": (_int, [_int, _f, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp]),
    "dicp_small_max_support": (_int, []),
    "dicp_small_workspace_bytes": (_sz, [_i64, _i64]),
    "dicp_small_rhs_step": (_int, [_int, _int, _f, _f, _i64, _i64, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "dicp_small_adj_step": (_int, [_int, _int, _f, _f, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "dicp_batch_rhs_step": (_int, [_int, _int, _f, _f, _int, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _f, _f, _vp, _vp,
                                   _vp, _sz, _vp]),
    "dicp_batch_adj_step": (_int, [_int, _int, _f, _f, _int, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _f, _f,
                                   _vp, _vp, _vp, _sz, _vp]),
    "dicp_batch_set_p": (_int, [_int, _int, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp]),
    "dicp_batch_quad_workspace_bytes": (_sz, [_int]),
    "dicp_batch_quad_loss": (_int, [_int, _int, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _sz, _vp]),
    "dicp_batch_closure_out": (_int, [_int, _int, _vp, _vp, _i64, _i64, _f, _vp, _vp, _vp, _vp, _i64, _int, _vp]),
    "dicp_batch_closure_cluster_rows": (_int, [_int, _f, _int, _i64, _i64, _int, _int]),
    "dicp_batch_closure_cluster": (_int, [_int, _int, _f, _f, _int, _vp, _vp, _i64, _i64, _i64, _int, _vp, _i64, _vp, _i64, _vp,
                                          _vp, _i64, _f, _vp, _i64, _int, _vp]),
    "dicp_lbfgs_dev_begin": (_int, [_ldp, _vp, _vp, _i64, _vp, _vp]),
    "dicp_lbfgs_dev_round": (_int, [_ldp] + _CLOSURE_ARGS + [_vp]),
    "dicp_lbfgs_dev_loop_create": (_vp, [_ldp] + _CLOSURE_ARGS + [_int]),
    "dicp_lbfgs_dev_loop_launch": (_int, [_vp, _vp]),
    "dicp_lbfgs_dev_loop_destroy": (None, [_vp]),
    "dicp_batch_coverage": (_int, [_int, _int, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _int, _f, _vp, _vp]),
    "dicp_lbfgs_create": (_vp, [_int, _vp, _i64, _int, _int, _int, ctypes.c_double, ctypes.c_double]),
    "dicp_lbfgs_destroy": (None, [_vp]),
    "dicp_lbfgs_set_x": (_int, [_vp, _int, _vp]),
    "dicp_lbfgs_get_x": (_int, [_vp, _int, _vp, _int]),
    "dicp_lbfgs_reset": (_int, [_vp, _int, _int]),
    "dicp_lbfgs_begin_step": (_int, [_vp, _vp]),
    "dicp_lbfgs_pending": (_int, [_vp, _vp, _vp]),
    "dicp_lbfgs_feed": (_int, [_vp, _vp, _vp]),
    "dicp_lbfgs_stats": (_int, [_vp, _int, _vp]),
    "dicp_lbfgs_get_all": (_int, [_vp, _vp, _int]),
    "dicp_lbfgs_stats_all": (_int, [_vp, _vp]),
    "dicp_min2_sqdist": (_int, [_int, _vp, _i64, _vp, _vp]),
    "dicp_decimate_workspace_bytes": (_sz, [_i64]),
    "dicp_decimate_steps": (_int, [_int, _vp, _i64, _f, _int, _int, _vp, _vp, _sz, _vp]),
    "dicp_decimate_status": (_int, [_vp, _i64, _vp, _vp]),
    "dicp_quad_loss": (_int, [_int, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dicp_axpy": (_int, [_i64, _vp, _vp, _f, _vp, _f, _vp, _vp]),
    "dicp_pipe_probe": (_int, [_int, _int, _int, _vp, _vp]),
}


class DicpError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises DicpError if it is absent: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise DicpError(
                f"{LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). diff_icp_b200 has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        names = {-1: "bad argument", -2: "unsupported configuration", -3: "workspace too small"}
        raise DicpError(f"{what}: {names.get(rc, rc)}")
    raise DicpError(f"{what}: CUDA error {rc} ({torch.cuda.get_device_name() if torch.cuda.is_available() else 'no device'})")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (raw handle: ~20x cheaper than building a
    torch.cuda.Stream object for every launch)."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:          # private binding renamed: the public, slower way
        return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    """Mirror of the reference's getspec consistency check (tools/spec.py:39-43), restricted to what the
    kernels support: fp32, one CUDA device."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("diff_icp_b200 computes on a CUDA device only (no CPU fallback): got a CPU tensor")
        if t.dtype != torch.float32:
            raise ValueError("diff_icp_b200 kernels are fp32 (as the reference, tools/spec.py:24-27)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError("the different input tensors to this function should be on the same device and use the same dtype !")
    return dev


# ---- workspace cache (per device, per stream) -----------------------------------------------------------
_ws = {}


def workspace(rows: int, cols: int, device):
    """A persistent, 256-byte aligned scratch buffer large enough for any pair kernel of that size.
    One buffer per (device, stream): calls on one stream are ordered, so sharing is safe."""
    lib = load()
    need = int(lib.dicp_pair_workspace_bytes(int(rows), int(cols)))
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr())
    buf = _ws.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf
