"""The reference-side binding of INTEGRATION.md (option B), as executable code: adds ``computversion="b200"`` to the
reference's two dispatch seams WITHOUT modifying the reference's files.

  * ``GenKernel.set_computversion`` (reference tools/kernel.py:91-110) rebinds the ten reduction attributes; the "b200"
    branch binds each of them to one ``dicp_ksum`` call of ``libdicp_b200.so`` through ctypes -- raw device pointers,
    caller-allocated output and workspace, the current CUDA stream (include/dicp_b200.h).
  * ``GaussianMixtureUnif.set_computversion`` (reference core/GMM.py:126-144) rebinds ``EM_step``; the "b200" branch
    binds it to ``dicp_em_rowpass`` / ``dicp_em_colstats`` / ``dicp_em_mstep`` through the drop-in GMM's EM step acting on
    the reference object's own state (mu, w, sigma, outliers, to_optimize), in the torch-twin ordering of the M step
    (what the reference executes when pykeops is absent) unless ``b200_ordering = "keops"`` is set on the object.

Usage (what tests/test_gpu_reference_seam.py does with the unmodified reference package):

    import diffICP.tools.kernel as rk, diffICP.core.GMM as rg
    from diff_icp_b200 import reference_seam
    reference_seam.install(rk, rg)
    LM = diffICP.core.LDDMM.LDDMMModel(sigma, D, lam, computversion="b200", spec=gpuspec)    # reference class, B200 kernels

The reductions bound here are forward-only (the reference's autograd walks its own torch / KeOps ops); the differentiable
path is the drop-in package itself (option A).
"""

from __future__ import annotations

import ctypes

import torch

from . import _lib

# (attribute, DICP_K_* selector, output slot, vector valued, argument adapter) -- include/dicp_b200.h, dicp_ksum
_REDUCTIONS = [
    ("KBase",        1,    0,  False, lambda x, y: (x, y, None, None, None)),
    ("KRedScal",     2,    1,  False, lambda x, y, d: (x, y, None, None, d.reshape(-1))),
    ("KRed",         4,    2,  True,  lambda x, y, b: (x, y, b, None, None)),
    ("GradKRed",     8,    3,  True,  lambda x, y: (x, y, None, None, None)),
    ("DDKRed",       16,   4,  True,  lambda x, y, b: (x, y, b, None, None)),
    ("GenDKRed",     32,   5,  True,  lambda x, y, b, c: (x, y, b, c, None)),
    ("HessKRed",     64,   6,  True,  lambda x, y, b, c: (x, y, b, c, None)),
    ("LapKRed",      128,  7,  False, lambda x, y: (x, y, None, None, None)),
    ("GradLapKRed",  256,  8,  True,  lambda x, y: (x, y, None, None, None)),
    # reduction over i:  sum_i gradK(x_i - y_j).d_i  = rows y, columns x, per-column vector d   (DICP_K_DOT)
    ("GradKRed_rev", 1024, 10, False, lambda x, y, d: (y, x, d, None, None)),
]


def _reduction(kernel, selector, slot, vector_valued, adapt):
    lib = _lib.load()

    def call(*args):
        x, y, b, c, d = adapt(*args)
        ts = [None if t is None else t.detach().contiguous() for t in (x, y, b, c, d)]
        x, y = ts[0], ts[1]
        if not x.is_cuda or x.dtype != torch.float32:
            raise ValueError("computversion='b200' computes on CUDA in fp32 only")
        M, D = x.shape
        N = y.shape[0]
        out = torch.empty((M, D) if vector_valued else (M,), dtype=torch.float32, device=x.device)
        ws = torch.empty(int(lib.dicp_pair_workspace_bytes(M, N)), dtype=torch.uint8, device=x.device)
        outs = [None] * 11
        outs[slot] = ctypes.c_void_p(out.data_ptr())
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(x.device):
            rc = lib.dicp_ksum(D, selector, float(kernel.sigma), p(ts[0]), M, p(ts[1]), N, p(ts[2]), p(ts[3]), p(ts[4]),
                               *outs, ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                               ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        if rc:
            raise RuntimeError(f"dicp_ksum failed: {rc}")
        return out
    return call


def _em_step_b200(gmm):
    from .core.GMM import GaussianMixtureUnif as B200GMM

    def EM_step(X, skip_M=False):
        ordering = getattr(gmm, "b200_ordering", "torch")
        twin = B200GMM(gmm.mu, sigma=gmm.sigma, use_outliers=gmm.outliers is not None,
                       spec={"device": X.device, "dtype": X.dtype}, computversion=ordering)
        twin.w = gmm.w
        twin.to_optimize = dict(gmm.to_optimize)
        twin.ensure_continuum = gmm.ensure_continuum
        if gmm.outliers is not None:
            twin.outliers = dict(gmm.outliers)
        Y, Cfe, FE = twin.EM_step(X, skip_M=skip_M)
        gmm.mu, gmm.w, gmm.sigma = twin.mu, twin.w, twin.sigma
        if gmm.outliers is not None:
            gmm.outliers.update(twin.outliers)
        return Y, Cfe, FE
    return EM_step


def install(kernel_module=None, gmm_module=None):
    """Wrap the two ``set_computversion`` methods of the (imported, unmodified) reference modules so that they accept
    "b200".  Idempotent."""
    if kernel_module is not None and not getattr(kernel_module.GenKernel, "_b200_seam", False):
        orig_k = kernel_module.GenKernel.set_computversion

        def set_computversion(self, version):
            if version != "b200":
                return orig_k(self, version)
            for name, sel, slot, vec, adapt in _REDUCTIONS:
                setattr(self, name, _reduction(self, sel, slot, vec, adapt))
            self.computversion = version
        kernel_module.GenKernel.set_computversion = set_computversion
        kernel_module.GenKernel._b200_seam = True
    if gmm_module is not None and not getattr(gmm_module.GaussianMixtureUnif, "_b200_seam", False):
        orig_g = gmm_module.GaussianMixtureUnif.set_computversion

        def set_computversion_gmm(self, version):
            if version != "b200":
                return orig_g(self, version)
            self.EM_step = _em_step_b200(self)
            self.computversion = version
            return self
        gmm_module.GaussianMixtureUnif.set_computversion = set_computversion_gmm
        gmm_module.GaussianMixtureUnif._b200_seam = True
