"""EM_optimization with the loop state on the device: the whole loop is ONE CUDA graph launch -- a WHILE conditional node whose
body is one EM step and whose condition the step's last kernel sets on the device -- and one host read at the end
(reference loop: core/GMM.py:330-357 around EM_step :402-496 / :236-325).

The host loop reads sigma' and the free-energy sums after every step (one synchronisation and ~10 small launches per step:
about 0.23 ms per step for 640 000 points x 50 components, of which the kernels are ~0.1 ms).  Here sigma, the normalisation
constant, the previous free energy, the step count and the stop flag live in a small device array (csrc/ops_em.cuh, EmState);
the loop stops at the step that met |FE - FE_prev| < tol |FE_prev| or at the step limit, exactly where the host loop returns, and
only the steps the host loop would execute are executed.  Same kernels and same arithmetic as the step-by-step loop (tests/test_gpu_em_psr.py compares
the two bit for bit).  Applies to <= 64 components without the outlier term, on one GPU.
"""

from __future__ import annotations

import math

import torch

from ._lib import check, load, ptr, stream_ptr

_LOG2E = 1.4426950408889634
# indices of csrc/ops_em.cuh's EmState
(ES_SIGMA, ES_KAPPA, ES_LGN, ES_DONE, ES_HAVE_LAST, ES_LAST_FE, ES_STEPS, ES_CFE, ES_FE, ES_N, ES_TOL, ES_SIGMA_NEW,
 ES_MAXIT) = range(13)
ES_COUNT = 16


def _kappa_f32(sigma):
    """gauss_const(float sigma).kappa of csrc/dispatch.cuh: sqrt(log2(e)/2) / sigma, sigma and the result rounded to fp32."""
    s32 = torch.tensor(sigma, dtype=torch.float32).item()
    return torch.tensor(math.sqrt(0.5 * _LOG2E) / s32, dtype=torch.float32).item()


class EMLoopGraph:
    def __init__(self, N, C, D, device, do_mu, do_w, sig_mode, keops_sem):
        self.N, self.C, self.D, self.device = int(N), int(C), int(D), device
        self.flags = (int(bool(do_mu)), int(bool(do_w)), int(sig_mode), int(bool(keops_sem)))
        f32 = dict(dtype=torch.float32, device=device)
        self.X = torch.empty(N, D, **f32)
        self.mu, self.mu_new = torch.empty(C, D, **f32), torch.empty(C, D, **f32)
        self.w, self.w_new, self.lpi, self.lpi_new, self.wl2 = (torch.empty(C, **f32) for _ in range(5))
        self.stats = torch.empty(C, D + 3, **f32)
        self.Y = torch.empty(N, D, **f32)
        self.T2 = torch.empty(N, **f32)
        self.scal4 = torch.zeros(4, **f32)
        self.state = torch.zeros(ES_COUNT, dtype=torch.float64, device=device)
        self.h_state = torch.zeros(ES_COUNT, dtype=torch.float64).pin_memory()
        nbytes = int(load().dicp_em_state_workspace_bytes(self.N, self.C))
        if nbytes == 0:
            raise ValueError("EMLoopGraph: sizes outside the few-component kernels")
        self.ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.graph = None

    def __del__(self):
        g, self.graph = getattr(self, "graph", None), None
        if g:
            load().dicp_em_loop_destroy(g)

    def _args(self):
        do_mu, do_w, sig_mode, keops_sem = self.flags
        return (self.D, ptr(self.X), self.N, self.C, ptr(self.mu), ptr(self.w), ptr(self.lpi), ptr(self.wl2), ptr(self.mu_new),
                ptr(self.w_new), ptr(self.lpi_new), ptr(self.stats), ptr(self.Y), ptr(self.T2), ptr(self.scal4), ptr(self.state),
                do_mu, do_w, sig_mode, keops_sem, ptr(self.ws), self.ws.numel(), stream_ptr())

    def run(self, X, mu, w, lpi, sigma, tol, max_iterations):
        """Runs up to max_iterations EM steps from (mu, w, sigma).  Returns (Y, mu, w, lpi, sigma, Cfe, FE, steps, stopped):
        fresh tensors / Python floats of the model after the last executed step."""
        from .shooting import ShootPlan
        D, lib = self.D, load()
        self.X.copy_(X)
        self.mu.copy_(mu)
        self.w.copy_(w)
        self.lpi.copy_(lpi)
        lgn = D * (math.log(sigma) + 0.5 * math.log(2 * math.pi))
        torch.mul(lpi - lgn, _LOG2E, out=self.wl2)
        h = self.h_state
        h.zero_()
        h[ES_SIGMA], h[ES_KAPPA], h[ES_LGN], h[ES_N] = sigma, _kappa_f32(sigma), lgn, float(self.N)
        h[ES_TOL] = -1.0 if tol is None else float(tol)
        h[ES_MAXIT] = float(max_iterations)
        self.state.copy_(h, non_blocking=True)
        if self.graph is None:
            # first call: the steps one by one (every kernel returns at once after the stop flag is set), which also loads the
            # kernels; then the graph is built for the following calls
            for _ in range(int(max_iterations)):
                check(lib.dicp_em_state_step(*self._args()), "dicp_em_state_step")
            torch.cuda.current_stream().synchronize()
            with ShootPlan._lock:
                self.graph = lib.dicp_em_loop_create(*self._args())
            if not self.graph:
                raise RuntimeError("dicp_em_loop_create failed")
        else:
            check(lib.dicp_em_loop_launch(self.graph, stream_ptr()), "dicp_em_loop_launch")
        h.copy_(self.state, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return (self.Y.clone(), self.mu.clone(), self.w.clone(), self.lpi.clone(), float(h[ES_SIGMA]), float(h[ES_CFE]),
                float(h[ES_FE]), int(h[ES_STEPS]), bool(h[ES_DONE] != 0))
