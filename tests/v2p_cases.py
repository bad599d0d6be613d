"""Shared body of the v2p / KpinvSolve parity checks (CPU emulation and GPU): every case of tests/golden/v2p.npz, which
tests/golden/make_golden.py wrote by running the reference's own LDDMMModel.v2p (core/LDDMM.py:235-253 ->
tools/kernel.py:227-232) in fp32 ("ref32") and fp64 ("gold")."""
import numpy as np
import torch

from conftest import relerr

# Tolerances (fp32 inputs, compared with the reference's fp64 run):
#   truncated pseudo-inverse with a cut-off well inside a spectral gap (rcond 1e-3, 1e-1): momenta 2e-4, fitted speeds 2e-5
#     (measured: dense branch 2e-5 / 7e-7, matrix-free branch 3e-5 / 7e-7; the reference's own fp32 run: 4e-6 / 2e-7);
#   rcond=None (machine-precision cut-off): the momenta are not determined (reference fp32 vs fp64: factor 38..250 apart),
#     only the fitted speeds are compared: 1e-3 of their scale.
TOL_P, TOL_V, TOL_V_NONE = 2e-4, 2e-5, 1e-3


def check_v2p_against_reference(g, to_dev, spec, dense_max):
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.tools.kernel import GaussKernel
    old = GaussKernel.DENSE_SOLVE_MAX
    GaussKernel.DENSE_SOLVE_MAX = dense_max
    try:
        for tag in g["cases"]:
            tag = str(tag)
            D, M, sig, lam = g[f"{tag}_meta"]
            LM = LDDMMModel(sigma=float(sig), D=int(D), lambd=float(lam), version=tag.split("_")[-1], spec=spec)
            q, v = to_dev(g[f"{tag}_in_q"]), to_dev(g[f"{tag}_ref32_v"])
            sv = g[f"{tag}_gold_svals"]
            for rc_tag, rc in (("rc3", 1e-3), ("rc1", 1e-1), ("rcNone", None)):
                p = LM.v2p(q, v, rcond=rc)
                assert p.shape == q.shape and p.dtype == q.dtype and p.device == q.device
                vb = LM.v(q, q, p)
                if rc is None:
                    assert relerr(vb.cpu().numpy(), g[f"{tag}_gold_vback_{rc_tag}"]) < TOL_V_NONE, (tag, rc_tag)
                    continue
                # the fixture's cut-offs sit inside a spectral gap (no singular value within 0.1 % of the threshold:
                # fp32 eigenvalues are good to ~1e-6 relative, so the retained set is the reference's)
                assert np.abs(sv / (rc * sv[0]) - 1).min() > 1e-3, (tag, rc_tag)
                assert relerr(p.cpu().numpy(), g[f"{tag}_gold_p_{rc_tag}"]) < TOL_P, (tag, rc_tag)
                assert relerr(vb.cpu().numpy(), g[f"{tag}_gold_vback_{rc_tag}"]) < TOL_V, (tag, rc_tag)
                # zero target speeds: what DiffPSR.initialize_a0 asks for (exactly 0 for eta = 0, eta*GradK for logdet)
                p0 = LM.v2p(q, torch.zeros_like(q), rcond=rc)
                gold0 = g[f"{tag}_gold_pzero_{rc_tag}"]
                if LM.eta == 0:
                    assert not bool(p0.any()) and not gold0.any()
                else:
                    assert relerr(p0.cpu().numpy(), gold0) < TOL_P, (tag, rc_tag)
                    scale = float(np.abs(g[f"{tag}_gold_v"]).max())
                    assert np.abs(LM.v(q, q, p0).cpu().numpy() - g[f"{tag}_gold_vzero_{rc_tag}"]).max() < 1e-4 * scale
            # ridge solve (tools/kernel.py:234-242): matrix-free conjugate gradients vs the reference's dense solve
            rhs = v + LM.eta * LM.Kernel.GradKRed(q, q) if LM.eta else v
            pr = LM.Kernel.KridgeSolve_keops(q, rhs, alpha=1e-2)
            assert relerr(pr.cpu().numpy(), g[f"{tag}_gold_pridge_a2"]) < 5e-3, tag
            pr2 = LM.v2p(q, v, alpha=1e-2, version="ridge_keops")
            assert torch.equal(pr, pr2)
    finally:
        GaussKernel.DENSE_SOLVE_MAX = old
