"""GPU: INTEGRATION.md option B EXECUTED -- the unmodified reference package (shipped as the git-ignored install
baseline/_ref, loaded through oracle/ref_loader.py) with diff_icp_b200.reference_seam installed: the reference's own
GaussKernel / LDDMMModel / GaussianMixtureUnif objects run with computversion="b200", i.e. with libdicp_b200.so behind
their two dispatch seams (tools/kernel.py:91-110, core/GMM.py:126-144), and reproduce the reference's fp64 outputs."""
import numpy as np
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def spec():
    return {"device": dev(), "dtype": torch.float32}


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(dev())


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_loader
    if ref_loader.find_root() is None:
        pytest.skip("reference install baseline/_ref not present (run __graft_entry__.build() in the build container)")
    r = ref_loader.load_reference(fix_coverage=True)
    from diff_icp_b200 import reference_seam
    reference_seam.install(r.kernel, r.GMM)
    return r


def test_reference_kernel_object_runs_on_the_c_abi(ref, golden):
    g = golden("kernels")
    for tag in ("a", "b", "c"):
        M, N, D, sig = g[f"{tag}_meta"]
        GK = ref.kernel.GaussKernel(float(sig), int(D), computversion="b200", spec=spec())      # the REFERENCE's class
        assert GK.computversion == "b200" and type(GK).__module__ == "diffICP.tools.kernel"
        x, y, b, c, d = (cu(g[f"{tag}_in_{n}"]) for n in "xybcd")
        res = {"KBase": GK.KBase(x, y), "KRedScal": GK.KRedScal(x, y, d), "KRed": GK.KRed(x, y, b),
               "GradKRed": GK.GradKRed(x, y), "DDKRed": GK.DDKRed(x, y, b), "GenDKRed": GK.GenDKRed(x, y, b, c),
               "HessKRed": GK.HessKRed(x, y, b, c), "LapKRed": GK.LapKRed(x, y), "GradLapKRed": GK.GradLapKRed(x, y),
               "GradKRed_rev": GK.GradKRed_rev(x, y, c)}
        for name, out in res.items():
            gold, r32 = g[f"{tag}_gold_{name}"], g[f"{tag}_ref32_{name}"]
            assert out.shape == gold.shape, (tag, name)
            assert relerr(out.cpu().numpy(), gold) < max(1e-5, 2 * relerr(r32, gold)), (tag, name)
        # unknown strings still raise like the reference; the reference's own strings still work
        with pytest.raises(ValueError):
            GK.set_computversion("nope")


def test_reference_shoot_forward_through_the_seam(ref, golden):
    """The reference's own LDDMMModel.Shoot (core/LDDMM.py:286-299: its Python integrator loop, its ODE composing
    v / GenDKRed / HessKRed / GradLapKRed / mdivsum) with every reduction served by dicp_ksum."""
    g = golden("lddmm")
    for tag in g["cases"]:
        tag = str(tag)
        D, Nq, Nx, nt, sig, lam = g[f"{tag}_meta"]
        _, version, scheme, xmode = tag.split("_")
        LM = ref.LDDMM.LDDMMModel(sigma=float(sig), D=int(D), lambd=float(lam), spec=spec(), version=version,
                                  computversion="b200", scheme=scheme, nt=int(nt))
        assert type(LM).__module__ == "diffICP.core.LDDMM" and LM.Kernel.computversion == "b200"
        q0, p0 = cu(g[f"{tag}_in_q0"]), cu(g[f"{tag}_in_p0"])
        x0 = cu(g[f"{tag}_in_x0"]) if xmode == "x" else None
        with torch.no_grad():
            sh = LM.Shoot(q0, p0, x0)
            H0 = LM.Hamiltonian(q0, p0)
        assert len(sh) == int(nt) + 1
        assert relerr(sh[-1][0].cpu().numpy(), g[f"{tag}_gold_q1"]) < 5e-5, tag
        assert relerr(sh[-1][1].cpu().numpy(), g[f"{tag}_gold_p1"]) < 5e-5, tag
        assert abs(float(sh[-1][2].sum()) - float(g[f"{tag}_gold_cost1"].sum())) < 5e-5 * max(1.0, abs(float(g[f"{tag}_gold_cost1"].sum()))), tag
        if x0 is not None:
            assert relerr(sh[-1][3].cpu().numpy(), g[f"{tag}_gold_x1"]) < 5e-5, tag
        assert abs(float(H0) - float(g[f"{tag}_gold_H0"])) < 2e-5 * abs(float(g[f"{tag}_gold_H0"])), tag


def test_reference_gmm_object_runs_em_on_the_c_abi(ref, golden):
    g = golden("gmm")
    for tag in ("2d_full", "3d_full", "3d_frozen", "2d_opt5", "3d_offset", "2d_skipM"):
        D, N, C, outl, skip, steps, sig0 = g[f"{tag}_meta"]
        G = ref.GMM.GaussianMixtureUnif(cu(g[f"{tag}_in_mu"]), sigma=float(sig0), spec=spec(), computversion="b200")
        assert type(G).__module__ == "diffICP.core.GMM" and G.computversion == "b200"
        G.w = cu(g[f"{tag}_in_w"])
        G.to_optimize = dict(zip(("mu", "sigma", "w", "eta0"), (bool(v) for v in g[f"{tag}_opt"])))
        X = cu(g[f"{tag}_in_X"])
        fes = []
        for _ in range(int(steps)):
            Y, Cfe, FE = G.EM_step(X, skip_M=bool(skip))
            fes.append(float(FE))

        def ok(a, key, base=2e-5):
            gold, r32 = g[f"{tag}_gold_{key}"], g[f"{tag}_ref32_{key}"]
            assert relerr(a, gold) < max(base, 2 * relerr(r32, gold)), (tag, key)
        ok(Y.cpu().numpy(), "Y")
        ok(G.mu.cpu().numpy(), "mu")
        ok(G.w.cpu().numpy(), "w", 5e-5)
        ok(np.array(float(G.sigma)), "sigma")
        ok(np.array(fes), "FE", 5e-5)
    # the reference's own EM_optimization loop (core/GMM.py:330-357) drives the seam
    tag = "2d_full"
    G = ref.GMM.GaussianMixtureUnif(cu(g[f"{tag}_in_mu"]), sigma=float(g[f"{tag}_meta"][6]), spec=spec(), computversion="b200")
    Y, Cfe, FE, i = G.EM_optimization(cu(g[f"{tag}_in_X"]), max_iterations=5, tol=1e-9)
    assert i == 5 and np.isfinite(float(FE))
