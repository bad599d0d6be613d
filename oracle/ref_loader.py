"""Loader for the UNMODIFIED reference package (TEST / BASELINE INFRASTRUCTURE, never on the product path).

Imports ``diffICP`` from ``/root/reference`` (build container) or from the git-ignored install ``baseline/_ref`` (the copy
that travels to the GPU box; see ``__graft_entry__.build``) exactly as SURVEY.md Appendix C describes:

  * stub modules for matplotlib / mpl_toolkits (not installed; only used for plotting: core/GMM.py:16-18,
    core/PSR.py:9-10, visualization/visu.py:9-14, api/*.py),
  * a pykeops-free module object registered as ``diffICP.tools.point_sets`` (the original hard-imports pykeops at its
    line 8): ``decimate`` and ``point_set_distance`` are the reference's OWN functions compiled from their source text,
    ``intrinsic_scale`` (one KeOps Kmin(2) reduction) is its dense equivalent,
  * pykeops stays absent, so every "keops" request falls back to the reference's own torch twin with a warning
    (tools/kernel.py:93-96, core/GMM.py:130-133),
  * the one-line fix of ``GaussKernel.check_coverage``'s broken torch branch (tools/kernel.py:328) is applied on request.

Used by tests/golden/make_golden.py (fixtures), tests/test_gpu_reference_seam.py (INTEGRATION.md option B executed) and
``bench.py --impl reference`` / its ``cpu_baseline`` leg (the reference itself as the CPU arm).
"""

from __future__ import annotations

import ast
import math
import os
import sys
import types
import warnings

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = ["/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]


def find_root():
    for root in CANDIDATES:
        if os.path.isfile(os.path.join(root, "diffICP", "tools", "kernel.py")):
            return root
    return None


def reference_function(root, relpath, name, namespace):
    """Compile ONE function of a reference module from its source text (for modules that cannot be imported because they
    import pykeops at the top) and return it; nothing of the source is written anywhere."""
    path = os.path.join(root, relpath)
    tree = ast.parse(open(path).read(), filename=path)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = dict(namespace)
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return _Anything("call")


_loaded = {}


def load_reference(root=None, fix_coverage=False):
    """Returns a namespace with the reference's modules: .kernel .LDDMM .GMM .PSR .ICP_atlas .ICP_two_set .spec .root"""
    root = root or find_root()
    if root is None:
        raise FileNotFoundError("reference package not found (looked in %s)" % ", ".join(CANDIDATES))
    if root not in _loaded:
        for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
                     "matplotlib.ticker", "mpl_toolkits", "mpl_toolkits.mplot3d"]:
            sys.modules.setdefault(name, _Anything(name))
        if root not in sys.path:
            sys.path.insert(0, root)
        warnings.filterwarnings("ignore", message=".*keops.*")
        import diffICP.tools  # noqa: F401
        ps = types.ModuleType("diffICP.tools.point_sets")

        def intrinsic_scale(x):          # dense equivalent of the Kmin(2) KeOps reduction (tools/point_sets.py:22-26)
            d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
            return float(d2.topk(2, dim=1, largest=False).values[:, 1].mean().sqrt())

        import diffICP.tools.kernel as rk
        ps.intrinsic_scale = intrinsic_scale
        ps.decimate = reference_function(root, "diffICP/tools/point_sets.py", "decimate", {"np": np, "torch": torch})
        ps.point_set_distance = reference_function(
            root, "diffICP/tools/point_sets.py", "point_set_distance",
            {"math": math, "warnings": warnings, "torch": torch, "intrinsic_scale": intrinsic_scale,
             "GaussKernel": lambda s, D: rk.GaussKernel(s, D, computversion="torch")})
        sys.modules["diffICP.tools.point_sets"] = ps
        import diffICP.api.ICP_atlas as ra
        import diffICP.api.ICP_two_set as rt
        import diffICP.core.GMM as rg
        import diffICP.core.LDDMM as rl
        import diffICP.core.PSR as rp
        import diffICP.tools.spec as rs
        _loaded[root] = types.SimpleNamespace(kernel=rk, LDDMM=rl, GMM=rg, PSR=rp, ICP_atlas=ra, ICP_two_set=rt, spec=rs,
                                              point_sets=ps, root=root)
    ref = _loaded[root]
    if fix_coverage:
        ref.kernel.GaussKernel.check_coverage = \
            lambda self, X, Y, R: ((X[:, None, :] - Y[None, :, :]) ** 2).sum(-1).min(dim=1).values > (R * self.sigma) ** 2
    return ref
