"""CPU: repository contract -- the C ABI library exports every symbol the header declares (no compute calls), the
product never imports the oracle, and it fails loudly without the CUDA path."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dicp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dicp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "diff_icp_b200", "libdicp_b200.so"))
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dicp_b200.h but not exported"
    assert lib.dicp_version() >= 100


def test_ctypes_table_matches_header():
    from diff_icp_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_product_never_imports_oracle_or_emulation():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "diff_icp_b200")):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                src = open(path).read()
                if re.search(r"^\s*(from|import)\s+(oracle|emu_backend)\b", src, flags=re.M) or "libdicp_hostemu" in src:
                    bad.append(path)
            elif f.endswith((".cu", ".cuh", ".h")):
                src = open(path).read()
                if re.search(r'#include\s+"[^"]*(tests|oracle)/', src):
                    bad.append(path)
    assert not bad, bad


def test_compute_fails_loudly_without_cuda():
    from diff_icp_b200.tools.kernel import GaussKernel
    K = GaussKernel(0.2, 2, spec={"device": "cpu", "dtype": torch.float32})
    with pytest.raises(ValueError):
        K.KRed(torch.rand(4, 2), torch.rand(5, 2), torch.rand(5, 2))
    with pytest.raises(ValueError):
        K.KRed(torch.rand(4, 2, dtype=torch.float64), torch.rand(5, 2, dtype=torch.float64), torch.rand(5, 2, dtype=torch.float64))
    with pytest.raises(ValueError):
        GaussKernel(0.2, 2, computversion="opencl")
