"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/smm_b200.h declares (and the ctypes table lists exactly those), and fails loudly -- no CPU
fallback -- when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

import socialmediamodel_b200 as smm
from socialmediamodel_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "smm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/smm_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototype table and header disagree"
    assert lib.smm_version() == 100


def test_struct_layout_matches_header():
    # smm_params: 3 int32 (+4 pad) + 10 double + 2*2 double + 1 + 2 + 1 + 2 doubles
    assert ctypes.sizeof(_lib.smm_params) == 16 + 8 * (10 + 4 + 6)
    assert ctypes.sizeof(_lib.smm_grid) == 8 + 4 * 8
    assert ctypes.sizeof(_lib.smm_noise) == 8 + 8 + 5 * 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    assert lib.smm_device_count() == 0
    with pytest.raises(smm.SmmError, match="no CUDA device|no CPU fallback"):
        smm.Context(0)
    with pytest.raises(smm.SmmError):
        smm.ABMsolve(5)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "socialmediamodel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"


def test_setting_constructors_match_reference_defaults():
    q = smm.parameters()
    assert (q.n, q.L, q.M, q.eta, q.a, q.b, q.c, q.sigma, q.frictionI, q.frictionM) == (250, 4, 2, 15.0, 1.0, 4.0, 2.0, 0.5, 10.0, 100.0)
    assert smm.parametersstronginf().eta == 1.0
    p = smm.ABMconstruct()
    assert p.X.shape == (89, 89) and abs(p.X[0, 0] + 2.2) < 1e-15 and abs(p.Y[0, -1] - 2.2) < 1e-12
    assert smm.PDEconstruct().N_x == 89 and smm.PDEconstructcoarse().N_x == 47 and smm.PDEconstructmeso().N_x == 61
    assert list(smm.savetimes_abm(200, 4)) == [0, 66, 133, 199]
    pc = smm.PDEconstructcoarse()
    W = pc.W_matrix
    assert W.shape == (47 * 47, 47 * 47) and np.all(np.diag(W) == 1)
    with pytest.raises(MemoryError):
        smm.PDEconstruct().K_matrix


def test_header_is_valid_c99_and_cxx():
    """include/smm_b200.h must be consumable by a C compiler (the boundary is a C ABI) and by C++."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "smm_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-Wextra", "-pedantic", "-x", "c", hdr])
    subprocess.check_call(["g++", "-std=c++11", "-fsyntax-only", "-x", "c++", hdr])
