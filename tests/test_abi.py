"""CPU-side checks: the C-ABI library loads and exports every symbol that
include/rufus_gpu.h declares; without a GPU rg_init fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from rufus_b200 import api


def _declared():
    txt = open(os.path.join(ROOT, "include", "rufus_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(api.LIB_PATH), "run `make lib` first"
    L = ctypes.CDLL(api.LIB_PATH)
    names = _declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_binding_covers_the_header():
    L = api.lib()
    for n in _declared():
        assert getattr(L, n).argtypes is not None, n


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.RufusGpuError, match="no CUDA device|no CPU fallback"):
        api.Context(0)


def test_host_helpers_match_reference_packing():
    L = api.lib()
    # Util.cpp:51-84
    assert L.rg_rufus_pack(b"ACGT", 4) == (0 | (2 << 2) | (1 << 4) | (3 << 6))
    assert L.rg_rufus_pack(b"xN", 2) == 0
    out = ctypes.create_string_buffer(16)
    n = L.rg_rufus_revcomp(b"AACGTNx", 7, out)
    assert out.raw[:n] == b"NACGTT"
    assert L.rg_histo_rows(1, 10000) == 10002     # histo_main.cc:64-68 with the defaults RUFUS uses
    assert L.rg_histo_rows(3, 10) == 10
    assert [L.rg_key_partition(k, 8) for k in range(100)] == [L.rg_key_partition(k, 8) for k in range(100)]
    assert all(0 <= L.rg_key_partition(k * 7919, 8) < 8 for k in range(1000))
