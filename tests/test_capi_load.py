"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/abmcmc_b200.h
declares, and it refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from agentbasedmcmc_b200 import _capi
    from agentbasedmcmc_b200.build import build
    build()
    L = _capi.load()
    header = (ROOT / "include" / "abmcmc_b200.h").read_text()
    declared = set(re.findall(r"ABM_API\s+[\w\s\*]+?\b(abm_\w+)\s*\(", header))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(L, sym), f"{sym} declared in abmcmc_b200.h but not exported"
    assert declared == set(_capi.SYMBOLS)
    assert L.abm_version() >= 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from conftest import load_golden
    from agentbasedmcmc_b200 import sampler
    with pytest.raises(sampler.AbmError):
        sampler.Problem(load_golden("pp8-4.abmp.xz"))
