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
    assert L.abm_version() >= 2   # 2: abm_problem_desc carries the agent_* tables


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from conftest import load_golden
    from agentbasedmcmc_b200 import sampler
    with pytest.raises(sampler.AbmError):
        sampler.Problem(load_golden("pp8-4.abmp.xz"))


def test_ctypes_structs_match_the_header(tmp_path):
    """The Python mirror of abm_problem_desc / abm_problem_info has the C layout (size and the offsets of the fields after which
    the struct grew): compiled from include/abmcmc_b200.h with the host compiler."""
    import subprocess
    from agentbasedmcmc_b200 import _capi
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "abmcmc_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(abm_problem_desc), offsetof(abm_problem_desc, kappa),\n'
                   '    offsetof(abm_problem_desc, sum_tree), offsetof(abm_problem_desc, n_agent_states), offsetof(abm_problem_desc, agent_cq),\n'
                   '    sizeof(abm_problem_info)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    D, I = _capi.ProblemDesc, _capi.ProblemInfo
    assert got == [C.sizeof(D), D.kappa.offset, D.sum_tree.offset, D.n_agent_states.offset, D.agent_cq.offset, C.sizeof(I)]
