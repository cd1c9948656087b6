"""The reference's own host code building its basis with the unmodified TableauNormMinimiser and with the drop-in
abmcmc::minimiseBasis (include/abmcmc_b200/basis_adaptor.hpp over `abm_basis_*`): the two TableauNormMinimiser objects
must be equal member by member (rows, cols, isBasic, basis, nAuxiliaryVars, F, L, U).  Host code only.
oracle/_ref/basis_check is built by `make -C oracle basis` where /root/reference exists (the build container)."""
import subprocess
from pathlib import Path

import pytest

EXE = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "basis_check"


@pytest.mark.parametrize("args", [("8", "4"), ("16", "8"), ("8", "4", "poisson"), ("16", "4", "poisson")])
def test_reference_host_code_gets_the_same_tableau(args):
    if not EXE.exists():
        pytest.skip("oracle/_ref/basis_check not built (needs /root/reference: make -C oracle basis)")
    res = subprocess.run([str(EXE), *args], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.startswith("OK")
