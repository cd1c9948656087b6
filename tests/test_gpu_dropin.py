"""Drop-in boundary, end to end in C++: the reference's own host code (PredPreyProblem -> TableauNormMinimiser, factor
closures, TrajectoryImportance) constructs `abmcmc::GpuSparseBasisSampler` through the reference constructor signature
(include/abmcmc_b200/gpu_sampler.hpp) next to the unmodified reference `SparseBasisSampler` on the same `Random::gen`
stream; every sample returned by operator() must be identical (oracle/ref/dropin_check.cpp, built by
`make -C oracle dropin` in the build container; the binary travels to the GPU box)."""
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
EXE = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "dropin_check"


# grid 3: the reference's validation experiment Experiments::PredPreySingleObservation (Experiments.cpp:49-65) with the GPU
# sampler constructed as `sampler(posterior, kappa)`; 8 and 16: the paper's problem generator
@pytest.mark.parametrize("grid,T,n", [(8, 4, 300), (16, 8, 60), (3, 2, 5000)])
def test_cpp_dropin_matches_reference_sample_for_sample(grid, T, n):
    if not EXE.exists():
        pytest.skip("oracle/_ref/dropin_check not built (needs /root/reference at build time)")
    out = subprocess.run([str(EXE), str(grid), str(T), str(n)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().startswith("OK:")
