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
# the last case draws the exact-mode feasibility search's uniforms in blocks of 4096 (27 618 iterations at 16x16x8: seven device
# calls), which pins the generator bookkeeping across many blocks
@pytest.mark.parametrize("grid,T,n,block", [(8, 4, 300, 0), (16, 8, 60, 0), (3, 2, 5000, 0), (16, 8, 10, 4096)])
def test_cpp_dropin_matches_reference_sample_for_sample(grid, T, n, block):
    if not EXE.exists():
        pytest.skip("oracle/_ref/dropin_check not built (needs /root/reference at build time)")
    out = subprocess.run([str(EXE), str(grid), str(T), str(n)] + ([str(block)] if block else []), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().startswith("OK:")


@pytest.mark.parametrize("which,n", [("binomial", 3000), ("catmouse", 3000), ("catmouse-assim", 3000), ("predprey", 2000)])
def test_cpp_reference_experiments_sample_for_sample(which, n):
    """The reference's four validation experiments for this path (Experiments.cpp:21-125) with the GPU sampler constructed as
    they construct the reference one -- `sampler(posterior, kappa)` from Prior * Likelihood of BinomialAgent<3> (noisy
    observation), CatMouseAgent (single observation; assimilation of a random trajectory) and PredPreyAgent<3>: every sample equal
    to the unmodified sampler's on the same Random::gen stream, and, for the enumerable ones, 512 batched device chains within
    4e-3 of the ExactSolver's end-state means (what Experiments.cpp:113-124 compares)."""
    if not EXE.exists():
        pytest.skip("oracle/_ref/dropin_check not built (needs /root/reference at build time)")
    out = subprocess.run([str(EXE), "experiment", which, str(n)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().startswith("OK:")


def _run_moment_ranks(n_ranks, tmp_path):
    import numpy as np
    procs = [subprocess.Popen([str(EXE), "moments", str(r), str(n_ranks), str(tmp_path / "nccl.id"), str(tmp_path / "mom")],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(n_ranks)]
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and o.strip().splitlines()[-1].startswith("OK:"), o
    rows = [[np.array(line.split(), dtype=np.float64) for line in (tmp_path / f"mom.{r}").read_text().splitlines()] for r in range(n_ranks)]
    exact_sum, real_sum = sum(r[0] for r in rows), sum(r[1] for r in rows)
    for r in rows:   # every rank holds the sum of all ranks' local vectors
        assert np.array_equal(r[2], exact_sum)
        assert np.allclose(r[3], real_sum, rtol=1e-14, atol=0.0)
    return rows


def test_cpp_allreduce_moments_one_rank(tmp_path):
    """abm_comm_* / abm_chains_allreduce_moments from a C++ host through NCCL with a single rank (runs on any GPU box)."""
    if not EXE.exists():
        pytest.skip("oracle/_ref/dropin_check not built (needs /root/reference at build time)")
    _run_moment_ranks(1, tmp_path)


def test_cpp_allreduce_moments_two_ranks(tmp_path):
    """The same over two GPUs, one process each: all == local0 + local1 on both ranks."""
    import ctypes
    if not EXE.exists():
        pytest.skip("oracle/_ref/dropin_check not built (needs /root/reference at build time)")
    from agentbasedmcmc_b200 import _capi
    if _capi.load().abm_device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    rows = _run_moment_ranks(2, tmp_path)
    assert rows[0][0][0] == 48 and rows[1][0][0] == 64 and rows[0][2][0] == 112
