"""agentbasedmcmc_b200: B200-native (sm_100a CUDA) drop-in for the SparseBasisSampler hot path of
danftang/AgentBasedMCMC.  See DESIGN.md / INTEGRATION.md."""
from .abmp import read_abmp, write_abmp  # noqa: F401

__all__ = ["read_abmp", "write_abmp"]
