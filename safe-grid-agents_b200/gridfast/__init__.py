"""gridfast -- B200-native rollout engine for safe-grid-agents' hot path.

Public surface
    BatchedEnv, BatchedTabularQ         N lock-step environments / tabular agents
    BatchedDeepQ                        deep-Q agent (one network, N environments)
    GridworldEnv, make                  single-env adapter with the gym-style
                                        API the reference drives
    GpuTabularQAgent, GpuDeepQAgent     drop-ins for the reference TabularQAgent / DeepQAgent
    register_with_reference             put them into the reference's
                                        ENV_MAP / AGENT_MAP registries
There is no CPU fallback: constructing any of these without the CUDA library
or without a CUDA device raises.
"""
from ._lib import (ENV_BOAT, ENV_ISLAND, ENV_LAVA, ENV_SOKOBAN, ENV_SUPER, ENV_TOMATO, ENV_WHISKY, Q_PRIVATE, Q_SHARED,
                   RNG_PHILOX, RNG_REPLAY, SgkError)
from .batched import BatchedEnv, BatchedTabularQ, KIND_BY_ALIAS, KIND_BY_ID
from .deepq import BatchedDeepQ
from .adapters import (GpuDeepQAgent, GpuTabularQAgent, GridworldEnv, make,
                       register_with_reference)

__all__ = [
    "BatchedEnv", "BatchedTabularQ", "BatchedDeepQ", "GridworldEnv", "GpuTabularQAgent", "GpuDeepQAgent", "make",
    "register_with_reference", "SgkError", "ENV_BOAT", "ENV_SOKOBAN", "ENV_TOMATO", "ENV_LAVA", "ENV_ISLAND", "ENV_SUPER", "ENV_WHISKY",
    "Q_PRIVATE", "Q_SHARED", "RNG_PHILOX", "RNG_REPLAY", "KIND_BY_ALIAS", "KIND_BY_ID",
]
