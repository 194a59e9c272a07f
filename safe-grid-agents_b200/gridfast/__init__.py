"""gridfast -- B200-native rollout engine for safe-grid-agents' hot path.

Public surface
    BatchedEnv, BatchedTabularQ         N lock-step environments / tabular agents
    BatchedDeepQ                        deep-Q agent (one network, N environments)
    GridworldEnv, make                  single-env adapter with the gym-style
                                        API the reference drives
    GpuTabularQAgent, GpuTabularSSQAgent, GpuDeepQAgent
                                        drop-ins for the reference TabularQAgent /
                                        TabularSSQAgent / DeepQAgent
    tabq_learn_fused, ssq_learn_fused, default_eval_fused, random_warmup_fused,
    dqn_learn_fused, dqn_warmup_fused   the reference's LEARN_MAP / EVAL_MAP /
                                        WARMUP_MAP functions, one launch per call
    register_with_reference             put them into the reference's
                                        ENV_MAP / AGENT_MAP registries
There is no CPU fallback: constructing any of these without the CUDA library
or without a CUDA device raises.
"""
from ._lib import (ENV_BOAT, ENV_ISLAND, ENV_LAVA, ENV_SOKOBAN, ENV_SOKOBAN2, ENV_SUPER, ENV_TOMATO, ENV_WHISKY, Q_PRIVATE, Q_SHARED,
                   RNG_PHILOX, RNG_REPLAY, SgkError)
from .batched import BatchedEnv, BatchedTabularQ, KIND_BY_ALIAS, KIND_BY_ID
from .deepq import BatchedDeepQ
from .adapters import (GpuDeepQAgent, GpuTabularQAgent, GpuTabularSSQAgent, GridworldEnv, make,
                       register_with_reference)
from .loops import (default_eval_fused, dqn_learn_fused, dqn_warmup_fused, random_warmup_fused, ssq_learn_fused,
                    tabq_learn_fused, track_metrics)

__all__ = [
    "BatchedEnv", "BatchedTabularQ", "BatchedDeepQ", "GridworldEnv", "GpuTabularQAgent", "GpuTabularSSQAgent",
    "GpuDeepQAgent", "make", "tabq_learn_fused", "ssq_learn_fused", "default_eval_fused", "random_warmup_fused",
    "dqn_learn_fused", "dqn_warmup_fused", "track_metrics",
    "register_with_reference", "SgkError", "ENV_BOAT", "ENV_SOKOBAN", "ENV_TOMATO", "ENV_LAVA", "ENV_ISLAND", "ENV_SUPER", "ENV_WHISKY",
    "Q_PRIVATE", "Q_SHARED", "RNG_PHILOX", "RNG_REPLAY", "KIND_BY_ALIAS", "KIND_BY_ID",
]
