"""Batched rollout collection for policy-gradient agents.

The reference's PPO agents collect experience with gather_rollout
(common/agents/policy_base.py:133-177): act with the old policy, env.step,
optionally learn from the hidden reward (--cheat), store state / action /
reward, and compute get_discounted_returns per episode.  The PPO agents
themselves are out of scope; this module is the
environment half of that loop on the batched engine: T lock-steps of N
environments with any torch policy, dense [T, N] outputs, finished episodes
reset in place, returns computed on the GPU with the reference's formula.
"""
import ctypes

import torch

from ._lib import check
from .batched import _p, _stream


def collect(env, policy, n_steps, discount=0.99, cheat=False):
    """policy(boards_u8 [N, HW] cuda) -> actions uint8 [N] cuda.

    Returns a dict of cuda tensors: states u8 [T, N, HW], actions u8 [T, N],
    rewards f64 [T, N] (hidden reward when `cheat`, None counted as 0),
    dones u8 [T, N], returns f32 [T, N]."""
    n, dev = env.n, env.device
    states = torch.empty(n_steps, n, env.hw, dtype=torch.uint8, device=dev)
    actions = torch.empty(n_steps, n, dtype=torch.uint8, device=dev)
    rewards = torch.empty(n_steps, n, dtype=torch.float64, device=dev)
    dones = torch.empty(n_steps, n, dtype=torch.uint8, device=dev)
    hidden = torch.empty(n, dtype=torch.float64, device=dev)
    nxt = torch.empty(n, env.hw, dtype=torch.uint8, device=dev)
    frame0 = ((env.core() >> 16) & 0xFF).to(torch.int32)
    boards = env.render()
    for t in range(n_steps):
        states[t].copy_(boards)
        with torch.no_grad():
            actions[t].copy_(policy(boards).to(torch.uint8).reshape(n))
        env.step(actions[t], out=(nxt, rewards[t], hidden, dones[t]))
        if cheat:
            rewards[t].copy_(torch.nan_to_num(hidden, nan=0.0))
        env.reset(mask=dones[t], want_boards=False)
        boards = env.render()
    returns = torch.empty(n_steps, n, dtype=torch.float32, device=dev)
    check(env.L.sgk_discounted_returns(env.h, _p(rewards), _p(dones), _p(frame0), n_steps, float(discount),
                                       _p(returns), _stream()))
    return {"states": states, "actions": actions, "rewards": rewards, "dones": dones, "returns": returns}
