"""Fused interaction loops with the reference's registry signatures.

The reference looks its per-agent loops up in plain dicts of plain functions
(LEARN_MAP common/learn.py:107-113, EVAL_MAP common/eval.py:59, WARMUP_MAP
common/warmup.py:31-33) and calls them from train.train (train.py:36-40,56,
70,77).  The functions here have exactly those signatures and return values,
but run the work of one call in ONE kernel launch instead of one Python
iteration (and >= 3 host<->device round trips) per environment step:

    tabq_learn_fused(agent, env, env_state, history, args) -> (env_state, history, eval_next)
        one whole episode of the tabq_learn body (learn.py:61-85) + whiler's
        bookkeeping (learn.py:13-24)
    ssq_learn_fused      the same for the SSRL agent -- the loop the reference never
                         wired up (no LEARN_MAP["tabular-ssq"]): the
                         tabq_learn body with TabularSSQAgent.learn, and at every
                         episode end, while budget remains, query_H + learn_C
    default_eval_fused(agent, env, eval_history, args) -> eval_history
        default_eval (eval.py:8-56) without the video
    random_warmup_fused(agent, env, history, args) -> (agent, env, history, args)
        ssrl.random_warmup (ssrl/warmup.py:4-35), in the argument order train.py:56 unpacks
    dqn_warmup_fused / dqn_learn_fused
        dqn_warmup (warmup.py:8-23) / dqn_learn (learn.py:29-58) for GpuDeepQAgent

`register_with_reference(learn_map=..., eval_map=..., warmup_map=...)`
(gridfast.adapters) installs them.  Scalars go to history["writer"] under the
reference's names and step indices: Train/epsilon once per step (learn.py:83),
the episode metrics through track_metrics (meters.py:66-108).
"""
import numpy as np

from .adapters import GpuDeepQAgent, GpuTabularQAgent, GpuTabularSSQAgent
from ._lib import SgkError


def track_metrics(history, env, eval=False, write=True):
    """common/utils/meters.py:66-108, restated so that the fused loops do not
    need the reference importable; the meters in `history` are the caller's
    (the reference's AverageMeter or anything with update/val/avg/max)."""
    _env = env._env if hasattr(env, "_env") else env
    ep = history["period"] if eval else history["episode"]
    history["returns"].update(_env.episode_return)
    safety = _env.get_last_performance()
    margin = None
    if safety is not None:
        history["safeties"].update(safety)
        margin = _env.episode_return - safety
        history["margins"].update(margin)
        if margin > 0:
            history["margins_support"].update(margin)
    if write:
        writer = history["writer"]
        if not eval:
            writer.add_scalar("Train/returns", history["returns"].val, ep)
            if safety is not None:
                writer.add_scalar("Train/safeties", safety, ep)
                writer.add_scalar("Train/margins", margin, ep)
                if margin > 0:
                    writer.add_scalar("Train/margins_support", margin, ep)
        else:
            for kw in ("returns", "safeties", "margins", "margins_support"):
                if safety is None and kw != "returns":
                    continue
                writer.add_scalars("Evaluation/%s" % kw, {"avg": history[kw].avg, "max": history[kw].max}, ep)
    return history


def _need(agent, cls, what):
    if not isinstance(agent, cls):
        raise SgkError("%s needs a gridfast %s (got %s)" % (what, cls.__name__, type(agent).__name__))


def tabq_learn_fused(agent, env, env_state, history, args):
    """LEARN_MAP["tabular-q"] (common/learn.py:8-26,61-85): one episode, one launch."""
    _need(agent, GpuTabularQAgent, "tabq_learn_fused")
    if agent.env is not env:
        raise SgkError("the agent was built for a different environment object")
    t = history["t"]
    k0 = agent._k
    obs, reward, done, info, steps = agent.run_episode(cheat=bool(getattr(args, "cheat", False)))
    writer = history["writer"]
    eps_at = agent.table.epsilon_at
    for j in range(steps):                       # learn.py:82-83, one scalar per step
        writer.add_scalar("Train/epsilon", eps_at(k0 + j + 1), t + j)
    history["t"] = t + steps
    history = track_metrics(history, env)
    eval_next = history["episode"] % args.eval_every == args.eval_every - 1
    return (obs, reward, done, info), history, eval_next


def ssq_learn_fused(agent, env, env_state, history, args):
    """The SSRL learning loop: tabq_learn's body with
    TabularSSQAgent.learn (ssrl/agents.py:34-42) and, when the episode ends
    and budget remains, safety = query_H(env); learn_C(return - safety > 0)
    (ssrl/agents.py:45-82) -- all inside the same launch."""
    _need(agent, GpuTabularSSQAgent, "ssq_learn_fused")
    agent._history = []
    return tabq_learn_fused(agent, env, env_state, history, args)


def default_eval_fused(agent, env, eval_history, args):
    """EVAL_MAP[...] (common/eval.py:8-56) for the tabular agents: greedy
    episodes until the first episode end at or after args.eval_timesteps, one
    launch.  The grid animation (eval.py:16-48, tensorboard video) is not
    produced: rendering is outside the hot path."""
    _need(agent, GpuTabularQAgent, "default_eval_fused")
    print("#### EVAL ####")
    env.reset()
    rows = agent.evaluate_episodes(int(args.eval_timesteps))
    inner = env._env
    outer = inner._outer
    for i, (ret, perf) in enumerate(rows):       # eval.py:22,50: per episode, the last one with write=True
        outer._episode_return, outer._last_performance = float(ret), float(perf)
        eval_history = track_metrics(eval_history, env, eval=True, write=(i == len(rows) - 1))
    eval_history["returns"].reset(reset_history=True)
    eval_history["safeties"].reset()
    eval_history["margins"].reset()
    eval_history["margins_support"].reset()
    eval_history["period"] += 1
    return eval_history


def random_warmup_fused(agent, env, history, args):
    """WARMUP_MAP["tabular-ssq"] (ssrl/warmup.py:4-35): int(budget * warmup)
    random-policy episodes, query_H + learn_C after each, one launch.  Returns
    in the order train.py:56 unpacks (the reference's own return statement has
    `args` and `history` swapped)."""
    _need(agent, GpuTabularSSQAgent, "random_warmup_fused")
    if getattr(args, "seed", None) and env._rng == "numpy":
        np.random.seed(args.seed)                # RandomAgent.__init__ (dummy.py:10-13)
    print("#### WARMUP ####\n")
    n_episodes = int(args.budget * args.warmup)
    if n_episodes > 0:
        limit = env.batched.max_iterations
        per_step = 1 + 26                          # randint: 1 word; the environment: <= 2 x 13
        env.reset()                                # the first `env.reset()` of the warm-up loop (ssrl/warmup.py:12);
        #                                            the later ones happen inside the kernel
        lent = env._lend_words(n_episodes * (limit * per_step + 26) + 64, always=True)
        steps = agent.table.ssrl_warmup(n_episodes, want_steps=True)
        env._settle_words(lent)
        env._after_fused(int(steps[0].item()), done=True)
    for name in ("returns", "safeties", "margins", "margins_support"):
        history[name].reset()                    # ssrl/warmup.py:30-33
    env.batched.clear_stats()
    return agent, env, history, args


def dqn_warmup_fused(agent, env, history, args):
    """WARMUP_MAP["deep-q"] (common/warmup.py:8-23): fill the replay ring with
    args.replay_capacity random-policy transitions, one launch."""
    _need(agent, GpuDeepQAgent, "dqn_warmup_fused")
    print("#### WARMUP ####\n")
    agent.net.warmup(int(args.replay_capacity))
    env._after_fused(int(args.replay_capacity), done=False)
    return agent, env, history, args


def dqn_learn_fused(agent, env, env_state, history, args):
    """LEARN_MAP["deep-q"] (common/learn.py:29-58): one episode of act_explore,
    env.step, replay.add, learn, update_epsilon and the periodic target sync.
    The lock-steps run device-side one launch group per step (captured graph);
    the host only reads the done flag."""
    _need(agent, GpuDeepQAgent, "dqn_learn_fused")
    t = history["t"]
    obs, reward, done, info, steps, losses = agent.run_episode(cheat=bool(getattr(args, "cheat", False)), t=t)
    writer = history["writer"]
    for j in range(steps):
        writer.add_scalar("Train/value_loss", losses[j], t + j)           # value.py:124
        writer.add_scalar("Train/epsilon", agent._epsilon_at(agent._k - steps + j + 1), t + j)   # learn.py:51-52
    history["t"] = t + steps
    history = track_metrics(history, env)
    eval_next = history["episode"] % args.eval_every == args.eval_every - 1
    return (obs, reward, done, info), history, eval_next
