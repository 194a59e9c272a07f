"""Metric export with the reference's tensorboard scalar names.

The reference logs Train/epsilon every step (common/learn.py:83), per-episode
Train/returns|safeties|margins|margins_support (common/utils/meters.py:86-95)
and per-evaluation Evaluation/<name> {avg,max} (meters.py:96-106).  The batched
path logs the same names once per rollout chunk: epsilon at the current
agent-step, and the episode metrics reduced over all environments.
`writer` is anything with add_scalar / add_scalars (tensorboardX.SummaryWriter).
"""
from .batched import summarize_totals


def log_train(writer, env, agent, previous_totals=None):
    """Log the chunk that just finished.  Returns the totals to pass as
    `previous_totals` next time so that each call reports only new episodes."""
    tot = env.totals()
    step = env.t
    writer.add_scalar("Train/epsilon", agent.epsilon_at(step), step)
    delta = dict(tot)
    if previous_totals is not None:
        for k in ("episodes", "sum_return", "sum_performance", "sum_margin_pos", "n_margin_pos"):
            delta[k] = tot[k] - previous_totals[k]
    summary = summarize_totals(delta)
    for name in ("returns", "safeties", "margins", "margins_support"):
        if name in summary:
            writer.add_scalar("Train/%s" % name, summary[name]["avg"], step)
    return tot


def log_eval(writer, summary, period):
    """`summary` is what BatchedTabularQ.evaluate returns."""
    for name in ("returns", "safeties", "margins", "margins_support"):
        if name in summary:
            writer.add_scalars("Evaluation/%s" % name,
                               {"avg": summary[name]["avg"], "max": summary[name]["max"]}, period)
