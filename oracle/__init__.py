"""CPU oracle for the rollout hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker or as the CPU
baseline being timed.  The product (``safe-grid-agents_b200/gridfast``) never
imports this package and fails loudly when its CUDA library is missing.

PARITY UNPINNED: the environment half of the hot path
(pycolab -> ai-safety-gridworlds -> safe-grid-gym) is an un-vendored, un-pinned
third-party dependency of the reference (setup.py:46) that is not on disk and
cannot be fetched here, and the reference ships no tests or golden vectors.
The env modules below therefore restate the *published* algorithm of those
packages from the rules written down in SURVEY.md section 8.1 plus the
interface facts the reference itself pins (observation shape/dtype, agent
value 2.0, info keys, ``_env.episode_return`` / ``get_last_performance``).
The agent half (``tabular.py``) restates code that IS in the reference
(safe_grid_agents/common/agents/value.py:15-58, ssrl/agents.py:9-86,
common/learn.py:8-85) and is pinned against the live reference classes by
``tests/golden/make_golden.py`` (run in the build container, fixtures
committed).

Layout
    colab.py                 pycolab engine subset (sprites, drapes, update
                             groups, z-order rendering)
    safety.py                ai_safety_gridworlds.shared.safety_game subset
    boat_race.py, side_effects_sokoban.py, tomato_watering.py
    gridworld_env.py         safe_grid_gym.GridworldEnv restatement + make()
    tabular.py               TabularQAgent / TabularSSQAgent + learn loops
    rng.py                   injectable RNG streams (numpy MT19937 passthrough,
                             Philox4x32-10 counter mode, word replay)
    cgrid.c / cgrid.py       plain-C batched restatement for large parity runs
"""
