"""The oracle's random streams: Philox known answers and numpy's legacy
word->value mapping (what value.py:38-39 and dummy.py:16 consume)."""
import numpy as np

from oracle import cgrid, rng

KAT = [  # Random123 kat_vectors, philox4x32 10 rounds
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_known_answers_python_and_c():
    for ctr, key, want in KAT:
        assert rng.philox4x32_10(ctr, key) == want
        c = np.array(ctr, np.uint32)
        k = np.array(key, np.uint32)
        out = np.zeros(4, np.uint32)
        cgrid.lib().cg_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        assert tuple(int(x) for x in out) == want


def test_numpy_legacy_word_mapping():
    for seed in (0, 5, 1234):
        words = rng.mt19937_words(seed, 64)
        replay = rng.ReplayWordsRng(words)
        np.random.seed(seed)
        for _ in range(6):
            assert np.random.sample() == replay.agent_uniform()
            assert np.random.choice(4) == replay.agent_choice(4)
            assert np.random.randint(0, 4) == replay.random_action(4)
            assert np.random.random() == replay.env_uniform(0)


def test_epsilon_schedule_matches_list_popping():
    from oracle.tabular import EpsilonSchedule

    for eps, anneal in ((0.01, 7), (0.01, 1), (0.3, 2), (0.01, 100000)):
        future = [1.0 - (1 - eps) * t / anneal for t in range(anneal)]
        future.pop(0)
        cur = 0.0  # value.py:27-28
        sched = EpsilonSchedule(eps, anneal)
        for k in range(12):
            assert sched.current == cur == cgrid.epsilon_at(eps, anneal, k)
            if future:
                cur = future.pop(0)
            sched.advance()
