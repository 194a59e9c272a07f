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


def test_environment_draw_layout_high_and_low_words_from_separate_calls():
    """Slot k of a step: high 27 bits from word k%4 of call 1+k//4, low 26 bits
    from word k%4 of call 16+k//4 (8+k//4 / 24+k//4 at a reset); the choice
    word of the whisky wrapper is word 2 of call 1."""
    seed, env_id, step = 0x1234_5678_9ABC_DEF0, 77, 4242
    stream = rng.PhiloxRng(seed, env_id=env_id)
    stream.set_context(env_id, step)
    key = (seed & 0xFFFFFFFF, seed >> 32)

    def call(c):
        return rng.philox4x32_10((env_id, 0, step, c), key)

    for slot in range(13):
        for at_reset, hi, lo in ((False, 1, 16), (True, 8, 24)):
            a = call(hi + slot // 4)[slot % 4]
            b = call(lo + slot // 4)[slot % 4]
            assert stream.env_uniform(slot, at_reset) == rng.words_to_double(a, b)
    assert stream.env_choice(4) == call(1)[2] & 3
    # the decision u < p needs the low word only when the high 27 bits sit on the threshold
    t = 450359962737049                       # u < 0.05  <=>  u53 <= t
    for hi27, lo26, expect in ((t >> 26) - 1, 0x3FFFFFF, True), ((t >> 26) + 1, 0, False), \
                              (t >> 26, t & 0x3FFFFFF, True), (t >> 26, (t & 0x3FFFFFF) + 1, False):
        u = ((hi27 << 26) | lo26) / 2.0 ** 53
        assert (u < 0.05) == expect
