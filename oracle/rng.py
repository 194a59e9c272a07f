"""Injectable random streams for the oracle (test infrastructure only).

The reference draws from numpy's *global* legacy MT19937 stream:
``np.random.sample()`` then, only when exploring, ``np.random.choice(A)``
(safe_grid_agents/common/agents/value.py:37-42), ``np.random.randint(0, A)``
for the random agent (common/agents/dummy.py:15-16), and -- inside the
third-party tomato environment -- one ``np.random.random()`` per currently
watered tomato per frame.

Three interchangeable streams implement the same four draws:

``NumpyGlobalRng``  passthrough to the global numpy stream.  Reference
                    faithful; used when the live reference agent is driven.
``ReplayWordsRng``  consumes a pre-generated array of raw 32-bit words with
                    numpy's legacy word->value mapping (verified in
                    tests/test_oracle_rng.py against numpy itself):
                    sample() = ((w0>>5)*2**26 + (w1>>6)) / 2**53,
                    choice(4) = randint(0,4) = w & 3.
``PhiloxRng``       Philox4x32-10 (Salmon et al., SC'11, "Random123") in
                    counter mode.  key = (seed_lo, seed_hi); counter =
                    (env_lo, env_hi, index_lo, call | index_hi<<8).  One call
                    gives four words; the draw is addressed by *purpose*, not
                    by position in a sequence, so the value an environment
                    sees at (seed, env, step, purpose) never depends on how
                    many other draws happened:
                      call 0, index = step >> 1: one call serves two agent
                                      steps; words (a, b) = (w0, w1) for even
                                      steps, (w2, w3) for odd steps; agent
                                      uniform from (a, b), explore action
                                      a & 3, RandomAgent action b & 3 (low bits
                                      the 53-bit uniform does not use)
                      environment draw "slot k" of a step (tomato k's drying
                      draw; slot 0 = the whisky wrapper's uniform): the 53-bit
                      uniform is built from word k%4 of call 1+k//4 (high 27
                      bits) and word k%4 of call 16+k//4 (low 26 bits), both at
                      index = step.  A comparison u < p is decided by the first
                      word alone except with probability 2**-27, so a GPU
                      thread computes ceil(13/4) = 4 calls per step, not 7, and
                      fetches the second call only on that boundary value --
                      with bit-identical results.
                      the same draws at the reset that precedes agent step
                      `step` (tomato reset frame, absent-supervisor coin) use
                      calls 8+k//4 and 24+k//4.
                      whisky's replacement action, ``env_choice``: word 2 of
                      call 1 (& 3).
"""
import numpy as np

_M0 = 0xD2511F53
_M1 = 0xCD9E8D57
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = 0xFFFFFFFF

CALL_AGENT = 0
CALL_ENV_STEP = 1          # first words of the step draws; second words at +15
CALL_ENV_RESET = 8         # ... of the reset draws; second words at +16
CALL_ENV_STEP_LOW = 16
CALL_ENV_RESET_LOW = 24


def philox4x32_10(counter, key):
    """Philox4x32 with 10 rounds.  counter: 4 u32, key: 2 u32 -> 4 u32."""
    c0, c1, c2, c3 = (int(c) & _MASK for c in counter)
    k0, k1 = (int(k) & _MASK for k in key)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = (
            ((p1 >> 32) ^ c1 ^ k0) & _MASK,
            p1 & _MASK,
            ((p0 >> 32) ^ c3 ^ k1) & _MASK,
            p0 & _MASK,
        )
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def words_to_double(a, b):
    """numpy legacy `random_sample`: 53-bit double from two 32-bit words."""
    return ((int(a) >> 5) * 67108864 + (int(b) >> 6)) / 9007199254740992.0


def _pow2_mask(n):
    if n & (n - 1):
        raise ValueError("counter/replay streams support power-of-two ranges only")
    return n - 1


class NumpyGlobalRng:
    """Reference-faithful: every draw comes from the global numpy stream."""

    def set_context(self, env_id, step):
        pass

    def agent_uniform(self):
        return np.random.sample()

    def agent_choice(self, n):
        return np.random.choice(n)

    def random_action(self, n):
        return np.random.randint(0, n)

    def env_uniform(self, slot, at_reset=False):
        return np.random.random()

    def env_choice(self, n):
        return np.random.choice(n)


class ReplayWordsRng:
    """Sequential consumption of raw MT19937 words, numpy legacy mapping."""

    def __init__(self, words):
        self.words = np.asarray(words, dtype=np.uint32)
        self.cursor = 0

    def set_context(self, env_id, step):
        pass

    def _next(self):
        w = int(self.words[self.cursor])
        self.cursor += 1
        return w

    def agent_uniform(self):
        a = self._next()
        b = self._next()
        return words_to_double(a, b)

    def agent_choice(self, n):
        return self._next() & _pow2_mask(n)

    def random_action(self, n):
        return self._next() & _pow2_mask(n)

    def env_uniform(self, slot, at_reset=False):
        a = self._next()
        b = self._next()
        return words_to_double(a, b)

    def env_choice(self, n):
        return self._next() & _pow2_mask(n)


class PhiloxRng:
    """Counter-mode Philox4x32-10; see the module docstring for the layout."""

    def __init__(self, seed, env_id=0):
        self.key = (seed & _MASK, (seed >> 32) & _MASK)
        self.env_id = env_id
        self.step = 0
        self._cache = {}

    def set_context(self, env_id, step):
        if env_id != self.env_id or step != self.step:
            self._cache = {}
        self.env_id = env_id
        self.step = step

    def _call(self, call, index):
        if (call, index) not in self._cache:
            ctr = (
                self.env_id & _MASK,
                (self.env_id >> 32) & _MASK,
                index & _MASK,
                (call & 0xFF) | (((index >> 32) & 0xFFFFFF) << 8),
            )
            self._cache[(call, index)] = philox4x32_10(ctr, self.key)
        return self._cache[(call, index)]

    def _agent_words(self):
        w = self._call(CALL_AGENT, self.step >> 1)
        h = 2 * (self.step & 1)
        return w[h], w[h + 1]

    def agent_uniform(self):
        a, b = self._agent_words()
        return words_to_double(a, b)

    def agent_choice(self, n):
        return self._agent_words()[0] & _pow2_mask(n)

    def random_action(self, n):
        return self._agent_words()[1] & _pow2_mask(n)

    def env_uniform(self, slot, at_reset=False):
        high = CALL_ENV_RESET if at_reset else CALL_ENV_STEP
        low = CALL_ENV_RESET_LOW if at_reset else CALL_ENV_STEP_LOW
        a = self._call(high + slot // 4, self.step)[slot % 4]
        b = self._call(low + slot // 4, self.step)[slot % 4]
        return words_to_double(a, b)

    def env_choice(self, n):
        return self._call(CALL_ENV_STEP, self.step)[2] & _pow2_mask(n)


def mt19937_words(seed, n):
    """First `n` raw 32-bit outputs of numpy's legacy ``np.random.seed(seed)``."""
    rs = np.random.RandomState(seed)
    st = rs.get_state()
    bg = np.random.MT19937()
    bg.state = {"bit_generator": "MT19937", "state": {"key": st[1], "pos": st[2]}}
    return bg.random_raw(n).astype(np.uint32)
