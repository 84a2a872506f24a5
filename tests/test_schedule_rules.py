"""Host restatement of two index rules of csrc/b2g_fused.cuh, checked on the CPU (no GPU needed):

* the in-place bitonic network of order_bucket_by_key ("flip" form: every compare-exchange puts the smaller key at the
  lower index, so the virtual +inf padding up to the next power of two never moves and pairs that reach into it are
  skipped) sorts any n, not only powers of two;
* the level schedule of a large serial bucket: walking the bucket in key order, a constraint takes the lowest of 96
  levels free on both of its bodies, else a chain level above them (one more than the highest chain level either body
  has reached).  Constraints of one level must share no body — they are solved side by side by one thread block."""
import random

import pytest

LEVELS_MAX = 96


def bitonic_flip_sort(a):
    n, m = len(a), 1
    while m < n:
        m <<= 1

    def exchange(i, l):
        if l < n and a[i] > a[l]:
            a[i], a[l] = a[l], a[i]
    k = 2
    while k <= m:
        hk = k >> 1
        for t in range(m >> 1):
            i = (t // hk) * k + (t % hk)
            exchange(i, i ^ (k - 1))
        j = k >> 2
        while j > 0:
            for t in range(m >> 1):
                i = (t // j) * 2 * j + (t % j)
                exchange(i, i + j)
            j >>= 1
        k <<= 1
    return a


@pytest.mark.parametrize("n", list(range(1, 40)) + [96, 97, 100, 127, 128, 129, 1000, 1023, 1025, 4097])
def test_bitonic_network_sorts_any_length(n):
    rng = random.Random(n)
    a = rng.sample(range(10 * n), n)
    assert bitonic_flip_sort(a[:]) == sorted(a)


def level_schedule(pairs, nbodies):
    """pairs: (bodyA, bodyB) per constraint in key order, -1 = a body outside the tile (static: not a conflict)"""
    used = [0] * nbodies        # 96-bit sets (T.pen, T.sleepMin, T.done)
    chain = [0] * nbodies       # chain counters (T.head)
    levels, top, chain_top, beyond = [], -1, -1, 0
    full = (1 << LEVELS_MAX) - 1
    for a, b in pairs:
        u = (used[a] if a >= 0 else 0) | (used[b] if b >= 0 else 0)
        free = ~u & full
        if free:
            L = (free & -free).bit_length() - 1
            for x in (a, b):
                if x >= 0:
                    used[x] |= 1 << L
            top = max(top, L)
        else:
            c = max(chain[a] if a >= 0 else 0, chain[b] if b >= 0 else 0)
            L = LEVELS_MAX + c
            for x in (a, b):
                if x >= 0:
                    chain[x] = c + 1
            chain_top, beyond = max(chain_top, c), beyond + 1
        levels.append(L)
    use_chain = beyond > 0 and (chain_top + 1) * 4 <= beyond
    n_levels = LEVELS_MAX + chain_top + 1 if use_chain else top + 1
    tail = 0 if use_chain else beyond
    engaged = top >= 0 and n_levels * 4 + tail <= len(pairs) * 3 // 4
    return levels, n_levels, tail, use_chain, engaged


@pytest.mark.parametrize("nbodies,n,seed", [(40, 600, 1), (400, 6500, 2), (100, 5000, 3), (12, 2000, 4), (750, 3000, 5)])
def test_constraints_of_one_level_share_no_body(nbodies, n, seed):
    rng = random.Random(seed)
    pairs = []
    for _ in range(n):
        a = rng.randrange(nbodies)
        b = rng.randrange(-1, nbodies)   # some constraints touch a static body
        while b == a:
            b = rng.randrange(-1, nbodies)
        pairs.append((a, b))
    levels, n_levels, tail, use_chain, engaged = level_schedule(pairs, nbodies)
    seen = set()
    for (a, b), L in zip(pairs, levels):
        if L >= LEVELS_MAX and not use_chain:
            continue  # the tail: one thread, key order
        for x in (a, b):
            if x >= 0:
                assert (x, L) not in seen
                seen.add((x, L))
    # chain levels respect the key order on every body (a later constraint of a body sits on a higher chain level)
    last = {}
    for (a, b), L in zip(pairs, levels):
        if L >= LEVELS_MAX:
            for x in (a, b):
                if x >= 0:
                    assert last.get(x, -1) < L
                    last[x] = L
    assert max(levels) < n_levels or not use_chain
    assert tail == sum(1 for L in levels if L >= LEVELS_MAX) or use_chain


def test_a_hub_chain_is_not_scheduled():
    """one body in every constraint (the tumbler's container): every constraint needs its own level, nothing to gain"""
    pairs = [(0, 1 + k) for k in range(200)]
    levels, n_levels, tail, use_chain, engaged = level_schedule(pairs, 201)
    assert sorted(levels[:96]) == list(range(96)) and not engaged
