"""GetSharedIDs restated per chunk (counts over the live sets, effective level, early stop, Q6) —
the formulation the lookup kernels implement — cross-checked on random families against the oracle's literal
restatement of util/bitset.go:308-411 + util/asm_amd64.s:121-509 (bit planes, register by register)."""
import numpy as np
import pytest

from oracle import pyoracle as po


def shared_ids_spec(sets, min_count, fast):
    """sets: lists of ascending distinct ids, none empty. Returns the ids GetSharedIDs emits."""
    if min_count > 24:
        fast = False
    n = len(sets)
    ends = [s[-1] >> 6 for s in sets]
    starts = [s[0] >> 6 for s in sets]
    lens = [e + 1 for e in ends]
    members = [set(s) for s in sets]
    order = list(range(n))          # column order: set index at column slot t
    live = n
    shortest = min(lens)
    if min_count >= 13:
        level, routine = min(min_count, 16), 16
    elif min_count >= 5:
        level, routine = min(min_count, 8), 8
    else:
        level, routine = max(min_count, 1), 4
    out = []
    for i in range(min(starts), max(ends) + 1):
        if shortest <= i:
            nxt = max(ends)
            t = 0
            while t < live:
                if lens[order[t]] <= i:
                    if live - 1 < min_count:
                        return out                      # the whole search ends (Q11)
                    order[t] = order[live - 1]          # swap with the last live set, re-test this slot
                    live -= 1
                else:
                    nxt = min(nxt, lens[order[t]])
                    t += 1
            shortest = nxt
        for c in range(64 * i, 64 * i + 64):
            has = [c in members[order[t]] for t in range(live)]
            cnt = sum(has)
            if cnt == 0:
                continue
            soft = cnt
            if routine == 16 and live >= 8 and has[7] and not any(has[:7]):
                soft -= 1                               # Q6: the 8th unrolled step forgets its carry
            if soft >= level and (fast or cnt >= min_count):
                out.append(c)
    return out


def random_family(rng, n, universe, density):
    sets = []
    for _ in range(n):
        k = max(1, int(rng.binomial(universe, density)))
        sets.append(sorted(int(x) for x in rng.choice(universe, size=k, replace=False)))
    return sets


@pytest.mark.parametrize("seed", range(8))
def test_count_formulation_matches_bit_planes(seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(25):
        n = int(rng.integers(6, 45))
        universe = int(rng.choice([64, 200, 700, 1500]))
        density = float(rng.choice([0.02, 0.1, 0.3, 0.6]))
        sets = random_family(rng, n, universe, density)
        # sets ending early exercise the drop order (and with it Q6's column 7) and the early stop
        for j in rng.choice(n, size=n // 3, replace=False):
            cut = int(rng.integers(1, universe))
            kept = [x for x in sets[int(j)] if x < cut]
            if kept:
                sets[int(j)] = kept
        osets = []
        for s in sets:
            o = po.IntSet()
            for x in s:
                o.add(x)
            osets.append(o)
        for min_count in sorted(set(int(x) for x in rng.integers(1, min(n, 30) + 1, size=6)) | {(n + 2) // 4}):
            if min_count in (5,) and n == 5:
                continue  # Q7: undefined in the reference
            for fast in (True, False):
                got = [int(x) for x in po.get_shared_ids(osets, min_count, fast)]
                want = shared_ids_spec(sets, min_count, fast)
                assert got == want, (seed, n, universe, density, min_count, fast)
