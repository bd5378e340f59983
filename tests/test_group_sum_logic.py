"""CPU restatement of AtomicSink::group_sum (csrc/sinks.cuh): lanes of a warp that add to the same scalar are summed
by pointer jumping along each group's lanes; after ceil(log2(largest group)) rounds the first lane of every group
holds the group's sum.  Integer values, so the check is exact whatever the association."""
import random


def group_sum(keys, vals, active):
    lanes = [l for l in range(32) if active >> l & 1]
    group = {l: sum(1 << m for m in lanes if keys[m] == keys[l]) for l in lanes}
    rounds = max(bin(group[l]).count("1") for l in lanes)
    nxt, s = {}, {}
    for l in lanes:
        above = group[l] & ~((2 << l) - 1) & 0xFFFFFFFF
        nxt[l] = (above & -above).bit_length() - 1 if above else -1
        s[l] = vals[l]
    steps = 0 if rounds < 2 else (rounds - 1).bit_length()          # 32 - clz(rounds - 1)
    for _ in range(steps):
        src = {l: (nxt[l] if nxt[l] >= 0 else l) for l in lanes}
        other = {l: s[src[l]] for l in lanes}                        # the shuffles read before anything is updated
        after = {l: nxt[src[l]] for l in lanes}
        for l in lanes:
            if nxt[l] >= 0:
                s[l] += other[l]
                nxt[l] = after[l]
    firsts = {l for l in lanes if (group[l] & -group[l]).bit_length() - 1 == l}
    return {keys[l]: s[l] for l in firsts}


def test_pointer_jumping_sums_every_group():
    rng = random.Random(3)
    for trial in range(2000):
        n_keys = rng.choice((1, 2, 3, 5, 9, 32))
        keys = [rng.randrange(n_keys) for _ in range(32)]
        vals = [rng.randrange(-10**6, 10**6) for _ in range(32)]
        active = rng.getrandbits(32) | 1 << rng.randrange(32)
        if trial % 7 == 0:
            active = 0xFFFFFFFF
        got = group_sum(keys, vals, active)
        want = {}
        for l in range(32):
            if active >> l & 1:
                want[keys[l]] = want.get(keys[l], 0) + vals[l]
        assert got == want
