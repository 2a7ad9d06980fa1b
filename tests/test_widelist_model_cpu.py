"""
CPU model check of wide.cuh's candidate list (sorted main array + 32 pending keys, one-pass merge by output position)
against the plain sorted list of beam.cuh it replaces -- the algorithm, not the CUDA code (tests/test_gpu_wide.py runs
that against the oracle).  Random streams of candidates with ties, duplicates ("re-evaluated" nodes) and interleaved
expansions must leave both lists in the same state after every step: same members, same far, same next entry to expand.
"""
import bisect
import random

PEND = 32


def key(dist, node):
    return (dist << 32) | (node << 1)          # ordered distance | id | expanded flag (bit 0)


class Plain:
    """beam.cuh: one sorted list of at most ef keys (list_insert / beam_search)."""
    def __init__(self, ef):
        self.ef, self.a = ef, []

    def far(self):
        return self.a[-1] >> 32 if len(self.a) == self.ef else 0xFFFFFFFF

    def offer(self, k):
        if not (len(self.a) < self.ef or (k >> 32) < self.far()):
            return
        pos = bisect.bisect_left(self.a, k)
        if pos < len(self.a) and self.a[pos] >> 1 == k >> 1:
            return                                  # duplicate (an expanded twin is k | 1)
        if pos >= self.ef:
            return
        self.a.insert(pos, k)
        del self.a[self.ef:]

    def select(self):
        for i, v in enumerate(self.a):
            if not v & 1:
                self.a[i] = v | 1
                return v
        return None

    def members(self):
        return list(self.a)


class Wide:
    """wide.cuh: WideList."""
    def __init__(self, ef):
        self.ef, self.main, self.pend, self.cursor, self.merges = ef, [], [], 0, 0

    def far(self):
        if len(self.main) + len(self.pend) < self.ef:
            return 0xFFFFFFFF
        mx = max(self.main[-1] if self.main else 0, self.pend[-1] if self.pend else 0)
        return mx >> 32

    def contains(self, k):                          # wl_contains: 32-ary search, then one bucket
        lo, ln = 0, len(self.main)
        while ln > 32:
            step = (ln + 31) >> 5
            below = sum(1 for lane in range(32) if self.main[min(lo + (lane + 1) * step, lo + ln) - 1] < k)
            nlo = lo + below * step
            ln = min(step, lo + ln - nlo)
            lo = nlo
        return any(v >> 1 == k >> 1 for v in self.main[lo:lo + max(ln, 0)])

    def merge(self):                                # wl_merge: output positions P_j = lower_bound_j + j
        if not self.pend:
            return
        self.merges += 1
        n, pc = len(self.main), len(self.pend)
        P = [bisect.bisect_left(self.main, p) + j for j, p in enumerate(self.pend)]
        out = [None] * (n + pc)
        slots = set(P)
        for c in range((n + pc - 1) >> 5, (P[0] >> 5) - 1, -1):            # top chunk first, as the kernel
            below = sum(1 for x in P if x < (c << 5))
            for lane in range(32):
                o = (c << 5) + lane
                if o >= n + pc or o < P[0] or o in slots:
                    continue
                src = o - below - sum(1 for x in P if (c << 5) <= x < o)
                out[o] = self.main[src]
        for o in range(P[0]):
            out[o] = self.main[o]
        for j, p in enumerate(self.pend):
            out[P[j]] = p
        assert None not in out and out == sorted(out)
        self.main, self.pend = out, []
        self.cursor = min(self.cursor, P[0])

    def offer(self, k):
        if not (len(self.main) + len(self.pend) < self.ef or (k >> 32) < self.far()):
            return
        if any(p >> 1 == k >> 1 for p in self.pend) or self.contains(k):
            return
        bisect.insort(self.pend, k)
        if len(self.main) + len(self.pend) > self.ef:                      # evict the largest of the set
            if not self.main or self.pend[-1] > self.main[-1]:
                self.pend.pop()
            else:
                self.main.pop()
                self.cursor = min(self.cursor, len(self.main))
        if len(self.pend) == PEND:
            self.merge()

    def select(self):
        sel = next((i for i in range(self.cursor, len(self.main)) if not self.main[i] & 1), None)
        self.cursor = sel if sel is not None else len(self.main)
        pl = next((j for j, p in enumerate(self.pend) if not p & 1), None)
        if sel is None and pl is None:
            return None
        mkey = self.main[sel] if sel is not None else (1 << 64)
        pkey = self.pend[pl] if pl is not None else (1 << 64)
        if pkey < mkey:
            self.pend[pl] = pkey | 1
            return pkey
        self.main[sel] = mkey | 1
        self.cursor = sel + 1
        return mkey

    def members(self):
        return sorted(self.main + self.pend)


def run(seed, ef, steps, dist_range, nodes):
    rng = random.Random(seed)
    a, b = Plain(ef), Wide(ef)
    seen = {}
    for _ in range(steps):
        if rng.random() < 0.15:                     # an expansion
            assert a.select() == b.select()
        else:
            node = rng.randrange(nodes)
            d = seen.setdefault(node, rng.randrange(dist_range))   # a re-evaluated node has the same distance
            k = key(d, node)
            a.offer(k)
            b.offer(k)
        assert a.members() == b.members()
        assert a.far() == b.far()
    while True:                                     # drain: same expansion order to the end
        x, y = a.select(), b.select()
        assert x == y
        if x is None:
            break
    b.merge()
    assert b.main == a.a
    return b.merges


def test_wide_list_equals_the_plain_sorted_list():
    merges = 0
    for seed in range(40):
        ef = [1, 2, 31, 32, 33, 64, 100, 500, 1500][seed % 9]
        # few distinct distances -> many exact ties; few nodes -> many duplicates; many of both -> the generic case
        dist_range, nodes = [(8, 50), (1 << 20, 5000), (300, 300), (1 << 30, 100000)][seed % 4]
        merges += run(seed, ef, 1500 if ef < 500 else 4000, dist_range, nodes)
    assert merges > 100                             # the merge path was exercised
