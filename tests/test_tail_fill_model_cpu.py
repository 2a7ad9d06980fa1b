"""Model check of the construction tail fill (csrc/build_kernels.cu) under random schedules -- CPU only.

compute-sanitizer is closed on the GPU pool, so the protocol between `build_select_kernel` and the programmatically dependent
`build_link_kernel` is replayed here as a small state machine and driven by random interleavings:

  * a select warp fetches a node from its batch's work counter, searches (reads rows of OLD nodes and of same-batch nodes that are
    already linked), writes the node's own row, then releases the node's ready flag (epoch of the batch);
  * a link warp may start as soon as every select warp of the batch has started; it fetches a (node, slot) pair, waits for the
    node's flag, reads that slot of the node's row and appends the node to the target's row under the target's lock unless the
    edge is already there (`dup`), pruning when the row is full;
  * the next batch's kernels start when the link kernel of this batch has ended.  Every select warp does one last, empty fetch
    from the work counter before it leaves, and the next batch resets that word: the link warp that draws the first empty index
    of the link queue therefore waits until the select counter shows `count` real fetches + one empty fetch per select warp.

Checked under schedules that hold back exactly those last fetches: no link item ever reads a row before it was written; every
node of every batch is inserted exactly once; no row ever holds an id twice or more than its capacity.  Negative controls: without
the wait a late select warp takes an index of the next batch away (a node is never inserted and its links wait forever) -- with
one counter word at once, with two alternating words as soon as a warp is held back for two batches (the first version of the
kernel alternated words; this model is what retired that).
"""
import random

import pytest

CAP = 6          # row capacity (MaxEdges0 of the model)
SEL = 3          # neighbours a new node selects


class Model:
    def __init__(self, rng, n_old, batches, select_warps, link_warps, alternate_counters=False, wait_for_last_fetches=True):
        self.rng = rng
        self.rows = {v: [] for v in range(n_old)}          # node -> row (list of ids)
        self.written = set(range(n_old))                   # nodes whose own row exists
        self.flag = {}                                      # node -> epoch
        self.counters = [0, 0]
        self.alternate = alternate_counters
        self.wait = wait_for_last_fetches
        self.inserted = []                                  # (epoch, node) in completion order
        self.violations = []
        self.next_id = n_old
        self.batches = batches
        self.nsel, self.nlink = select_warps, link_warps

    # ---- one batch: returns when its link kernel has ended; stragglers of the select kernel are handed to the caller
    def run_batch(self, epoch, count, stragglers):
        nodes = list(range(self.next_id, self.next_id + count))
        self.next_id += count
        for v in nodes:
            self.rows[v] = None                             # allocated, not written
        word = (epoch & 1) if self.alternate else 0
        self.counters[word] = 0                             # the memset in front of the select kernel
        sel = [{"state": "fetch", "word": word, "nodes": nodes, "epoch": epoch} for _ in range(self.nsel)]
        link = [{"state": "idle"} for _ in range(self.nlink)]
        link_counter = 0
        pairs = [(v, s) for v in nodes for s in range(CAP)]
        started = 0
        link_done = 0
        steps = 0
        while link_done < self.nlink or (self.wait and any(w["state"] != "exit" for w in sel)):
            steps += 1
            if steps > 20000:   # (on the GPU: the bounded spin of the link kernel gives up and the Add fails loudly)
                self.violations.append(("link kernel waits for a node that nobody inserts", epoch))
                break
            actors = [("s", w) for w in sel if w["state"] != "exit"] + [("s", w) for w in stragglers if w["state"] != "exit"]
            if started == self.nsel:                        # dependents may launch once every primary warp has started
                actors += [("l", w) for w in link if w["state"] != "exit"]
            if not actors:
                break
            # adversarial bias: a select warp whose batch is completely flagged is about to do its final, empty fetch -- hold most
            # of those back so that they become stragglers of the next batch
            late = [a for a in actors if a[0] == "s" and a[1]["state"] == "fetch" and all(self.flag.get(v) == a[1]["epoch"] for v in a[1]["nodes"])]
            if late and len(late) < len(actors) and self.rng.random() < 0.97:
                actors = [a for a in actors if a not in late]
            kind, w = self.rng.choice(actors)
            if kind == "s":
                if w["state"] == "fetch":
                    if w.get("fresh", True) and w in sel:
                        w["fresh"] = False
                        started += 1
                    wi = self.counters[w["word"]]
                    self.counters[w["word"]] += 1
                    if wi >= len(w["nodes"]):
                        w["state"] = "exit"
                    else:
                        w["node"] = w["nodes"][wi]
                        w["state"] = "search"
                        w["steps"] = self.rng.randint(0, 4)
                elif w["state"] == "search":
                    if w["steps"] > 0:
                        w["steps"] -= 1
                    else:
                        v = w["node"]
                        reachable = [u for u in self.written if u != v and (u not in w["nodes"] or any(u in (r or []) for r in self.rows.values()))]
                        if self.rows[v] is not None:
                            self.violations.append(("node inserted twice", w["epoch"], v))
                        self.rows[v] = self.rng.sample(reachable, min(SEL, len(reachable)))
                        self.written.add(v)
                        w["state"] = "flag"
                elif w["state"] == "flag":
                    self.flag[w["node"]] = w["epoch"]
                    self.inserted.append((w["epoch"], w["node"]))
                    w["state"] = "fetch"
            else:
                if w["state"] == "idle":
                    if link_counter >= len(pairs):
                        w["state"] = "exit"
                        link_done += 1
                    else:
                        w["pair"] = pairs[link_counter]
                        link_counter += 1
                        w["state"] = "wait"
                elif w["state"] == "wait":
                    v, s = w["pair"]
                    if self.flag.get(v) == epoch:
                        if self.rows[v] is None:
                            self.violations.append(("row read before it was written", epoch, v))
                        row = self.rows[v] or []
                        n = row[s] if s < len(row) else -1
                        if n >= 0:
                            tgt = self.rows[n]
                            if tgt is None:
                                self.violations.append(("link target without a row", epoch, n))
                            elif v in tgt:
                                pass                        # dup: the edge is there already
                            elif len(tgt) < CAP:
                                tgt.append(v)
                            else:                           # overflow prune: keep a subset, never more than CAP, never twice
                                keep = self.rng.sample(tgt + [v], self.rng.randint(SEL, CAP))
                                tgt[:] = keep
                        w["state"] = "idle"
        # select warps that have not yet done their final fetch keep running into the next batch
        return [w for w in sel if w["state"] != "exit"], nodes

    def run(self):
        stragglers = []
        expected = []
        for epoch, count in enumerate(self.batches, start=1):
            left, nodes = self.run_batch(epoch, count, stragglers)
            expected += nodes
            stragglers = [w for w in stragglers if w["state"] != "exit"] + left
        guard = 0
        while any(w["state"] != "exit" for w in stragglers) and guard < 10000:   # drain
            guard += 1
            for w in stragglers:
                if w["state"] == "fetch":
                    wi = self.counters[w["word"]]
                    self.counters[w["word"]] += 1
                    w["state"] = "exit" if wi >= len(w["nodes"]) else "lost"
                    if w["state"] == "lost":
                        w["state"] = "exit"
        return expected


def _check(model, expected):
    done = [v for _, v in model.inserted]
    missing = sorted(set(expected) - set(done))
    twice = sorted({v for v in done if done.count(v) > 1})
    for v, row in model.rows.items():
        if row is None:
            continue
        assert len(row) <= CAP, (v, row)
        assert len(set(row)) == len(row), ("duplicate edge", v, row)
        assert v not in row, ("self edge", v)
    return missing, twice


@pytest.mark.parametrize("seed", range(60))
def test_random_schedules_insert_every_node_once_and_never_read_an_unwritten_row(seed):
    rng = random.Random(seed)
    batches = [rng.randint(1, 9) for _ in range(rng.randint(2, 6))]
    m = Model(rng, n_old=8, batches=batches, select_warps=rng.randint(1, 5), link_warps=rng.randint(1, 4))
    expected = m.run()
    missing, twice = _check(m, expected)
    assert not m.violations, m.violations
    assert not missing and not twice, (missing, twice)


@pytest.mark.parametrize("alternate", [False, True])
def test_without_the_wait_a_late_select_warp_steals_from_a_later_batch(alternate):
    bad = 0
    for seed in range(300):
        rng = random.Random(seed)
        batches = [rng.randint(1, 6) for _ in range(6)]
        m = Model(rng, n_old=8, batches=batches, select_warps=4, link_warps=2, alternate_counters=alternate, wait_for_last_fetches=False)
        expected = m.run()
        missing, twice = _check(m, expected)
        bad += bool(missing) or bool(twice) or bool(m.violations)
    assert bad > 0
