"""Randomised model check of the mbarrier protocol of exact_tc5ws_scan_kernel (csrc/exact_tf32_kernels.cu): producers,
the MMA issuer (one producer thread), the tensor core (asynchronous) and the consumers are stepped in random order; the
model asserts that no wait is satisfied by the wrong phase (parity aliasing), that the operand stage, the two
accumulators and the three norm slots are never overwritten while still in use, and that every run terminates."""
import random
import sys


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.done = count, count, 0  # done = completed phases

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending, self.done = self.count, self.done + 1

    def test(self, parity):  # mbarrier.try_wait.parity
        return (self.done & 1) != parity


def run(ntiles, seed):
    rng = random.Random(seed)
    mma_done, tmem_free = [Bar(1), Bar(1)], [Bar(1), Bar(1)]  # consumers modelled as one arriving agent
    stage_tile = -1          # tile whose operands are in the stage
    mma_reading = None       # tile the tensor core is reading the stage for
    acc_tile = [-1, -1]      # tile held by each accumulator
    acc_read_done = [True, True]
    xa_slot = [-1, -1, -1]
    P = {"t": 0, "pc": "wait_stage"}          # producer warps (non-issuer part runs with the issuer up to the named barrier)
    I = {"t": -1, "pc": "idle"}               # issuer: picks up tile t after the producers' barrier
    T = []                                    # queued MMAs: (tile)
    C = {"t": 0, "pc": "wait_mma"}
    consumer_reading = None
    steps = 0
    while P["t"] < ntiles or I["pc"] != "idle" or T or C["t"] < ntiles:
        steps += 1
        assert steps < 100000, "no progress: deadlock"
        who = rng.choice("PITC")
        if who == "P" and P["t"] < ntiles:
            t = P["t"]
            if P["pc"] == "wait_stage":
                if t == 0 or mma_done[(t - 1) & 1].test(((t - 1) >> 1) & 1):
                    if t > 0:
                        assert mma_done[(t - 1) & 1].done == ((t - 1) >> 1) + 1, "producer wait satisfied by the wrong phase"
                    assert mma_reading is None, "stage overwritten while the tensor core reads it"
                    stage_tile = t
                    assert consumer_reading is None or consumer_reading % 3 != t % 3, "norm slot overwritten while read"
                    xa_slot[t % 3] = t
                    P["pc"] = "barrier"
            elif P["pc"] == "barrier":
                if I["pc"] == "idle":  # the issuer thread is one of the producers: it must have left its previous issue
                    I.update(t=t, pc="wait_tmem")
                    P.update(t=t + 1, pc="wait_stage")
        elif who == "I" and I["pc"] != "idle":
            t = I["t"]
            if I["pc"] == "wait_tmem":
                if t < 2 or tmem_free[t & 1].test(((t - 2) >> 1) & 1):
                    if t >= 2:
                        assert tmem_free[t & 1].done == ((t - 2) >> 1) + 1, "issuer wait satisfied by the wrong phase"
                    assert acc_read_done[t & 1], "accumulator overwritten before it was read"
                    assert stage_tile == t
                    T.append(t)
                    mma_reading = t
                    acc_tile[t & 1] = t
                    acc_read_done[t & 1] = False
                    I["pc"] = "idle"
        elif who == "T" and T:
            t = T.pop(0)           # MMAs complete in order; tcgen05.commit arrives on mma_done
            mma_reading = None
            mma_done[t & 1].arrive()
        elif who == "C" and C["t"] < ntiles:
            t = C["t"]
            if C["pc"] == "wait_mma":
                if mma_done[t & 1].test((t >> 1) & 1):
                    assert mma_done[t & 1].done == (t >> 1) + 1, "consumer wait satisfied by the wrong phase"
                    assert acc_tile[t & 1] == t and xa_slot[t % 3] == t
                    consumer_reading = t
                    C["pc"] = "epilogue"
            else:
                assert acc_tile[t & 1] == t and xa_slot[t % 3] == t, "accumulator / norm slot changed under the epilogue"
                consumer_reading = None
                acc_read_done[t & 1] = True
                tmem_free[t & 1].arrive()
                C.update(t=t + 1, pc="wait_mma")
    return steps


if __name__ == "__main__":
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    total = 0
    for seed in range(runs):
        total += run(random.Random(seed).choice([0, 1, 2, 3, 4, 5, 9, 17]), seed)
    print("ok:", runs, "random schedules,", total, "steps")
