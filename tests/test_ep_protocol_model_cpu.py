"""Model check of the expert-parallel exchange protocol (csrc/ep.cu) on the CPU.

The exchange areas are SINGLE-buffered and nothing is ever reset: safety rests on the ordering argument in the header of
csrc/ep.cu (a rank overwrites its row block / its combine slot in a peer's area only after events that imply the peer
has finished reading the previous call's data).  This test states that protocol as a small transition system — G ranks,
each a stream-ordered sequence of kernels per MoE call

    dispatch_push   store my rows into EVERY peer's area (row block = my rank), then raise flag_disp[me] = epoch there
    wait_sort       wait until the G dispatch flags of MY area show the epoch
    experts         READ the gathered rows of all ranks from MY area
    combine_push    store my partial sums of rank q's rows into q's area (slot = my rank), then raise flag_comb[me]
    reduce_finalize wait for the G combine flags of MY area, READ the G slots, epoch += 1

— and runs it under randomly interleaved schedules (any rank whose next kernel is not blocked may advance; the two
"store ... then raise flag" kernels are split into their stores and their flag raise, and the stores / flag raises
towards different peers land in random order, as NVLink traffic does).  Every store tags its cell with the writer's
call number; every read asserts it sees exactly the CURRENT call's tag: never data of the next call (overwritten too
early) and never data of the previous one (read too early).  A deliberately broken variant (finalize not waiting for the
combine flags) must be caught, which shows the model can see such hazards."""
import random

import pytest


class Rank:
    def __init__(self, r, G):
        self.r, self.G = r, G
        self.epoch = 0                      # device-side epoch counter of this rank's area
        self.rows = [None] * G              # area: row block of source rank s  (tag = call number of the data)
        self.slots = [None] * G             # area: combine slot of source rank s
        self.flag_disp = [0] * G
        self.flag_comb = [0] * G
        self.pc = 0                         # index into the per-call kernel list
        self.call = 1                       # call number of the kernel at pc (== epoch + 1 while the call runs)
        self.pending = []                   # remote effects of the running push kernel, applied one at a time


KERNELS = ("dispatch_push", "wait_sort", "experts", "combine_push", "reduce_finalize")


def step(ranks, r, rng, broken=False):
    """Advances rank r by one micro-step if it can; returns False if r is blocked."""
    me = ranks[r]
    G = me.G
    k = KERNELS[me.pc]
    if me.pending:  # a push kernel in flight: one more of its remote effects lands
        stores = [e for e in me.pending if e[0] in ("row", "slot")]
        pool = stores if stores else me.pending   # fence.sys: every store has landed before any flag is raised
        eff = pool[rng.randrange(len(pool))]
        me.pending.remove(eff)
        kind, dst, tag = eff
        tgt = ranks[dst]
        if kind == "row":
            tgt.rows[r] = tag
        elif kind == "slot":
            tgt.slots[r] = tag
        elif kind == "flag_disp":
            tgt.flag_disp[r] = tag
        elif kind == "flag_comb":
            tgt.flag_comb[r] = tag
        if not me.pending:
            me.pc += 1
        return True
    n = me.epoch + 1
    assert n == me.call
    if k == "dispatch_push":
        me.pending = [("row", d, n) for d in range(G)] + [("flag_disp", d, n) for d in range(G)]
        return step(ranks, r, rng, broken)
    if k == "wait_sort":
        if any(f < n for f in me.flag_disp):
            return False
        me.pc += 1
        return True
    if k == "experts":
        for s in range(G):
            assert me.rows[s] == n, f"rank {r} call {n}: rows of rank {s} carry call {me.rows[s]}"
        me.pc += 1
        return True
    if k == "combine_push":
        me.pending = [("slot", d, n) for d in range(G)] + [("flag_comb", d, n) for d in range(G)]
        return step(ranks, r, rng, broken)
    if k == "reduce_finalize":
        if not broken and any(f < n for f in me.flag_comb):
            return False
        for s in range(G):
            assert me.slots[s] == n, f"rank {r} call {n}: combine slot of rank {s} carries call {me.slots[s]}"
        me.epoch += 1
        me.call += 1
        me.pc = 0
        return True
    raise AssertionError(k)


def run(G, calls, seed, broken=False):
    rng = random.Random(seed)
    ranks = [Rank(r, G) for r in range(G)]
    # bias: some schedules let one rank run far ahead, others keep the ranks close
    weights = [rng.choice((1, 1, 1, 8, 30)) for _ in range(G)]
    while any(rk.epoch < calls for rk in ranks):
        live = [r for r in range(G) if ranks[r].epoch < calls]
        order = sorted(live, key=lambda r: rng.random() / weights[r])
        for r in order:
            if step(ranks, r, rng, broken):
                break
        else:
            raise AssertionError("deadlock: every rank is blocked")
    assert all(rk.epoch == calls for rk in ranks)


@pytest.mark.parametrize("G", [2, 3, 4, 8])
def test_single_buffered_exchange_is_safe_under_random_schedules(G):
    for seed in range(400 if G <= 4 else 100):
        run(G, calls=6, seed=seed)


def test_the_model_sees_a_missing_wait():
    caught = 0
    for seed in range(200):
        try:
            run(3, calls=4, seed=seed, broken=True)
        except AssertionError:
            caught += 1
    assert caught > 150  # finalize without its wait reads stale / missing slots in almost every schedule
