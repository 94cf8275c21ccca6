"""Exhaustive interleaving check of the peer-memory all-reduce protocol of finalize_kernel (csrc/finalize.cu exchange_sum):
every rank, for step s = 1, 2, ...: (1) writes its data into slot [s & 1][me] of EVERY rank's inbox, (2) publishes s in
every rank's flag word [s & 1][me], (3) waits until all flags [s & 1][*] in its OWN buffer show s, (4) reads the slots
[s & 1][*] of its own inbox.  Claim: two slots by sequence parity suffice, i.e. no rank can overwrite a slot a peer
has not read yet, although there is no second barrier.  The model explores ALL interleavings of these atomic actions for 2
and 3 ranks over several steps (sequentially consistent memory; the kernel's fences/release-acquire provide the ordering of
data before flag) and asserts that every read sees exactly the data of the step being summed."""
import itertools
from collections import deque


def explore(nranks, nsteps, nslots):
    # per-rank program counter: (step, phase, sub) with phases 0 push, 1 flag, 2 wait, 3 read; memory: inbox[r][slot][src] = step tag
    def program(me):
        acts = []
        for s in range(1, nsteps + 1):
            acts += [("push", s, r) for r in range(nranks)]
            acts += [("flag", s, r) for r in range(nranks)]
            acts += [("wait", s, None)]
            acts += [("read", s, r) for r in range(nranks)]
        return acts

    progs = [program(r) for r in range(nranks)]
    zero = tuple(tuple(tuple(0 for _ in range(nranks)) for _ in range(nslots)) for _ in range(nranks))
    start = (tuple(0 for _ in range(nranks)), zero, zero)          # pcs, inbox data tags, flags
    seen, todo, states = {start}, deque([start]), 0
    while todo:
        pcs, data, flags = todo.popleft()
        states += 1
        for me in range(nranks):
            if pcs[me] >= len(progs[me]):
                continue
            kind, s, r = progs[me][pcs[me]]
            slot = s % nslots
            nd, nf = data, flags
            if kind == "push":
                nd = tuple(tuple(tuple(s if (rr == r and sl == slot and src == me) else data[rr][sl][src] for src in range(nranks))
                                 for sl in range(nslots)) for rr in range(nranks))
            elif kind == "flag":
                nf = tuple(tuple(tuple(s if (rr == r and sl == slot and src == me) else flags[rr][sl][src] for src in range(nranks))
                                 for sl in range(nslots)) for rr in range(nranks))
            elif kind == "wait":
                if any(flags[me][slot][src] < s for src in range(nranks)):
                    continue                                                 # blocked: not an enabled action
            elif kind == "read":
                if data[me][slot][r] != s:
                    return False, states                                     # a peer's later step overwrote (or never wrote) the slot
            npcs = tuple(pcs[i] + (1 if i == me else 0) for i in range(nranks))
            st = (npcs, nd, nf)
            if st not in seen:
                seen.add(st)
                todo.append(st)
    return True, states


def test_two_parity_slots_are_enough():
    ok, n = explore(nranks=2, nsteps=4, nslots=2)
    assert ok and n > 200          # all reachable states visited (253 for this configuration)
    ok, n = explore(nranks=3, nsteps=2, nslots=2)
    assert ok


def test_a_single_slot_would_not_be():
    """The model has teeth: with one slot a fast rank's next push overwrites data a slow peer has not summed yet."""
    ok, _ = explore(nranks=2, nsteps=2, nslots=1)
    assert not ok
