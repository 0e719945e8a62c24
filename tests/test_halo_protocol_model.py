"""Model check of the peer-memory halo protocol of csrc/comm.cu (k_halo_p2p) on the CPU.

Each rank runs, per exchange s = 1, 2, ...:   wait ACK_from(nb) >= s-2  ->  write its planes into slot s&1 of the
neighbour's inbox  ->  raise ARR at the neighbour  ->  wait own ARR >= s  ->  read slot s&1 of its inbox  ->  raise
ACK at the neighbour.  The model interleaves the ranks' steps at random (a rank only blocks in its waits) and
asserts what the kernel relies on: a slot is never overwritten before its reader has read it, every read sees the
planes of the matching exchange, and no schedule deadlocks."""
import random

import pytest


def run(world, n_exchanges, rng):
    # inbox[r][d][slot] = (sender, seq) ; d = 0: from below, 1: from above
    inbox = [[[None, None], [None, None]] for _ in range(world)]
    arr = [[0, 0] for _ in range(world)]      # arr[r][d]: newest seq arrived from below / above
    ack = [[0, 0] for _ in range(world)]      # ack[r][d]: newest seq the rank below / above has consumed from ITS inbox
    unread = [[[False, False], [False, False]] for _ in range(world)]
    # program counter per rank: (seq, phase) ; phases: 0 wait-ack, 1 send, 2 announce, 3 wait-arr, 4 read, 5 ack
    pc = [[1, 0] for _ in range(world)]
    got = [[] for _ in range(world)]
    nbrs = lambda r: [(0, r - 1) for _ in [0] if r > 0] + [(1, r + 1) for _ in [0] if r < world - 1]
    steps = 0
    while any(p[0] <= n_exchanges for p in pc):
        runnable = []
        for r in range(world):
            s, ph = pc[r]
            if s > n_exchanges:
                continue
            if ph == 0 and not all(s <= 2 or ack[r][d] >= s - 2 for d, _ in nbrs(r)):
                continue
            if ph == 3 and not all(arr[r][d] >= s for d, _ in nbrs(r)):
                continue
            runnable.append(r)
        assert runnable, f"deadlock at {pc}"
        r = rng.choice(runnable)
        s, ph = pc[r]
        sl = s & 1
        if ph == 1:
            for d, nb in nbrs(r):
                dd = 1 - d                      # I am "above" my lower neighbour and "below" my upper one
                assert not unread[nb][dd][sl], f"rank {r} overwrites an unread slot of rank {nb} at seq {s}"
                inbox[nb][dd][sl] = (r, s)
                unread[nb][dd][sl] = True
        elif ph == 2:
            for d, nb in nbrs(r):
                arr[nb][1 - d] = s
        elif ph == 4:
            for d, nb in nbrs(r):
                assert inbox[r][d][sl] == (nb, s), f"rank {r} reads {inbox[r][d][sl]} instead of ({nb}, {s})"
                got[r].append((nb, s))
                unread[r][d][sl] = False
        elif ph == 5:
            for d, nb in nbrs(r):
                ack[nb][1 - d] = s
        pc[r] = [s, ph + 1] if ph < 5 else [s + 1, 0]
        steps += 1
    return got, steps


@pytest.mark.parametrize("world", [2, 3, 8])
def test_no_overwrite_no_deadlock_random_schedules(world):
    rng = random.Random(1234 + world)
    for _ in range(60):
        got, _ = run(world, 9, rng)
        for r in range(world):
            want = [(nb, s) for s in range(1, 10) for nb in ([r - 1] if r > 0 else []) + ([r + 1] if r < world - 1 else [])]
            assert got[r] == want


def test_greedy_rank_cannot_overrun_its_neighbour():
    """Most adversarial schedule for slot reuse: rank 0 runs whenever it is not blocked (it gets as far ahead as the
    s-2 acknowledgement rule lets it: one full exchange); the invariants inside run() must still hold."""
    rng = random.Random(7)

    class Greedy(random.Random):
        def choice(self, seq):
            return 0 if 0 in seq else rng.choice(seq)

    for world in (2, 4):
        for _ in range(20):
            got, _ = run(world, 8, Greedy())
            assert [s for nb, s in got[0]] == list(range(1, 9))
