#!/usr/bin/env python
"""Discrete-event model of the barrier protocol of conv_tcp_kernel / wgrad_tcp_kernel (csrc/conv_tc.cu): a TMA producer,
one MMA thread and four epilogue warps walking the same strided tile list, an operand ring of STAGES slots
(full/empty mbarriers, count 1) and a double-buffered accumulator (acc_full count 1, acc_empty count 4).

mbarrier semantics modelled: a barrier has a phase bit and a pending-arrival counter; `wait(parity)` returns once the
phase with that parity has COMPLETED (i.e. current phase bit != parity); the last arrival flips the phase and re-arms the
counter.  Threads are generators scheduled in random order; asynchronous completions (TMA landing, tcgen05.commit) are
events with random delays.  Checked: no deadlock, every ring slot is consumed exactly once per fill and never overwritten
before its MMAs retired, every accumulator buffer is drained by all four warps before the next tile accumulates into
it, the epilogue of tile i sees exactly the MMAs of tile i.  Run: python tools/sim_persistent_pipeline.py"""
import random


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, 'more arrivals than the barrier expects in one phase'
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def done(self, parity):            # has the phase with this parity completed?
        return self.phase != parity


def simulate(stages, tiles_iters, seed, skip_empty=False):
    rng = random.Random(seed)
    full = [MBar(1) for _ in range(stages)]
    empty = [MBar(1) for _ in range(stages)]
    acc_full = [MBar(1), MBar(1)]
    acc_empty = [MBar(4), MBar(4)]
    events = []                        # (time, fn)
    now = [0]
    slot_state = ['free'] * stages     # free -> loading -> ready -> reading -> free
    slot_tag = [None] * stages
    acc_state = [('idle', None), ('idle', None)]     # (state, tile)
    acc_mmas = [0, 0]
    drained = [0, 0]
    log = {'epilogues': 0}

    def later(dt, fn):
        events.append((now[0] + dt, rng.random(), fn))

    def producer():
        s, par = 0, 0
        for t, iters in enumerate(tiles_iters):
            for it in range(iters):
                if s == stages:
                    s, par = 0, par ^ 1
                while not empty[s].done(par ^ 1):
                    yield
                assert slot_state[s] == 'free', ('overwrite of a ring slot', s, slot_state[s])
                slot_state[s], slot_tag[s] = 'loading', (t, it)

                def land(s=s):
                    slot_state[s] = 'ready'
                    full[s].arrive()             # expect_tx + complete_tx collapsed into the one arrival
                later(rng.randint(1, 30), land)
                s += 1
                yield

    def mma():
        s, par, i = 0, 0, 0
        for t, iters in enumerate(tiles_iters):
            if skip_empty and iters == 0:
                continue
            buf = i & 1
            while not acc_empty[buf].done(((i >> 1) & 1) ^ 1):
                yield
            assert acc_state[buf][0] == 'idle', ('accumulating into a buffer that is not drained', acc_state[buf])
            acc_state[buf], acc_mmas[buf] = ('accumulating', t), 0
            for it in range(iters):
                if s == stages:
                    s, par = 0, par ^ 1
                while not full[s].done(par):
                    yield
                assert slot_state[s] == 'ready' and slot_tag[s] == (t, it), ('MMA reads the wrong data', slot_tag[s], (t, it))
                slot_state[s] = 'reading'

                def retire(s=s, buf=buf):
                    slot_state[s] = 'free'
                    acc_mmas[buf] += 1
                    empty[s].arrive()            # tcgen05.commit -> empty[s]
                later(rng.randint(1, 20), retire)
                s += 1
                yield

            def acc_done(buf=buf, t=t, iters=iters):
                assert acc_mmas[buf] == iters, 'commit fired before all MMAs of the tile retired'
                acc_state[buf] = ('full', t)
                acc_full[buf].arrive()
            # tcgen05.commit completes after all previously issued MMAs: model with a delay beyond every retire
            later(25, acc_done)
            i += 1
            yield

    def epilogue(q):
        i = 0
        for t, iters in enumerate(tiles_iters):
            if skip_empty and iters == 0:
                continue
            buf = i & 1
            while not acc_full[buf].done((i >> 1) & 1):
                yield
            assert acc_state[buf] == ('full', t), ('epilogue sees the wrong accumulator', acc_state[buf], t)
            for _ in range(rng.randint(1, 4)):   # tcgen05.ld chunks
                yield
            drained[buf] += 1
            if drained[buf] == 4:
                drained[buf] = 0
                acc_state[buf] = ('idle', None)
                log['epilogues'] += 1
            acc_empty[buf].arrive()
            for _ in range(rng.randint(0, 6)):   # global stores after the buffer was handed back
                yield
            i += 1

    threads = [producer(), mma()] + [epilogue(q) for q in range(4)]
    alive = list(threads)
    idle_rounds = 0
    while alive:
        progressed = False
        # fire due events
        events.sort()
        while events and events[0][0] <= now[0]:
            events.pop(0)[2]()
            progressed = True
        rng.shuffle(alive)
        for th in list(alive):
            try:
                next(th)
            except StopIteration:
                alive.remove(th)
                progressed = True
        now[0] += 1
        idle_rounds = 0 if (progressed or events) else idle_rounds + 1
        assert idle_rounds < 2000, 'deadlock'
    while events:                      # drain outstanding completions
        events.sort()
        now[0] = events[0][0]
        events.pop(0)[2]()
    n_tiles = sum(1 for it in tiles_iters if not (skip_empty and it == 0))
    assert log['epilogues'] == n_tiles, (log, n_tiles)
    assert all(st == 'free' for st in slot_state)


if __name__ == '__main__':
    rng = random.Random(0)
    runs = 0
    for stages in (2, 3, 4, 6):
        for trial in range(150):
            n = rng.randint(1, 12)
            tiles = [rng.randint(1, 20) for _ in range(n)]
            simulate(stages, tiles, seed=trial)
            runs += 1
            tiles0 = [rng.choice([0, 0, 3, 8, 1]) for _ in range(n)]        # wgrad: empty k-splits are skipped
            simulate(stages, tiles0, seed=trial, skip_empty=True)
            runs += 1
    print('persistent pipeline protocol: %d randomized schedules, no deadlock / overwrite / mismatch' % runs)
