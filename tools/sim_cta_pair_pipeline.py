#!/usr/bin/env python
"""Discrete-event model of the barrier protocol of conv_tc2_kernel (CTA pairs, csrc/conv_tc.cu).

Two CTAs; each has a TMA producer and a ring of STAGES slots with its own `empty` barriers (count 1); only the LEADER's
`full` barriers are used: the leader's producer arms them with arrive.expect_tx(bytes of BOTH CTAs), both producers'
TMA loads complete_tx on them; the leader's MMA thread waits `full`, issues the pair MMA (reads both CTAs' slots) and
commits — multicast — to `empty[s]` of both CTAs, finally to `accum_full` of both, which the two epilogues wait on.
mbarrier model: phase completes when pending arrivals == 0 AND tx-count == 0 (tx may go negative before the expect_tx
of the same phase, as on hardware).  Checked over randomized schedules: no deadlock, no slot overwritten before the
pair MMA that reads it retired, the MMA never starts before both halves landed.  Run: python tools/sim_cta_pair_pipeline.py"""
import random


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def arrive(self, expect=0):
        self.tx += expect
        self.pending -= 1
        assert self.pending >= 0
        self._check()

    def complete_tx(self, n):
        self.tx -= n
        self._check()

    def done(self, parity):
        return self.phase != parity


def simulate(stages, iters, seed, bytes_per_cta=3):
    rng = random.Random(seed)
    full = [MBar(1) for _ in range(stages)]                       # leader's
    empty = [[MBar(1) for _ in range(stages)] for _ in range(2)]
    accum_full = [MBar(1), MBar(1)]
    slot = [['free'] * stages for _ in range(2)]
    tag = [[None] * stages for _ in range(2)]
    events, now = [], [0]
    done_epi = [0]

    def later(dt, fn):
        events.append((now[0] + dt, rng.random(), fn))

    def producer(c):
        s, par = 0, 0
        for it in range(iters):
            if s == stages:
                s, par = 0, par ^ 1
            while not empty[c][s].done(par ^ 1):
                yield
            assert slot[c][s] == 'free', ('overwrite', c, s, slot[c][s])
            slot[c][s], tag[c][s] = 'loading', it
            if c == 0:
                full[s].arrive(expect=2 * bytes_per_cta)
            for _ in range(bytes_per_cta):                        # the A box and the B half land separately

                def land(c=c, s=s):
                    full[s].complete_tx(1)
                later(rng.randint(1, 40), land)

            def landed(c=c, s=s, it=it):
                if slot[c][s] == 'loading' and tag[c][s] == it:   # bookkeeping only (the MMA may already own the slot)
                    slot[c][s] = 'ready'
            later(41, landed)
            s += 1
            yield

    def mma():
        s, par = 0, 0
        for it in range(iters):
            if s == stages:
                s, par = 0, par ^ 1
            while not full[s].done(par):
                yield
            for c in range(2):
                # all bytes of both CTAs have landed when `full` completes; the 'ready' flag trails by design
                assert tag[c][s] == it and slot[c][s] in ('loading', 'ready'), ('pair MMA reads the wrong slot', c, s, tag[c][s], it)
                slot[c][s] = 'reading'

            def retire(s=s):
                for c in range(2):
                    slot[c][s] = 'free'
                    empty[c][s].arrive()                          # multicast commit
            later(rng.randint(1, 20), retire)
            s += 1
            yield

        def acc():
            for c in range(2):
                accum_full[c].arrive()
        later(25, acc)

    def epilogue(c):
        while not accum_full[c].done(0):
            yield
        done_epi[0] += 1

    alive = [producer(0), producer(1), mma(), epilogue(0), epilogue(1)]
    idle = 0
    while alive:
        progressed = False
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
        idle = 0 if (progressed or events) else idle + 1
        assert idle < 2000, 'deadlock'
    assert done_epi[0] == 2


if __name__ == '__main__':
    rng = random.Random(1)
    n = 0
    for stages in (2, 4, 6, 8):
        for trial in range(200):
            simulate(stages, rng.randint(1, 60), seed=trial)
            n += 1
    print('CTA-pair pipeline protocol: %d randomized schedules, no deadlock / overwrite / early MMA' % n)
