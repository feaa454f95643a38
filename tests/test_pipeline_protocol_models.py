"""Host-side models of the mbarrier protocols of the experimental kernels (tools/sim_*.py): randomized schedules must
finish without deadlock, slot overwrite or accumulator mix-up, and the models must catch the bugs they were built for."""
import importlib.util
import os
import random

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, 'tools', name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_persistent_kernel_protocol_model():
    sim = _load('sim_persistent_pipeline')
    rng = random.Random(3)
    for stages in (2, 4, 6):
        for trial in range(25):
            tiles = [rng.randint(1, 20) for _ in range(rng.randint(1, 10))]
            sim.simulate(stages, tiles, seed=trial)
            sim.simulate(stages, [rng.choice([0, 0, 3, 8, 1]) for _ in tiles], seed=trial, skip_empty=True)


def test_persistent_model_detects_a_double_release():
    sim = _load('sim_persistent_pipeline')

    class TwiceMBar(sim.MBar):             # an accumulator-release barrier that every warp hits twice
        def arrive(self):
            super().arrive()
            if self.count == 4:
                super().arrive()
    orig = sim.MBar
    sim.MBar = TwiceMBar
    try:
        with pytest.raises(AssertionError):
            for trial in range(10):
                sim.simulate(4, [5, 7, 3, 9, 4, 6], seed=trial)
    finally:
        sim.MBar = orig


def test_cta_pair_protocol_model():
    sim = _load('sim_cta_pair_pipeline')
    rng = random.Random(4)
    for stages in (2, 6, 8):
        for trial in range(30):
            sim.simulate(stages, rng.randint(1, 60), seed=trial)
