"""Persistent (one CTA per SM, double-buffered TMEM accumulator) variants of the conv and wgrad kernels — EXPERIMENTAL and
off by default (SG_CONV_PERSIST=1, SG_WGRAD_PERSIST=1).  Written at the end of round 1 without hardware access: the checks only run on request
(SG_TEST_PERSIST=1).  They re-run the conv / module / compact suites and the short-K timing script with the switch on."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ON = os.environ.get('SG_TEST_PERSIST') == '1'


@pytest.mark.skipif(not ON, reason='experimental kernel variant: set SG_TEST_PERSIST=1 to run')
def test_persistent_conv_short_k_shapes():
    for flag in ('0', '1'):
        env = dict(os.environ, SG_CONV_PERSIST=flag)
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'persist_check.py')], env=env, capture_output=True,
                           text=True, timeout=240, cwd=ROOT)
        print(r.stdout[-3000:], r.stderr[-1500:])
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1500:]


@pytest.mark.skipif(not ON, reason='experimental kernel variant: set SG_TEST_PERSIST=1 to run')
def test_conv_and_module_suites_with_persistent_kernel():
    env = dict(os.environ, SG_CONV_PERSIST='1', SG_WGRAD_PERSIST='1')
    env.pop('SG_TEST_PERSIST', None)
    r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-x', '-m', 'gpu', 'tests/test_gpu_conv_tc.py',
                        'tests/test_gpu_modules.py', 'tests/test_gpu_compact.py', 'tests/test_gpu_train_step.py'],
                       env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-1500:])
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-1500:]
