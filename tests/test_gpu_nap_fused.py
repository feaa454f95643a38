"""Single-kernel InstanceNorm backward for small maps (SG_NAP_FUSED=1) — EXPERIMENTAL, off by default, written at the
end of round 1 without hardware access.  On request (SG_TEST_NAP_FUSED=1) the elementwise / module / train-step suites
are re-run with the switch on."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(os.environ.get('SG_TEST_NAP_FUSED') != '1', reason='experimental kernel variant: set SG_TEST_NAP_FUSED=1 to run')
def test_suites_with_fused_instance_norm_backward():
    env = dict(os.environ, SG_NAP_FUSED='1')
    env.pop('SG_TEST_NAP_FUSED', None)
    r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-x', '-m', 'gpu', 'tests/test_gpu_elementwise.py',
                        'tests/test_gpu_modules.py', 'tests/test_gpu_train_step.py'],
                       env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-1500:])
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-1500:]
