"""The cluster/multicast variant of the conv kernel is selected by an environment switch read once per process:
check it in a child process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multicast_conv_variant_matches_reference():
    env = dict(os.environ, SG_CONV_MULTICAST='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'mc_check.py')], env=env, capture_output=True, text=True,
                       timeout=240, cwd=ROOT)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
