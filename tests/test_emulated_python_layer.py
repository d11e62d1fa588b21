"""The product's Python layer (fgs_b200._C, torch_bindings, loss, densification) run UNCHANGED on host tensors against the
emulated build of the whole library (tests/cuda_emu 'all', selected with FGS_B200_LIB), in a subprocess so that this
process keeps the real libfgs_b200.so.  See tests/emu_python_layer.py for what is compared; part "gpu_tests" executes the
bodies of the `-m gpu` tests of the widened rows with device "cpu".  The mini training loop (the
code of tests/test_widen_training.py at a toy size, ~3 min) only runs with FGS_TEST_SLOW=1."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

SCRIPT = Path(__file__).resolve().parent / 'emu_python_layer.py'
# two subprocesses (each pays the interpreter + torch start-up once)
PARTS = [('rasterizer', 'reference_bindings', 'loss', 'peer_adam', 'deferred_sh'), ('densification', 'fourd', 'gpu_tests')] + \
    ([('training',)] if os.environ.get('FGS_TEST_SLOW') == '1' else [])


@pytest.mark.parametrize('parts', PARTS, ids=lambda p: '+'.join(p))
def test_python_layer_on_emulated_library(parts):
    env = {k: v for k, v in os.environ.items() if k != 'FGS_B200_LIB'}
    r = subprocess.run([sys.executable, str(SCRIPT), *parts], capture_output=True, text=True, timeout=1700, env=env)
    assert r.returncode == 0 and 'ALL OK' in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
