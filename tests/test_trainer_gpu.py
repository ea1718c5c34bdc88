"""End-to-end PS training on one B200: worker + GPU server co-located, nvl van."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("wire", ["bf16", "fp8"])
def test_joint_training_matches_local_adamw(wire):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = subprocess.run([sys.executable, os.path.join(HERE, "helpers", "train_joint.py"), wire, "8"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PASS" in r.stdout


def test_smoke_entry():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    root = os.path.dirname(HERE)
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=root,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "smoke ok" in r.stdout
