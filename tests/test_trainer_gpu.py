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


def test_nvl_van_with_a_peer_on_another_host():
    """values in HBM, worker and server on different "hosts" (127.0.0.2 / 127.0.0.1): nothing can be mapped,
    so pushes are staged device -> host -> frame and pull replies frame -> host -> device; the bytes that come
    back must be the bytes that were pushed (apps/test_foreign_host.cc checks every element)"""
    import re

    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sys.path.insert(0, os.path.join(HERE, "helpers"))
    from foreign_host import run_foreign_host

    rcs, out = run_foreign_host({"PS_VAN_TYPE": "nvl", "TEST_FOREIGN_GPU": 1, "PS_CUDA_DEVICE": 0,
                                 "TEST_FOREIGN_LEN": 1 << 20}, timeout=240)
    if rcs is None:
        pytest.skip(out)
    assert rcs == [0, 0, 0], out[-3000:]
    m = re.search(r"PASSED: one-sided copies (\d+), staged copies (\d+)", out)
    assert m and int(m.group(1)) == 0 and int(m.group(2)) == 4, out[-3000:]
