"""Launcher / tracker / lint / baseline-arm tests (CPU only)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tracker"))


def test_lint_clean():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "lint.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]


def test_tracker_topology():
    import tracker

    for n in (1, 2, 5, 8, 13):
        tree, parent, ring = tracker.link_map(n)
        assert sorted(tree) == list(range(n))
        assert parent[0] == -1
        # ring is a single cycle 0 -> 1 -> ... -> n-1 -> 0 after relabelling
        for r in range(n):
            assert ring[r] == ((r - 1) % n, (r + 1) % n)
        # tree edges are symmetric and connect everything
        for r, nbrs in tree.items():
            for x in nbrs:
                assert r in tree[x]
        seen, stack = {0}, [0]
        while stack:
            for x in tree[stack.pop()]:
                if x not in seen:
                    seen.add(x)
                    stack.append(x)
        assert len(seen) == n


def test_dmlc_local_launcher(built_native_tree):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tracker", "dmlc_local.py"), "-n", "2", "-s", "2",
                        os.path.join(built_native_tree, "test_kv_app"), "300", "2", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and out.count("test_kv_app PASSED") == 2, out[-3000:]


def test_dmlc_ssh_and_mpi_dry_run(tmp_path):
    hosts = tmp_path / "hosts"
    hosts.write_text("nodeA:2200\n# comment\nnodeB\n")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tracker", "dmlc_ssh.py"), "-n", "2", "-s", "1",
                        "-H", str(hosts), "--dry-run", "echo", "hi"], capture_output=True, text=True)
    assert r.returncode == 0
    lines = [line for line in r.stdout.splitlines() if line.startswith("ssh ")]
    assert len(lines) == 3 and "nodeA -p 2200" in lines[0] and "DMLC_ROLE=server" in lines[0]
    assert "DMLC_ROLE=worker" in lines[1] and "nodeB -p 22" in lines[1]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tracker", "dmlc_mpi.py"), "-n", "4", "-s", "2",
                        "--dry-run", "./app"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.count("mpirun -n") == 2


def test_reference_arm_runs_unmodified_reference():
    """bench.py --impl reference builds baseline/_ref and runs the reference's own test_benchmark."""
    if not os.path.isdir(os.environ.get("PS_REFERENCE_SRC", "/root/reference")) and not os.path.isdir(
            os.path.join(ROOT, "baseline", "_ref", "src")):
        pytest.skip("reference sources not available")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4",
                        "--warmup", "1", "--len", "256000", "--keys-per-server", "4"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert "unavailable" in line or line["value"] > 0
