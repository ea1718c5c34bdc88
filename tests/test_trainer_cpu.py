"""The full PS training stack without a GPU: tiny Llama workers + the server engine on its host
backend (shards in host memory, CPU twins of the kernels), shm one-sided van, under torchrun.
Same helper and same checks as the multi-GPU tests (tests/test_multigpu.py): the loss must drop
and every worker must end with identical parameters. What differs from a GPU run is only the
MemDomain (shared memory instead of HBM) and the kernel implementations."""
import os
import subprocess
import sys

import pytest

from pslite_b200.utils.env import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HELPER = os.path.join(ROOT, "tests", "helpers", "train_multi.py")


def _run(nproc, topo, wire, steps, **env_extra):
    env = dict(os.environ)
    env.update({"PSLITE_NO_AUTOBUILD": "1", "OMP_NUM_THREADS": "1", "CUDA_VISIBLE_DEVICES": ""})
    env.update({k: str(v) for k, v in env_extra.items()})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), HELPER, topo, wire, str(steps)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=280, cwd=ROOT)
    out = p.stdout + p.stderr
    assert p.returncode == 0 and "PASS" in out, out[-3000:]
    return out


_cache = {}


def _run_cached(nproc, topo, wire, steps, **env_extra):
    key = (nproc, topo, wire, steps, tuple(sorted(env_extra.items())))
    if key not in _cache:
        _cache[key] = _run(nproc, topo, wire, steps, **env_extra)
    return _cache[key]


def _curve(out):
    import re

    m = re.search(r"rank 0: engine=\w+ losses \[(.*?)\]\.\.\[(.*?)\]", out)
    return m.group(1) + " .. " + m.group(2)


def _all_losses(out):
    import re

    m = re.search(r"rank 0: all losses \[(.*?)\]", out)
    return [float(x.strip("' ")) for x in m.group(1).split(",")]


@pytest.mark.timeout(300)
def test_two_sided_fallback_with_fp8_wire(native):
    """plain torch tensors cannot be exported: gradients are fp8-encoded on the host and travel
    in the frames, pull replies come back two-sided; 2 workers + 2 servers, co-located"""
    out = _run(2, "joint", "fp8", 6)
    assert "engine=host" in out and "fused=0 " in out


PLAIN = dict(PSLITE_TEST_EXPORTABLE_PARAMS=1, PS_SHM_ASYNC=1)


@pytest.mark.timeout(300)
def test_zero_copy_fused_fanout_many_peers_async(native):
    """parameters in shared memory, copies completing asynchronously (as on a CUDA stream):
    every update writes the new bf16 parameters straight into all 4 workers' buffers"""
    out = _run_cached(4, "joint", "fp8", 5, **PLAIN)
    assert "engine=host" in out and "fused=0 " not in out


@pytest.mark.timeout(300)
def test_all_transport_options_keep_the_loss_curve(native):
    """launch coalescing (corked handlers and batch calls), fused push-pull (one message pair per
    chunk) and lazy per-module waits instead of a blocking step(), all at once: the numbers
    must be those of the plain run"""
    plain = _run_cached(4, "joint", "fp8", 5, **PLAIN)
    tuned = _run(4, "joint", "fp8", 5, PS_COALESCE_LAUNCHES=1, PSLITE_TEST_PUSHPULL=1, PSLITE_TEST_LAZY=1, **PLAIN)
    assert _curve(plain) == _curve(tuned)
    assert "fused=0 " not in tuned


@pytest.mark.timeout(300)
def test_ranks_on_different_hosts_keep_the_loss_curve(native):
    """every rank announces another host name (and the arena pretends to be device memory): each worker
    reaches its own server one-sidedly and the three others through socket frames, gradients staged and
    fp8-encoded on the host, pull replies copied into the parameters on arrival — the numbers must be those
    of the run in which everybody shares a host"""
    plain = _run_cached(4, "joint", "fp8", 5, **PLAIN)
    spread = _run(4, "joint", "fp8", 5, PS_TEST_HOST_PER_RANK=1, PS_TEST_STAGE_ARENA=1, **PLAIN)
    assert _curve(plain) == _curve(spread)


@pytest.mark.timeout(300)
def test_split_topology_and_remote_learning_rate_control(native):
    """dedicated server processes (2 workers + 2 servers, bf16 wire); opt.set_lr() reaches them
    (CMD_SET_LR): with the rate set to 0 after step 3 the parameters — and, on a fixed batch,
    the loss — stop changing"""
    out = _run(4, "split", "bf16", 7, PSLITE_TEST_FREEZE_AFTER=3, PSLITE_TEST_EXPORTABLE_PARAMS=1)
    losses = _all_losses(out)
    assert losses[2] < losses[0]                    # it was learning
    # the update of step 3's gradients already ran with rate 0: from step 4 on the loss is frozen
    assert losses[4] == losses[5] == losses[6], losses


@pytest.mark.timeout(300)
@pytest.mark.parametrize("wire", ["bf16", "fp8"])
def test_host_engine_matches_local_adamw_and_checkpoints(native, wire):
    """one worker + the host engine in one process: the loss curve tracks a local fp32-master AdamW
    run, and a save / train on / load cycle restores the saved fp32 master bit for bit"""
    env = dict(os.environ)
    env.update({"PSLITE_NO_AUTOBUILD": "1", "CUDA_VISIBLE_DEVICES": ""})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "train_joint.py"), wire, "8"],
                       env=env, capture_output=True, text=True, timeout=280, cwd=ROOT)
    out = p.stdout + p.stderr
    assert p.returncode == 0 and "PASS" in out and "restored_equal=True" in out, out[-3000:]


@pytest.mark.timeout(300)
def test_asynchronous_sgd(native):
    """async_updates: each push is applied on arrival as its own AdamW step and pulls never wait
    for other workers (the reference's asynchronous mode, docs/overview.md)"""
    env = dict(os.environ)
    env.update({"PSLITE_NO_AUTOBUILD": "1", "OMP_NUM_THREADS": "1", "CUDA_VISIBLE_DEVICES": "",
                "PSLITE_TEST_EXPORTABLE_PARAMS": "1"})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), HELPER, "joint", "bf16", "8", "async"]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=280, cwd=ROOT)
    out = p.stdout + p.stderr
    assert p.returncode == 0 and "PASS" in out, out[-3000:]


@pytest.mark.timeout(300)
def test_llama_example_script_runs(native):
    """examples/train_llama_ps.py (the tutorial as a script) on the host engine, two ranks"""
    env = dict(os.environ)
    env.update({"PSLITE_NO_AUTOBUILD": "1", "OMP_NUM_THREADS": "1", "CUDA_VISIBLE_DEVICES": ""})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "examples", "train_llama_ps.py"), "--model", "tiny", "--steps", "4"]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=280, cwd=ROOT)
    out = p.stdout + p.stderr
    assert p.returncode == 0 and "step 3: loss" in out and "host engine" in out, out[-3000:]


@pytest.mark.timeout(400)
def test_checkpoint_resume_across_jobs(native, tmp_path):
    """job 1 trains 6 steps and every server saves its shards after step 4; job 2 loads them before
    its workers initialise and reproduces steps 5 and 6 of job 1 exactly (fp32 masters, Adam
    moments, step count all restored)"""
    first = _run(2, "joint", "bf16", 6, PSLITE_CKPT_DIR=str(tmp_path), PSLITE_CKPT_AT=4)
    assert (tmp_path / "server0.ckpt").exists() and (tmp_path / "server1.ckpt").exists()
    second = _run(2, "joint", "bf16", 2, PSLITE_CKPT_DIR=str(tmp_path))
    assert "checkpoint resumed and saved" in second
    assert _all_losses(second)[:2] == _all_losses(first)[4:6], (_all_losses(second), _all_losses(first))
