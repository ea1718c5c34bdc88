"""C++ unit tests and multi-process functional tests of the native runtime (CPU only).

Mirrors the reference's test strategy: role-from-env binaries launched as
scheduler + S servers + W workers on localhost, plus fault injection (PS_DROP_MSG with
PS_RESEND), instance groups, joint role, ipc:// mode, heartbeats and the shm one-sided van.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOCAL = os.path.join(ROOT, "scripts", "local.sh")


def launch(build_dir, servers, workers, app, *args, env=None, timeout=120):
    e = dict(os.environ)
    e.pop("DMLC_RANK", None)
    if env:
        e.update({k: str(v) for k, v in env.items()})
    cmd = [LOCAL, str(servers), str(workers), os.path.join(build_dir, app), *map(str, args)]
    r = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    return r.returncode, r.stdout + r.stderr


def test_cpp_unit_tests(built_native_tree):
    for t in ("test_foundation", "test_inproc_cluster", "test_api_surface"):
        r = subprocess.run([os.path.join(built_native_tree, "cpp_tests", t)], capture_output=True,
                           text=True, timeout=120)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("s,w", [(1, 1), (2, 2), (3, 2)])
def test_kv_app_tcp(built_native_tree, s, w):
    rc, out = launch(built_native_tree, s, w, "test_kv_app")
    assert rc == 0 and out.count("test_kv_app PASSED") == w, out[-3000:]


@pytest.mark.parametrize("async_copies", [0, 1])
def test_kv_app_one_sided_van_with_plain_buffers(built_native_tree, async_copies):
    """std::vector buffers cannot be exported: the shm van falls back to two-sided transfers for
    them and Pull() must still stitch the slices into the caller's vector."""
    rc, out = launch(built_native_tree, 2, 2, "test_kv_app",
                     env={"PS_VAN_TYPE": "shm", "PS_SHM_ASYNC": async_copies})
    assert rc == 0 and out.count("test_kv_app PASSED") == 2, out[-3000:]


def test_benchmark_one_sided_async_copies_many_peers(built_native_tree):
    """3 workers x 3 servers, exportable buffers, copies completing asynchronously."""
    env = {"PS_VAN_TYPE": "shm", "PS_SHM_ASYNC": 1, "TEST_EXPORTABLE_VALS": 1, "NUM_KEY_PER_SERVER": 10,
           "TOTAL_DURATION": 4, "LOG_DURATION": 2}
    rc, out = launch(built_native_tree, 3, 3, "test_benchmark", 262144, 100, 1, env=env)
    assert rc == 0 and "goodput" in out, out[-3000:]


@pytest.mark.parametrize("async_copies", [0, 1])
def test_descriptors_gated_by_copy_engine(built_native_tree, async_copies):
    """same-host peers: the descriptor of a one-sided transfer waits in the ring for the completion
    word its copy stores (no ticket, no completion thread); PS_GATED_FRAMES=0 restores the tickets"""
    import re

    env = {"PS_VAN_TYPE": "shm", "PS_SHM_ASYNC": async_copies, "TEST_EXPORTABLE_VALS": 1, "NUM_KEY_PER_SERVER": 6,
           "TOTAL_DURATION": 4, "LOG_DURATION": 2, "PS_VERBOSE": 1}
    rc, out = launch(built_native_tree, 2, 2, "test_benchmark", 65536, 50, 1, env=env)
    assert rc == 0 and "goodput" in out, out[-3000:]
    gated = [int(x) for x in re.findall(r"(\d+) descriptors gated by the copy engine", out)]
    copies = [int(x) for x in re.findall(r"(\d+) one-sided copies", out)]
    # (the scheduler's van reports 0 / 0)
    assert len([g for g in gated if g > 0]) == 4 and gated == copies, out[-2000:]
    env["PS_GATED_FRAMES"] = 0
    rc, out = launch(built_native_tree, 2, 2, "test_benchmark", 65536, 50, 1, env=env)
    assert rc == 0 and "goodput" in out, out[-3000:]
    assert all(int(x) == 0 for x in re.findall(r"(\d+) descriptors gated by the copy engine", out))


@pytest.mark.parametrize("van", ["zmq", "shm"])
def test_declined_ring_offer_falls_back_to_socket(built_native_tree, van):
    """a peer that cannot map the offered shared-memory ring (same IP, private /dev/shm) declines it: the
    job runs over the sockets, and a one-sided van takes the verdict for what it means — no shared memory
    with that peer — and sends frames instead of exchanging regions it could not map"""
    import re

    env = {"PS_VAN_TYPE": van, "PS_TEST_DECLINE_PIPE": 1, "PS_VERBOSE": 1, "TEST_EXPORTABLE_VALS": 1}
    rc, out = launch(built_native_tree, 2, 2, "test_kv_app", env=env)
    assert rc == 0 and out.count("test_kv_app PASSED") == 2, out[-3000:]
    assert "stay on the socket" in out
    assert all(int(x) == 0 for x in re.findall(r"(\d+) descriptors gated by the copy engine", out))
    assert all(int(x) == 0 for x in re.findall(r"(\d+) one-sided copies", out))


def test_ipc_benchmark_symmetric_buffer_and_mixed_mode(built_native_tree):
    """test_ipc_benchmark: values in a symmetric buffer (IPC_NVLS_PULL; on the shm van the replies stay
    unicast), and the reference's mixed mode — 2 co-located + 1 plain server, keys spread by its formula"""
    env = {"PS_VAN_TYPE": "shm", "JOINT": 1, "IPC_NVLS_PULL": 1, "IPC_VERIFY": 1, "NUM_KEY_PER_SERVER": 4}
    rc, out = launch(built_native_tree, 2, 2, "test_ipc_benchmark", 65536, 10, env=env)
    assert rc == 0 and out.count("VERIFIED") == 2 and "symmetric buffer" in out, out[-3000:]
    env = {"PS_VAN_TYPE": "shm", "JOINT": 1, "BYTEPS_ENABLE_MIXED_MODE": 1, "IPC_VERIFY": 1, "NUM_KEY_PER_SERVER": 6}
    rc, out = launch(built_native_tree, 3, 2, "test_ipc_benchmark", 65536, 10, env=env)
    assert rc == 0 and out.count("VERIFIED") == 2 and "mixed mode" in out, out[-3000:]


def test_large_in_frame_payloads_with_inline_dispatch_and_message_loss(built_native_tree):
    """plain heap buffers on the shm van travel INSIDE the frames; 4 joint nodes stream megabytes at each
    other while 5 % of the messages are dropped and retransmitted. Receive threads must not handle such
    messages inline (two of them streaming replies into each other's full ring would never drain their
    own): this configuration deadlocked until the rings' 60 s write timeout fired"""
    env = {"PS_VAN_TYPE": "shm", "JOINT": 1, "BENCHMARK_NTHREAD": 2, "DEBUG_MODE": 1, "PS_DROP_MSG": 5,
           "PS_RESEND": 1, "PS_RESEND_TIMEOUT": 1000}
    rc, out = launch(built_native_tree, 4, 4, "test_benchmark_stress", 2048000, 3, env=env, timeout=200)
    assert rc == 0 and out.count("test_benchmark_stress PASSED") == 4, out[-3000:]


def test_tutorial_example_runs(built_native_tree):
    """examples/kv_hello.cc is the program printed in docs/tutorials.md"""
    rc, out = launch(built_native_tree, 2, 2, "kv_hello")
    assert rc == 0 and out.count("kv_hello PASSED") == 2, out[-3000:]


@pytest.mark.parametrize("van", ["zmq", "shm"])
def test_in_process_handoff(built_native_tree, van):
    """PS_LOCAL_HANDOFF=1: vans of one process (joint role) pass data messages — and the one-sided
    van's rendezvous messages — directly, bypassing wire format and receive thread"""
    env = {"PS_LOCAL_HANDOFF": 1, "PS_VAN_TYPE": van, "JOINT": 1, "PS_VERBOSE": 0}
    rc, out = launch(built_native_tree, 2, 2, "test_kv_app", env=env)
    assert rc == 0 and out.count("test_kv_app PASSED") == 2, out[-3000:]
    env.update({"NUM_KEY_PER_SERVER": 8, "TOTAL_DURATION": 50, "LOG_DURATION": 25})
    rc, out = launch(built_native_tree, 2, 2, "test_ipc_benchmark", 65536, 50, env=env)
    assert rc == 0 and "goodput" in out, out[-3000:]


def test_kv_app_ipc_sockets(built_native_tree):
    rc, out = launch(built_native_tree, 2, 2, "test_kv_app", env={"DMLC_LOCAL": 1})
    assert rc == 0 and out.count("PASSED") == 2, out[-3000:]


def test_kv_app_lockless_queue_and_direct_dispatch(built_native_tree):
    rc, out = launch(built_native_tree, 2, 1, "test_kv_app", env={"DMLC_LOCKLESS_QUEUE": 1})
    assert rc == 0 and "PASSED" in out, out[-3000:]


def test_kv_app_survives_message_drops_with_resend(built_native_tree):
    env = {"PS_RESEND": 1, "PS_RESEND_TIMEOUT": 100, "PS_DROP_MSG": 10}
    # 40 pushes + the rest: ~100 messages, so "no message was dropped" has probability < 1e-4
    rc, out = launch(built_native_tree, 1, 1, "test_kv_app", 200, 2, 40, env=env, timeout=180)
    assert "test_kv_app PASSED" in out, out[-3000:]   # every push/pull survived 10 % loss
    assert "Drop message" in out                      # the injector really fired
    if rc != 0:  # teardown with messages still being retransmitted is best-effort: retry once
        rc, out = launch(built_native_tree, 1, 1, "test_kv_app", 200, 2, 40, env=env, timeout=180)
    assert rc == 0, out[-3000:]


def test_resend_survives_many_requests_without_false_duplicates(built_native_tree):
    """more than 32767 requests of one customer with PS_RESEND on: the timestamp that equals the
    'no timestamp' sentinel is skipped (the reference aborts there), message identities do not
    wrap, and an ACK that overtakes the sender's own bookkeeping is not lost — so a loss-free run
    shows no duplicate at all"""
    env = {"PS_RESEND": 1, "PS_RESEND_TIMEOUT": 1000, "NUM_KEY_PER_SERVER": 40, "TOTAL_DURATION": 1000,
           "LOG_DURATION": 500}
    rc, out = launch(built_native_tree, 1, 1, "test_benchmark", 1024, 10, 1, env=env, timeout=240)
    assert rc == 0 and "goodput" in out, out[-3000:]
    assert "Duplicated message" not in out, out[-2000:]


def test_kv_app_instance_groups(built_native_tree):
    rc, out = launch(built_native_tree, 1, 1, "test_kv_app", 500, 2, 3,
                     env={"DMLC_GROUP_SIZE": 2, "SET_RANKS": 1})
    assert rc == 0 and "PASSED" in out, out[-3000:]


def test_kv_app_joint_role(built_native_tree):
    rc, out = launch(built_native_tree, 2, 2, "test_kv_app", env={"JOINT": 1})
    assert rc == 0 and out.count("PASSED") == 2, out[-3000:]


def test_kv_app_preferred_ranks(built_native_tree):
    rc, out = launch(built_native_tree, 2, 2, "test_connection", env={"SET_RANKS": 1})
    assert rc == 0 and out.count("test_connection PASSED") == 5, out[-3000:]


def test_simple_app(built_native_tree):
    rc, out = launch(built_native_tree, 2, 2, "test_simple_app")
    assert rc == 0 and out.count("test_simple_app PASSED") == 2, out[-3000:]


def test_connection_with_heartbeats(built_native_tree):
    rc, out = launch(built_native_tree, 1, 2, "test_connection",
                     env={"PS_HEARTBEAT_INTERVAL": 1, "PS_HEARTBEAT_TIMEOUT": 5})
    assert rc == 0 and out.count("test_connection PASSED") == 4, out[-3000:]


@pytest.mark.parametrize("van", ["zmq", "shm"])
def test_benchmark_small_and_large(built_native_tree, van):
    for length in (1024, 1024000):
        rc, out = launch(built_native_tree, 2, 2, "test_benchmark", length, 5, 1,
                         env={"PS_VAN_TYPE": van, "NUM_KEY_PER_SERVER": 4, "TOTAL_DURATION": 10,
                              "LOG_DURATION": 5})
        assert rc == 0 and "Application goodput" in out, out[-3000:]


def test_benchmark_registered_recv_buffers_shm(built_native_tree):
    rc, out = launch(built_native_tree, 2, 2, "test_benchmark", 256000, 5, 1,
                     env={"PS_VAN_TYPE": "shm", "ENABLE_RECV_BUFFER": 1, "NUM_KEY_PER_SERVER": 4,
                          "TOTAL_DURATION": 10, "LOG_DURATION": 5})
    assert rc == 0 and "Application goodput" in out, out[-3000:]


def test_benchmark_registered_recv_buffers_tcp(built_native_tree):
    rc, out = launch(built_native_tree, 1, 1, "test_benchmark", 256000, 5, 1,
                     env={"ENABLE_RECV_BUFFER": 1, "NUM_KEY_PER_SERVER": 4, "TOTAL_DURATION": 10,
                          "LOG_DURATION": 5})
    assert rc == 0 and "Application goodput" in out, out[-3000:]


@pytest.mark.parametrize("mode", [0, 2, 3])
def test_benchmark_modes(built_native_tree, mode):
    rc, out = launch(built_native_tree, 1, 1, "test_benchmark", 64000, 3, mode,
                     env={"NUM_KEY_PER_SERVER": 4, "TOTAL_DURATION": 10, "LOG_DURATION": 5})
    assert rc == 0, out[-3000:]
    assert ("total_time" in out) if mode == 0 else ("Application goodput" in out), out[-2000:]


def test_van_profiling_log(built_native_tree, tmp_path):
    prefix = str(tmp_path / "prof")
    rc, out = launch(built_native_tree, 1, 1, "test_benchmark", 4096, 3, 1,
                     env={"ENABLE_PROFILING": 1, "PROFILE_PATH": prefix, "NUM_KEY_PER_SERVER": 2,
                          "TOTAL_DURATION": 4, "LOG_DURATION": 2})
    assert rc == 0, out[-2000:]
    server_log = open(prefix + "_van_server").read().strip().splitlines()
    assert server_log and server_log[0].split("\t")[1] in ("server_van_recv_push", "server_van_recv_pull")


@pytest.mark.parametrize("van", ["zmq", "shm", "shm-onesided"])
def test_dead_node_is_replaced_by_late_registration(built_native_tree, van):
    """heartbeats -> dead-node detection -> a late worker inherits the dead worker's id. "shm-onesided":
    values in exportable memory, so the survivors hold mappings of the dead process's regions under the
    node id the replacement takes over — its pull replies must land in ITS memory."""
    import random
    import time

    app = os.path.join(built_native_tree, "test_recovery")
    env = dict(os.environ)
    env.update({"DMLC_NUM_SERVER": "1", "DMLC_NUM_WORKER": "2", "DMLC_PS_ROOT_URI": "127.0.0.1",
                "DMLC_PS_ROOT_PORT": str(21000 + random.randrange(10000)), "DMLC_NODE_HOST": "127.0.0.1",
                "PS_HEARTBEAT_INTERVAL": "1", "PS_HEARTBEAT_TIMEOUT": "2", "PS_VAN_TYPE": van.split("-")[0]})
    if van.endswith("-onesided"):
        env["RECOVERY_ONESIDED"] = "1"
    env.pop("DMLC_RANK", None)

    def spawn(role, **extra):
        e = dict(env, DMLC_ROLE=role, **extra)
        return subprocess.Popen([app], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)

    procs = [spawn("scheduler"), spawn("server"), spawn("worker"), spawn("worker", RECOVERY_CRASH="1")]
    time.sleep(5)  # > heartbeat timeout: the crashed worker is now considered dead
    procs.append(spawn("worker", RECOVERY_LATE="1"))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=60)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            o, _ = p.communicate()
        outs.append(o)
    joined = "\n".join(outs)
    assert "test_recovery PASSED" in joined, joined[-4000:]
    assert "recovery worker adopted rank" in joined  # whichever rank the dead worker had
    assert all(p.returncode == 0 for p in procs), [p.returncode for p in procs]


@pytest.mark.parametrize("staged", [0, 1])
def test_onesided_van_with_a_peer_on_another_host(built_native_tree, staged):
    """a peer whose host name differs cannot map memory: the one-sided van sends frames instead, and with
    device-like memory (PS_TEST_STAGE_ARENA: the arena pretends to be HBM) stages through the host on both
    ends; plain pull and fused push-pull must return the pushed bytes exactly"""
    import re
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
    from foreign_host import run_foreign_host

    rcs, out = run_foreign_host({"PS_VAN_TYPE": "shm", "PS_TEST_STAGE_ARENA": staged})
    assert rcs == [0, 0, 0], out[-3000:]
    m = re.search(r"PASSED: one-sided copies (\d+), staged copies (\d+)", out)
    assert m, out[-3000:]
    assert int(m.group(1)) == 0  # nothing was written into the peer's memory
    assert int(m.group(2)) == (4 if staged else 0)  # push, pull, and both halves of the push-pull


@pytest.mark.parametrize("staged", [0, 1])
def test_registered_receive_buffers_with_a_sender_on_another_host(built_native_tree, staged):
    """KVServer::RegisterRecvBuffer promises the handler the values IN the registered buffer; a push from
    another host arrives in a frame, so the van copies it there (host to device for a buffer in HBM; the
    arena plays HBM with PS_TEST_STAGE_ARENA). test_benchmark checks the pointer on every push."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
    from foreign_host import run_foreign_host

    rcs, out = run_foreign_host({"PS_VAN_TYPE": "shm", "TEST_EXPORTABLE_VALS": 1, "ENABLE_RECV_BUFFER": 1,
                                 "NUM_KEY_PER_SERVER": 4, "TOTAL_DURATION": 20, "LOG_DURATION": 10,
                                 "PS_TEST_STAGE_ARENA": staged}, app="test_benchmark", args=(65536, 20, 1))
    assert rcs == [0, 0, 0] and "goodput" in out, out[-3000:]
