"""Run apps/test_foreign_host with the worker on another "host": scheduler and server announce
127.0.0.1, the worker 127.0.0.2 — every 127/8 address reaches this box, and host names are all the
vans compare to decide whether a peer's memory can be mapped."""
import os
import random
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_foreign_host(env_extra, timeout=90, app="test_foreign_host", args=()):
    name = app
    app = os.path.join(ROOT, "build", name)
    if not os.path.exists(app):
        return None, f"build/{name} not built (make)"
    env = dict(os.environ)
    env.update({"DMLC_NUM_SERVER": "1", "DMLC_NUM_WORKER": "1", "DMLC_PS_ROOT_URI": "127.0.0.1",
                "DMLC_PS_ROOT_PORT": str(21000 + random.randrange(10000))})
    env.update({k: str(v) for k, v in env_extra.items()})
    env.pop("DMLC_RANK", None)
    procs = [subprocess.Popen([app, *[str(a) for a in args]], env=dict(env, DMLC_ROLE=role, DMLC_NODE_HOST=host), cwd=ROOT,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for role, host in (("scheduler", "127.0.0.1"), ("server", "127.0.0.1"), ("worker", "127.0.0.2"))]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            o, _ = p.communicate()
        outs.append(o)
    return [p.returncode for p in procs], "\n".join(outs)
