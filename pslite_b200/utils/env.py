"""DMLC_* environment construction (parity: reference docs/env.md, tracker/dmlc_local.py)."""
from __future__ import annotations

import os
import socket


def free_port() -> int:
    s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def ps_env(role: str, num_workers: int, num_servers: int, root_uri: str, root_port: int,
           van: str = "zmq", extra: dict | None = None) -> dict:
    """Environment block for one process of a ps job."""
    env = {
        "DMLC_ROLE": role,
        "DMLC_NUM_WORKER": str(num_workers),
        "DMLC_NUM_SERVER": str(num_servers),
        "DMLC_PS_ROOT_URI": root_uri,
        "DMLC_PS_ROOT_PORT": str(root_port),
        "DMLC_NODE_HOST": os.environ.get("DMLC_NODE_HOST", "127.0.0.1"),
        "PS_VAN_TYPE": van,
    }
    if extra:
        env.update({k: str(v) for k, v in extra.items()})
    return env
