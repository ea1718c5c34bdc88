"""Small host-side utilities: environment plumbing, device timing, clock sampling."""
from .env import ps_env, free_port
from .timing import CudaTimer, ClockSampler, l2_flush_buffer

__all__ = ["ps_env", "free_port", "CudaTimer", "ClockSampler", "l2_flush_buffer"]
