"""Device-side timing helpers used by bench.py (CUDA events, L2 flush, clock sampling)."""
from __future__ import annotations

import statistics
import subprocess
import threading
import time


class CudaTimer:
    """CUDA-event stopwatch on the current stream."""

    def __init__(self):
        import torch

        self._torch = torch
        self.start_ev = torch.cuda.Event(enable_timing=True)
        self.stop_ev = torch.cuda.Event(enable_timing=True)

    def start(self):
        self.start_ev.record()

    def stop(self) -> float:
        """milliseconds between start() and now, after synchronising the stop event"""
        self.stop_ev.record()
        self.stop_ev.synchronize()
        return self.start_ev.elapsed_time(self.stop_ev)


def l2_flush_buffer(device=None, mib: int = 256):
    """A buffer larger than B200's 126 MB L2; writing it evicts whatever was cached."""
    import torch

    return torch.empty(mib << 20, dtype=torch.uint8, device=device or "cuda")


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while a benchmark runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0, period_ms: int = 200):
        self.gpu_index = gpu_index
        self.period_ms = period_ms
        self.samples: list[dict] = []
        self._proc = None
        self._thread = None

    def start(self):
        try:
            self._proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu_index), "-lms", str(self.period_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self._proc = None
            return self
        self._thread = threading.Thread(target=self._pump, daemon=True)
        self._thread.start()
        return self

    def _pump(self):
        assert self._proc is not None and self._proc.stdout is not None
        for line in self._proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                self.samples.append({
                    "t": time.time(), "sm": float(f[1]), "sm_max": float(f[2]),
                    "power": float(f[3]) if f[3] not in ("[N/A]", "N/A") else 0.0,
                    "hw_slowdown": f[5], "hw_thermal": f[6], "sw_thermal": f[7],
                    "sw_power_cap": f[8]})
            except ValueError:
                continue

    def stop(self) -> dict:
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self._proc.kill()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [s for s in self.samples if s["power"] > 200] or self.samples
        reasons = set()
        for s in self.samples:
            for k in ("hw_slowdown", "hw_thermal", "sw_thermal", "sw_power_cap"):
                if s[k].lower().startswith("active"):
                    reasons.add({"hw_thermal": "hw_thermal_slowdown",
                                 "sw_thermal": "sw_thermal_slowdown"}.get(k, k))
        return {"sm_mhz": statistics.median(s["sm"] for s in busy),
                "sm_max_mhz": max(s["sm_max"] for s in self.samples),
                "reasons": sorted(reasons), "samples": len(self.samples)}
