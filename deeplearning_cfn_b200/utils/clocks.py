"""GPU clock / throttle sampling during a timed region (B200_PROFILING.md "clocks line")."""
from __future__ import annotations

import shutil
import statistics
import subprocess
import tempfile

_FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
           "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
           "clocks_event_reasons.sw_power_cap")
_REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")


class ClockSampler:
    """Runs `nvidia-smi -lms 200` for the duration of a `with` block and summarises it."""

    def __init__(self, gpu_index: int | None = None, period_ms: int = 200):
        self.gpu_index = gpu_index
        self.period_ms = period_ms
        self._proc = None
        self._file = None

    def __enter__(self):
        exe = shutil.which("nvidia-smi")
        if exe is None:
            return self
        self._file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        cmd = [exe, f"--query-gpu={_FIELDS}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms)]
        if self.gpu_index is not None:
            cmd += ["-i", str(self.gpu_index)]
        self._proc = subprocess.Popen(cmd, stdout=self._file, stderr=subprocess.DEVNULL)
        return self

    def __exit__(self, *exc):
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self._proc.kill()
        return False

    def summary(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self._file is None:
            return out
        self._file.flush()
        sm, mx, seen = [], [], set()
        with open(self._file.name) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(_REASONS, parts[4:8]):
                    if val.lower().startswith("active"):
                        seen.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(seen), samples=len(sm))
        return out
