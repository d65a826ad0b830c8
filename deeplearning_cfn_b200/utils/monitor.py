"""`Monitor` (reference MX/python/mxnet/monitor.py): every `interval` batches collect a statistic of every layer
output (and optionally parameters / gradients) whose name matches `pattern` - the `--monitor N` flag of the symbolic
example (common/fit.py:108, 236).  Implemented with forward hooks; statistics stay on the device until `toc()`."""
from __future__ import annotations

import logging
import math
import re

import torch


def _default_stat(x: torch.Tensor) -> torch.Tensor:
    """|x|_2 / sqrt(size) - the reference's default (monitor.py:62-64)."""
    return x.detach().float().norm() / math.sqrt(max(1, x.numel()))


class Monitor:
    def __init__(self, interval: int, stat_func=None, pattern: str = ".*", sort: bool = False, monitor_all: bool = False):
        self.interval, self.stat_func = interval, stat_func or _default_stat
        self.re_prog, self.sort, self.monitor_all = re.compile(pattern), sort, monitor_all
        self.step, self.activated = 0, False
        self.queue: list = []
        self._handles: list = []
        self._modules: list = []

    def install(self, module: torch.nn.Module) -> None:
        """Hook every named sub-module (an 'executor' in the reference)."""
        self._modules.append(module)
        for name, m in module.named_modules():
            if not name or not self.re_prog.match(name + "_output"):
                continue

            def hook(_m, _inp, out, name=name):
                if self.activated and isinstance(out, torch.Tensor):
                    self.queue.append((self.step, name + "_output", self.stat_func(out)))
            self._handles.append(m.register_forward_hook(hook))

    def tic(self) -> None:
        if self.step % self.interval == 0:
            self.queue = []
            self.activated = True
        self.step += 1

    def toc(self) -> list:
        if not self.activated:
            return []
        self.activated = False
        if self.monitor_all:
            for module in self._modules:
                for name, p in module.named_parameters():
                    if self.re_prog.match(name):
                        self.queue.append((self.step, name, self.stat_func(p)))
                    if p.grad is not None and self.re_prog.match(name + "_grad"):
                        self.queue.append((self.step, name + "_grad", self.stat_func(p.grad)))
        res = [(n, k, str(float(v))) for n, k, v in self.queue]
        if self.sort:
            res.sort(key=lambda t: t[1])
        self.queue = []
        return res

    def toc_print(self) -> None:
        for n, k, v in self.toc():
            logging.info("Batch: %7d %30s %s", n, k, v)

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
