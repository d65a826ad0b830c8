"""Training callbacks and checkpoint helpers.

Reference: `MX/python/mxnet/callback.py` - `Speedometer` logs "Epoch[%d] Batch [%d]\tSpeed: %.2f samples/sec"
every `frequent` batches (:120-182; `benchmark.py` regex-scrapes exactly that line, :156-166), `do_checkpoint`
saves every `period` epochs (:55-90); `model.py:383-440` `save_checkpoint` / `load_checkpoint` write
`prefix-symbol.json` + `prefix-%04d.params`; non-zero ranks save under `prefix-<rank>` (fit.py:57-74)."""
from __future__ import annotations

import json
import logging
import math
import os
import time

import torch


class Speedometer:
    def __init__(self, batch_size: int, frequent: int = 50, auto_reset: bool = True, rank: int | None = None):
        self.batch_size, self.frequent, self.auto_reset = batch_size, frequent, auto_reset
        self.init, self.tic, self.last_count = False, 0.0, 0
        self.prefix = f"Node[{rank}] " if rank is not None else ""
        self.last_speed = None

    def __call__(self, epoch: int, nbatch: int, metrics: dict | None = None) -> None:
        if self.last_count > nbatch:
            self.init = False
        self.last_count = nbatch
        if self.init:
            if nbatch % self.frequent == 0:
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                speed = self.frequent * self.batch_size / (time.time() - self.tic)
                self.last_speed = speed
                msg = f"{self.prefix}Epoch[{epoch}] Batch [{nbatch}]\tSpeed: {speed:.2f} samples/sec"
                for k, v in _metric_pairs(metrics):
                    msg += f"\t{k}={v:f}"
                logging.info(msg)
                if self.auto_reset and hasattr(metrics, "reset"):
                    metrics.reset()
                self.tic = time.time()
        else:
            self.init = True
            self.tic = time.time()


def save_checkpoint(prefix: str, epoch: int, model: torch.nn.Module, trainer=None, rank: int = 0, symbol: dict | None = None) -> str:
    """`prefix-symbol.json` (architecture description) + `prefix-%04d.params` (+ `.states` for the optimizer)."""
    if rank > 0:
        prefix = f"{prefix}-{rank}"
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(f"{prefix}-symbol.json", "w") as f:
        json.dump(symbol or {"class": type(model).__name__, "repr": repr(model)[:20000]}, f)
    path = f"{prefix}-{epoch:04d}.params"
    sd = {}
    for k, v in model.state_dict().items():
        kind = "aux" if ("running_" in k or "num_batches_tracked" in k) else "arg"
        sd[f"{kind}:{k}"] = v.detach().cpu()
    torch.save(sd, path)
    if trainer is not None and hasattr(trainer, "state_dict"):
        torch.save(trainer.state_dict(), f"{prefix}-{epoch:04d}.states")
    logging.info('Saved checkpoint to "%s"', path)
    return path


def load_checkpoint(prefix: str, epoch: int, model: torch.nn.Module | None = None, trainer=None, rank: int = 0):
    if rank > 0 and os.path.exists(f"{prefix}-{rank}-{epoch:04d}.params"):
        prefix = f"{prefix}-{rank}"
    raw = torch.load(f"{prefix}-{epoch:04d}.params", map_location="cpu")
    arg = {k[4:]: v for k, v in raw.items() if k.startswith("arg:")}
    aux = {k[4:]: v for k, v in raw.items() if k.startswith("aux:")}
    if model is not None:
        model.load_state_dict({**arg, **aux})
    st = f"{prefix}-{epoch:04d}.states"
    if trainer is not None and os.path.exists(st):
        trainer.load_state_dict(torch.load(st, map_location="cpu", weights_only=False))
    return arg, aux


def load_params_file(fname: str, model: torch.nn.Module, strict: bool = True):
    """Load a parameter file into `model` whichever way it was written: the checkpoint format (`arg:` / `aux:` prefixes,
    save_checkpoint / Module.save_params) or a plain state dict (`net.save_parameters`, the `-best.params` of the gluon example)."""
    raw = torch.load(fname, map_location="cpu")
    if raw and all(isinstance(k, str) and k[:4] in ("arg:", "aux:") for k in raw):
        raw = {k[4:]: v for k, v in raw.items()}
    return model.load_state_dict(raw, strict=strict)


def latest_checkpoint(prefix: str) -> int | None:
    d, b = os.path.split(os.path.abspath(prefix))
    if not os.path.isdir(d):
        return None
    eps = [int(f[len(b) + 1:-7]) for f in os.listdir(d) if f.startswith(b + "-") and f.endswith(".params") and f[len(b) + 1:-7].isdigit()]
    return max(eps) if eps else None


def do_checkpoint(prefix: str, period: int = 1, model=None, trainer=None, rank: int = 0):
    period = int(max(1, period))

    def _cb(epoch: int, *_):
        if (epoch + 1) % period == 0:
            save_checkpoint(prefix, epoch + 1, model, trainer, rank)
    return _cb


# --------------------------------------------------------------------------- remaining reference callbacks
def _metric_pairs(metrics) -> list:
    """dict | EvalMetric | None -> [(name, value)]"""
    if metrics is None:
        return []
    if isinstance(metrics, dict):
        return list(metrics.items())
    if hasattr(metrics, "get_name_value"):
        return metrics.get_name_value()
    return []


def module_checkpoint(mod, prefix: str, period: int = 1, save_optimizer_states: bool = False):
    """Epoch-end callback that checkpoints a `Module` (callback.py:27-53)."""
    period = int(max(1, period))

    def _cb(epoch: int, *_):
        if (epoch + 1) % period == 0:
            mod.save_checkpoint(prefix, epoch + 1, save_optimizer_states)
    return _cb


def log_train_metric(period: int, auto_reset: bool = False):
    """Batch-end callback: log the training metric every `period` batches (callback.py:93-117)."""
    def _cb(epoch: int, nbatch: int, metrics=None):
        if nbatch % period == 0 and metrics is not None:
            for name, value in _metric_pairs(metrics):
                logging.info("Iter[%d] Batch[%d] Train-%s=%f", epoch, nbatch, name, value)
            if auto_reset and hasattr(metrics, "reset"):
                metrics.reset()
    return _cb


class ProgressBar:
    """Batch-end callback drawing a text progress bar (callback.py:185-210)."""

    def __init__(self, total: int, length: int = 80):
        self.bar_len, self.total = length, total

    def __call__(self, epoch: int, nbatch: int, metrics=None) -> None:
        filled = int(round(self.bar_len * nbatch / float(self.total)))
        pct = math.ceil(100.0 * nbatch / float(self.total))
        logging.info("[%s] %s%s\r", "=" * filled + "-" * (self.bar_len - filled), pct, "%")


class LogValidationMetricsCallback:
    """Epoch-end evaluation callback: log every validation metric (callback.py:213-222)."""

    def __call__(self, epoch: int, metrics=None, *_) -> None:
        for name, value in _metric_pairs(metrics):
            logging.info("Epoch[%d] Validation-%s=%f", epoch, name, value)
