"""Synchronous-replica training with backup workers (TF `SyncReplicasOptimizer` semantics) on the dlcfn
parameter server.

Reference workload W6: `TFM/inception/inception/inception_distributed_train.py:197` and
`TFM/slim/train_image_classifier.py:526` wrap their optimizer in `tf.train.SyncReplicasOptimizer`
(`TFV/python/training/sync_replicas_optimizer.py`): every variable has a conditional accumulator on its parameter
server; a worker pushes (gradient, local_step); gradients whose local_step is older than the global step are
DROPPED; when `replicas_to_aggregate` fresh gradients have arrived their MEAN is applied, the global step advances
and `total_num_replicas` tokens carrying the new step are released; a worker proceeds to its next step only after
taking a token.  With replicas_to_aggregate < total_num_replicas the slowest workers are "backup workers" whose
late gradients are simply discarded.

Here the same protocol runs on the native parameter server of csrc/rt in `dist_async` mode (every push is handed to
the server-side updater at once, kvstore_dist_server.h:348,715-723):

  * `SyncReplicasOptimizer` is an ordinary `kv.optimizer.Optimizer` that wraps the real one.  `set_optimizer`
    pickles it to the servers (kvstore.py:480-497), where its `update()` IS the conditional accumulator: it reads the
    worker's local step from the last element of the pushed vector, drops stale pushes, accumulates fresh ones,
    and on the R-th fresh push applies the inner optimizer to the mean and bumps the step slot stored as the last
    element of the weight vector.
  * the token queue becomes "pull until the step slot is newer than my local step" - the released token IS the new
    global step, exactly what TF's sync_token_q carries.
  * parameters are packed into flat chunks smaller than the server's big-array bound, each with its own step slot,
    so a chunk lives on exactly one server ((key * 9973) % S) and chunks spread over all servers.
"""
from __future__ import annotations

import os
import time

import torch

from ..kv import optimizer as opt


@opt.register
class SyncReplicasOptimizer(opt.Optimizer):
    """Server-side conditional accumulator + token release around `inner` (see module docstring)."""

    def __init__(self, inner=None, replicas_to_aggregate: int = 1, total_num_replicas: int | None = None, **kw):
        super().__init__(**kw)
        self.inner = opt.create(inner) if inner is not None else opt.create("sgd")
        self.replicas_to_aggregate = int(replicas_to_aggregate)
        self.total_num_replicas = int(total_num_replicas or replicas_to_aggregate)
        if self.replicas_to_aggregate < 1:
            raise ValueError("replicas_to_aggregate must be >= 1")
        self.dropped = 0          # stale gradients discarded on this server (observability)
        self.applied = 0

    def create_state(self, index, weight):
        n = weight.numel() - 1
        return {"acc": torch.zeros(n, dtype=torch.float32), "count": 0,
                "inner": self.inner.create_state(index, weight.reshape(-1)[:n].float())}

    def update(self, index, weight, grad, state):
        w = weight.reshape(-1)
        g = grad.reshape(-1)
        n = w.numel() - 1
        local_step = int(round(float(g[n])))
        global_step = int(round(float(w[n])))
        if local_step < global_step:                      # stale: computed against an older model
            self.dropped += 1
            return
        state["acc"] += g[:n].float()
        state["count"] += 1
        if state["count"] < self.replicas_to_aggregate:
            return
        mean = state["acc"] / float(state["count"])
        body = w[:n]
        if body.dtype != torch.float32:
            tmp = body.float()
            self.inner.update(index, tmp, mean, state["inner"])
            body.copy_(tmp.to(body.dtype))
        else:
            self.inner.update(index, body, mean, state["inner"])
        w[n] = float(global_step + 1)                     # releases the tokens: every worker sees the new step
        state["acc"].zero_()
        state["count"] = 0
        self.applied += 1


class SyncReplicasTrainer:
    """Worker side.  `kv` is a `dist_async` KVStore; `params` the tensors to train (updated in place by `step`).

        kv = deeplearning_cfn_b200.kv.create("dist_async")
        t = SyncReplicasTrainer(kv, list(model.parameters()), "sgd", {"learning_rate": 0.1},
                                replicas_to_aggregate=kv.num_workers - 1)
        loss.backward(); t.step()
    """

    def __init__(self, kv, params, optimizer="sgd", optimizer_params: dict | None = None,
                 replicas_to_aggregate: int | None = None, total_num_replicas: int | None = None,
                 key_base: int = 7000, poll_interval: float = 0.0005, timeout: float = 300.0):
        if "async" not in kv.type:
            raise ValueError("SyncReplicasTrainer needs a dist_async kvstore (the accumulator does the synchronisation)")
        self.kv = kv
        self.params = [p for p in params]
        self.total = total_num_replicas or kv.num_workers
        self.R = replicas_to_aggregate or self.total
        if self.R > self.total:
            raise ValueError("replicas_to_aggregate cannot exceed total_num_replicas")
        self.poll, self.timeout = poll_interval, timeout
        bound = int(os.environ.get("MXNET_KVSTORE_BIGARRAY_BOUND", "1000000"))
        chunk = max(1, bound - 2)
        self.numel = sum(p.numel() for p in self.params)
        self.chunks = [(lo, min(lo + chunk, self.numel)) for lo in range(0, self.numel, chunk)]
        self.keys = [key_base + i for i in range(len(self.chunks))]
        flat = torch.cat([p.detach().reshape(-1).float().cpu() for p in self.params]) if self.params else torch.zeros(0)
        for k, (lo, hi) in zip(self.keys, self.chunks):
            kv.init(k, torch.cat([flat[lo:hi], torch.zeros(1)]))            # [weights..., global_step]
        inner = opt.create(optimizer, **(optimizer_params or {}))
        kv.set_optimizer(SyncReplicasOptimizer(inner=inner, replicas_to_aggregate=self.R, total_num_replicas=self.total))
        self.local_step = 0
        self.steps_taken = 0

    @property
    def global_step(self) -> int:
        return self.local_step

    def _flat_grad(self) -> torch.Tensor:
        return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).detach().reshape(-1).float().cpu()
                          for p in self.params])

    def step(self) -> int:
        """Push this replica's gradients stamped with its local step, then wait for a token (= the global step moving
        past the local step) and load the fresh weights.  Returns the new global step."""
        g = self._flat_grad()
        stamp = torch.tensor([float(self.local_step)])
        for k, (lo, hi) in zip(self.keys, self.chunks):
            self.kv.push(k, torch.cat([g[lo:hi], stamp]))
        new_flat = torch.empty(self.numel)
        new_step = None
        self.last_steps = []
        deadline = time.time() + self.timeout
        for k, (lo, hi) in zip(self.keys, self.chunks):
            buf = torch.zeros(hi - lo + 1)
            while True:
                self.kv.pull(k, out=buf)
                s = int(round(float(buf[-1])))
                if s > self.local_step:
                    break
                if time.time() > deadline:
                    raise TimeoutError(f"no sync token for key {k}: global step stuck at {s} "
                                       f"(fewer than {self.R} live replicas?)")
                time.sleep(self.poll)
            new_flat[lo:hi] = buf[:-1]
            self.last_steps.append(s)
            new_step = s if new_step is None else min(new_step, s)
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                p.copy_(new_flat[off:off + n].view(p.shape).to(p.dtype))
                off += n
        self.local_step = new_step if new_step is not None else self.local_step + 1
        self.steps_taken += 1
        return self.local_step
