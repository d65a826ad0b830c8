"""`Trainer` - applies an optimizer to a set of parameters through a KVStore.

Reference: `mxnet.gluon.Trainer` (MX/python/mxnet/gluon/trainer.py:62-434): constructor
`Trainer(params, optimizer, optimizer_params=None, kvstore='device', compression_params=None,
update_on_kvstore=None)`; `_init_kvstore` decides `update_on_kvstore` (forced for dist async, :158-218);
`step(batch_size)` = `allreduce_grads()` (push every gradient with priority=-index, pull it back unless the
update happens on the kvstore, :315-323) + `update()` (pull weights, or run local updaters, :358-385);
`set_learning_rate`, `learning_rate`, `save_states` / `load_states` (:387-434).
Parameters are torch `nn.Parameter`s; the per-key path below is API parity, the bucketed fused path for
throughput is `parallel.trainer.DataParallelTrainer`.
"""
from __future__ import annotations

import pickle

import torch

from . import optimizer as opt
from .kvstore import KVStore, KVStoreError, create


class Trainer:
    def __init__(self, params, optimizer, optimizer_params: dict | None = None, kvstore="device",
                 compression_params: dict | None = None, update_on_kvstore: bool | None = None):
        if hasattr(params, "named_parameters"):
            params = dict(params.named_parameters())
        if not isinstance(params, dict):
            params = list(params)
            if params and isinstance(params[0], tuple):
                params = dict(params)
        if isinstance(params, dict):
            self._names = list(params.keys())
            self._params = [params[k] for k in self._names]
        else:
            self._params = list(params)
            self._names = [str(i) for i in range(len(self._params))]
        self._params = [p for p in self._params if p.requires_grad]
        optimizer_params = dict(optimizer_params or {})
        self._scale = optimizer_params.get("rescale_grad", 1.0)
        if isinstance(optimizer, opt.Optimizer):
            if optimizer_params:
                raise ValueError("optimizer_params must be None if optimizer is an Optimizer instance")
            self._optimizer = optimizer
        else:
            self._optimizer = opt.create(optimizer, param_idx2name=dict(enumerate(self._names)), **optimizer_params)
        self._updaters = [opt.get_updater(self._optimizer)]
        self._compression_params = compression_params
        self._kvstore_arg = kvstore
        self._update_on_kvstore_arg = update_on_kvstore
        self._kv_initialized = False
        self._kvstore: KVStore | None = None
        self._update_on_kvstore = False

    # ------------------------------------------------------------------ kvstore
    def _init_kvstore(self) -> None:
        kvs = self._kvstore_arg
        if isinstance(kvs, str):
            kvs = create(kvs) if kvs else None
        update_on_kv = self._update_on_kvstore_arg
        if kvs is not None:
            if self._compression_params:
                kvs.set_gradient_compression(self._compression_params)
            if "dist" in kvs.type:
                if update_on_kv is None:
                    update_on_kv = True
                if "_async" in kvs.type and update_on_kv is False:
                    raise ValueError("Please set update_on_kvstore to true when training in async mode")
            elif update_on_kv is None:
                update_on_kv = False
            if kvs.num_workers == 1 and "dist" not in kvs.type and len(self._params) and not self._compression_params \
                    and update_on_kv is False and self._kvstore_arg in ("device", "local"):
                # single process, single copy of every parameter: nothing to aggregate
                kvs = None
        self._kvstore = kvs
        self._update_on_kvstore = bool(update_on_kv) and kvs is not None
        if kvs is not None:
            for i, p in enumerate(self._params):
                kvs.init(i, p.data)
            if self._update_on_kvstore:
                kvs.set_optimizer(self._optimizer)
            for i, p in enumerate(self._params):       # first pull: every worker starts from rank 0's values
                kvs.pull(i, out=p.data, priority=-i)
        self._kv_initialized = True
        self._shipped = self._optimizer_signature()

    def _optimizer_signature(self):
        o = self._optimizer
        return (float(o.rescale_grad), float(o.learning_rate), float(getattr(o, "wd", 0.0)))

    def _sync_optimizer_to_kvstore(self) -> None:
        """With update_on_kvstore the optimizer lives on the servers: a change of rescale_grad (batch size), learning rate
        or weight decay made on the worker's copy is re-sent (upstream raises in _check_and_rescale_grad; the native
        servers keep their per-key states across a config update, so re-sending is safe).  Every worker calls step() /
        set_learning_rate() with the same values, which keeps the set_optimizer barrier symmetric."""
        if not (self._update_on_kvstore and self._kvstore is not None):
            return
        sig = self._optimizer_signature()
        if sig != self._shipped:
            self._kvstore.set_optimizer(self._optimizer)
            self._shipped = sig

    @property
    def learning_rate(self) -> float:
        return self._optimizer.learning_rate

    def set_learning_rate(self, lr: float) -> None:
        self._optimizer.set_learning_rate(lr)

    # ------------------------------------------------------------------ step
    def step(self, batch_size: int, ignore_stale_grad: bool = False) -> None:
        # rescale_grad BEFORE the store is initialised: _init_kvstore ships the optimizer to the servers (trainer.py:192-197)
        self._optimizer.rescale_grad = self._scale / batch_size
        if not self._kv_initialized:
            self._init_kvstore()
        self._sync_optimizer_to_kvstore()
        self._allreduce_grads()
        self._update(ignore_stale_grad)

    def allreduce_grads(self) -> None:
        if not self._kv_initialized:
            self._init_kvstore()
        if self._update_on_kvstore:
            raise AssertionError("allreduce_grads() when parameters are updated on kvstore is not supported. "
                                 "Try setting `update_on_kvstore` to False when creating trainer.")
        self._allreduce_grads()

    def _allreduce_grads(self) -> None:
        if self._kvstore is None:
            return
        for i, p in enumerate(self._params):
            if p.grad is None:
                continue
            self._kvstore.push(i, p.grad, priority=-i)
            if not self._update_on_kvstore:
                self._kvstore.pull(i, out=p.grad, priority=-i)

    def update(self, batch_size: int, ignore_stale_grad: bool = False) -> None:
        if not self._kv_initialized:
            self._init_kvstore()
        if self._update_on_kvstore:
            raise AssertionError("update() when parameters are updated on kvstore is not supported. "
                                 "Try setting `update_on_kvstore` to False when creating trainer.")
        self._optimizer.rescale_grad = self._scale / batch_size
        self._update(ignore_stale_grad)

    @torch.no_grad()
    def _update(self, ignore_stale_grad: bool = False) -> None:
        for i, p in enumerate(self._params):
            if p.grad is None:
                if ignore_stale_grad:
                    continue
                raise UserWarning(f"Gradient of Parameter `{self._names[i]}` has not been updated by backward since "
                                  "last `step`; call step with ignore_stale_grad=True to suppress this")
            if self._update_on_kvstore:
                self._kvstore.pull(i, out=p.data, priority=-i)
            else:
                self._updaters[0](i, p.grad, p.data)

    # ------------------------------------------------------------------ states
    def save_states(self, fname: str) -> None:
        if not self._kv_initialized:
            self._init_kvstore()
        if self._update_on_kvstore:
            self._kvstore.save_optimizer_states(fname, dump_optimizer=True)
        else:
            with open(fname, "wb") as f:
                f.write(self._updaters[0].get_states(dump_optimizer=True))

    def load_states(self, fname: str) -> None:
        if not self._kv_initialized:
            self._init_kvstore()
        if self._update_on_kvstore:
            self._kvstore.load_optimizer_states(fname)
            self._optimizer = self._kvstore._updater.optimizer
        else:
            with open(fname, "rb") as f:
                self._updaters[0].set_states(f.read())
            self._optimizer = self._updaters[0].optimizer
