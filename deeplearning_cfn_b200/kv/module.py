"""`Module` - the symbolic-API training wrapper (reference Y6: MX/python/mxnet/module/module.py,
base_module.py): `bind` -> `init_params` -> `init_optimizer(kvstore, optimizer, optimizer_params)` (rescale_grad =
1/(batch_size * num_workers) when the kvstore is `dist_sync`, module.py:474-556) -> `fit(train_data, num_epoch,
batch_end_callback, epoch_end_callback, ...)` whose inner loop is `forward_backward` + `update`
(base_module.py:516-529; `update` = `_update_params_on_kvstore` push/pull with priority -index, or pull grads +
local updater, model.py:145-176), plus `save_checkpoint` / `load` / `score`."""
from __future__ import annotations

import logging
import time

import torch
import torch.nn.functional as F

from ..ops.nn import softmax_cross_entropy
from . import metric as metric_mod
from . import optimizer as opt
from .kvstore import KVStore, create


def decide_slices(batch_size: int, work_load_list: list) -> list[slice]:
    """Batch rows per device, proportional to its workload (reference executor_group.py:281-305 `_split_input_slice`:
    rounded shares, the last device takes what is left; an empty slice is an error)."""
    total = float(sum(work_load_list))
    sizes = [int(round(batch_size * w / total)) for w in work_load_list]
    sizes[-1] = batch_size - sum(sizes[:-1])
    out, start = [], 0
    for n in sizes:
        if n <= 0:
            raise ValueError(f"batch of {batch_size} rows cannot be sliced over {len(work_load_list)} devices: too many slices")
        out.append(slice(start, start + n))
        start += n
    return out


class Module:
    """`context` may be ONE device or a LIST of devices.  With a list the module is the reference's
    DataParallelExecutorGroup (module/executor_group.py:143-572): one replica of the network per device, every batch sliced
    by `work_load_list`, forward / backward per slice, gradients of parameter i pushed as a per-device list (the kvstore
    reduces them), then either the updated weight is pulled into every replica (update_on_kvstore) or the summed gradient
    is pulled back and every device runs its own updater with state index `i * num_device + k` (model.py:157-176)."""

    def __init__(self, symbol: torch.nn.Module, data_names=("data",), label_names=("softmax_label",), context=None,
                 loss=None, work_load_list=None):
        self.symbol = symbol
        ctxs = list(context) if isinstance(context, (list, tuple)) else ([context] if context is not None else [])
        ctxs = [torch.device(c) for c in ctxs]
        self.device = ctxs[0] if ctxs else next(symbol.parameters()).device
        self._contexts = ctxs if len(ctxs) > 1 else [self.device]
        self._work_load = list(work_load_list) if work_load_list is not None else [1] * len(self._contexts)
        if len(self._work_load) != len(self._contexts):
            raise ValueError("work_load_list needs one entry per context")
        self.loss = loss or (lambda out, y: softmax_cross_entropy(out, y, reduction="sum"))
        self.binded = self.params_initialized = self.optimizer_initialized = False
        self._kv: KVStore | None = None
        self._update_on_kvstore = False
        if len(self._contexts) > 1:
            import copy
            self._replicas = [symbol.to(self._contexts[0])] + [copy.deepcopy(symbol).to(c) for c in self._contexts[1:]]
        else:
            self._replicas = [symbol]
        self._rparams = [[p for p in r.parameters() if p.requires_grad] for r in self._replicas]
        self._params = self._rparams[0]
        self._batch_size = None
        self._outputs = None

    @property
    def multi_device(self) -> bool:
        return len(self._replicas) > 1

    def _sync_replicas_from_first(self) -> None:
        sd = self.symbol.state_dict()
        for r in self._replicas[1:]:
            r.load_state_dict(sd)

    def bind(self, data_shapes, label_shapes=None, for_training: bool = True, **_):
        self._batch_size = data_shapes[0][1][0] if isinstance(data_shapes[0], tuple) else data_shapes[0]
        self.binded = True

    def init_params(self, initializer=None, arg_params=None, aux_params=None, allow_missing=False, force_init=False):
        if arg_params or aux_params:
            self.symbol.load_state_dict({**(arg_params or {}), **(aux_params or {})}, strict=not allow_missing)
        elif initializer is not None:
            from . import initializer as init_mod
            init = init_mod.create(initializer) if isinstance(initializer, str) else initializer
            if hasattr(init, "init_module"):
                init.init_module(self.symbol)              # by parameter name, like the reference's Initializer.__call__
            else:
                for p in self._params:
                    init(p)
        self._sync_replicas_from_first()
        self.params_initialized = True

    def init_optimizer(self, kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),),
                       force_init: bool = False):
        assert self.binded and self.params_initialized
        if kvstore is None and self.multi_device:
            kvstore = "device" if self.device.type == "cuda" else "local"     # several devices always aggregate (model.py:77-90)
        kv = create(kvstore) if isinstance(kvstore, str) else kvstore
        params = dict(optimizer_params)
        batch = self._batch_size or 1
        if kv is not None and "dist" in kv.type and "_sync" in kv.type:
            batch *= kv.num_workers
        params.setdefault("rescale_grad", 1.0 / batch)
        self._optimizer = opt.create(optimizer, **params) if isinstance(optimizer, str) else optimizer
        self._kv = kv
        # update_on_kvstore: true for distributed stores, false when there is nothing to aggregate (model.py:77-114)
        self._update_on_kvstore = kv is not None and "dist" in kv.type
        self._updater = opt.get_updater(self._optimizer)
        if kv is not None:
            for i, p in enumerate(self._params):
                kv.init(i, p.data)
            if self._update_on_kvstore:
                kv.set_optimizer(self._optimizer)
            for i, p in enumerate(self._params):
                kv.pull(i, out=[rp[i].data for rp in self._rparams] if self.multi_device else p.data, priority=-i)
        self.optimizer_initialized = True

    def _forward_slices(self, x, y, train: bool):
        """Slice the batch over the replicas, run each slice on its device; outputs merged on the first device."""
        slices = decide_slices(x.shape[0], self._work_load)
        outs, labels = [], []
        for r, (net, ctx, sl) in enumerate(zip(self._replicas, self._contexts, slices)):
            xr = x[sl].to(ctx, non_blocking=True)
            labels.append(y[sl].to(ctx, non_blocking=True) if y is not None else None)
            if train:
                net.zero_grad(set_to_none=True)
                outs.append(net(xr))
            else:
                with torch.no_grad():
                    outs.append(net(xr))
        self._slice_outputs, self._slice_labels = outs, labels
        self._outputs = torch.cat([o.detach().to(self.device) for o in outs])      # merge_multi_context
        return outs, labels

    def forward_backward(self, data_batch):
        x, y = data_batch
        if self.multi_device:
            outs, labels = self._forward_slices(x, y, True)
            losses = [self.loss(o, l) for o, l in zip(outs, labels)]
            for l in losses:
                l.backward()
            self._loss = sum(l.detach().to(self.device) for l in losses)
            return
        self.symbol.zero_grad(set_to_none=True)
        self._outputs = self.symbol(x)
        self._loss = self.loss(self._outputs, y)
        self._loss.backward()

    @torch.no_grad()
    def _update_multi_device(self):
        """model.py:145-176 with several executors: push the per-device gradient list of every parameter."""
        kv, nd = self._kv, len(self._replicas)
        for i in range(len(self._params)):
            grads = [rp[i].grad for rp in self._rparams]
            if any(g is None for g in grads):
                continue
            kv.push(i, grads, priority=-i)
            if self._update_on_kvstore:
                kv.pull(i, out=[rp[i].data for rp in self._rparams], priority=-i)
            else:
                kv.pull(i, out=grads, priority=-i)                       # every device receives the sum
                for k, rp in enumerate(self._rparams):
                    self._updater(i * nd + k, rp[i].grad, rp[i].data)

    @torch.no_grad()
    def update(self):
        kv = self._kv
        if self.multi_device:
            return self._update_multi_device()
        if kv is not None and "nccl" in kv.type:
            return self._update_params_on_kvstore_nccl()
        for i, p in enumerate(self._params):
            if p.grad is None:
                continue
            if kv is not None:
                kv.push(i, p.grad, priority=-i)
                if self._update_on_kvstore:
                    kv.pull(i, out=p.data, priority=-i)
                else:
                    kv.pull(i, out=p.grad, priority=-i)
                    self._updater(i, p.grad, p.data)
            else:
                self._updater(i, p.grad, p.data)

    def _update_params_on_kvstore_nccl(self):
        """nccl kvstore: keys go to the store in groups of MXNET_UPDATE_AGGREGATION_SIZE (default 16) so that one grouped
        launch carries many small tensors (reference model.py:125-143, kvstore_nccl.h:221-307)."""
        import os
        kv = self._kv
        size = int(os.environ.get("MXNET_UPDATE_AGGREGATION_SIZE", "16"))
        idx = [i for i, p in enumerate(self._params) if p.grad is not None]
        for s0 in range(0, len(idx), size):
            batch = idx[s0:s0 + size]
            kv.push(batch, [self._params[i].grad for i in batch], priority=-s0)
            if self._update_on_kvstore:
                kv.pull(batch, out=[self._params[i].data for i in batch], priority=-s0)
            else:
                kv.pull(batch, out=[self._params[i].grad for i in batch], priority=-s0)
                for i in batch:
                    self._updater(i, self._params[i].grad, self._params[i].data)

    def get_outputs(self, merge_multi_context: bool = True):
        """[merged output] - or, with merge_multi_context=False on several devices, [[output of device 0, device 1, ...]]."""
        if not merge_multi_context and self.multi_device:
            return [list(self._slice_outputs)]
        return [self._outputs]

    # ---- the finer-grained steps of BaseModule (base_module.py: forward / backward / update_metric / get_params ...)
    def forward(self, data_batch, is_train: bool | None = None) -> None:
        x, y = _unpack(data_batch)
        train = self.symbol.training if is_train is None else is_train
        for r in self._replicas:
            r.train(train)
        self._label = y
        if self.multi_device:
            self._forward_slices(x, y, train)
            return
        if train:
            self.symbol.zero_grad(set_to_none=True)
            self._outputs = self.symbol(x)
        else:
            with torch.no_grad():
                self._outputs = self.symbol(x)

    def backward(self, out_grads=None) -> None:
        if self.multi_device:
            if out_grads is not None:
                g = out_grads[0] if isinstance(out_grads, (list, tuple)) else out_grads
                for o, sl in zip(self._slice_outputs, decide_slices(g.shape[0], self._work_load)):
                    o.backward(g[sl].to(o.device))
            else:
                losses = [self.loss(o, l) for o, l in zip(self._slice_outputs, self._slice_labels)]
                for l in losses:
                    l.backward()
                self._loss = sum(l.detach().to(self.device) for l in losses)
            return
        if out_grads is not None:
            self._outputs.backward(out_grads[0] if isinstance(out_grads, (list, tuple)) else out_grads)
        else:
            self._loss = self.loss(self._outputs, self._label)
            self._loss.backward()

    def update_metric(self, eval_metric, labels, pre_sliced: bool = False) -> None:
        eval_metric.update(labels if isinstance(labels, (list, tuple)) else [labels], [self._outputs])

    def get_params(self):
        """(arg_params, aux_params): learnable parameters and auxiliary states (BatchNorm running statistics)."""
        arg = {n: p.detach() for n, p in self.symbol.named_parameters()}
        aux = {n: b.detach() for n, b in self.symbol.named_buffers()}
        return arg, aux

    def set_params(self, arg_params, aux_params=None, allow_missing: bool = False, force_init: bool = True,
                   allow_extra: bool = False) -> None:
        sd = {**(arg_params or {}), **(aux_params or {})}
        res = self.symbol.load_state_dict(sd, strict=False)
        if res.missing_keys and not allow_missing:
            raise RuntimeError(f"missing parameters: {res.missing_keys[:5]}")
        if res.unexpected_keys and not allow_extra:
            raise ValueError(f"unexpected parameters: {res.unexpected_keys[:5]}")
        self._sync_replicas_from_first()
        self.params_initialized = True

    def save_params(self, fname: str) -> None:
        arg, aux = self.get_params()
        sd = {f"arg:{k}": v.cpu() for k, v in arg.items()}
        sd.update({f"aux:{k}": v.cpu() for k, v in aux.items()})
        torch.save(sd, fname)

    def load_params(self, fname: str) -> None:
        raw = torch.load(fname, map_location="cpu")
        self.set_params({k[4:]: v for k, v in raw.items() if k.startswith("arg:")},
                        {k[4:]: v for k, v in raw.items() if k.startswith("aux:")})

    def install_monitor(self, mon) -> None:
        mon.install(self.symbol)

    def fit(self, train_data, eval_data=None, eval_metric="acc", epoch_end_callback=None, batch_end_callback=None,
            kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),), initializer=None,
            arg_params=None, aux_params=None, begin_epoch: int = 0, num_epoch: int | None = None, batch_size=None,
            validation_metric=None, monitor=None, eval_end_callback=None, eval_batch_end_callback=None):
        assert num_epoch is not None, "please specify number of epochs"
        if monitor is not None:
            monitor.install(self.symbol)
        if not self.binded:
            self.bind([("data", (batch_size or 1,))])
        if not self.params_initialized:
            self.init_params(initializer, arg_params, aux_params)
        if not self.optimizer_initialized:
            self.init_optimizer(kvstore, optimizer, optimizer_params)
        train_metric = metric_mod.create(eval_metric)
        for epoch in range(begin_epoch, num_epoch):
            tic = time.time()
            train_metric.reset()
            for r in self._replicas:
                r.train()
            if hasattr(train_data, "reset"):
                train_data.reset()
            it = train_data() if callable(train_data) else train_data
            for nbatch, batch in enumerate(it):
                x, y = _unpack(batch)
                if monitor is not None:
                    monitor.tic()
                self.forward_backward((x, y))
                self.update()
                train_metric.update([y], [self._outputs])
                if monitor is not None:
                    monitor.toc_print()
                for cb in _as_list(batch_end_callback):
                    cb(epoch, nbatch, train_metric)
            for name, val in train_metric.get_name_value():
                logging.info("Epoch[%d] Train-%s=%f", epoch, name, val)
            logging.info("Epoch[%d] Time cost=%.3f", epoch, time.time() - tic)
            for cb in _as_list(epoch_end_callback):
                cb(epoch, self.symbol, None, None)
            if eval_data is not None:
                res = self.score(eval_data, validation_metric or eval_metric, epoch=epoch,
                                 batch_end_callback=eval_batch_end_callback, score_end_callback=eval_end_callback)
                for name, val in res:
                    logging.info("Epoch[%d] Validation-%s=%f", epoch, name, val)

    @torch.no_grad()
    def score(self, eval_data, eval_metric="acc", num_batch: int | None = None, batch_end_callback=None,
              score_end_callback=None, reset: bool = True, epoch: int = 0):
        """Run `eval_data` through the network and return `eval_metric.get_name_value()` (base_module.py:171-240).
        For the default accuracy metric the plain float is also available as `score(...)[0][1]`."""
        self.symbol.eval()
        m = metric_mod.create(eval_metric)
        m.reset()
        if reset and hasattr(eval_data, "reset"):
            eval_data.reset()
        it = eval_data() if callable(eval_data) else eval_data
        for nbatch, batch in enumerate(it):
            if num_batch is not None and nbatch == num_batch:
                break
            x, y = _unpack(batch)
            m.update([y], [self.symbol(x)])
            for cb in _as_list(batch_end_callback):
                cb(epoch, nbatch, m)
        for cb in _as_list(score_end_callback):
            cb(epoch, m)
        return m.get_name_value()

    def predict(self, eval_data, num_batch: int | None = None, merge_batches: bool = True, reset: bool = True):
        """Forward only; outputs concatenated over batches with the padding of the last batch removed."""
        self.symbol.eval()
        outs = []
        if reset and hasattr(eval_data, "reset"):
            eval_data.reset()
        it = eval_data() if callable(eval_data) else eval_data
        with torch.no_grad():
            for nbatch, batch in enumerate(it):
                if num_batch is not None and nbatch == num_batch:
                    break
                x, _ = _unpack(batch)
                o = self.symbol(x)
                pad = getattr(batch, "pad", 0)
                outs.append(o[:o.shape[0] - pad] if pad else o)
        return torch.cat(outs) if merge_batches else outs

    def save_checkpoint(self, prefix: str, epoch: int, save_optimizer_states: bool = False):
        from ..utils.callback import save_checkpoint
        save_checkpoint(prefix, epoch, self.symbol)
        if save_optimizer_states:
            with open(f"{prefix}-{epoch:04d}.states", "wb") as f:
                f.write(self._updater.get_states(dump_optimizer=True))

    @staticmethod
    def load(prefix: str, epoch: int, symbol: torch.nn.Module, **kw) -> "Module":
        from ..utils.callback import load_checkpoint
        load_checkpoint(prefix, epoch, symbol)
        m = Module(symbol, **kw)
        m.params_initialized = True
        return m


def _as_list(x):
    if x is None:
        return []
    return list(x) if isinstance(x, (list, tuple)) else [x]


def _unpack(batch):
    """(x, y) tuple or an io.DataBatch."""
    if hasattr(batch, "data") and hasattr(batch, "label"):
        return batch.data[0], (batch.label[0] if batch.label else None)
    return batch[0], batch[1]
