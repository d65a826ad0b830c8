"""Evaluation metrics (reference `mx.metric`, MX/python/mxnet/metric.py): the objects that `Module.fit`,
`Module.score`, `Speedometer` and the example scripts pass around as `eval_metric`.

    m = metric.create("acc")                 # or ["acc", "ce"], or metric.TopKAccuracy(5), or a callable
    m.update(labels, preds); m.get() -> ("accuracy", 0.93); m.get_name_value(); m.reset()

Sums are accumulated on the device that holds the predictions and only read back in `get()` - `update()` never
synchronises the stream, so metrics can be updated every batch inside an asynchronous training loop.
"""
from __future__ import annotations

import math

import torch

_REGISTRY: dict = {}


def register(*names):
    def deco(cls):
        for n in names or (cls.__name__.lower(),):
            _REGISTRY[n.lower()] = cls
        return cls
    return deco


def _as_list(x):
    if x is None:
        return []
    return list(x) if isinstance(x, (list, tuple)) else [x]


class EvalMetric:
    def __init__(self, name: str, output_names=None, label_names=None):
        self.name = name
        self.output_names, self.label_names = output_names, label_names
        self.reset()

    def reset(self) -> None:
        self.num_inst = 0
        self.sum_metric = 0.0          # python float or 0-dim tensor on the prediction's device

    def update(self, labels, preds) -> None:
        raise NotImplementedError

    def update_dict(self, label: dict, pred: dict) -> None:
        preds = [pred[n] for n in self.output_names] if self.output_names else list(pred.values())
        labels = [label[n] for n in self.label_names] if self.label_names else list(label.values())
        self.update(labels, preds)

    def _value(self) -> float:
        s = self.sum_metric
        return float(s.item() if isinstance(s, torch.Tensor) else s)

    def get(self):
        if self.num_inst == 0:
            return self.name, float("nan")
        return self.name, self._value() / self.num_inst

    def get_name_value(self):
        name, value = self.get()
        return list(zip(_as_list(name), _as_list(value)))

    def get_config(self) -> dict:
        return {"metric": type(self).__name__, "name": self.name}

    def __str__(self):
        return f"EvalMetric: {dict(self.get_name_value())}"


@register("acc", "accuracy")
class Accuracy(EvalMetric):
    def __init__(self, axis: int = -1, name: str = "accuracy", **kw):
        self.axis = axis
        super().__init__(name, **kw)

    def update(self, labels, preds) -> None:
        for y, p in zip(_as_list(labels), _as_list(preds)):
            p = p.detach()
            if p.dim() > y.dim():
                p = p.argmax(self.axis)
            y = y.detach().to(p.device).reshape(-1)
            self.sum_metric = self.sum_metric + (p.reshape(-1).to(torch.int64) == y.to(torch.int64)).sum()
            self.num_inst += y.numel()


@register("top_k_accuracy", "top_k_acc")
class TopKAccuracy(EvalMetric):
    def __init__(self, top_k: int = 1, name: str = "top_k_accuracy", **kw):
        assert top_k > 1, "use Accuracy for top_k = 1"
        self.top_k = top_k
        super().__init__(f"{name}_{top_k}", **kw)

    def update(self, labels, preds) -> None:
        for y, p in zip(_as_list(labels), _as_list(preds)):
            p = p.detach().float()
            k = min(self.top_k, p.shape[-1])
            top = p.topk(k, dim=-1).indices
            y = y.detach().to(p.device).reshape(-1, 1).to(torch.int64)
            self.sum_metric = self.sum_metric + (top.reshape(-1, k) == y).any(-1).sum()
            self.num_inst += y.numel()


class _Regression(EvalMetric):
    def _err(self, y, p):
        raise NotImplementedError

    def update(self, labels, preds) -> None:
        for y, p in zip(_as_list(labels), _as_list(preds)):
            p = p.detach().float()
            y = y.detach().to(p.device).float().reshape(p.shape)
            self.sum_metric = self.sum_metric + self._err(y, p)
            self.num_inst += 1


@register("mae")
class MAE(_Regression):
    def __init__(self, name: str = "mae", **kw):
        super().__init__(name, **kw)

    def _err(self, y, p):
        return (y - p).abs().mean()


@register("mse")
class MSE(_Regression):
    def __init__(self, name: str = "mse", **kw):
        super().__init__(name, **kw)

    def _err(self, y, p):
        return ((y - p) ** 2).mean()


@register("rmse")
class RMSE(_Regression):
    def __init__(self, name: str = "rmse", **kw):
        super().__init__(name, **kw)

    def _err(self, y, p):
        return ((y - p) ** 2).mean().sqrt()


@register("ce", "cross-entropy", "crossentropy")
class CrossEntropy(EvalMetric):
    """-log p[label]; `preds` are probabilities (the reference feeds SoftmaxOutput) unless from_logits=True."""

    def __init__(self, eps: float = 1e-12, name: str = "cross-entropy", from_logits: bool = False, **kw):
        self.eps, self.from_logits = eps, from_logits
        super().__init__(name, **kw)

    def update(self, labels, preds) -> None:
        for y, p in zip(_as_list(labels), _as_list(preds)):
            p = p.detach().float().reshape(-1, p.shape[-1])
            y = y.detach().to(p.device).reshape(-1).to(torch.int64)
            if self.from_logits:
                p = torch.softmax(p, -1)
            prob = p.gather(1, y[:, None]).reshape(-1)
            self.sum_metric = self.sum_metric + (-(prob + self.eps).log()).sum()
            self.num_inst += y.numel()


@register("nll_loss", "nll-loss")
class NegativeLogLikelihood(CrossEntropy):
    def __init__(self, eps: float = 1e-12, name: str = "nll-loss", **kw):
        super().__init__(eps=eps, name=name, **kw)


@register("perplexity")
class Perplexity(CrossEntropy):
    def __init__(self, ignore_label: int | None = None, name: str = "perplexity", **kw):
        self.ignore_label = ignore_label
        super().__init__(name=name, **kw)

    def update(self, labels, preds) -> None:
        for y, p in zip(_as_list(labels), _as_list(preds)):
            p = p.detach().float().reshape(-1, p.shape[-1])
            y = y.detach().to(p.device).reshape(-1).to(torch.int64)
            if self.from_logits:
                p = torch.softmax(p, -1)
            keep = torch.ones_like(y, dtype=torch.bool) if self.ignore_label is None else y != self.ignore_label
            prob = p.gather(1, y.clamp(min=0)[:, None]).reshape(-1)
            self.sum_metric = self.sum_metric + (-(prob + self.eps).log() * keep).sum()
            self.num_inst += int(keep.sum())

    def get(self):
        if self.num_inst == 0:
            return self.name, float("nan")
        return self.name, math.exp(self._value() / self.num_inst)


@register("f1")
class F1(EvalMetric):
    """Binary F1 of class 1; average='macro' averages per-batch F1 (reference default), 'micro' pools the counts."""

    def __init__(self, name: str = "f1", average: str = "macro", **kw):
        self.average = average
        super().__init__(name, **kw)

    def reset(self) -> None:
        super().reset()
        self.tp = self.fp = self.fn = 0.0

    def update(self, labels, preds) -> None:
        for y, p in zip(_as_list(labels), _as_list(preds)):
            p = p.detach()
            if p.dim() > y.dim():
                if p.shape[-1] != 2:
                    raise ValueError("F1 currently only supports binary classification")
                p = p.argmax(-1)
            y = y.detach().to(p.device).reshape(-1)
            p = p.reshape(-1)
            tp = float(((p == 1) & (y == 1)).sum())
            fp = float(((p == 1) & (y == 0)).sum())
            fn = float(((p == 0) & (y == 1)).sum())
            if self.average == "macro":
                prec = tp / (tp + fp) if tp + fp else 0.0
                rec = tp / (tp + fn) if tp + fn else 0.0
                self.sum_metric += 2 * prec * rec / (prec + rec) if prec + rec else 0.0
                self.num_inst += 1
            else:
                self.tp += tp
                self.fp += fp
                self.fn += fn
                prec = self.tp / (self.tp + self.fp) if self.tp + self.fp else 0.0
                rec = self.tp / (self.tp + self.fn) if self.tp + self.fn else 0.0
                self.sum_metric = 2 * prec * rec / (prec + rec) if prec + rec else 0.0
                self.num_inst = 1


@register("loss")
class Loss(EvalMetric):
    """Mean of whatever the network outputs as its loss (labels ignored)."""

    def __init__(self, name: str = "loss", **kw):
        super().__init__(name, **kw)

    def update(self, _labels, preds) -> None:
        for p in _as_list(preds):
            p = p.detach().float()
            self.sum_metric = self.sum_metric + p.sum()
            self.num_inst += p.numel()


class CustomMetric(EvalMetric):
    """feval(label, pred) -> value or (sum, count); `metric.np(f)` in the reference."""

    def __init__(self, feval, name: str | None = None, allow_extra_outputs: bool = False, **kw):
        self._feval, self._allow_extra = feval, allow_extra_outputs
        super().__init__(name or f"custom({getattr(feval, '__name__', 'feval')})", **kw)

    def update(self, labels, preds) -> None:
        labels, preds = _as_list(labels), _as_list(preds)
        if not self._allow_extra and len(labels) != len(preds):
            raise ValueError(f"{len(labels)} labels vs {len(preds)} predictions")
        for y, p in zip(labels, preds):
            r = self._feval(y, p)
            if isinstance(r, tuple):
                self.sum_metric += r[0]
                self.num_inst += r[1]
            else:
                self.sum_metric += float(r)
                self.num_inst += 1


def np(feval, name: str | None = None, allow_extra_outputs: bool = False) -> CustomMetric:   # noqa: A001
    return CustomMetric(feval, name, allow_extra_outputs)


@register("composite")
class CompositeEvalMetric(EvalMetric):
    def __init__(self, metrics=None, name: str = "composite", **kw):
        self.metrics = [create(m) for m in (metrics or [])]
        super().__init__(name, **kw)

    def add(self, metric) -> None:
        self.metrics.append(create(metric))

    def get_metric(self, index: int) -> EvalMetric:
        return self.metrics[index]

    def reset(self) -> None:
        for m in getattr(self, "metrics", []):
            m.reset()

    def update(self, labels, preds) -> None:
        for m in self.metrics:
            m.update(labels, preds)

    def update_dict(self, label, pred) -> None:
        for m in self.metrics:
            m.update_dict(label, pred)

    def get(self):
        names, values = [], []
        for m in self.metrics:
            n, v = m.get()
            names += _as_list(n)
            values += _as_list(v)
        return names, values


def create(metric, *args, **kwargs) -> EvalMetric:
    """name | EvalMetric | callable | list of those (-> CompositeEvalMetric)."""
    if isinstance(metric, EvalMetric):
        return metric
    if callable(metric):
        return CustomMetric(metric, *args, **kwargs)
    if isinstance(metric, (list, tuple)):
        return CompositeEvalMetric(list(metric))
    try:
        return _REGISTRY[str(metric).lower()](*args, **kwargs)
    except KeyError:
        raise ValueError(f"unknown metric {metric!r}; known: {sorted(_REGISTRY)}")
