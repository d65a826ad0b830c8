"""Parameter initializers with the `mx.init` names (reference MX/python/mxnet/initializer.py): callables
`init(name, tensor)` (or `init(tensor)`), selected by name pattern the way the reference's `Initializer.__call__` does -
`*_bias` / `*beta` / moving means -> 0, `*gamma` / moving variances -> 1, everything else by the concrete rule - plus `Mixed`,
`Load` and the `create('xavier', ...)` registry the `--initializer` flag of the symbolic example uses (fit.py:200-215)."""
from __future__ import annotations

import math
import re

import torch

_REGISTRY: dict = {}


def register(cls):
    _REGISTRY[cls.__name__.lower()] = cls
    return cls


def _fans(t: torch.Tensor):
    if t.dim() < 2:
        return t.numel(), t.numel()
    rf = 1
    for d in t.shape[2:]:
        rf *= d
    return t.shape[1] * rf, t.shape[0] * rf          # fan_in, fan_out (weights are [out, in, ...])


class Initializer:
    def __init__(self, **kwargs):
        self._kwargs = kwargs

    def __call__(self, name, tensor: torch.Tensor | None = None):
        if tensor is None:                   # init(tensor)
            name, tensor = "weight", name
        name = str(name)
        with torch.no_grad():
            if name.endswith("bias") or name.endswith("beta") or name.endswith("running_mean") or name.endswith("moving_mean"):
                tensor.zero_()
            elif name.endswith("gamma") or name.endswith("running_var") or name.endswith("moving_var") or (
                    name.endswith("weight") and tensor.dim() == 1):          # BatchNorm / LayerNorm scale
                tensor.fill_(1.0)
            elif name.endswith("num_batches_tracked"):
                pass
            else:
                self._init_weight(name, tensor)
        return tensor

    def _init_weight(self, name, t):
        raise NotImplementedError

    def init_module(self, module: torch.nn.Module) -> None:
        """Initialise every parameter and buffer of `module` by its qualified name."""
        for n, p in module.named_parameters():
            self(n, p.data)
        for n, b in module.named_buffers():
            if b.dtype.is_floating_point:
                self(n, b)

    def dumps(self) -> str:
        import json
        return json.dumps([type(self).__name__.lower(), self._kwargs])


@register
class Zero(Initializer):
    def _init_weight(self, _, t):
        t.zero_()


@register
class One(Initializer):
    def _init_weight(self, _, t):
        t.fill_(1.0)


@register
class Constant(Initializer):
    def __init__(self, value=0.0):
        super().__init__(value=value)
        self.value = value

    def _init_weight(self, _, t):
        t.fill_(self.value)


@register
class Uniform(Initializer):
    def __init__(self, scale=0.07):
        super().__init__(scale=scale)
        self.scale = scale

    def _init_weight(self, _, t):
        t.uniform_(-self.scale, self.scale)


@register
class Normal(Initializer):
    def __init__(self, sigma=0.01):
        super().__init__(sigma=sigma)
        self.sigma = sigma

    def _init_weight(self, _, t):
        t.normal_(0.0, self.sigma)


@register
class Orthogonal(Initializer):
    def __init__(self, scale=1.414, rand_type="uniform"):
        super().__init__(scale=scale, rand_type=rand_type)
        self.scale, self.rand_type = scale, rand_type

    def _init_weight(self, _, t):
        rows, cols = t.shape[0], t.numel() // t.shape[0]
        a = torch.empty(rows, cols).uniform_(-1, 1) if self.rand_type == "uniform" else torch.randn(rows, cols)
        u, _s, vh = torch.linalg.svd(a, full_matrices=False)
        q = u if u.shape == (rows, cols) else vh
        t.copy_((self.scale * q).reshape(t.shape).to(t.dtype))


@register
class Xavier(Initializer):
    """rnd_type uniform|gaussian, factor_type avg|in|out, magnitude (initializer.py Xavier)."""

    def __init__(self, rnd_type="uniform", factor_type="avg", magnitude=3.0):
        super().__init__(rnd_type=rnd_type, factor_type=factor_type, magnitude=magnitude)
        self.rnd_type, self.factor_type, self.magnitude = rnd_type, factor_type, float(magnitude)

    def _init_weight(self, name, t):
        if t.dim() < 2:
            raise ValueError(f"Xavier initializer cannot be applied to vector {name}; it needs at least 2 dimensions")
        fan_in, fan_out = _fans(t)
        factor = {"avg": (fan_in + fan_out) / 2.0, "in": fan_in, "out": fan_out}[self.factor_type]
        scale = math.sqrt(self.magnitude / factor)
        if self.rnd_type == "uniform":
            t.uniform_(-scale, scale)
        elif self.rnd_type == "gaussian":
            t.normal_(0.0, scale)
        else:
            raise ValueError("unknown random type")


@register
class MSRAPrelu(Xavier):
    def __init__(self, factor_type="avg", slope=0.25):
        super().__init__("gaussian", factor_type, 2.0 / (1 + slope ** 2))
        self._kwargs = {"factor_type": factor_type, "slope": slope}


@register
class Bilinear(Initializer):
    """Bilinear up-sampling kernels for deconvolution weights."""

    def _init_weight(self, _, t):
        kh, kw = t.shape[-2], t.shape[-1]
        f = math.ceil(kw / 2.0)
        c = (2 * f - 1 - f % 2) / (2.0 * f)
        ys = 1 - (torch.arange(kh, dtype=torch.float32) / f - c).abs()
        xs = 1 - (torch.arange(kw, dtype=torch.float32) / f - c).abs()
        t.copy_((ys[:, None] * xs[None, :]).expand_as(t).to(t.dtype))


class Load:
    """Initialise from saved parameters (dict or a `.params` file written by save_checkpoint), fall back to `default_init`."""

    def __init__(self, param, default_init: Initializer | None = None, verbose: bool = False):
        if isinstance(param, str):
            raw = torch.load(param, map_location="cpu")
            param = {(k[4:] if k.startswith(("arg:", "aux:")) else k): v for k, v in raw.items()}
        self.param, self.default_init, self.verbose = param, default_init, verbose

    def __call__(self, name, tensor):
        if name in self.param:
            if tuple(self.param[name].shape) != tuple(tensor.shape):
                raise ValueError(f"Parameter {name} cannot be initialized from loading: shape mismatch")
            with torch.no_grad():
                tensor.copy_(self.param[name].to(tensor.dtype))
        elif self.default_init is not None:
            self.default_init(name, tensor)
        else:
            raise ValueError(f"Cannot initialize {name}: not found in the loaded params and no default initializer")
        return tensor

    init_module = Initializer.init_module


class Mixed:
    """patterns[i] (regex) -> initializers[i]; the first match wins (initializer.py Mixed)."""

    def __init__(self, patterns, initializers):
        assert len(patterns) == len(initializers)
        self.map = [(re.compile(p), i) for p, i in zip(patterns, initializers)]

    def __call__(self, name, tensor):
        for prog, init in self.map:
            if prog.match(str(name)):
                return init(name, tensor)
        raise ValueError(f"Parameter name {name} did not match any pattern; add a '.*' pattern at the end")

    init_module = Initializer.init_module


def create(name, **kwargs):
    if isinstance(name, (Initializer, Mixed, Load)):
        return name
    try:
        return _REGISTRY[str(name).lower()](**kwargs)
    except KeyError:
        raise ValueError(f"unknown initializer {name!r}; known: {sorted(_REGISTRY)}")
