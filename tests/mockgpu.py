"""A stand-in GPU for the host test tier.

Every Python wrapper around a CUDA kernel (`ops/*`, the trainer's fused optimizer ...) takes its CUDA branch only when its
tensors report `is_cuda`, which they never do on the CPU-only test box - so a wrong argument count, a tensor constructor
that does not exist, or a shape bug in such a branch would first show up on the GPU box.  `mock_gpu()` makes the host tier
walk those branches:

  * `torch.Tensor.is_cuda` reports True,
  * `native.lib()` returns a proxy whose `dlcfn_*` entries do NOT launch anything: they push every argument through the
    declared ctypes `argtypes` (count and convertibility, exactly what ctypes would check on the real call), log the call and
    return 0 (success),
  * `native.stream_ptr()` returns the null stream.

Outputs of the "kernels" are whatever `torch.empty` left in the buffers; the tests built on this check control flow, shapes,
dtypes, strides and which kernels a code path launches - numerics stay with the real GPU tests.
"""
from __future__ import annotations

import contextlib
from unittest import mock

import torch


class _Checker:
    def __init__(self, real, name, log):
        object.__setattr__(self, "real", real)
        object.__setattr__(self, "name", name)
        object.__setattr__(self, "log", log)

    def __setattr__(self, k, v):            # argtypes / restype declarations land on the real function object
        setattr(self.real, k, v)

    def __getattr__(self, k):
        return getattr(self.real, k)

    def __call__(self, *args):
        at = self.real.argtypes
        assert at is not None, f"{self.name}: called without an argtypes declaration"
        assert len(args) == len(at), f"{self.name}: {len(args)} arguments passed, {len(at)} declared"
        for i, (a, t) in enumerate(zip(args, at)):
            try:
                t.from_param(a)
            except Exception as e:        # noqa: BLE001
                raise AssertionError(f"{self.name}: argument {i} = {a!r} is not convertible to {t.__name__}: {e}") from e
        self.log.append(self.name)
        return 0


class _FakeLib:
    def __init__(self, real, log):
        object.__setattr__(self, "_real", real)
        object.__setattr__(self, "_log", log)
        object.__setattr__(self, "_cache", {})

    def __getattr__(self, name):
        real = getattr(self._real, name)
        if not name.startswith("dlcfn_"):
            return real
        if name not in self._cache:
            self._cache[name] = _Checker(real, name, self._log)
        return self._cache[name]

    def __setattr__(self, name, v):
        setattr(self._real, name, v)


@contextlib.contextmanager
def mock_gpu():
    """Yields the list of `dlcfn_*` functions "launched" inside the block, in order."""
    from deeplearning_cfn_b200.ops import native
    log: list[str] = []
    fake = _FakeLib(native.lib(), log)
    with mock.patch.object(native, "lib", lambda: fake), \
            mock.patch.object(native, "stream_ptr", lambda stream=None: 0), \
            mock.patch.object(torch.Tensor, "is_cuda", property(lambda self: True)):
        yield log
