"""Minimal torch-backed stand-in for the `paddle` package (see ../README.md)."""
import builtins

import numpy as _np
import torch as _torch

bool = _torch.bool
int32 = _torch.int32
int64 = _torch.int64
float32 = _torch.float32
float64 = _torch.float64
float16 = _torch.float16
_DT = {"bool": bool, "int32": int32, "int64": int64, "float32": float32, "float64": float64, "int": int64, "float": float32}


def _dt(d, default=None):
    if d is None:
        return default
    return _DT[d] if isinstance(d, str) else d


class Tensor(_torch.Tensor):
    @staticmethod
    def __new__(cls, data):
        return _torch.Tensor._make_subclass(cls, data.detach() if isinstance(data, _torch.Tensor) else _torch.as_tensor(data))

    # ---- paddle-style methods ----
    def transpose(self, perm, *more):
        if isinstance(perm, (list, tuple)):
            return self.permute(*perm)
        return _torch.Tensor.transpose(self, perm, *more)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (list, tuple)):
            shape = list(shape[0])
        shape = [int(self.shape[i]) if s == 0 else int(s) for i, s in enumerate(shape)]  # paddle: 0 = keep this dim
        return _torch.Tensor.reshape(self, shape)

    def equal(self, other):
        return self == other

    def astype(self, dtype):
        return self.to(_dt(dtype))

    cast = astype

    def numpy(self):
        return self.detach().as_subclass(_torch.Tensor).numpy()

    def broadcast_to(self, shape):
        return _torch.Tensor.expand(self, *[int(s) for s in shape])

    def expand(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (list, tuple)):
            shape = shape[0]
        return _torch.Tensor.expand(self, *[int(s) for s in shape])

    def sum(self, axis=None, dtype=None, keepdim=False):
        return _torch.Tensor.sum(self) if axis is None else _torch.Tensor.sum(self, dim=axis, keepdim=keepdim)

    def all(self, axis=None, keepdim=False):
        return _torch.Tensor.all(self) if axis is None else _torch.Tensor.all(self, dim=axis, keepdim=keepdim)

    def max(self, axis=None, keepdim=False):
        return _torch.Tensor.max(self) if axis is None else _torch.Tensor.max(self, dim=axis, keepdim=keepdim).values

    def squeeze(self, axis=None):
        return _torch.Tensor.squeeze(self) if axis is None else _torch.Tensor.squeeze(self, axis)

    def split(self, num_or_sections, axis=0):
        return split(self, num_or_sections, axis)

    def repeat_interleave(self, repeats, axis=None):
        return _torch.Tensor.repeat_interleave(self, repeats, dim=axis)

    def tile(self, repeat_times):
        return _torch.Tensor.repeat(self, *repeat_times)

    @property
    def stop_gradient(self):
        return True

    @stop_gradient.setter
    def stop_gradient(self, v):
        pass


def _T(x):
    return x if isinstance(x, Tensor) else Tensor(x)


def to_tensor(data, dtype=None, place=None, stop_gradient=True):
    if isinstance(data, _torch.Tensor):
        t = data
    else:
        a = _np.asarray(data)
        if a.dtype == _np.float64 and dtype is None:
            a = a.astype(_np.float32)
        t = _torch.from_numpy(_np.ascontiguousarray(a)) if a.ndim else _torch.tensor(a.item(), dtype=_torch.from_numpy(a.reshape(1)).dtype)
    if dtype is not None:
        t = t.to(_dt(dtype))
    return _T(t)


def _shape(s):
    return [int(v) for v in (s.tolist() if isinstance(s, _torch.Tensor) else s)]


def zeros(shape, dtype=None):
    return _T(_torch.zeros(_shape(shape), dtype=_dt(dtype, float32)))


def ones(shape, dtype=None):
    return _T(_torch.ones(_shape(shape), dtype=_dt(dtype, float32)))


def empty(shape, dtype=None):
    return _T(_torch.zeros(_shape(shape), dtype=_dt(dtype, float32)))


def full(shape, fill_value, dtype=None):
    d = _dt(dtype, None)
    if d is None:
        d = float32 if isinstance(fill_value, float) else (bool if isinstance(fill_value, builtins.bool) else int64)
    return _T(_torch.full(_shape(shape), fill_value, dtype=d))


def full_like(x, fill_value, dtype=None):
    return _T(_torch.full_like(x, fill_value, dtype=_dt(dtype, None)))


def arange(start=0, end=None, step=1, dtype=None):
    if end is None:
        start, end = 0, start
    d = _dt(dtype, None)
    if d is None:
        d = int64 if all(isinstance(v, int) for v in (start, end, step)) else float32
    return _T(_torch.arange(start, end, step, dtype=d))


def concat(x, axis=0):
    return _T(_torch.cat(list(x), dim=int(axis)))


def stack(x, axis=0):
    return _T(_torch.stack(list(x), dim=axis))


def split(x, num_or_sections, axis=0):
    if isinstance(num_or_sections, int):
        return [_T(t) for t in _torch.chunk(x, num_or_sections, dim=axis)]
    return [_T(t) for t in _torch.split(x, list(num_or_sections), dim=axis)]


def matmul(x, y, transpose_x=False, transpose_y=False):
    if transpose_x:
        x = _torch.Tensor.transpose(x, -1, -2)
    if transpose_y:
        y = _torch.Tensor.transpose(y, -1, -2)
    return _T(_torch.matmul(x, y))


def where(cond, x, y):
    return _T(_torch.where(cond, x, y))


def shape(x):
    return _T(_torch.tensor(list(x.shape), dtype=_torch.int32))


def squeeze(x, axis=None):
    return _T(x).squeeze(axis)


def unsqueeze(x, axis):
    return _T(_torch.unsqueeze(x, axis))


def repeat_interleave(x, repeats, axis=None):
    return _T(_torch.repeat_interleave(x, repeats, dim=axis))


def tril(x, diagonal=0):
    return _T(_torch.tril(x, diagonal))


def flip(x, axis):
    return _T(_torch.flip(x, axis if isinstance(axis, (list, tuple)) else [axis]))


def argmax(x, axis=None, keepdim=False):
    return _T(_torch.argmax(x, dim=axis, keepdim=keepdim))


def sum(x, axis=None, keepdim=False):
    return _T(x).sum(axis, keepdim=keepdim)


def randint(low, high=None, shape=(1,), dtype=None):
    return _T(_torch.randint(low, high, _shape(shape)))


exp = lambda x: _T(_torch.exp(x))
sin = lambda x: _T(_torch.sin(x))
cos = lambda x: _T(_torch.cos(x))
log = lambda x: _T(_torch.log(x))
no_grad = _torch.no_grad


class ParamAttr:
    def __init__(self, *a, **k):
        pass


class _Stub:
    def __getattr__(self, name):
        return lambda *a, **k: (a[0] if a and callable(a[0]) else None)


static = _Stub()
jit = _Stub()
static.InputSpec = lambda *a, **k: None

from . import nn  # noqa: E402,F401
