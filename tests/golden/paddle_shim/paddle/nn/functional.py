import torch as _torch
import torch.nn.functional as _F

from .. import Tensor, _T


def softmax(x, axis=-1, dtype=None):
    return _T(_F.softmax(x, dim=axis))


def log_softmax(x, axis=-1):
    return _T(_F.log_softmax(x, dim=axis))


def glu(x, axis=-1):
    return _T(_F.glu(x, dim=axis))


def relu(x):
    return _T(_F.relu(x))


def pad(x, pad, mode="constant", value=0.0, data_format="NCHW"):
    """paddle.nn.functional.pad: `pad` lists (left, right) pairs starting from the LAST spatial dimension; for channel-last
    layouts ('NLC') the spatial dimension sits before the channels."""
    pad = [int(v) for v in (pad.tolist() if isinstance(pad, _torch.Tensor) else pad)]
    if data_format in ("NLC", "NHWC") and x.dim() == 3:
        return _T(_F.pad(x.transpose(1, 2) if not isinstance(x, Tensor) else _torch.Tensor.transpose(x, 1, 2), pad, mode=mode,
                         value=value).transpose(1, 2))
    if len(pad) == 2 * x.dim():  # full form: pairs from the FIRST dimension
        pairs = [pad[2 * i:2 * i + 2] for i in range(x.dim())][::-1]
        pad = [v for pr in pairs for v in pr]
    return _T(_F.pad(x, pad, mode=mode, value=value))
