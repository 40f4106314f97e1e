import torch as _torch
import torch.nn.functional as _F

from .. import Tensor, _T, float32
from . import functional, initializer  # noqa: F401


class Layer(_torch.nn.Module):
    def create_parameter(self, shape, attr=None, dtype=None, is_bias=False, default_initializer=None):
        return _torch.nn.Parameter(_torch.zeros([int(s) for s in shape], dtype=float32), requires_grad=False)

    def add_parameter(self, name, p):
        self.register_parameter(name, p)

    def sublayers(self):
        return list(self.modules())[1:]

    def set_state_dict(self, sd):
        own = self.state_dict()
        missing = [k for k in own if k not in sd]
        extra = [k for k in sd if k not in own]
        assert not missing and not extra, (missing[:5], extra[:5])
        with _torch.no_grad():
            for k, v in own.items():
                v.copy_(_torch.as_tensor(sd[k]).reshape(v.shape))

    def __call__(self, *a, **k):
        out = super().__call__(*a, **k)
        return out


Module = Layer
LayerList = _torch.nn.ModuleList
Sequential = _torch.nn.Sequential


class Identity(Layer):
    def forward(self, x):
        return x


class Dropout(Layer):
    def __init__(self, p=0.5, *a, **k):
        super().__init__()

    def forward(self, x):
        return x


class _Act(Layer):
    fn = None

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x):
        return _T(type(self).fn(x))


class ReLU(_Act):
    fn = staticmethod(_F.relu)


class Swish(_Act):
    fn = staticmethod(_F.silu)


Silu = Swish


class Tanh(_Act):
    fn = staticmethod(_torch.tanh)


class GELU(_Act):
    fn = staticmethod(_F.gelu)


class SELU(_Act):
    fn = staticmethod(_F.selu)


class ELU(_Act):
    fn = staticmethod(_F.elu)


class LeakyReLU(_Act):
    fn = staticmethod(_F.leaky_relu)


class ReLU6(_Act):
    fn = staticmethod(_F.relu6)


class Hardtanh(_Act):
    fn = staticmethod(_F.hardtanh)


class Hardswish(_Act):
    fn = staticmethod(_F.hardswish)


class Hardshrink(_Act):
    fn = staticmethod(_F.hardshrink)


class Linear(Layer):
    """paddle.nn.Linear: weight [in, out], y = x W + b."""

    def __init__(self, in_features, out_features, weight_attr=None, bias_attr=None, name=None):
        super().__init__()
        self.weight = self.create_parameter([in_features, out_features])
        self.bias = None if bias_attr is False else self.create_parameter([out_features])

    def forward(self, x):
        y = _torch.matmul(x, self.weight)
        return _T(y if self.bias is None else y + self.bias)


class _ConvNd(Layer):
    nd = 1

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, padding_mode="zeros",
                 weight_attr=None, bias_attr=None, data_format=None):
        super().__init__()
        k = (kernel_size,) * self.nd if isinstance(kernel_size, int) else tuple(kernel_size)
        self.weight = self.create_parameter([out_channels, in_channels // groups, *k])
        self.bias = None if bias_attr is False else self.create_parameter([out_channels])
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups

    def forward(self, x):
        f = _F.conv1d if self.nd == 1 else _F.conv2d
        return _T(f(x, self.weight, self.bias, stride=self.stride, padding=self.padding, dilation=self.dilation, groups=self.groups))


class Conv1D(_ConvNd):
    nd = 1


class Conv2D(_ConvNd):
    nd = 2


class LayerNorm(Layer):
    def __init__(self, normalized_shape, epsilon=1e-05, weight_attr=None, bias_attr=None, name=None):
        super().__init__()
        shp = [normalized_shape] if isinstance(normalized_shape, int) else list(normalized_shape)
        self.shp, self.eps = shp, epsilon
        self.weight = self.create_parameter(shp)
        self.bias = self.create_parameter(shp)

    def forward(self, x):
        return _T(_F.layer_norm(x, self.shp, self.weight, self.bias, self.eps))


class BatchNorm1D(Layer):
    def __init__(self, num_features, momentum=0.9, epsilon=1e-05, weight_attr=None, bias_attr=None, data_format="NCL", name=None):
        super().__init__()
        self.eps = epsilon
        self.weight = self.create_parameter([num_features])
        self.bias = self.create_parameter([num_features])
        self.register_buffer("_mean", _torch.zeros(num_features))
        self.register_buffer("_variance", _torch.ones(num_features))

    def forward(self, x):  # inference statistics
        return _T(_F.batch_norm(x, self._mean, self._variance, self.weight, self.bias, training=False, eps=self.eps))


class AvgPool1D(Layer):
    def __init__(self, kernel_size, stride=None, padding=0, exclusive=True, ceil_mode=False, name=None):
        super().__init__()
        self.k, self.s, self.p, self.ex, self.ceil = kernel_size, stride or kernel_size, padding, exclusive, ceil_mode

    def forward(self, x):
        return _T(_F.avg_pool1d(x, self.k, self.s, self.p, ceil_mode=self.ceil, count_include_pad=not self.ex))


class _RNN(Layer):
    kind = None

    def __init__(self, input_size, hidden_size, num_layers=1, direction="forward", time_major=False, dropout=0.0, **k):
        super().__init__()
        assert num_layers == 1 and not time_major
        self.bi = direction in ("bidirect", "bidirectional")
        # kept out of the module tree so that only paddle's flat names (weight_ih_l0[_reverse], ...) appear in state_dict()
        object.__setattr__(self, "_impl", self.kind(input_size, hidden_size, 1, batch_first=True, bidirectional=self.bi))
        for n, p_ in list(self._impl.named_parameters()):
            p_.requires_grad_(False)
            self.register_parameter(n, p_)

    def forward(self, x, initial_states=None, sequence_length=None):
        lens = sequence_length.to(_torch.int64).cpu() if sequence_length is not None else _torch.full((x.shape[0],), x.shape[1])
        pk = _torch.nn.utils.rnn.pack_padded_sequence(x, lens, batch_first=True, enforce_sorted=False)
        out, st = self._impl(pk, initial_states)
        out, _ = _torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=x.shape[1])
        st = tuple(_T(s) for s in st) if isinstance(st, tuple) else _T(st)
        return _T(out), st


class LSTM(_RNN):
    kind = _torch.nn.LSTM


class GRU(_RNN):
    kind = _torch.nn.GRU


class CTCLoss(Layer):
    def __init__(self, *a, **k):
        super().__init__()


class Embedding(Layer):
    def __init__(self, num_embeddings, embedding_dim, padding_idx=None, sparse=False, weight_attr=None, name=None):
        super().__init__()
        self.weight = self.create_parameter([num_embeddings, embedding_dim])

    def forward(self, x):
        return _T(_F.embedding(x, self.weight))


class KLDivLoss(Layer):  # constructed by the training-only LabelSmoothingLoss; never called on the inference path
    def __init__(self, *a, **k):
        super().__init__()
