"""Parameter containers for the hot path: synthetic initialisation and file I/O.

Parameters are kept as a dict  name -> float32 numpy array  using the reference's own (Paddle)
parameter names and layouts (`Linear.weight` is [in, out], `Conv*.weight` is [out, in/groups, k...];
SURVEY.md Appendix A), i.e. exactly what `paddle.load("model.pdparams")` yields after `.numpy()`.
The C library repacks them (ppasr_b200_finalize).

No pretrained PPASR weights are reachable offline, so tests and the benchmark use seeded synthetic
weights drawn with the reference's initialisers (ppasr/model_utils/utils/base.py:58-138: Kaiming-uniform
with negative_slope sqrt(5) => U(-1/sqrt(fan_in), 1/sqrt(fan_in)); LayerNorm 1/0; pos_bias_u/v
Xavier-uniform, conformer/attention.py:193-196). The CTC projection is scaled by `ctc_gain` so the
posteriors are peaked (arg-max margins far above bf16 noise), as documented in SURVEY.md §8d.
"""
import json
import math
import os
import pickle
from typing import Dict

import numpy as np


class ConformerConfig:
    """Inference-relevant keys of configs/conformer.yml (`encoder_conf`, `streaming`, n_mels)."""

    def __init__(self, input_dim=80, vocab_size=4233, output_size=256, attention_heads=4, linear_units=2048,
                 num_blocks=12, cnn_module_kernel=15, streaming=True, cnn_module_norm="layer_norm", max_len=5000,
                 **_ignored):
        self.input_dim = int(input_dim)
        self.vocab_size = int(vocab_size)
        self.output_size = int(output_size)
        self.attention_heads = int(attention_heads)
        self.linear_units = int(linear_units)
        self.num_blocks = int(num_blocks)
        self.cnn_module_kernel = int(cnn_module_kernel)
        self.streaming = bool(streaming)
        self.causal = bool(streaming)            # conformer/model.py:35-39
        self.use_dynamic_chunk = bool(streaming)
        self.cnn_module_norm = cnn_module_norm
        self.max_len = int(max_len)

    def to_dict(self):
        return dict(input_dim=self.input_dim, vocab_size=self.vocab_size, output_size=self.output_size,
                    attention_heads=self.attention_heads, linear_units=self.linear_units,
                    num_blocks=self.num_blocks, cnn_module_kernel=self.cnn_module_kernel, streaming=self.streaming,
                    cnn_module_norm=self.cnn_module_norm, max_len=self.max_len)


def conformer_param_shapes(cfg: ConformerConfig) -> Dict[str, tuple]:
    D, FF, K, V, F = cfg.output_size, cfg.linear_units, cfg.cnn_module_kernel, cfg.vocab_size, cfg.input_dim
    H = cfg.attention_heads
    f2 = ((F - 1) // 2 - 1) // 2
    s = {
        "encoder.global_cmvn.mean": (F,), "encoder.global_cmvn.istd": (F,),
        "encoder.embed.conv.0.weight": (D, 1, 3, 3), "encoder.embed.conv.0.bias": (D,),
        "encoder.embed.conv.2.weight": (D, D, 3, 3), "encoder.embed.conv.2.bias": (D,),
        "encoder.embed.out.0.weight": (D * f2, D), "encoder.embed.out.0.bias": (D,),
        "encoder.after_norm.weight": (D,), "encoder.after_norm.bias": (D,),
        "ctc.ctc_lo.weight": (D, V), "ctc.ctc_lo.bias": (V,),
    }
    for i in range(cfg.num_blocks):
        p = f"encoder.encoders.{i}."
        for n in ("norm_ff_macaron", "norm_mha", "norm_conv", "norm_ff", "norm_final"):
            s[p + n + ".weight"] = (D,)
            s[p + n + ".bias"] = (D,)
        for n in ("feed_forward_macaron", "feed_forward"):
            s[p + n + ".w_1.weight"] = (D, FF)
            s[p + n + ".w_1.bias"] = (FF,)
            s[p + n + ".w_2.weight"] = (FF, D)
            s[p + n + ".w_2.bias"] = (D,)
        for n in ("linear_q", "linear_k", "linear_v", "linear_out"):
            s[p + f"self_attn.{n}.weight"] = (D, D)
            s[p + f"self_attn.{n}.bias"] = (D,)
        s[p + "self_attn.linear_pos.weight"] = (D, D)
        s[p + "self_attn.pos_bias_u"] = (H, D // H)
        s[p + "self_attn.pos_bias_v"] = (H, D // H)
        s[p + "conv_module.pointwise_conv1.weight"] = (2 * D, D, 1)
        s[p + "conv_module.pointwise_conv1.bias"] = (2 * D,)
        s[p + "conv_module.depthwise_conv.weight"] = (D, 1, K)
        s[p + "conv_module.depthwise_conv.bias"] = (D,)
        s[p + "conv_module.norm.weight"] = (D,)
        s[p + "conv_module.norm.bias"] = (D,)
        if cfg.cnn_module_norm == "batch_norm":
            s[p + "conv_module.norm._mean"] = (D,)
            s[p + "conv_module.norm._variance"] = (D,)
        s[p + "conv_module.pointwise_conv2.weight"] = (D, D, 1)
        s[p + "conv_module.pointwise_conv2.bias"] = (D,)
    return s


def _fan_in(name, shape):
    if len(shape) == 1:
        return shape[0]
    if len(shape) == 2:  # Paddle Linear [in, out]
        return shape[0]
    return int(np.prod(shape[1:]))  # conv [out, in/groups, k...]


def init_conformer_weights(cfg: ConformerConfig, seed: int = 1000, ctc_gain: float = 8.0,
                           perturb_norms: bool = True) -> Dict[str, np.ndarray]:
    """Seeded synthetic parameters (seed 1000 mirrors ppasr/trainer.py:518).

    perturb_norms: LayerNorm gains/biases are 1/0 at initialisation in the reference; a trained model has
    non-trivial values, so by default they are perturbed (gain U(0.8,1.2), bias U(-0.1,0.1)) to make the
    parity tests sensitive to them."""
    rng = np.random.RandomState(seed)
    w = {}
    for name, shape in conformer_param_shapes(cfg).items():
        if name.endswith("global_cmvn.mean"):
            a = rng.uniform(-1.0, 1.0, shape) + 10.0  # fbank log-mel energies sit around 5..15
        elif name.endswith("global_cmvn.istd"):
            a = rng.uniform(0.2, 0.5, shape)
        elif ".norm" in name and name.endswith("_mean"):
            a = rng.uniform(-0.1, 0.1, shape)
        elif ".norm" in name and name.endswith("_variance"):
            a = rng.uniform(0.5, 1.5, shape)
        elif ("norm" in name.split(".")[-2]) and name.endswith(".weight"):
            a = rng.uniform(0.8, 1.2, shape) if perturb_norms else np.ones(shape)
        elif ("norm" in name.split(".")[-2]) and name.endswith(".bias"):
            a = rng.uniform(-0.1, 0.1, shape) if perturb_norms else np.zeros(shape)
        elif name.endswith("pos_bias_u") or name.endswith("pos_bias_v"):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            a = rng.uniform(-lim, lim, shape)
        elif name.startswith("ctc.ctc_lo"):
            if name.endswith("weight"):  # paddle.nn.Linear default: Xavier-uniform, zero bias
                lim = math.sqrt(6.0 / (shape[0] + shape[1])) * ctc_gain
                a = rng.uniform(-lim, lim, shape)
            else:
                a = np.zeros(shape)
        else:
            lim = 1.0 / math.sqrt(_fan_in(name, shape))
            a = rng.uniform(-lim, lim, shape)
        w[name] = np.ascontiguousarray(a, dtype=np.float32)
    return w


class EfficientConformerConfig(ConformerConfig):
    """configs/efficient_conformer.yml `encoder_conf`; the nested `efficient_conf` block is swallowed by **kwargs in the
    reference constructor (efficient_conformer/encoder.py:55), so its defaults apply unless given at the top level."""
    model_type = "efficient_conformer"

    def __init__(self, stride_layer_idx=3, stride=2, group_layer_idx=(0, 1, 2, 3), group_size=3, stride_kernel=True,
                 efficient_conf=None, **kw):
        super().__init__(**kw)
        if isinstance(stride_layer_idx, (list, tuple)):
            assert len(stride_layer_idx) == 1, "one stride layer is supported"
            stride_layer_idx = stride_layer_idx[0]
        if isinstance(stride, (list, tuple)):
            stride = stride[0]
        assert stride == 2 and group_size == 3, "stride 2 / group size 3 (the reference defaults) are supported"
        self.stride_layer_idx = None if stride_layer_idx is None else int(stride_layer_idx)
        self.stride = int(stride)
        self.group_layer_idx = tuple(int(i) for i in group_layer_idx)
        self.group_size = int(group_size)
        self.stride_kernel = bool(stride_kernel)

    def to_dict(self):
        d = super().to_dict()
        d.update(stride_layer_idx=self.stride_layer_idx, stride=self.stride, group_layer_idx=list(self.group_layer_idx),
                 group_size=self.group_size, stride_kernel=self.stride_kernel)
        return d

    def layer_kernel(self, i):
        k = self.cnn_module_kernel
        if self.stride_layer_idx is None or i <= self.stride_layer_idx:
            return k
        return k // self.stride if self.stride_kernel else k


def efficient_conformer_param_shapes(cfg: EfficientConformerConfig) -> Dict[str, tuple]:
    s = conformer_param_shapes(cfg)
    D, H = cfg.output_size, cfg.attention_heads
    for i in range(cfg.num_blocks):
        p = f"encoder.encoders.{i}."
        s[p + "conv_module.depthwise_conv.weight"] = (D, 1, cfg.layer_kernel(i))
        if i in cfg.group_layer_idx:  # efficient_conformer/attention.py:31-38
            s[p + "self_attn.linear_pos.bias"] = (D,)
            s[p + "self_attn.pos_bias_u"] = (H, D // H * cfg.group_size)
            s[p + "self_attn.pos_bias_v"] = (H, D // H * cfg.group_size)
    return s


def init_efficient_conformer_weights(cfg: EfficientConformerConfig, seed: int = 1000, ctc_gain: float = 8.0,
                                     perturb_norms: bool = True) -> Dict[str, np.ndarray]:
    rng = np.random.RandomState(seed)
    w = {}
    for name, shape in efficient_conformer_param_shapes(cfg).items():
        if name.endswith("global_cmvn.mean"):
            a = rng.uniform(-1.0, 1.0, shape) + 10.0
        elif name.endswith("global_cmvn.istd"):
            a = rng.uniform(0.2, 0.5, shape)
        elif ".norm" in name and name.endswith("_mean"):
            a = rng.uniform(-0.1, 0.1, shape)
        elif ".norm" in name and name.endswith("_variance"):
            a = rng.uniform(0.5, 1.5, shape)
        elif ("norm" in name.split(".")[-2]) and name.endswith(".weight"):
            a = rng.uniform(0.8, 1.2, shape) if perturb_norms else np.ones(shape)
        elif ("norm" in name.split(".")[-2]) and name.endswith(".bias"):
            a = rng.uniform(-0.1, 0.1, shape) if perturb_norms else np.zeros(shape)
        elif name.endswith("pos_bias_u") or name.endswith("pos_bias_v"):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            a = rng.uniform(-lim, lim, shape)
        elif name.startswith("ctc.ctc_lo"):
            if name.endswith("weight"):
                lim = math.sqrt(6.0 / (shape[0] + shape[1])) * ctc_gain
                a = rng.uniform(-lim, lim, shape)
            else:
                a = np.zeros(shape)
        else:
            lim = 1.0 / math.sqrt(_fan_in(name, shape))
            a = rng.uniform(-lim, lim, shape)
        w[name] = np.ascontiguousarray(a, dtype=np.float32)
    return w


class SqueezeformerConfig:
    """Inference-relevant keys of configs/squeezeformer.yml (`encoder_conf`, `streaming`, n_mels); model switches per
    ppasr/model_utils/squeezeformer/model.py:35-41."""
    model_type = "squeezeformer"

    def __init__(self, input_dim=80, vocab_size=4233, encoder_dim=256, output_size=256, attention_heads=4, num_blocks=12,
                 reduce_idx=5, recover_idx=11, feed_forward_expansion_factor=8, cnn_module_kernel=31,
                 cnn_norm_type="layer_norm", adaptive_scale=True, streaming=True, max_len=5000, **_ignored):
        self.input_dim = int(input_dim)
        self.vocab_size = int(vocab_size)
        self.encoder_dim = int(encoder_dim)
        self.output_size = int(output_size)
        self.attention_heads = int(attention_heads)
        self.num_blocks = int(num_blocks)
        self.reduce_idx = reduce_idx
        self.recover_idx = recover_idx
        self.feed_forward_expansion_factor = int(feed_forward_expansion_factor)
        self.linear_units = self.encoder_dim * self.feed_forward_expansion_factor
        self.cnn_module_kernel = int(cnn_module_kernel)
        self.cnn_norm_type = cnn_norm_type
        self.cnn_module_norm = cnn_norm_type
        self.adaptive_scale = bool(adaptive_scale)
        self.streaming = bool(streaming)
        self.causal = bool(streaming)
        self.time_reduction_layer_type = "stream" if streaming else "conv1d"
        self.max_len = int(max_len)

    def to_dict(self):
        return dict(input_dim=self.input_dim, vocab_size=self.vocab_size, encoder_dim=self.encoder_dim,
                    output_size=self.output_size, attention_heads=self.attention_heads, num_blocks=self.num_blocks,
                    reduce_idx=self.reduce_idx, recover_idx=self.recover_idx,
                    feed_forward_expansion_factor=self.feed_forward_expansion_factor,
                    cnn_module_kernel=self.cnn_module_kernel, cnn_norm_type=self.cnn_norm_type,
                    adaptive_scale=self.adaptive_scale, streaming=self.streaming, max_len=self.max_len)


def squeezeformer_param_shapes(cfg: SqueezeformerConfig) -> Dict[str, tuple]:
    """Parameter names/layouts of SqueezeformerModel.encoder + ctc (squeezeformer/encoder.py:129-169,
    attention.py:24-38, positionwise.py:31-43, convolution.py:36-81, subsampling.py:33-44, time_reduction.py)."""
    D, FF, K, V, F = cfg.encoder_dim, cfg.linear_units, cfg.cnn_module_kernel, cfg.vocab_size, cfg.input_dim
    H = cfg.attention_heads
    f2 = ((F - 1) // 2 - 1) // 2
    kr = 1 if cfg.time_reduction_layer_type == "stream" else 5
    s = {
        "encoder.global_cmvn.mean": (F,), "encoder.global_cmvn.istd": (F,),
        "encoder.embed.pw_conv.weight": (D, 1, 3, 3), "encoder.embed.pw_conv.bias": (D,),
        "encoder.embed.dw_conv.weight": (D, D, 3, 3), "encoder.embed.dw_conv.bias": (D,),
        "encoder.embed.input_proj.0.weight": (D * f2, D), "encoder.embed.input_proj.0.bias": (D,),
        "encoder.preln.weight": (D,), "encoder.preln.bias": (D,),
        "encoder.time_reduction_layer.dw_conv.weight": (D, 1, kr), "encoder.time_reduction_layer.dw_conv.bias": (D,),
        "encoder.time_reduction_layer.pw_conv.weight": (D, D, 1), "encoder.time_reduction_layer.pw_conv.bias": (D,),
        "encoder.time_recover_layer.weight": (D, D), "encoder.time_recover_layer.bias": (D,),
        "ctc.ctc_lo.weight": (D, V), "ctc.ctc_lo.bias": (V,),
    }
    for i in range(cfg.num_blocks):
        p = f"encoder.encoders.{i}."
        for n in ("layer_norm1", "layer_norm2", "layer_norm3", "layer_norm4"):
            s[p + n + ".weight"] = (D,)
            s[p + n + ".bias"] = (D,)
        for n in ("ffn1", "ffn2"):
            s[p + n + ".w_1.weight"] = (D, FF)
            s[p + n + ".w_1.bias"] = (FF,)
            s[p + n + ".w_2.weight"] = (FF, D)
            s[p + n + ".w_2.bias"] = (D,)
            s[p + n + ".ada_scale"] = (1, 1, D)
            s[p + n + ".ada_bias"] = (1, 1, D)
        for n in ("linear_q", "linear_k", "linear_v", "linear_out", "linear_pos"):
            s[p + f"self_attn.{n}.weight"] = (D, D)
            s[p + f"self_attn.{n}.bias"] = (D,)
        s[p + "self_attn.pos_bias_u"] = (H, D // H)
        s[p + "self_attn.pos_bias_v"] = (H, D // H)
        s[p + "self_attn.ada_scale"] = (1, 1, D)
        s[p + "self_attn.ada_bias"] = (1, 1, D)
        s[p + "conv_module.ada_scale"] = (1, 1, D)
        s[p + "conv_module.ada_bias"] = (1, 1, D)
        s[p + "conv_module.pointwise_conv1.weight"] = (2 * D, D, 1)
        s[p + "conv_module.pointwise_conv1.bias"] = (2 * D,)
        s[p + "conv_module.depthwise_conv.weight"] = (D, 1, K)
        s[p + "conv_module.depthwise_conv.bias"] = (D,)
        s[p + "conv_module.norm.weight"] = (D,)
        s[p + "conv_module.norm.bias"] = (D,)
        if cfg.cnn_norm_type == "batch_norm":
            s[p + "conv_module.norm._mean"] = (D,)
            s[p + "conv_module.norm._variance"] = (D,)
        s[p + "conv_module.pointwise_conv2.weight"] = (D, D, 1)
        s[p + "conv_module.pointwise_conv2.bias"] = (D,)
    return s


def init_squeezeformer_weights(cfg: SqueezeformerConfig, seed: int = 1000, ctc_gain: float = 8.0,
                               perturb_norms: bool = True) -> Dict[str, np.ndarray]:
    """Seeded synthetic parameters; LayerNorm gains/biases and the adaptive scale/bias vectors are perturbed around their
    reference initial values (1 / 0) so that the parity tests are sensitive to them."""
    rng = np.random.RandomState(seed)
    w = {}
    for name, shape in squeezeformer_param_shapes(cfg).items():
        leaf = name.split(".")[-1]
        parent = name.split(".")[-2]
        if name.endswith("global_cmvn.mean"):
            a = rng.uniform(-1.0, 1.0, shape) + 10.0
        elif name.endswith("global_cmvn.istd"):
            a = rng.uniform(0.2, 0.5, shape)
        elif leaf == "_mean":
            a = rng.uniform(-0.1, 0.1, shape)
        elif leaf == "_variance":
            a = rng.uniform(0.5, 1.5, shape)
        elif leaf == "ada_scale":
            a = rng.uniform(0.8, 1.2, shape) if perturb_norms else np.ones(shape)
        elif leaf == "ada_bias":
            a = rng.uniform(-0.1, 0.1, shape) if perturb_norms else np.zeros(shape)
        elif ("norm" in parent or parent == "preln") and leaf == "weight":
            a = rng.uniform(0.8, 1.2, shape) if perturb_norms else np.ones(shape)
        elif ("norm" in parent or parent == "preln") and leaf == "bias":
            a = rng.uniform(-0.1, 0.1, shape) if perturb_norms else np.zeros(shape)
        elif leaf in ("pos_bias_u", "pos_bias_v"):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            a = rng.uniform(-lim, lim, shape)
        elif name.startswith("ctc.ctc_lo"):
            if leaf == "weight":
                lim = math.sqrt(6.0 / (shape[0] + shape[1])) * ctc_gain
                a = rng.uniform(-lim, lim, shape)
            else:
                a = np.zeros(shape)
        else:
            lim = 1.0 / math.sqrt(_fan_in(name, shape))
            a = rng.uniform(-lim, lim, shape)
        w[name] = np.ascontiguousarray(a, dtype=np.float32)
    return w


class DeepSpeech2Config:
    """configs/deepspeech2.yml `encoder_conf` + the streaming switch (deepspeech2/model.py:40: 'forward' LSTM when streaming,
    'bidirect' otherwise)."""
    model_type = "deepspeech2"

    def __init__(self, input_dim=80, vocab_size=4233, num_rnn_layers=5, rnn_size=1024, use_gru=False, streaming=False,
                 **_ignored):
        self.input_dim = int(input_dim)
        self.vocab_size = int(vocab_size)
        self.num_rnn_layers = int(num_rnn_layers)
        self.rnn_size = int(rnn_size)
        self.use_gru = bool(use_gru)
        self.streaming = bool(streaming)
        self.num_directions = 1 if streaming else 2

    def to_dict(self):
        return dict(input_dim=self.input_dim, vocab_size=self.vocab_size, num_rnn_layers=self.num_rnn_layers,
                    rnn_size=self.rnn_size, use_gru=self.use_gru, streaming=self.streaming)


def deepspeech2_param_shapes(cfg: DeepSpeech2Config) -> Dict[str, tuple]:
    """deepspeech2/conv.py:8-12, encoder.py:29-55, model.py:36-44 (the CTC module is named `decoder`); RNN parameters use
    paddle.nn.LSTM/GRU's flat names weight_ih_l0 / weight_hh_l0 / bias_ih_l0 / bias_hh_l0 (+ `_reverse`)."""
    F, V, H = cfg.input_dim, cfg.vocab_size, cfg.rnn_size
    f2 = ((F - 1) // 2 - 1) // 2
    G = 3 if cfg.use_gru else 4
    nd = cfg.num_directions
    s = {
        "encoder.global_cmvn.mean": (F,), "encoder.global_cmvn.istd": (F,),
        "encoder.conv.conv.0.weight": (32, 1, 3, 3), "encoder.conv.conv.0.bias": (32,),
        "encoder.conv.conv.2.weight": (32, 32, 3, 3), "encoder.conv.conv.2.bias": (32,),
        "decoder.ctc_lo.weight": (H * nd, V), "decoder.ctc_lo.bias": (V,),
    }
    for i in range(cfg.num_rnn_layers):
        isz = 32 * f2 if i == 0 else H * nd
        for suf in ([""] if nd == 1 else ["", "_reverse"]):
            s[f"encoder.rnn.{i}.weight_ih_l0{suf}"] = (G * H, isz)
            s[f"encoder.rnn.{i}.weight_hh_l0{suf}"] = (G * H, H)
            s[f"encoder.rnn.{i}.bias_ih_l0{suf}"] = (G * H,)
            s[f"encoder.rnn.{i}.bias_hh_l0{suf}"] = (G * H,)
        s[f"encoder.layernorm_list.{i}.weight"] = (H * nd,)
        s[f"encoder.layernorm_list.{i}.bias"] = (H * nd,)
    return s


def init_deepspeech2_weights(cfg: DeepSpeech2Config, seed: int = 1000, ctc_gain: float = 8.0,
                             perturb_norms: bool = True) -> Dict[str, np.ndarray]:
    """Seeded synthetic parameters: RNN weights/biases U(+-1/sqrt(H)) (paddle.nn.LSTM default), conv Kaiming-uniform."""
    rng = np.random.RandomState(seed)
    w = {}
    H = cfg.rnn_size
    for name, shape in deepspeech2_param_shapes(cfg).items():
        leaf = name.split(".")[-1]
        if name.endswith("global_cmvn.mean"):
            a = rng.uniform(-1.0, 1.0, shape) + 10.0
        elif name.endswith("global_cmvn.istd"):
            a = rng.uniform(0.2, 0.5, shape)
        elif "layernorm_list" in name and leaf == "weight":
            a = rng.uniform(0.8, 1.2, shape) if perturb_norms else np.ones(shape)
        elif "layernorm_list" in name and leaf == "bias":
            a = rng.uniform(-0.1, 0.1, shape) if perturb_norms else np.zeros(shape)
        elif ".rnn." in name:
            lim = 1.0 / math.sqrt(H)
            a = rng.uniform(-lim, lim, shape)
        elif name.startswith("decoder.ctc_lo"):
            if leaf == "weight":
                lim = math.sqrt(6.0 / (shape[0] + shape[1])) * ctc_gain
                a = rng.uniform(-lim, lim, shape)
            else:
                a = np.zeros(shape)
        else:
            lim = 1.0 / math.sqrt(_fan_in(name, shape))
            a = rng.uniform(-lim, lim, shape)
        w[name] = np.ascontiguousarray(a, dtype=np.float32)
    return w


def save_npz(path: str, weights: Dict[str, np.ndarray], cfg: ConformerConfig = None):
    extra = {}
    if cfg is not None:
        extra["__config_json__"] = np.frombuffer(json.dumps(cfg.to_dict()).encode(), dtype=np.uint8)
    np.savez(path, **weights, **extra)


def load_npz(path: str):
    z = np.load(path)
    cfg = None
    w = {}
    for k in z.files:
        if k == "__config_json__":
            cfg = ConformerConfig(**json.loads(bytes(z[k]).decode()))
        else:
            w[k] = np.asarray(z[k], dtype=np.float32)
    return w, cfg


def load_pdparams(path: str) -> Dict[str, np.ndarray]:
    """`model.pdparams` as written by paddle.save (ppasr/trainer.py:311): a pickle of name -> ndarray
    (Paddle >= 2.1 stores plain numpy arrays; tuples (name, ndarray) from older versions are unwrapped)."""
    with open(path, "rb") as f:
        obj = pickle.load(f, encoding="latin1")
    out = {}
    for k, v in obj.items():
        if isinstance(v, tuple) and len(v) == 2:
            v = v[1]
        if isinstance(v, np.ndarray):
            out[k] = np.asarray(v, dtype=np.float32)
    return out


# ---- exported inference model: model.pdiparams (+ model.pdiparams.info), written by paddle.jit.save (ppasr/trainer.py:675-681)
# and read by the reference through paddle.inference (infer_utils/inference_predictor.py:41-45). Format restated from the
# Paddle 2.5 sources (paddle/fluid/framework/lod_tensor.cc SerializeToStream, tensor_util.cc TensorToStream, save_combine_op):
# tensors concatenated in the order of the sorted program variable names; per tensor
#   uint32 lod-tensor version (0) | uint64 number of LoD levels | per level: uint64 byte size + data
#   uint32 tensor version (0) | int32 size of the TensorDesc protobuf | TensorDesc { 1: data_type (varint), 2: dims (int64) }
#   raw little-endian data.
# `model.pdiparams.info` is a pickle {program variable name: {'structured_name': 'encoder.embed.conv.0.weight', ...}} that maps the
# auto-generated variable names back to the state-dict names. UNVERIFIED against a real file (no Paddle in this sandbox);
# `save_pdiparams` writes the same layout so the reader is at least round-trip tested.
_PD_DTYPES = {0: np.bool_, 1: np.int16, 2: np.int32, 3: np.int64, 4: np.float16, 5: np.float32, 6: np.float64, 20: np.uint8,
              21: np.int8}


def _read_varint(buf, pos):
    val, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _parse_tensor_desc(buf):
    pos, dtype, dims = 0, None, []
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            v, pos = _read_varint(buf, pos)
            if v >= 1 << 63:
                v -= 1 << 64
            if field == 1:
                dtype = v
            elif field == 2:
                dims.append(v)
        elif wire == 2:  # packed repeated int64
            n, pos = _read_varint(buf, pos)
            end = pos + n
            while pos < end:
                v, pos = _read_varint(buf, pos)
                if field == 2:
                    dims.append(v - (1 << 64) if v >= 1 << 63 else v)
        else:
            raise ValueError("unexpected wire type in TensorDesc")
    return dtype, dims


def load_pdiparams(path: str, info_path: str = None) -> Dict[str, np.ndarray]:
    """Reads `model.pdiparams`; names come from `model.pdiparams.info` (structured names) when present, else tensors are
    returned as 'param_%d' in file order."""
    import struct
    data = open(path, "rb").read()
    pos, tensors = 0, []
    while pos < len(data):
        (ver,) = struct.unpack_from("<I", data, pos)
        pos += 4
        if ver != 0:
            raise ValueError(f"unsupported LoDTensor version {ver} at byte {pos - 4}")
        (nlod,) = struct.unpack_from("<Q", data, pos)
        pos += 8
        for _ in range(nlod):
            (nbytes,) = struct.unpack_from("<Q", data, pos)
            pos += 8 + nbytes
        (tver,) = struct.unpack_from("<I", data, pos)
        pos += 4
        if tver != 0:
            raise ValueError(f"unsupported Tensor version {tver}")
        (dsz,) = struct.unpack_from("<i", data, pos)
        pos += 4
        dtype, dims = _parse_tensor_desc(data[pos:pos + dsz])
        pos += dsz
        if dtype not in _PD_DTYPES:
            raise ValueError(f"unsupported Paddle dtype enum {dtype}")
        npdt = np.dtype(_PD_DTYPES[dtype])
        n = int(np.prod(dims)) if dims else 1
        arr = np.frombuffer(data, dtype=npdt, count=n, offset=pos).reshape(dims)
        pos += n * npdt.itemsize
        tensors.append(arr)
    if info_path is None and os.path.exists(path + ".info"):
        info_path = path + ".info"
    names = None
    if info_path is not None:
        with open(info_path, "rb") as f:
            info = pickle.load(f, encoding="latin1")
        keys = sorted(k for k in info if isinstance(info[k], dict))
        if len(keys) == len(tensors):
            names = [info[k].get("structured_name", k) for k in keys]
    if names is None:
        names = [f"param_{i}" for i in range(len(tensors))]
    return {n: np.asarray(t, dtype=np.float32) if t.dtype.kind == "f" else np.asarray(t) for n, t in zip(names, tensors)}


def save_pdiparams(path: str, weights: Dict[str, np.ndarray], write_info: bool = True):
    """Writes the layout `load_pdiparams` reads (variable names = 'param_%05d' in the order of the structured names)."""
    import struct

    def varint(v):
        if v < 0:
            v += 1 << 64
        out = bytearray()
        while True:
            b = v & 0x7F
            v >>= 7
            out.append(b | (0x80 if v else 0))
            if not v:
                return bytes(out)

    rev = {np.dtype(v): k for k, v in _PD_DTYPES.items()}
    info = {}
    with open(path, "wb") as f:
        for i, (name, arr) in enumerate(weights.items()):
            arr = np.ascontiguousarray(arr)
            desc = b"\x08" + varint(rev[arr.dtype]) + b"".join(b"\x10" + varint(int(d)) for d in arr.shape)
            f.write(struct.pack("<IQ", 0, 0))
            f.write(struct.pack("<Ii", 0, len(desc)))
            f.write(desc)
            f.write(arr.tobytes())
            info[f"param_{i:05d}"] = {"structured_name": name, "stop_gradient": False, "trainable": True}
    if write_info:
        with open(path + ".info", "wb") as f:
            pickle.dump(info, f, protocol=2)


def make_vocab(vocab_size: int):
    """Synthetic vocabulary in the reference's order (ppasr/trainer.py:479-487): <blank>, <unk>, chars, <eos>."""
    chars = [chr(0x4E00 + i) for i in range(vocab_size - 3)]
    return ["<blank>", "<unk>"] + chars + ["<eos>"]


def read_vocab_file(path: str):
    """ppasr/data_utils/featurizer/text_featurizer.py:52-59: one `token\\tcount` per line."""
    with open(path, "r", encoding="utf-8") as f:
        return [line.split("\t")[0].replace("\n", "") for line in f.readlines()]


def read_mean_istd(path: str, eps: float = 1e-20):
    """ppasr/data_utils/normalizer.py:27-41."""
    with open(path, "r", encoding="utf-8") as f:
        data = json.load(f)
    mean = np.array(data["mean"], dtype=np.float32)
    istd = np.maximum(np.array(data["istd"], dtype=np.float32), eps)
    return mean, istd


def synthetic_fbank(batch: int, frames: int, n_mels: int = 80, seed: int = 1234) -> np.ndarray:
    """Seeded stand-in for kaldi fbank features of 16 kHz audio (log-mel energies, roughly 5..15 with
    temporal and spectral correlation). Shapes follow audio_featurizer.py:125-136: frames = 1 + (N-400)//160."""
    rng = np.random.RandomState(seed)
    t = np.arange(frames, dtype=np.float32)[None, :, None]
    f = np.arange(n_mels, dtype=np.float32)[None, None, :]
    phase = rng.uniform(0, 2 * np.pi, (batch, 1, 1)).astype(np.float32)
    x = 10.0 + 2.0 * np.sin(0.05 * t + phase) * np.cos(0.11 * f + phase) + 1.5 * np.sin(0.31 * t + 0.07 * f)
    x = x + rng.normal(0.0, 1.0, (batch, frames, n_mels))
    return np.ascontiguousarray(x, dtype=np.float32)
