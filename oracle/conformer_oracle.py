"""CPU oracle for the Conformer hot path (TEST INFRASTRUCTURE ONLY).

A PyTorch-CPU restatement (fp32 by default, fp64 optional) of the reference inference math:
fbank features -> GlobalCMVN -> Conv2dSubsampling4 -> RelPositionalEncoding -> N x
ConformerEncoderLayer -> after_norm -> CTC projection -> softmax. Every function cites the
reference file:line it follows (paths relative to /root/reference, yeyupiaoling/PPASR @ c8bb3b96).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker / reported CPU baseline. The product path (ppasr_b200/) never
imports it and has no CPU fallback.

PARITY STATUS: **pinned against the reference's own model code** (not against Paddle's kernels). The
reference ships no tests or golden vectors and PaddlePaddle 2.5.1 is not installable
here, so tests/golden/make_encoder_golden.py imports the UNMODIFIED reference classes (ConformerEncoder,
CTCLoss, GlobalCMVN, masks, embedding, attention, convolution ...) from /root/reference on top of
tests/golden/paddle_shim -- a torch-CPU stand-in for the `paddle` package implementing only the primitive
calls those files make -- and commits their outputs for seeded weights (tests/golden/encoder_golden_*.npz).
tests/test_encoder_golden_cpu.py checks this restatement against them: offline logits/probabilities of a
ragged batch incl. padded frames, a forward_chunk chain and its final att/cnn caches; max |diff| 1.2e-5 on
logits of magnitude ~30 (fp32 summation order). What stays unpinned is only the arithmetic of Paddle's
primitive ops themselves (matmul, conv, layer_norm, softmax: standard definitions). Further checks: the
docstring known-answer masks, chunked == chunk-masked offline forward, batch invariance, an fp64 re-run
as noise floor. The greedy decoder is pinned bit-for-bit (oracle/decoders_oracle.py against golden vectors
produced by the reference's own NumPy code).

Weights are a dict name -> numpy array using the reference's Paddle parameter names and layouts
(`Linear.weight` is [in, out]; `Conv*.weight` is [out, in/groups, k...]).
"""
import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F


class ConformerConf:
    """Subset of configs/conformer.yml encoder_conf + model switches that affect inference."""

    def __init__(self, input_dim=80, vocab_size=4233, output_size=256, attention_heads=4, linear_units=2048,
                 num_blocks=12, cnn_module_kernel=15, streaming=True, cnn_module_norm="layer_norm",
                 max_len=5000):
        self.input_dim = input_dim
        self.vocab_size = vocab_size
        self.output_size = output_size
        self.attention_heads = attention_heads
        self.linear_units = linear_units
        self.num_blocks = num_blocks
        self.cnn_module_kernel = cnn_module_kernel
        self.streaming = streaming
        # ppasr/model_utils/conformer/model.py:35-39: streaming => dynamic chunk + causal conv
        self.use_dynamic_chunk = bool(streaming)
        self.causal = bool(streaming)
        self.cnn_module_norm = cnn_module_norm
        self.max_len = max_len


# ------------------------------------------------------------------------------------------------
# masks  (ppasr/model_utils/utils/mask.py)
# ------------------------------------------------------------------------------------------------
def make_pad_mask(lengths: torch.Tensor) -> torch.Tensor:
    """mask.py:22-43 -- True at padded positions, shape (B, max(lengths))."""
    batch_size = int(lengths.shape[0])
    max_len = int(lengths.max())
    seq_range = torch.arange(0, max_len, dtype=torch.int64)
    seq_range_expand = seq_range.unsqueeze(0).expand(batch_size, max_len)
    seq_length_expand = lengths.to(torch.int64).unsqueeze(-1)
    return seq_range_expand >= seq_length_expand


def make_non_pad_mask(lengths: torch.Tensor) -> torch.Tensor:
    """mask.py:46-68."""
    return make_pad_mask(lengths).logical_not()


def subsequent_chunk_mask(size: int, chunk_size: int, num_left_chunks: int = -1) -> torch.Tensor:
    """mask.py:93-122."""
    ret = torch.zeros(size, size, dtype=torch.bool)
    for i in range(size):
        if num_left_chunks < 0:
            start = 0
        else:
            start = max(0, (i // chunk_size - num_left_chunks) * chunk_size)
        ending = min(size, (i // chunk_size + 1) * chunk_size)
        ret[i, start:ending] = True
    return ret


def add_optional_chunk_mask(xs, masks, use_dynamic_chunk, decoding_chunk_size, static_chunk_size,
                            num_decoding_left_chunks):
    """mask.py:125-186, decoding branches only (decoding_chunk_size != 0)."""
    if use_dynamic_chunk:
        max_len = xs.shape[1]
        if decoding_chunk_size < 0:
            chunk_size = max_len
            num_left_chunks = -1
        elif decoding_chunk_size > 0:
            chunk_size = decoding_chunk_size
            num_left_chunks = num_decoding_left_chunks
        else:
            raise ValueError("training-time random chunk is out of scope for the inference oracle")
        chunk_masks = subsequent_chunk_mask(xs.shape[1], chunk_size, num_left_chunks).unsqueeze(0)
        chunk_masks = masks.logical_and(chunk_masks)  # (B, L, L)
    elif static_chunk_size > 0:
        chunk_masks = subsequent_chunk_mask(xs.shape[1], static_chunk_size, num_decoding_left_chunks).unsqueeze(0)
        chunk_masks = masks.logical_and(chunk_masks)
    else:
        chunk_masks = masks
    return chunk_masks


# ------------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------------
def positional_table(max_len: int, d_model: int, dtype) -> torch.Tensor:
    """ppasr/model_utils/conformer/embedding.py:41-53 -- [1, max_len, d] sin/cos table.

    The table is built in fp32 exactly like the reference and then cast."""
    pe = torch.zeros(1, max_len, d_model, dtype=torch.float32)
    position = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe[:, :, 0::2] = torch.sin(position * div_term)
    pe[:, :, 1::2] = torch.cos(position * div_term)
    return pe.to(dtype)


def linear(x, w, b=None):
    """Paddle Linear: y = x @ W[in,out] + b."""
    y = x @ w
    return y if b is None else y + b


def layer_norm(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def swish(x):
    return x * torch.sigmoid(x)


class ConformerOracle:
    def __init__(self, conf: ConformerConf, weights: Dict[str, np.ndarray], dtype=torch.float32):
        self.conf = conf
        self.dtype = dtype
        self.w = {k: torch.from_numpy(np.asarray(v)).to(dtype) for k, v in weights.items()}
        self.pe = positional_table(conf.max_len, conf.output_size, dtype)
        self.xscale = math.sqrt(conf.output_size)

    # -- ppasr/model_utils/utils/cmvn.py:21-32 ---------------------------------------------------
    def global_cmvn(self, x):
        x = x - self.w["encoder.global_cmvn.mean"]
        return x * self.w["encoder.global_cmvn.istd"]

    # -- ppasr/model_utils/conformer/subsampling.py:96-115 + embedding.py:102-115 -----------------
    def embed(self, x, x_mask, offset=0):
        w = self.w
        x = x.unsqueeze(1)  # (b, 1, t, f)
        x = F.relu(F.conv2d(x, w["encoder.embed.conv.0.weight"], w["encoder.embed.conv.0.bias"], stride=2))
        x = F.relu(F.conv2d(x, w["encoder.embed.conv.2.weight"], w["encoder.embed.conv.2.bias"], stride=2))
        b, c, t, f = x.shape
        x = linear(x.permute(0, 2, 1, 3).reshape(b, t, c * f), w["encoder.embed.out.0.weight"],
                   w["encoder.embed.out.0.bias"])
        assert offset + x.shape[1] < self.conf.max_len  # embedding.py:110-112
        x = x * self.xscale
        pos_emb = self.pe[:, offset:offset + x.shape[1]]
        return x, pos_emb, x_mask[:, :, :-2:2][:, :, :-2:2]

    def position_encoding(self, offset, size):
        """embedding.py:71-85."""
        assert offset + size < self.conf.max_len
        return self.pe[:, offset:offset + size]

    # -- ppasr/model_utils/conformer/positionwise.py:30-39 ----------------------------------------
    def ffn(self, prefix, x):
        w = self.w
        h = swish(linear(x, w[prefix + ".w_1.weight"], w[prefix + ".w_1.bias"]))
        return linear(h, w[prefix + ".w_2.weight"], w[prefix + ".w_2.bias"])

    # -- ppasr/model_utils/conformer/attention.py:198-262 (+ 56-126) ------------------------------
    def rel_mha(self, prefix, x, mask, pos_emb, cache):
        w = self.w
        H = self.conf.attention_heads
        D = self.conf.output_size
        dk = D // H
        B = x.shape[0]
        q = linear(x, w[prefix + ".linear_q.weight"], w[prefix + ".linear_q.bias"]).reshape(B, -1, H, dk)
        k = linear(x, w[prefix + ".linear_k.weight"], w[prefix + ".linear_k.bias"]).reshape(B, -1, H, dk)
        v = linear(x, w[prefix + ".linear_v.weight"], w[prefix + ".linear_v.bias"]).reshape(B, -1, H, dk)
        q = q.transpose(1, 2)
        k = k.transpose(1, 2)
        v = v.transpose(1, 2)
        if cache is not None and cache.shape[0] > 0:  # attention.py:225-229
            key_cache, value_cache = torch.split(cache, cache.shape[-1] // 2, dim=-1)
            k = torch.cat([key_cache, k], dim=2)
            v = torch.cat([value_cache, v], dim=2)
        new_cache = torch.cat((k, v), dim=-1)  # attention.py:232
        p = linear(pos_emb, w[prefix + ".linear_pos.weight"]).reshape(pos_emb.shape[0], -1, H, dk).transpose(1, 2)
        q_u = q + w[prefix + ".pos_bias_u"].unsqueeze(1)  # (B,H,T1,dk) + (H,1,dk)
        q_v = q + w[prefix + ".pos_bias_v"].unsqueeze(1)
        matrix_ac = q_u @ k.transpose(-2, -1)
        matrix_bd = q_v @ p.transpose(-2, -1)  # no rel_shift (attention.py:256-258)
        scores = (matrix_ac + matrix_bd) / math.sqrt(dk)
        # forward_attention, attention.py:86-126
        if mask is not None and mask.shape[2] > 0:
            m = mask.unsqueeze(1).eq(0)
            m = m[:, :, :, :scores.shape[-1]]
            scores = scores.masked_fill(m, -float("inf"))
            attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)
        else:
            attn = torch.softmax(scores, dim=-1)
        x = attn @ v
        x = x.transpose(1, 2).reshape(B, -1, D)
        return linear(x, w[prefix + ".linear_out.weight"], w[prefix + ".linear_out.bias"]), new_cache

    # -- ppasr/model_utils/conformer/convolution.py:82-143 ----------------------------------------
    def conv_module(self, prefix, x, mask_pad, cache):
        """mask_pad: bool (B,1,T) with True = frames to ZERO (the caller already inverted it,
        encoder.py:191-192 + convolution.py:104-106), or None for the fake (0,0,0) mask."""
        w = self.w
        conf = self.conf
        lorder = conf.cnn_module_kernel - 1 if conf.causal else 0
        x = x.transpose(1, 2)  # (B, C, T)
        if mask_pad is not None and mask_pad.shape[2] > 0:
            x = x.masked_fill(mask_pad, 0.0)
        if lorder > 0:
            if cache is None or cache.shape[-1] == 0:
                x = F.pad(x, (lorder, 0), "constant", 0.0)
            else:
                assert cache.shape[0] == x.shape[0] and cache.shape[1] == x.shape[1]
                x = torch.cat((cache, x), dim=2)
            assert x.shape[2] > lorder
            new_cache = x[:, :, -lorder:]
        else:
            new_cache = torch.zeros(0, 0, 0, dtype=x.dtype)
        x = F.conv1d(x, w[prefix + ".pointwise_conv1.weight"], w[prefix + ".pointwise_conv1.bias"])
        x = F.glu(x, dim=1)
        pad = 0 if conf.causal else (conf.cnn_module_kernel - 1) // 2
        x = F.conv1d(x, w[prefix + ".depthwise_conv.weight"], w[prefix + ".depthwise_conv.bias"], padding=pad,
                     groups=x.shape[1])
        if conf.cnn_module_norm == "layer_norm":
            x = layer_norm(x.transpose(1, 2), w[prefix + ".norm.weight"], w[prefix + ".norm.bias"]).transpose(1, 2)
        else:  # BatchNorm1D in eval mode, epsilon 1e-5
            x = F.batch_norm(x, w[prefix + ".norm._mean"], w[prefix + ".norm._variance"], w[prefix + ".norm.weight"],
                             w[prefix + ".norm.bias"], training=False, eps=1e-5)
        x = swish(x)
        x = F.conv1d(x, w[prefix + ".pointwise_conv2.weight"], w[prefix + ".pointwise_conv2.bias"])
        if mask_pad is not None and mask_pad.shape[2] > 0:
            x = x.masked_fill(mask_pad, 0.0)
        return x.transpose(1, 2), new_cache

    # -- ppasr/model_utils/conformer/encoder.py:346-431 -------------------------------------------
    def layer(self, i, x, mask, pos_emb, mask_pad=None, att_cache=None, cnn_cache=None):
        w = self.w
        p = f"encoder.encoders.{i}"

        def ln(name, t):
            return layer_norm(t, w[f"{p}.{name}.weight"], w[f"{p}.{name}.bias"])

        x = x + 0.5 * self.ffn(p + ".feed_forward_macaron", ln("norm_ff_macaron", x))
        x_att, new_att_cache = self.rel_mha(p + ".self_attn", ln("norm_mha", x), mask, pos_emb, att_cache)
        x = x + x_att
        x_conv, new_cnn_cache = self.conv_module(p + ".conv_module", ln("norm_conv", x), mask_pad, cnn_cache)
        x = x + x_conv
        x = x + 0.5 * self.ffn(p + ".feed_forward", ln("norm_ff", x))
        x = ln("norm_final", x)
        return x, new_att_cache, new_cnn_cache

    # -- ppasr/model_utils/conformer/encoder.py:164-206 -------------------------------------------
    def encoder_forward(self, xs, xs_lens, decoding_chunk_size=-1, num_decoding_left_chunks=-1):
        conf = self.conf
        xs = xs.to(self.dtype)
        T = xs.shape[1]
        masks = make_non_pad_mask(xs_lens)
        if masks.shape[1] < T:  # lengths shorter than the padded tensor
            masks = F.pad(masks, (0, T - masks.shape[1]), value=False)
        masks = masks.unsqueeze(1)  # (B,1,T)
        xs = self.global_cmvn(xs)
        xs, pos_emb, masks = self.embed(xs, masks, offset=0)
        mask_pad = ~masks  # encoder.py:191-192 (True = PAD => zeroed inside the conv module)
        chunk_masks = add_optional_chunk_mask(xs, masks, conf.use_dynamic_chunk, decoding_chunk_size, 0,
                                              num_decoding_left_chunks)
        for i in range(conf.num_blocks):
            xs, _, _ = self.layer(i, xs, chunk_masks, pos_emb, mask_pad)
        xs = layer_norm(xs, self.w["encoder.after_norm.weight"], self.w["encoder.after_norm.bias"])
        return xs, masks

    # -- ppasr/model_utils/conformer/encoder.py:208-283 -------------------------------------------
    def encoder_forward_chunk(self, xs, offset, required_cache_size, att_cache, cnn_cache):
        conf = self.conf
        assert xs.shape[0] == 1
        xs = self.global_cmvn(xs.to(self.dtype))
        tmp_masks = torch.ones(1, 1, xs.shape[1], dtype=torch.bool)
        xs, pos_emb, _ = self.embed(xs, tmp_masks, offset=offset)
        cache_t1 = att_cache.shape[2] if att_cache.dim() == 4 else 0
        chunk_size = xs.shape[1]
        attention_key_size = cache_t1 + chunk_size
        pos_emb = self.position_encoding(offset=offset - cache_t1, size=attention_key_size)
        if required_cache_size < 0:
            next_cache_start = 0
        elif required_cache_size == 0:
            next_cache_start = attention_key_size
        else:
            next_cache_start = max(attention_key_size - required_cache_size, 0)
        r_att, r_cnn = [], []
        for i in range(conf.num_blocks):
            ac = att_cache[i:i + 1] if att_cache.numel() > 0 else None
            cc = cnn_cache[i] if cnn_cache.numel() > 0 else None
            xs, new_att, new_cnn = self.layer(i, xs, None, pos_emb, None, ac, cc)
            r_att.append(new_att[:, :, next_cache_start:, :])
            r_cnn.append(new_cnn)
        xs = layer_norm(xs, self.w["encoder.after_norm.weight"], self.w["encoder.after_norm.bias"])
        return xs, torch.cat(r_att, dim=0), torch.stack(r_cnn, dim=0)

    # -- ppasr/model_utils/loss/ctc.py:27,62-70 ---------------------------------------------------
    def ctc_logits(self, hs):
        return linear(hs, self.w["ctc.ctc_lo.weight"], self.w["ctc.ctc_lo.bias"])

    def ctc_softmax(self, hs):
        return torch.softmax(self.ctc_logits(hs), dim=2)

    # -- ppasr/model_utils/conformer/model.py:148-184 ---------------------------------------------
    @torch.no_grad()
    def get_encoder_out(self, speech, speech_lengths, return_logits=False):
        enc, _ = self.encoder_forward(speech, speech_lengths, decoding_chunk_size=-1, num_decoding_left_chunks=-1)
        return self.ctc_logits(enc) if return_logits else self.ctc_softmax(enc)

    @torch.no_grad()
    def get_encoder_out_chunk(self, speech, offset, required_cache_size, att_cache, cnn_cache, return_logits=False):
        xs, att_cache, cnn_cache = self.encoder_forward_chunk(speech, offset, required_cache_size, att_cache,
                                                              cnn_cache)
        out = self.ctc_logits(xs) if return_logits else self.ctc_softmax(xs)
        return out, att_cache, cnn_cache


def stream_windows(num_frames: int, is_end: bool = False, decoding_chunk_size=16, context=7, subsampling=4):
    """Window [start, end) pairs used by PPASRPredictor.predict_stream (ppasr/predict.py:277-300):
    window = (chunk-1)*sub + context = 67 frames, stride = sub*chunk = 64; the last (short) window
    is only consumed when is_end (left_frames = context)."""
    decoding_window = (decoding_chunk_size - 1) * subsampling + context
    stride = subsampling * decoding_chunk_size
    left = context if is_end else decoding_window
    return [(cur, min(cur + decoding_window, num_frames)) for cur in range(0, num_frames - left + 1, stride)]
