"""CPU oracle for the Squeezeformer hot path (TEST INFRASTRUCTURE ONLY -- see oracle/conformer_oracle.py header).

PyTorch-CPU restatement of the reference inference math of ppasr/model_utils/squeezeformer/ (paths relative to
/root/reference, yeyupiaoling/PPASR @ c8bb3b96): GlobalCMVN -> DepthwiseConv2DSubsampling4 -> preln ->
12 x SqueezeformerEncoderLayer (post-norm MHA / FFN / Conv / FFN, adaptive scale on every module input) with the
time-reduction layer before block `reduce_idx` and the recover step before block `recover_idx` -> CTC projection
-> softmax. No final norm, no final_proj when output_size == encoder_dim (encoder.py:167-169).

PARITY STATUS: **pinned against the reference's own model code** run on tests/golden/paddle_shim (see the Conformer oracle
header): tests/golden/encoder_golden_squeezeformer_{stream,offline}.npz come from the unmodified SqueezeformerEncoder
(stream = TimeReductionLayerStream + causal + dynamic chunk, offline = TimeReductionLayer1D), and
tests/test_encoder_golden_cpu.py checks offline logits (padded frames included), a forward_chunk chain and its final
caches to 5e-5 (observed 1.1e-5). Plus self-consistency properties (tests/test_oracle_cpu.py) and an fp64 re-run.
"""
import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from oracle.conformer_oracle import (ConformerOracle, add_optional_chunk_mask, layer_norm, linear, make_non_pad_mask,
                                     positional_table)


class SqueezeformerConf:
    """Inference-relevant keys of configs/squeezeformer.yml encoder_conf + the model switches of
    squeezeformer/model.py:35-41 (streaming => causal conv, dynamic chunk, 'stream' time reduction)."""

    def __init__(self, input_dim=80, vocab_size=4233, encoder_dim=256, output_size=256, attention_heads=4, num_blocks=12,
                 reduce_idx=5, recover_idx=11, feed_forward_expansion_factor=8, cnn_module_kernel=31,
                 cnn_norm_type="layer_norm", adaptive_scale=True, streaming=True, max_len=5000, **_ignored):
        assert output_size == encoder_dim, "final_proj (output_size != encoder_dim) is not used by the shipped config"
        self.input_dim = input_dim
        self.vocab_size = vocab_size
        self.encoder_dim = encoder_dim
        self.output_size = output_size
        self.attention_heads = attention_heads
        self.num_blocks = num_blocks
        self.reduce_idx = reduce_idx
        self.recover_idx = recover_idx
        self.feed_forward_expansion_factor = feed_forward_expansion_factor
        self.linear_units = encoder_dim * feed_forward_expansion_factor
        self.cnn_module_kernel = cnn_module_kernel
        self.cnn_module_norm = cnn_norm_type
        self.adaptive_scale = adaptive_scale
        self.streaming = streaming
        self.causal = bool(streaming)
        self.use_dynamic_chunk = bool(streaming)
        self.time_reduction_layer_type = "stream" if streaming else "conv1d"
        self.max_len = max_len


class SqueezeformerOracle(ConformerOracle):
    def __init__(self, conf: SqueezeformerConf, weights: Dict[str, np.ndarray], dtype=torch.float32):
        self.conf = conf
        self.dtype = dtype
        self.w = {k: torch.from_numpy(np.asarray(v)).to(dtype) for k, v in weights.items()}
        self.pe = positional_table(conf.max_len, conf.encoder_dim, dtype)
        self.xscale = math.sqrt(conf.encoder_dim)

    # -- squeezeformer/subsampling.py:53-68 (+ conformer/embedding.py:102-115 applied to the 4864-wide tensor) ----
    def embed(self, x, x_mask, offset=0):
        w = self.w
        x = x.unsqueeze(1)
        x = F.relu(F.conv2d(x, w["encoder.embed.pw_conv.weight"], w["encoder.embed.pw_conv.bias"], stride=2))
        x = F.relu(F.conv2d(x, w["encoder.embed.dw_conv.weight"], w["encoder.embed.dw_conv.bias"], stride=2))
        b, c, t, f = x.shape
        x = x.permute(0, 2, 1, 3).reshape(b, t, c * f)
        assert offset + x.shape[1] < self.conf.max_len
        x = x * self.xscale  # pos_enc on the flattened conv output, before input_proj
        pos_emb = self.pe[:, offset:offset + x.shape[1]]
        x = linear(x, w["encoder.embed.input_proj.0.weight"], w["encoder.embed.input_proj.0.bias"])
        return x, pos_emb, x_mask[:, :, :-2:2][:, :, :-2:2]

    def ada(self, prefix, x):
        if not self.conf.adaptive_scale:
            return x
        return self.w[prefix + ".ada_scale"] * x + self.w[prefix + ".ada_bias"]

    # -- squeezeformer/positionwise.py:55-65 ------------------------------------------------------------------
    def ffn(self, prefix, x):
        return super().ffn(prefix, self.ada(prefix, x))

    # -- squeezeformer/attention.py:96-162 (linear_pos WITH bias, ada on q/k/v inputs) -----------------------------
    def rel_mha(self, prefix, x, mask, pos_emb, cache):
        w = self.w
        H = self.conf.attention_heads
        D = self.conf.encoder_dim
        dk = D // H
        B = x.shape[0]
        x = self.ada(prefix, x)
        q = linear(x, w[prefix + ".linear_q.weight"], w[prefix + ".linear_q.bias"]).reshape(B, -1, H, dk).transpose(1, 2)
        k = linear(x, w[prefix + ".linear_k.weight"], w[prefix + ".linear_k.bias"]).reshape(B, -1, H, dk).transpose(1, 2)
        v = linear(x, w[prefix + ".linear_v.weight"], w[prefix + ".linear_v.bias"]).reshape(B, -1, H, dk).transpose(1, 2)
        if cache is not None and cache.numel() > 0:
            key_cache, value_cache = torch.split(cache, cache.shape[-1] // 2, dim=-1)
            k = torch.cat([key_cache, k], dim=2)
            v = torch.cat([value_cache, v], dim=2)
        new_cache = torch.cat((k, v), dim=-1)
        p = linear(pos_emb, w[prefix + ".linear_pos.weight"], w[prefix + ".linear_pos.bias"])
        p = p.reshape(pos_emb.shape[0], -1, H, dk).transpose(1, 2)
        q_u = q + w[prefix + ".pos_bias_u"].unsqueeze(1)
        q_v = q + w[prefix + ".pos_bias_v"].unsqueeze(1)
        scores = (q_u @ k.transpose(-2, -1) + q_v @ p.transpose(-2, -1)) / math.sqrt(dk)
        if mask is not None and mask.shape[2] > 0:
            m = mask.unsqueeze(1).eq(0)[:, :, :, :scores.shape[-1]]
            scores = scores.masked_fill(m, -float("inf"))
            attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)
        else:
            attn = torch.softmax(scores, dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, -1, D)
        return linear(x, w[prefix + ".linear_out.weight"], w[prefix + ".linear_out.bias"]), new_cache

    # -- squeezeformer/convolution.py:102-163: ada, then zero the PAD frames (mask_pad True = VALID, negated inside) ---
    def conv_module(self, prefix, x, mask_pad, cache):
        x = self.ada(prefix, x)
        inv = None if (mask_pad is None or mask_pad.shape[2] == 0) else ~mask_pad
        return super().conv_module(prefix, x, inv, cache)

    # -- squeezeformer/encoder.py:435-506 (normalize_before False: post-norm) ---------------------------------
    def layer(self, i, x, mask, pos_emb, mask_pad=None, att_cache=None, cnn_cache=None):
        w = self.w
        p = f"encoder.encoders.{i}"

        def ln(name, t):
            return layer_norm(t, w[f"{p}.{name}.weight"], w[f"{p}.{name}.bias"])

        x_att, new_att_cache = self.rel_mha(p + ".self_attn", x, mask, pos_emb, att_cache)
        x = ln("layer_norm1", x + x_att)
        x = ln("layer_norm2", x + self.ffn(p + ".ffn1", x))
        x_conv, new_cnn_cache = self.conv_module(p + ".conv_module", x, mask_pad, cnn_cache)
        x = ln("layer_norm3", x + x_conv)
        x = ln("layer_norm4", x + self.ffn(p + ".ffn2", x))
        return x, new_att_cache, new_cnn_cache

    # -- squeezeformer/time_reduction.py:183-206 (stream) / :61-84 (conv1d) -------------------------------------
    def time_reduction(self, xs, mask, mask_pad):
        w = self.w
        conf = self.conf
        xs = xs.transpose(1, 2)
        xs = xs.masked_fill(mask_pad.eq(0), 0.0)
        wd = w["encoder.time_reduction_layer.dw_conv.weight"]
        K = wd.shape[-1]
        pad = 0 if conf.time_reduction_layer_type == "stream" else max(0, K - 2)
        xs = F.conv1d(xs, wd, w["encoder.time_reduction_layer.dw_conv.bias"], stride=2, padding=pad, groups=xs.shape[1])
        xs = F.conv1d(xs, w["encoder.time_reduction_layer.pw_conv.weight"], w["encoder.time_reduction_layer.pw_conv.bias"])
        xs = xs.transpose(1, 2)
        B, T, D = xs.shape
        mask = mask[:, ::2, ::2]
        mask_pad = mask_pad[:, :, ::2]
        L = mask_pad.shape[-1]
        if L - T < 0:
            xs = xs[:, :L - T, :]
        else:
            xs = torch.cat([xs, torch.zeros(B, L - T, D, dtype=xs.dtype)], dim=1)
        return xs, mask, mask_pad

    # -- squeezeformer/encoder.py:172-236 ------------------------------------------------------------------------
    def encoder_forward(self, xs, xs_lens, decoding_chunk_size=-1, num_decoding_left_chunks=-1):
        conf = self.conf
        xs = xs.to(self.dtype)
        T = xs.shape[1]
        masks = make_non_pad_mask(xs_lens)
        if masks.shape[1] < T:
            masks = F.pad(masks, (0, T - masks.shape[1]), value=False)
        masks = masks.unsqueeze(1)
        xs = self.global_cmvn(xs)
        xs, pos_emb, masks = self.embed(xs, masks, offset=0)
        mask_pad = masks  # True = valid; the conv module negates it (encoder.py:204, convolution.py:127)
        chunk_masks = add_optional_chunk_mask(xs, masks, conf.use_dynamic_chunk, decoding_chunk_size, 0,
                                              num_decoding_left_chunks)
        xs = layer_norm(xs, self.w["encoder.preln.weight"], self.w["encoder.preln.bias"])
        recover = None
        for i in range(conf.num_blocks):
            if conf.reduce_idx is not None and i == conf.reduce_idx:
                recover = (xs, chunk_masks, pos_emb, mask_pad)
                xs, chunk_masks, mask_pad = self.time_reduction(xs, chunk_masks, mask_pad)
                pos_emb = pos_emb[:, ::2, :]
            if conf.recover_idx is not None and i == conf.recover_idx and recover is not None:
                rec_x, rec_masks, rec_pos, rec_pad = recover
                xs = torch.repeat_interleave(xs, 2, dim=1)
                xs = linear(xs, self.w["encoder.time_recover_layer.weight"], self.w["encoder.time_recover_layer.bias"])
                xs = rec_x + xs[:, :rec_x.shape[1], :]
                chunk_masks, pos_emb, mask_pad = rec_masks, rec_pos, rec_pad
            xs, _, _ = self.layer(i, xs, chunk_masks, pos_emb, mask_pad)
        return xs, masks

    # -- squeezeformer/encoder.py:246-258 ------------------------------------------------------------------------
    def calculate_downsampling_factor(self, i):
        conf = self.conf
        if conf.reduce_idx is None:
            return 1
        reduce_exp = 1 if i >= conf.reduce_idx else 0
        recover_exp = 1 if (conf.recover_idx is not None and i >= conf.recover_idx) else 0
        return int(2 ** (reduce_exp - recover_exp))

    # -- squeezeformer/encoder.py:260-383 ------------------------------------------------------------------------
    def encoder_forward_chunk(self, xs, offset, required_cache_size, att_cache, cnn_cache):
        """att_cache (L, H, t, 2 dk) stored at the FULL frame rate for every block (the half-rate blocks write each key twice,
        encoder.py:368, and read every second one, :351); cnn_cache (L, 1, D, lorder)."""
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
        mask_pad = torch.ones(1, 1, xs.shape[1], dtype=torch.bool)
        att_mask = torch.ones(0, 0, 0, dtype=torch.bool)
        recover = None
        max_att_len = 0
        xs = layer_norm(xs, self.w["encoder.preln.weight"], self.w["encoder.preln.bias"])
        for i in range(conf.num_blocks):
            if conf.reduce_idx is not None and i == conf.reduce_idx:
                recover = (xs, att_mask, pos_emb, mask_pad)
                xs, _m, mask_pad = self.time_reduction(xs, torch.ones(1, xs.shape[1], xs.shape[1], dtype=torch.bool), mask_pad)
                pos_emb = pos_emb[:, ::2, :]
            if conf.recover_idx is not None and i == conf.recover_idx and recover is not None:
                rec_x, att_mask, rec_pos, rec_pad = recover
                xs = torch.repeat_interleave(xs, 2, dim=1)
                xs = linear(xs, self.w["encoder.time_recover_layer.weight"], self.w["encoder.time_recover_layer.bias"])
                xs = rec_x + xs[:, :rec_x.shape[1], :]
                pos_emb, mask_pad = rec_pos, rec_pad
            factor = self.calculate_downsampling_factor(i)
            ac = None
            if att_cache.numel() > 0:
                ac = att_cache[i:i + 1][:, :, ::factor, :][:, :, :pos_emb.shape[1] - xs.shape[1], :]
            cc = cnn_cache[i] if cnn_cache.numel() > 0 else None
            xs, new_att, new_cnn = self.layer(i, xs, None, pos_emb, None, ac, cc)
            cached_att = new_att[:, :, next_cache_start // factor:, :]
            cached_att = torch.repeat_interleave(cached_att, factor, dim=2)
            if i == 0:
                max_att_len = cached_att.shape[2]
            r_att.append(cached_att[:, :, :max_att_len, :])
            r_cnn.append(new_cnn.unsqueeze(0))
        return xs, torch.cat(r_att, dim=0), torch.cat(r_cnn, dim=0)

    @torch.no_grad()
    def get_encoder_out_chunk(self, speech, offset, required_cache_size, att_cache, cnn_cache, return_logits=False):
        xs, att_cache, cnn_cache = self.encoder_forward_chunk(speech, offset, required_cache_size, att_cache, cnn_cache)
        out = self.ctc_logits(xs) if return_logits else self.ctc_softmax(xs)
        return out, att_cache, cnn_cache
