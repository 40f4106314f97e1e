"""CPU oracle for the Efficient-Conformer hot path (TEST INFRASTRUCTURE ONLY -- see oracle/conformer_oracle.py header).

PyTorch-CPU restatement of ppasr/model_utils/efficient_conformer/ (paths relative to /root/reference, yeyupiaoling/PPASR
@ c8bb3b96): Conv2dSubsampling4 -> 12 pre-norm macaron blocks (conformer/encoder.py:346-431) where
  * blocks in group_layer_idx use GroupedRelPositionMultiHeadedAttention (attention.py:40-79,128-193: q/k/v/p padded to
    a multiple of group_size frames and re-viewed as T/3 tokens of 4 x 192 features, scale 1/sqrt(192), mask [::3, ::3],
    linear_pos WITH bias),
  * block stride_layer_idx is a StrideConformerEncoderLayer (encoder.py:455-548): its depthwise conv has stride 2 and the
    residual goes through AvgPool1D(k2, s2, ceil_mode) (encoder.py:171-172,523-526); afterwards masks / pos_emb are
    strided (encoder.py:255-260) and the conv kernel shrinks 15 -> 7 (encoder.py:123-128),
-> after_norm -> CTC. `efficient_conf` in configs/efficient_conformer.yml is swallowed by **kwargs (encoder.py:55), so the
constructor defaults (stride_layer_idx 3, stride 2, group_layer_idx 0..3, group_size 3, stride_kernel True) apply.

PARITY STATUS: **pinned against the reference's own model code** run on tests/golden/paddle_shim (see the Conformer oracle
header): tests/golden/encoder_golden_efficient_conformer_{offline,stream}.npz come from the unmodified
EfficientConformerEncoder (grouped attention in blocks 0-1, stride block 1, batch_norm / layer_norm conv norm);
tests/test_encoder_golden_cpu.py checks offline logits (padded frames included) and the forward_chunk chain with its caches
to 5e-5 (observed 1e-5). The reference's StrideConformerEncoderLayer also owns an unused `concat_linear` parameter
(encoder.py:453, read only when concat_after=True): it appears in .pdparams files and is ignored by name here.
"""
import math
import torch
import torch.nn.functional as F

from oracle.conformer_oracle import (ConformerConf, ConformerOracle, add_optional_chunk_mask, layer_norm, linear,
                                     make_non_pad_mask, swish)


class EfficientConformerConf(ConformerConf):
    def __init__(self, stride_layer_idx=3, stride=2, group_layer_idx=(0, 1, 2, 3), group_size=3, stride_kernel=True, **kw):
        super().__init__(**kw)
        self.stride_layer_idx = stride_layer_idx
        self.stride = stride
        self.group_layer_idx = tuple(group_layer_idx)
        self.group_size = group_size
        self.stride_kernel = stride_kernel
        # encoder.py:123-128
        k0 = self.cnn_module_kernel
        self.cnn_module_kernels = [k0, (k0 // stride) if stride_kernel else k0]


class EfficientConformerOracle(ConformerOracle):
    def layer_kernel(self, i):
        conf = self.conf
        return conf.cnn_module_kernels[1 if (conf.stride_layer_idx is not None and i > conf.stride_layer_idx) else 0]

    # -- efficient_conformer/attention.py:128-193 (+ pad4group :40-79, forward_attention :81-126) -----------------
    def grouped_mha(self, prefix, x, mask, pos_emb):
        w = self.w
        H = self.conf.attention_heads
        D = self.conf.output_size
        dk = D // H
        gs = self.conf.group_size
        B = x.shape[0]
        q = linear(x, w[prefix + ".linear_q.weight"], w[prefix + ".linear_q.bias"])
        k = linear(x, w[prefix + ".linear_k.weight"], w[prefix + ".linear_k.bias"])
        v = linear(x, w[prefix + ".linear_v.weight"], w[prefix + ".linear_v.bias"])
        p = linear(pos_emb, w[prefix + ".linear_pos.weight"], w[prefix + ".linear_pos.bias"])  # (1, T2, D)
        T = q.shape[1]
        pad = (gs - T % gs) % gs
        # (B,T,D) == (B,T,H,dk) flattened: padding frames and re-viewing as (B, T/gs, H, dk*gs) is a plain reshape
        qg = F.pad(q, (0, 0, 0, pad)).reshape(B, -1, H, dk * gs).transpose(1, 2)
        kg = F.pad(k, (0, 0, 0, pad)).reshape(B, -1, H, dk * gs).transpose(1, 2)
        vg = F.pad(v, (0, 0, 0, pad)).reshape(B, -1, H, dk * gs).transpose(1, 2)
        padp = (gs - p.shape[1] % gs) % gs
        pg = F.pad(p, (0, 0, 0, padp)).reshape(p.shape[0], -1, H, dk * gs).transpose(1, 2)
        if mask is not None and mask.shape[2] > 0:
            mask = mask[:, ::gs, ::gs]
        q_u = qg + w[prefix + ".pos_bias_u"].unsqueeze(1)  # (B,H,Tg,192) + (H,1,192)
        q_v = qg + w[prefix + ".pos_bias_v"].unsqueeze(1)
        scores = (q_u @ kg.transpose(-2, -1) + q_v @ pg.transpose(-2, -1)) / math.sqrt(dk * gs)
        if mask is not None and mask.shape[2] > 0:
            m = mask.unsqueeze(1).eq(0)[:, :, :, :scores.shape[-1]]
            scores = scores.masked_fill(m, -float("inf"))
            attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)
        else:
            attn = torch.softmax(scores, dim=-1)
        o = (attn @ vg).transpose(1, 2).reshape(B, -1, D)
        o = o[:, :o.shape[1] - pad]
        return linear(o, w[prefix + ".linear_out.weight"], w[prefix + ".linear_out.bias"])

    # plain rel-pos MHA of the later blocks: conformer attention (linear_pos without bias, conformer/attention.py:192)
    # -- efficient_conformer/convolution.py:80-138 ---------------------------------------------------------------------
    def eff_conv_module(self, prefix, x, mask_pad, K, stride):
        w = self.w
        conf = self.conf
        lorder = K - 1 if conf.causal else 0
        x = x.transpose(1, 2)
        x = x.masked_fill(~mask_pad, 0.0)
        if lorder > 0:
            x = F.pad(x, (lorder, 0), "constant", 0.0)
        x = F.conv1d(x, w[prefix + ".pointwise_conv1.weight"], w[prefix + ".pointwise_conv1.bias"])
        x = F.glu(x, dim=1)
        pad = 0 if conf.causal else (K - 1) // 2
        x = F.conv1d(x, w[prefix + ".depthwise_conv.weight"], w[prefix + ".depthwise_conv.bias"], stride=stride, padding=pad,
                     groups=x.shape[1])
        if conf.cnn_module_norm == "layer_norm":
            x = layer_norm(x.transpose(1, 2), w[prefix + ".norm.weight"], w[prefix + ".norm.bias"]).transpose(1, 2)
        else:
            x = F.batch_norm(x, w[prefix + ".norm._mean"], w[prefix + ".norm._variance"], w[prefix + ".norm.weight"],
                             w[prefix + ".norm.bias"], training=False, eps=1e-5)
        x = swish(x)
        x = F.conv1d(x, w[prefix + ".pointwise_conv2.weight"], w[prefix + ".pointwise_conv2.bias"])
        if mask_pad.shape[2] != x.shape[2]:
            mask_pad = mask_pad[:, :, ::stride]
        x = x.masked_fill(~mask_pad, 0.0)
        return x.transpose(1, 2)

    # -- conformer/encoder.py:346-431 and efficient_conformer/encoder.py:455-548 ----------------------------------------
    def eff_layer(self, i, x, mask, pos_emb, mask_pad):
        w = self.w
        conf = self.conf
        p = f"encoder.encoders.{i}"

        def ln(name, t):
            return layer_norm(t, w[f"{p}.{name}.weight"], w[f"{p}.{name}.bias"])

        x = x + 0.5 * self.ffn(p + ".feed_forward_macaron", ln("norm_ff_macaron", x))
        if i in conf.group_layer_idx:
            x = x + self.grouped_mha(p + ".self_attn", ln("norm_mha", x), mask, pos_emb)
        else:
            x = x + self.rel_mha(p + ".self_attn", ln("norm_mha", x), mask, pos_emb, None)[0]
        strided = conf.stride_layer_idx is not None and i == conf.stride_layer_idx
        residual = x
        xc = self.eff_conv_module(p + ".conv_module", ln("norm_conv", x), mask_pad, self.layer_kernel(i),
                                  conf.stride if strided else 1)
        if strided:  # AvgPool1D(kernel 2, stride 2, ceil_mode=True, exclusive) on the residual
            residual = F.avg_pool1d(residual.transpose(1, 2), conf.stride, conf.stride, 0, ceil_mode=True,
                                    count_include_pad=False).transpose(1, 2)
        x = residual + xc
        x = x + 0.5 * self.ffn(p + ".feed_forward", ln("norm_ff", x))
        return ln("norm_final", x)

    # -- efficient_conformer/encoder.py:212-264 ------------------------------------------------------------------------------
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
        mask_pad = masks
        chunk_masks = add_optional_chunk_mask(xs, masks, conf.use_dynamic_chunk, decoding_chunk_size, 0,
                                              num_decoding_left_chunks)
        for i in range(conf.num_blocks):
            xs = self.eff_layer(i, xs, chunk_masks, pos_emb, mask_pad)
            if conf.stride_layer_idx is not None and i == conf.stride_layer_idx:
                s = conf.stride
                masks = masks[:, :, ::s]
                chunk_masks = chunk_masks[:, ::s, ::s]
                mask_pad = masks
                pos_emb = pos_emb[:, ::s, :]
        xs = layer_norm(xs, self.w["encoder.after_norm.weight"], self.w["encoder.after_norm.bias"])
        return xs, masks

    # -- efficient_conformer/encoder.py:205-210 -------------------------------------------------------------------------
    def calculate_downsampling_factor(self, i):
        conf = self.conf
        return conf.stride if (conf.stride_layer_idx is not None and i > conf.stride_layer_idx) else 1

    # grouped attention with a K/V cache (attention.py:153-160): the cache is concatenated BEFORE pad4group, so the groups of
    # the keys are re-formed from the first cached frame for every chunk
    def grouped_mha_chunk(self, prefix, x, pos_emb, cache):
        w = self.w
        H, D = self.conf.attention_heads, self.conf.output_size
        dk, gs = D // H, self.conf.group_size
        B = x.shape[0]
        q = linear(x, w[prefix + ".linear_q.weight"], w[prefix + ".linear_q.bias"]).reshape(B, -1, H, dk).transpose(1, 2)
        k = linear(x, w[prefix + ".linear_k.weight"], w[prefix + ".linear_k.bias"]).reshape(B, -1, H, dk).transpose(1, 2)
        v = linear(x, w[prefix + ".linear_v.weight"], w[prefix + ".linear_v.bias"]).reshape(B, -1, H, dk).transpose(1, 2)
        p = linear(pos_emb, w[prefix + ".linear_pos.weight"], w[prefix + ".linear_pos.bias"])
        if cache is not None and cache.numel() > 0:
            kc, vc = torch.split(cache, cache.shape[-1] // 2, dim=-1)
            k = torch.cat([kc, k], dim=2)
            v = torch.cat([vc, v], dim=2)
        new_cache = torch.cat((k, v), dim=-1)

        def group(t):  # (B,H,T,dk) -> pad T to a multiple of gs -> (B,H,T/gs,dk*gs) through the (B,T,H*dk) view
            T = t.shape[2]
            pad = (gs - T % gs) % gs
            t = F.pad(t, (0, 0, 0, pad))
            return t.transpose(1, 2).reshape(B, -1, H, dk * gs).transpose(1, 2), pad

        qg, pad_q = group(q)
        kg, _ = group(k)
        vg, _ = group(v)
        padp = (gs - p.shape[1] % gs) % gs
        pg = F.pad(p, (0, 0, 0, padp)).reshape(p.shape[0], -1, H, dk * gs).transpose(1, 2)
        q_u = qg + w[prefix + ".pos_bias_u"].unsqueeze(1)
        q_v = qg + w[prefix + ".pos_bias_v"].unsqueeze(1)
        scores = (q_u @ kg.transpose(-2, -1) + q_v @ pg.transpose(-2, -1)) / math.sqrt(dk * gs)
        attn = torch.softmax(scores, dim=-1)  # fake (0,0,0) mask in forward_chunk
        o = (attn @ vg).transpose(1, 2).reshape(B, -1, D)
        o = o[:, :o.shape[1] - pad_q]
        return linear(o, w[prefix + ".linear_out.weight"], w[prefix + ".linear_out.bias"]), new_cache

    # efficient_conformer/convolution.py:80-138 with a cache (cache[:, :, -lorder:], :108)
    def eff_conv_module_chunk(self, prefix, x, cache, K, stride):
        w = self.w
        conf = self.conf
        lorder = K - 1
        x = x.transpose(1, 2)
        if cache is None or cache.numel() == 0:
            x = F.pad(x, (lorder, 0), "constant", 0.0)
        else:
            x = torch.cat((cache[:, :, -lorder:], x), dim=2)
        new_cache = x[:, :, -lorder:]
        x = F.conv1d(x, w[prefix + ".pointwise_conv1.weight"], w[prefix + ".pointwise_conv1.bias"])
        x = F.glu(x, dim=1)
        x = F.conv1d(x, w[prefix + ".depthwise_conv.weight"], w[prefix + ".depthwise_conv.bias"], stride=stride, groups=x.shape[1])
        if conf.cnn_module_norm == "layer_norm":
            x = layer_norm(x.transpose(1, 2), w[prefix + ".norm.weight"], w[prefix + ".norm.bias"]).transpose(1, 2)
        else:
            x = F.batch_norm(x, w[prefix + ".norm._mean"], w[prefix + ".norm._variance"], w[prefix + ".norm.weight"],
                             w[prefix + ".norm.bias"], training=False, eps=1e-5)
        x = swish(x)
        x = F.conv1d(x, w[prefix + ".pointwise_conv2.weight"], w[prefix + ".pointwise_conv2.bias"])
        return x.transpose(1, 2), new_cache

    # -- efficient_conformer/encoder.py:266-394 (streaming model only: causal conv) ----------------------------------------
    @torch.no_grad()
    def encoder_forward_chunk(self, xs, offset, required_cache_size, att_cache, cnn_cache):
        """att_cache (L, H, t, 2 dk) at the full frame rate (half-rate blocks repeat every key, :368, and read every second
        one, :351); cnn_cache (L, 1, D, cnn_module_kernel - 1), left-padded with zeros for the kernel-7 blocks (:370-372).
        `offset` counts OUTPUT frames (8 per 16-frame chunk) and is scaled back to the full rate first (:304)."""
        conf = self.conf
        w = self.w
        assert xs.shape[0] == 1 and conf.causal
        has_stride = conf.stride_layer_idx is not None
        offset = offset * (conf.stride if has_stride else 1)
        xs = self.global_cmvn(xs.to(self.dtype))
        tmp_masks = torch.ones(1, 1, xs.shape[1], dtype=torch.bool)
        xs, pos_emb, _ = self.embed(xs, tmp_masks, offset=offset)
        cache_t1 = att_cache.shape[2] if att_cache.dim() == 4 else 0
        attention_key_size = cache_t1 + xs.shape[1]
        pos_emb = self.position_encoding(offset=offset - cache_t1, size=attention_key_size)
        if required_cache_size < 0:
            next_cache_start = 0
        elif required_cache_size == 0:
            next_cache_start = attention_key_size
        else:
            next_cache_start = max(attention_key_size - required_cache_size, 0)
        r_att, r_cnn = [], []
        kmax = conf.cnn_module_kernel
        for i in range(conf.num_blocks):
            p = f"encoder.encoders.{i}"
            factor = self.calculate_downsampling_factor(i)
            ac = att_cache[i:i + 1, :, ::factor, :] if att_cache.numel() > 0 else None
            cc = cnn_cache[i] if cnn_cache.numel() > 0 else None

            def ln(name, t):
                return layer_norm(t, w[f"{p}.{name}.weight"], w[f"{p}.{name}.bias"])

            xs = xs + 0.5 * self.ffn(p + ".feed_forward_macaron", ln("norm_ff_macaron", xs))
            if i in conf.group_layer_idx:
                x_att, new_att = self.grouped_mha_chunk(p + ".self_attn", ln("norm_mha", xs), pos_emb, ac)
            else:
                x_att, new_att = self.rel_mha(p + ".self_attn", ln("norm_mha", xs), None, pos_emb, ac)
            xs = xs + x_att
            strided = has_stride and i == conf.stride_layer_idx
            residual = xs
            xc, new_cnn = self.eff_conv_module_chunk(p + ".conv_module", ln("norm_conv", xs), cc, self.layer_kernel(i),
                                                     conf.stride if strided else 1)
            if strided:
                residual = F.avg_pool1d(residual.transpose(1, 2), conf.stride, conf.stride, 0, ceil_mode=True,
                                        count_include_pad=False).transpose(1, 2)
                pos_emb = pos_emb[:, ::conf.stride, :]
            xs = residual + xc
            xs = xs + 0.5 * self.ffn(p + ".feed_forward", ln("norm_ff", xs))
            xs = ln("norm_final", xs)
            new_att = new_att[:, :, next_cache_start // factor:, :]
            new_att = torch.repeat_interleave(new_att, factor, dim=2)
            new_cnn = F.pad(new_cnn.unsqueeze(0), (kmax - 1 - new_cnn.shape[2], 0))
            r_att.append(new_att)
            r_cnn.append(new_cnn)
        xs = layer_norm(xs, w["encoder.after_norm.weight"], w["encoder.after_norm.bias"])
        return xs, torch.cat(r_att, dim=0), torch.cat(r_cnn, dim=0)

    @torch.no_grad()
    def get_encoder_out_chunk(self, speech, offset, required_cache_size, att_cache, cnn_cache, return_logits=False):
        xs, att_cache, cnn_cache = self.encoder_forward_chunk(speech, offset, required_cache_size, att_cache, cnn_cache)
        out = self.ctc_logits(xs) if return_logits else self.ctc_softmax(xs)
        return out, att_cache, cnn_cache
