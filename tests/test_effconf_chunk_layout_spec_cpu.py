"""Executable specification of the DEVICE LAYOUT planned for the Efficient-Conformer `forward_chunk` GPU path,
checked against the reference code's own streaming outputs (tests/golden/encoder_golden_efficient_conformer_stream.npz).

The oracle restates the reference literally (caches in the reference layout, keys re-grouped from the first cached frame for
every chunk, half-rate caches stored repeated x2). A GPU implementation wants something else: caches that are only ever
APPENDED to, in the operand layout of the existing kernels. This file states that layout in NumPy/torch and proves the claims
it rests on, so that the kernels can be written against it:
  * grouped blocks keep K and V in the `EpiQKVGrouped` layout [H][group][192] with group = absolute_frame // 3 -- valid because
    PPASRPredictor streams with an unlimited history (required_cache_size < 0), so the first cached frame is always frame 0;
    a new frame f is written at group f // 3, feature offset (f % 3) * 256 + c, and the partially filled last group reads as
    zeros for its missing frames (caches start zeroed);
  * the positional operand of a grouped block is linear_pos(pe[j * rate]) for the t_total keys, ZERO beyond t_total inside the
    last group (pad4group pads p after the projection, attention.py:73-77) -- it must be rebuilt (or masked) per chunk;
  * queries are grouped from the chunk start, independent of the key grouping, zero-padded BEFORE pos_bias_u/v are added;
  * blocks after the stride block keep their caches at half rate with key j at full-rate position 2 j; conv caches are the last
    K-1 rows of the conv-module input (post norm_conv, pre pointwise_conv1), 14 rows before / 6 rows after the stride block;
  * the stride block's causal stride-2 conv over [cache | chunk] is aligned with the offline run because chunks start at even
    frames, and its residual goes through AvgPool1D(2, 2, ceil).
"""
import ast
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle.conformer_oracle import layer_norm, linear, stream_windows, swish
from oracle.efficient_conformer_oracle import EfficientConformerConf, EfficientConformerOracle
from ppasr_b200 import weights as W

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class DeviceLayoutSpec(EfficientConformerOracle):
    """Streaming state the way the GPU would hold it; only the attention / conv cache handling differs from the oracle."""

    def reset(self, group_cap=64, frame_cap=192):
        c = self.conf
        H, dk = c.attention_heads, c.output_size // c.attention_heads
        self.frames = [0] * c.num_blocks            # keys cached per block, at the block's own rate
        self.kg = [torch.zeros(H, group_cap, 3 * dk) for _ in range(c.num_blocks)]   # grouped blocks: [H'][group][192]
        self.vg = [torch.zeros(H, group_cap, 3 * dk) for _ in range(c.num_blocks)]
        self.kk = [torch.zeros(H, frame_cap, dk) for _ in range(c.num_blocks)]       # plain blocks: [H][frame][64]
        self.vv = [torch.zeros(H, frame_cap, dk) for _ in range(c.num_blocks)]
        self.cnn = [torch.zeros(self.layer_kernel(i) - 1, c.output_size) for i in range(c.num_blocks)]
        self.offset = 0                              # full-rate subsampled frames consumed so far

    # -- the EpiQKVGrouped scatter: frame f, feature c of a [frames, 256] matrix -> [H'][f // 3][192] ----------------------
    @staticmethod
    def scatter_grouped(dst, rows, frame0):
        D = rows.shape[1]
        for i in range(rows.shape[0]):
            f = frame0 + i
            row768 = torch.zeros(3 * D)
            row768[(f % 3) * D:(f % 3 + 1) * D] = rows[i]
            dst[:, f // 3, :] += row768.reshape(-1, dst.shape[2])   # += into zeros == write of this frame's 256 features

    def grouped_step(self, i, prefix, y, rate):
        w, c = self.w, self.conf
        H, D = c.attention_heads, c.output_size
        n, t0 = y.shape[0], self.frames[i]
        q = linear(y, w[prefix + ".linear_q.weight"], w[prefix + ".linear_q.bias"])
        k = linear(y, w[prefix + ".linear_k.weight"], w[prefix + ".linear_k.bias"])
        v = linear(y, w[prefix + ".linear_v.weight"], w[prefix + ".linear_v.bias"])
        self.scatter_grouped(self.kg[i], k, t0)      # append in place
        self.scatter_grouped(self.vg[i], v, t0)
        T2 = t0 + n
        G2 = (T2 + 2) // 3
        # positional operand: rows of the keys' absolute positions, zero beyond T2 inside the last group
        pe = self.pe[0, torch.arange(T2) * rate]
        p = linear(pe, w[prefix + ".linear_pos.weight"], w[prefix + ".linear_pos.bias"])
        pg = torch.zeros(H, G2, 3 * D // H)
        self.scatter_grouped(pg, p, 0)
        # queries: grouped from the chunk start, zero padded, THEN the biases
        Gq = (n + 2) // 3
        qg = torch.zeros(H, Gq, 3 * D // H)
        self.scatter_grouped(qg, q, 0)
        q_u = qg + w[prefix + ".pos_bias_u"].unsqueeze(1)
        q_v = qg + w[prefix + ".pos_bias_v"].unsqueeze(1)
        kg, vg = self.kg[i][:, :G2], self.vg[i][:, :G2]
        s = (q_u @ kg.transpose(-2, -1) + q_v @ pg.transpose(-2, -1)) / math.sqrt(3 * D // H)
        o = torch.softmax(s, dim=-1) @ vg                                 # [H', Gq, 192]
        o = o.transpose(0, 1).reshape(Gq * 3, D)[:n]                      # back to frames, padded query frames dropped
        self.frames[i] = T2
        return linear(o, w[prefix + ".linear_out.weight"], w[prefix + ".linear_out.bias"])

    def plain_step(self, i, prefix, y, rate):
        w, c = self.w, self.conf
        H, D = c.attention_heads, c.output_size
        dk = D // H
        n, t0 = y.shape[0], self.frames[i]
        q = linear(y, w[prefix + ".linear_q.weight"], w[prefix + ".linear_q.bias"]).reshape(n, H, dk).transpose(0, 1)
        k = linear(y, w[prefix + ".linear_k.weight"], w[prefix + ".linear_k.bias"]).reshape(n, H, dk).transpose(0, 1)
        v = linear(y, w[prefix + ".linear_v.weight"], w[prefix + ".linear_v.bias"]).reshape(n, H, dk).transpose(0, 1)
        self.kk[i][:, t0:t0 + n] = k
        self.vv[i][:, t0:t0 + n] = v
        T2 = t0 + n
        pe = self.pe[0, torch.arange(T2) * rate]
        p = linear(pe, w[prefix + ".linear_pos.weight"]).reshape(T2, H, dk).transpose(0, 1)   # no bias (conformer/attention.py:192)
        q_u = q + w[prefix + ".pos_bias_u"].unsqueeze(1)
        q_v = q + w[prefix + ".pos_bias_v"].unsqueeze(1)
        s = (q_u @ self.kk[i][:, :T2].transpose(-2, -1) + q_v @ p.transpose(-2, -1)) / math.sqrt(dk)
        o = (torch.softmax(s, dim=-1) @ self.vv[i][:, :T2]).transpose(0, 1).reshape(n, D)
        self.frames[i] = T2
        return linear(o, w[prefix + ".linear_out.weight"], w[prefix + ".linear_out.bias"])

    def conv_step(self, i, prefix, yc, stride):
        w, c = self.w, self.conf
        K = self.layer_kernel(i)
        cat = torch.cat([self.cnn[i], yc], 0)                      # [K-1 + n, D]
        self.cnn[i] = cat[-(K - 1):].clone()
        x = cat.t().unsqueeze(0)
        x = F.glu(F.conv1d(x, w[prefix + ".pointwise_conv1.weight"], w[prefix + ".pointwise_conv1.bias"]), dim=1)
        x = F.conv1d(x, w[prefix + ".depthwise_conv.weight"], w[prefix + ".depthwise_conv.bias"], stride=stride, groups=x.shape[1])
        if c.cnn_module_norm == "layer_norm":
            x = layer_norm(x.transpose(1, 2), w[prefix + ".norm.weight"], w[prefix + ".norm.bias"]).transpose(1, 2)
        else:
            x = F.batch_norm(x, w[prefix + ".norm._mean"], w[prefix + ".norm._variance"], w[prefix + ".norm.weight"],
                             w[prefix + ".norm.bias"], training=False, eps=1e-5)
        x = F.conv1d(swish(x), w[prefix + ".pointwise_conv2.weight"], w[prefix + ".pointwise_conv2.bias"])
        return x[0].t()

    @torch.no_grad()
    def step(self, feats):
        """One window of fbank frames [t, 80] -> logits of its output frames."""
        c, w = self.conf, self.w
        xs = self.global_cmvn(torch.from_numpy(feats)[None].to(self.dtype))
        xs, _, _ = self.embed(xs, torch.ones(1, 1, xs.shape[1], dtype=torch.bool), offset=self.offset)
        x = xs[0]
        n_full = x.shape[0]
        assert self.offset % 2 == 0 or c.stride_layer_idx is None   # chunk starts stay aligned with the stride-2 grid
        rate = 1
        for i in range(c.num_blocks):
            p = f"encoder.encoders.{i}"

            def ln(name, t):
                return layer_norm(t, w[f"{p}.{name}.weight"], w[f"{p}.{name}.bias"])

            x = x + 0.5 * self.ffn(p + ".feed_forward_macaron", ln("norm_ff_macaron", x))
            y = ln("norm_mha", x)
            x = x + (self.grouped_step(i, p + ".self_attn", y, rate) if i in c.group_layer_idx
                     else self.plain_step(i, p + ".self_attn", y, rate))
            strided = c.stride_layer_idx is not None and i == c.stride_layer_idx
            xc = self.conv_step(i, p + ".conv_module", ln("norm_conv", x), c.stride if strided else 1)
            if strided:
                x = F.avg_pool1d(x.t().unsqueeze(0), c.stride, c.stride, 0, ceil_mode=True, count_include_pad=False)[0].t()
                rate *= c.stride
            x = x + xc
            x = x + 0.5 * self.ffn(p + ".feed_forward", ln("norm_ff", x))
            x = ln("norm_final", x)
        self.offset += n_full
        x = layer_norm(x, w["encoder.after_norm.weight"], w["encoder.after_norm.bias"])
        return self.ctc_logits(x[None])[0]


def test_device_layout_spec_reproduces_reference_streaming_outputs():
    g = np.load(os.path.join(GOLDEN, "encoder_golden_efficient_conformer_stream.npz"))
    cfgd = ast.literal_eval(str(g["cfg"]))
    cfg = W.EfficientConformerConfig(**cfgd)
    assert cfg.group_layer_idx and cfg.stride_layer_idx is not None       # the fixture exercises both mechanisms
    spec = DeviceLayoutSpec(EfficientConformerConf(**cfgd), W.init_efficient_conformer_weights(cfg, seed=int(g["seed"])))
    spec.reset()
    cf = g["chunk_feats"]
    outs = [spec.step(cf[a:b]).numpy() for (a, b) in stream_windows(cf.shape[0], is_end=True)]
    got = np.concatenate(outs, 0)
    ref = g["chunk_logits"]
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=0, atol=5e-5)
    # the caches only ever grew: grouped blocks hold ceil(frames / 3) groups, later blocks half as many frames
    n_full = spec.offset
    for i in range(cfg.num_blocks):
        expect = n_full if i <= cfg.stride_layer_idx else (n_full + 1) // 2
        assert spec.frames[i] == expect
    # conv caches in the reference layout: last K-1 rows, the kernel-7 blocks left-padded with zeros to 14 (encoder.py:370-372)
    ref_cnn = g["chunk_cnn_cache"]                      # [L, 1, D, 14]
    for i in range(cfg.num_blocks):
        mine = spec.cnn[i].t().numpy()                  # [D, K-1]
        pad = ref_cnn.shape[3] - mine.shape[1]
        np.testing.assert_allclose(ref_cnn[i, 0, :, pad:], mine, rtol=0, atol=1e-5)
        assert np.all(ref_cnn[i, 0, :, :pad] == 0)
