"""CPU oracle for the DeepSpeech2 hot path (TEST INFRASTRUCTURE ONLY -- see oracle/conformer_oracle.py header).

PyTorch-CPU restatement of ppasr/model_utils/deepspeech2/ (paths relative to /root/reference, yeyupiaoling/PPASR @
c8bb3b96): GlobalCMVN -> Conv2dSubsampling4Pure (2 x Conv2d k3 s2 + ReLU, 32 channels, conv.py:8-21) -> num_rnn_layers x
{paddle.nn.LSTM | GRU (1 layer, 'forward' when streaming else 'bidirect', run with sequence_length = subsampled lengths)
-> LayerNorm} (encoder.py:37-55,86-93) -> CTC Linear + softmax (model.py:62-65; the CTC module is called `decoder`).

paddle.nn.LSTM / GRU are third-party arithmetic (PaddlePaddle 2.5.1, docs/install.md:5-8; not installable here). Their
published cell equations (python/paddle/nn/layer/rnn.py LSTMCell / GRUCell) are restated explicitly below:
    LSTM  gates = W_ih x + b_ih + W_hh h + b_hh, chunks (i, f, g, o):  c' = sig(f) c + sig(i) tanh(g);  h' = sig(o) tanh(c')
    GRU   x_g = W_ih x + b_ih, h_g = W_hh h + b_hh, chunks (r, z, c):  r = sig(x_r + h_r); z = sig(x_z + h_z);
          c = tanh(x_c + r * h_c);  h' = (h - c) * z + c
and the sequence_length rule of paddle.nn.RNN: steps t >= len keep the state and emit zeros; the reverse direction walks
t = len-1 .. 0. PARITY STATUS: the WIRING (conv reshape, subsampled lengths, per-layer state-box split/concat, LayerNorm
placement, CTC head) is pinned against the reference's own CRNNEncoder run on tests/golden/paddle_shim
(tests/golden/encoder_golden_deepspeech2_*.npz: LSTM/GRU x offline-bidirectional/streaming-forward, chunk chains with
their final state boxes; tests/test_encoder_golden_cpu.py, 5e-5). The CELL arithmetic of paddle.nn.LSTM/GRU stays unpinned
(third-party, not installable): the shim and tests/test_oracle_cpu.py both use torch.nn.LSTM/GRU with packed sequences,
which implement the same published equations and gate order.
"""
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from oracle.conformer_oracle import layer_norm, linear


class DeepSpeech2Conf:
    """configs/deepspeech2.yml encoder_conf + the streaming switch (deepspeech2/model.py:40)."""

    def __init__(self, input_dim=80, vocab_size=4233, num_rnn_layers=5, rnn_size=1024, use_gru=False, streaming=False,
                 **_ignored):
        self.input_dim = input_dim
        self.vocab_size = vocab_size
        self.num_rnn_layers = num_rnn_layers
        self.rnn_size = rnn_size
        self.use_gru = bool(use_gru)
        self.streaming = bool(streaming)
        self.rnn_direction = "forward" if streaming else "bidirect"
        self.num_directions = 1 if streaming else 2


def lstm_cell(x_g, h, c, w_hh, b_hh):
    g = x_g + h @ w_hh.t() + b_hh
    i, f, gg, o = g.chunk(4, dim=-1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    return torch.sigmoid(o) * torch.tanh(c2), c2


def gru_cell(x_g, h, w_hh, b_hh):
    h_g = h @ w_hh.t() + b_hh
    x_r, x_z, x_c = x_g.chunk(3, dim=-1)
    h_r, h_z, h_c = h_g.chunk(3, dim=-1)
    r = torch.sigmoid(x_r + h_r)
    z = torch.sigmoid(x_z + h_z)
    cand = torch.tanh(x_c + r * h_c)
    return (h - cand) * z + cand


class DeepSpeech2Oracle:
    def __init__(self, conf: DeepSpeech2Conf, weights: Dict[str, np.ndarray], dtype=torch.float32):
        self.conf = conf
        self.dtype = dtype
        self.w = {k: torch.from_numpy(np.asarray(v)).to(dtype) for k, v in weights.items()}

    # -- deepspeech2/conv.py:16-21 ----------------------------------------------------------------------------
    def conv(self, x, x_len):
        w = self.w
        x = x.unsqueeze(1)
        x = F.relu(F.conv2d(x, w["encoder.conv.conv.0.weight"], w["encoder.conv.conv.0.bias"], stride=2))
        x = F.relu(F.conv2d(x, w["encoder.conv.conv.2.weight"], w["encoder.conv.conv.2.bias"], stride=2))
        b, c, t, f = x.shape
        x = x.permute(0, 2, 1, 3).reshape(b, t, c * f)
        return x, ((x_len - 1) // 2 - 1) // 2

    # -- one direction of paddle.nn.LSTM / GRU with sequence_length -----------------------------------------------
    def run_direction(self, prefix, suffix, x, lens, h0, c0, reverse):
        w = self.w
        B, T, _ = x.shape
        H = self.conf.rnn_size
        w_ih, w_hh = w[f"{prefix}.weight_ih_l0{suffix}"], w[f"{prefix}.weight_hh_l0{suffix}"]
        b_ih, b_hh = w[f"{prefix}.bias_ih_l0{suffix}"], w[f"{prefix}.bias_hh_l0{suffix}"]
        xg = x @ w_ih.t() + b_ih
        h = torch.zeros(B, H, dtype=x.dtype) if h0 is None else h0.clone()
        c = torch.zeros(B, H, dtype=x.dtype) if c0 is None else c0.clone()
        out = torch.zeros(B, T, H, dtype=x.dtype)
        steps = range(T - 1, -1, -1) if reverse else range(T)
        for t in steps:
            m = (t < lens).to(x.dtype).unsqueeze(1)
            if self.conf.use_gru:
                h2 = gru_cell(xg[:, t], h, w_hh, b_hh)
                c2 = c
            else:
                h2, c2 = lstm_cell(xg[:, t], h, c, w_hh, b_hh)
            h = m * h2 + (1 - m) * h
            c = m * c2 + (1 - m) * c
            out[:, t] = m * h2
        return out, h, c

    # -- deepspeech2/encoder.py:61-104 ------------------------------------------------------------------------------
    @torch.no_grad()
    def encoder_forward(self, x, x_lens, init_state_h_box: Optional[torch.Tensor] = None,
                        init_state_c_box: Optional[torch.Tensor] = None):
        conf = self.conf
        w = self.w
        x = x.to(self.dtype)
        x = (x - w["encoder.global_cmvn.mean"]) * w["encoder.global_cmvn.istd"]
        x, x_lens = self.conv(x, x_lens)
        nd = conf.num_directions
        hs, cs = [], []
        for i in range(conf.num_rnn_layers):
            outs = []
            for d in range(nd):
                h0 = None if init_state_h_box is None else init_state_h_box[i * nd + d].to(self.dtype)
                c0 = None if (init_state_c_box is None or conf.use_gru) else init_state_c_box[i * nd + d].to(self.dtype)
                o, h, c = self.run_direction(f"encoder.rnn.{i}", "_reverse" if d == 1 else "", x, x_lens, h0, c0, d == 1)
                outs.append(o)
                hs.append(h)
                cs.append(c)
            x = torch.cat(outs, dim=-1)
            x = layer_norm(x, w[f"encoder.layernorm_list.{i}.weight"], w[f"encoder.layernorm_list.{i}.bias"])
        h_box = torch.stack(hs, dim=0)
        c_box = init_state_c_box if conf.use_gru else torch.stack(cs, dim=0)
        return x, x_lens, h_box, c_box

    # -- deepspeech2/model.py:62-72 + loss/ctc.py:27,62-70 ----------------------------------------------------------
    def ctc_logits(self, hs):
        return linear(hs, self.w["decoder.ctc_lo.weight"], self.w["decoder.ctc_lo.bias"])

    @torch.no_grad()
    def get_encoder_out(self, speech, speech_lengths, return_logits=False):
        eouts, _, _, _ = self.encoder_forward(speech, speech_lengths)
        lg = self.ctc_logits(eouts)
        return lg if return_logits else torch.softmax(lg, dim=2)

    @torch.no_grad()
    def get_encoder_out_chunk(self, speech, speech_lengths, init_state_h_box=None, init_state_c_box=None,
                              return_logits=False):
        eouts, lens, h, c = self.encoder_forward(speech, speech_lengths, init_state_h_box, init_state_c_box)
        lg = self.ctc_logits(eouts)
        return (lg if return_logits else torch.softmax(lg, dim=2)), lens, h, c
