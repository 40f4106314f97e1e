"""GPU parity tests (pytest -m gpu): every kernel and the whole hot path, called through the C-ABI,
against the CPU oracle / golden vectors. Tolerances: bit-exact for ids, texts and the greedy score;
bf16 compute => logits within 1e-2 of max|logit| (north_star: 1e-2 rel for bf16), op-level 2e-2."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold_vocab(V):
    v = ["<blank>", "<unk>"] + [chr(0x4E00 + i) for i in range(V - 4)] + ["<space>", "<eos>"]
    return v[:V]


def rel_err(got, ref):
    return (got.float() - ref.float()).abs().max().item() / max(ref.float().abs().max().item(), 1e-9)


# ------------------------------------------------------------------------------------------------
# tcgen05 GEMM + epilogues
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,epi,act,bn", [
    (128, 128, 64, 0, 0, 128), (300, 512, 128, 0, 2, 256), (1000, 768, 256, 0, 1, 128), (16, 768, 256, 0, 0, 256),
    (777, 256, 2048, 1, 0, 128), (500, 256, 4864, 1, 0, 256), (640, 512, 256, 2, 0, 256), (333, 4233, 256, 3, 0, 128),
])
def test_gemm_epilogues(lib, cuda, M, N, K, epi, act, bn):
    from ppasr_b200 import _lib as L
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device=cuda) * 0.5).to(torch.bfloat16)
    npad = (N + bn - 1) // bn * bn
    w = torch.zeros(npad, K, device=cuda, dtype=torch.bfloat16)
    w[:N] = (torch.randn(N, K, device=cuda) / K ** 0.5).to(torch.bfloat16)
    bias = torch.zeros(npad, device=cuda)
    bias[:N] = torch.randn(N, device=cuda)
    ref = a.float() @ w[:N].float().t() + bias[:N]
    alpha, residual, lens, T = 1.0, 0, None, 0
    if epi == 0:
        out = torch.zeros(M, N, device=cuda, dtype=torch.bfloat16)
        ref = torch.relu(ref) if act == 1 else (ref * torch.sigmoid(ref) if act == 2 else ref)
        ldo, tol = N, 2e-2
    elif epi == 1:
        out = torch.randn(M, N, device=cuda)
        alpha, residual = 0.5, 1
        ref = out.clone() + alpha * ref
        ldo, tol = N, 1e-4
    elif epi == 2:
        out = torch.zeros(M, N // 2, device=cuda, dtype=torch.bfloat16)
        ref = ref[:, 0::2] * torch.sigmoid(ref[:, 1::2])
        ldo, tol = N // 2, 2e-2
    else:
        ldo = (N + 3) // 4 * 4
        out = torch.zeros(M, ldo, device=cuda)
        tol = 1e-5
    L.check(lib.ppasr_b200_op_linear(L.ptr(a), K, L.ptr(w), npad, L.ptr(bias), L.ptr(out), ldo, M, N, K, epi, act,
                                     alpha, residual, None, T, bn, L.stream_ptr()))
    torch.cuda.synchronize()
    got = out[:, :N] if epi == 3 else out
    assert rel_err(got, ref) < tol


# ------------------------------------------------------------------------------------------------
# fused feed-forward block: one CTA per row tile (split 0) and 2-CTA cluster with the hidden split (split 1)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("split", [0, 1, 2])
@pytest.mark.parametrize("M,FF,dbl", [(7936, 2048, 0), (7936, 2048, 1), (1000, 2048, 0), (77, 512, 1), (128, 1024, 0)])
def test_fused_ffn_op_both_variants(lib, cuda, split, M, FF, dbl):
    """positionwise.py:30-39 + residual + LayerNorm(s) (encoder.py:380-386,419-429) on raw pointers; tolerance 2e-2 of
    max|.| on the bf16 y output (bf16 hidden activation), 5e-3 on the fp32 residual stream."""
    from ppasr_b200 import _lib as L
    torch.manual_seed(M + FF + dbl)
    y = torch.randn(M, 256, device=cuda).to(torch.bfloat16)
    w1 = (torch.randn(FF, 256, device=cuda) / 16).to(torch.bfloat16)
    w2 = (torch.randn(256, FF, device=cuda) / FF ** 0.5 * 0.5).to(torch.bfloat16)
    b1 = torch.randn(FF, device=cuda) * 0.1
    b2 = torch.randn(256, device=cuda) * 0.1
    g1 = torch.rand(256, device=cuda) + 0.5
    bn1 = torch.randn(256, device=cuda) * 0.1
    g2 = torch.rand(256, device=cuda) + 0.5
    bn2 = torch.randn(256, device=cuda) * 0.1
    x0 = torch.randn(M, 256, device=cuda)
    x = x0.clone()
    yo = torch.zeros(M, 256, device=cuda, dtype=torch.bfloat16)
    L.check(lib.ppasr_b200_set_ffn_split(split))
    try:
        L.check(lib.ppasr_b200_op_fused_ffn(L.ptr(y), L.ptr(w1), L.ptr(w2), L.ptr(x), L.ptr(yo), L.ptr(b1), L.ptr(b2),
                                            L.ptr(g1), L.ptr(bn1), L.ptr(g2) if dbl else None, L.ptr(bn2) if dbl else None,
                                            M, FF, 1e-5, L.stream_ptr()))
        torch.cuda.synchronize()
    finally:
        lib.ppasr_b200_set_ffn_split(1)
    h = y.float() @ w1.float().t() + b1
    h = (h * torch.sigmoid(h)).to(torch.bfloat16).float()
    xr = x0 + h @ w2.float().t() + b2
    if dbl:
        xr = F.layer_norm(xr, (256,), g1, bn1, 1e-5)
        yr = F.layer_norm(xr, (256,), g2, bn2, 1e-5)
    else:
        yr = F.layer_norm(xr, (256,), g1, bn1, 1e-5)
    assert rel_err(x, xr) < 5e-3
    assert rel_err(yo, yr) < 2e-2


def test_fused_attn_out_variants_agree(lib, cuda):
    """fused_attn_out with and without the serial residual preload differ only in the fp32 summation order of
    x + bo + Wo.att (a flipped bf16 rounding of the LayerNorm output is 0.4 %): whole-model logits within 5e-3 of max|logit|,
    ragged batch (pad rows zeroed in the conv-module input)."""
    from ppasr_b200 import engine as E, weights as W
    cfg = W.ConformerConfig(num_blocks=3, vocab_size=301)
    w = W.init_conformer_weights(cfg)
    feats = torch.from_numpy(W.synthetic_fbank(3, 523)).cuda()
    lens = [523, 3, 260]
    outs = []
    eng = E.ConformerEngine(cfg, w, device=0)
    for v in (0, 1):
        eng.set_option("attn_out_v2", v)
        try:
            eng.encode(feats, lens)
            outs.append(eng.ctc_logits().float().cpu())
        finally:
            eng.set_option("attn_out_v2", 1)
    eng.close()
    assert rel_err(outs[1], outs[0]) < 5e-3


@pytest.mark.parametrize("family", ["conformer", "squeezeformer"])
def test_fused_ffn_split_matches_single_cta_model_level(lib, cuda, family):
    """The kernel variants differ only in fp32 summation order (which can flip a bf16 rounding of the LayerNorm output,
    1 ulp = 0.4 %): whole-model logits (plain + chained + post-norm chained modes are all exercised by these two
    families) agree to 5e-3 of max|logit|, i.e. inside the bf16 noise the oracle comparison allows (1e-2)."""
    from ppasr_b200 import engine as E, weights as W
    if family == "conformer":
        cfg = W.ConformerConfig(num_blocks=3, vocab_size=301)
        w = W.init_conformer_weights(cfg)
        mk = lambda: E.ConformerEngine(cfg, w, device=0)
    else:
        cfg = W.SqueezeformerConfig(num_blocks=4, vocab_size=301, reduce_idx=1, recover_idx=3)
        w = W.init_squeezeformer_weights(cfg)
        mk = lambda: E.ConformerEngine(cfg, w, device=0)
    feats = torch.from_numpy(W.synthetic_fbank(3, 523)).cuda()
    lens = [523, 3, 260]
    outs = []
    for split in (0, 1, 2):
        lib.ppasr_b200_set_ffn_split(split)
        try:
            eng = mk()
            eng.encode(feats, lens)
            outs.append(eng.ctc_logits().float().cpu())
        finally:
            lib.ppasr_b200_set_ffn_split(1)
    assert rel_err(outs[1], outs[0]) < 5e-3
    assert rel_err(outs[2], outs[0]) < 5e-3


def test_gemm_residual_row_mask(lib, cuda):
    from ppasr_b200 import _lib as L
    B, T, N, K = 4, 248, 256, 256
    M = B * T
    a = torch.randn(M, K, device=cuda).to(torch.bfloat16)
    w = (torch.randn(N, K, device=cuda) / 16).to(torch.bfloat16)
    bias = torch.randn(N, device=cuda)
    lens = torch.tensor([100, 248, 7, 0], device=cuda, dtype=torch.int32)
    x0 = torch.randn(M, N, device=cuda)
    out = x0.clone()
    L.check(lib.ppasr_b200_op_linear(L.ptr(a), K, L.ptr(w), N, L.ptr(bias), L.ptr(out), N, M, N, K, 1, 0, 1.0, 1,
                                     L.ptr(lens), T, 128, L.stream_ptr()))
    ref = x0 + a.float() @ w.float().t() + bias
    rows = torch.arange(M, device=cuda)
    masked = (rows % T) >= lens[rows // T]
    ref = torch.where(masked[:, None], x0, ref)
    assert rel_err(out, ref) < 1e-4
    assert torch.equal(out[masked], x0[masked])  # padded rows untouched, bit-exact


# ------------------------------------------------------------------------------------------------
# CUDA-core kernels
# ------------------------------------------------------------------------------------------------
def test_layernorm_variants(lib, cuda):
    from ppasr_b200 import _lib as L
    M, D = 1003, 256
    x = torch.randn(M, D, device=cuda) * 3 + 0.5
    g1, b1, g2, b2 = [torch.randn(D, device=cuda) for _ in range(4)]
    y = torch.zeros(M, D, device=cuda, dtype=torch.bfloat16)
    xs = x.clone()
    L.check(lib.ppasr_b200_op_layernorm(L.ptr(xs), L.ptr(y), L.ptr(g1), L.ptr(b1), None, None, None, 0, M, D, 1e-5,
                                        L.stream_ptr()))
    assert rel_err(y, F.layer_norm(x, (D,), g1, b1, 1e-5)) < 1e-2
    assert torch.equal(xs, x)
    L.check(lib.ppasr_b200_op_layernorm(L.ptr(xs), L.ptr(y), L.ptr(g1), L.ptr(b1), L.ptr(g2), L.ptr(b2), None, 0, M,
                                        D, 1e-5, L.stream_ptr()))
    r1 = F.layer_norm(x, (D,), g1, b1, 1e-5)
    assert rel_err(xs, r1) < 1e-5
    assert rel_err(y, F.layer_norm(r1, (D,), g2, b2, 1e-5)) < 1e-2


@pytest.mark.parametrize("K,causal", [(15, True), (15, False), (31, False), (7, True)])
def test_dwconv_norm_swish(lib, cuda, K, causal):
    from ppasr_b200 import _lib as L
    B, T, C = 3, 77, 256
    torch.manual_seed(K)
    g = torch.randn(B, T, C, device=cuda).to(torch.bfloat16)
    w = torch.randn(C, K, device=cuda) / K ** 0.5
    bias = torch.randn(C, device=cuda) * 0.1
    gam, bet = torch.rand(C, device=cuda) + 0.5, torch.randn(C, device=cuda) * 0.1
    pad = torch.randn(C, device=cuda).to(torch.bfloat16).float()
    out = torch.zeros(B, T, C, device=cuda, dtype=torch.bfloat16)
    lpad = K - 1 if causal else (K - 1) // 2
    L.check(lib.ppasr_b200_op_dwconv(L.ptr(g), L.ptr(w), L.ptr(bias), L.ptr(pad) if causal else None, L.ptr(gam),
                                     L.ptr(bet), 1, L.ptr(out), B, T, T, C, K, lpad, 1e-5, L.stream_ptr()))
    gi = g.float().transpose(1, 2)
    if causal:
        gi = torch.cat([pad.view(1, C, 1).expand(B, C, K - 1), gi], 2)
        cv = F.conv1d(gi, w.view(C, 1, K), bias, groups=C)
    else:
        cv = F.conv1d(gi, w.view(C, 1, K), bias, groups=C, padding=(K - 1) // 2)
    r = F.layer_norm(cv.transpose(1, 2), (C,), gam, bet, 1e-5)
    assert rel_err(out, r * torch.sigmoid(r)) < 1e-2


def test_softmax_rows(lib, cuda):
    from ppasr_b200 import _lib as L
    for V in (97, 4233, 6000):
        M = 301
        ld = (V + 3) // 4 * 4
        lg = torch.randn(M, ld, device=cuda) * 4
        pr = torch.zeros(M, V, device=cuda)
        L.check(lib.ppasr_b200_op_softmax(L.ptr(lg), ld, L.ptr(pr), M, V, L.stream_ptr()))
        assert rel_err(pr, torch.softmax(lg[:, :V], -1)) < 1e-5


@pytest.mark.parametrize("B,H,T1,T2", [(2, 4, 248, 248), (3, 4, 100, 100), (1, 4, 16, 80), (2, 4, 300, 300),
                                       (1, 4, 1, 1), (1, 4, 129, 129)])
def test_rel_attention(lib, cuda, B, H, T1, T2):
    from ppasr_b200 import _lib as L
    torch.manual_seed(T1 * 7 + T2)
    q = torch.randn(B, H, T1, 64, device=cuda)
    k = torch.randn(B, H, T2, 64, device=cuda)
    v = torch.randn(B, H, T2, 64, device=cuda)
    pos_rows, Lc, l, row0 = 400, 2, 1, 7
    pos = torch.randn(pos_rows, Lc * H * 64, device=cuda).to(torch.bfloat16)
    u = torch.randn(H, 64, device=cuda) * 0.3
    vb = torch.randn(H, 64, device=cuda) * 0.3
    klens = torch.randint(max(1, T2 // 2), T2 + 1, (B,), device=cuda, dtype=torch.int32)
    klens[0] = T2
    q2 = torch.cat([(q + u[None, :, None, :]), (q + vb[None, :, None, :])], -1).to(torch.bfloat16).contiguous()
    kb = k.to(torch.bfloat16).contiguous()
    T2p = (T2 + 63) // 64 * 64
    vt = torch.zeros(B, H, 64, T2p, device=cuda, dtype=torch.bfloat16)
    vt[..., :T2] = v.to(torch.bfloat16).transpose(-1, -2)
    out = torch.zeros(B * T1, H * 64, device=cuda, dtype=torch.bfloat16)
    L.check(lib.ppasr_b200_op_attention(L.ptr(q2), L.ptr(kb), L.ptr(vt), T2p, L.ptr(pos), pos_rows, Lc * H * 64, row0,
                                        l * H * 64, L.ptr(out), B, H, T1, T2, L.ptr(klens), L.stream_ptr()))
    p = pos[row0:row0 + T2, l * H * 64:(l + 1) * H * 64].float().view(T2, H, 64).transpose(0, 1)
    q2f = q2.float()
    s = (q2f[..., :64] @ kb.float().transpose(-1, -2) + q2f[..., 64:] @ p[None].transpose(-1, -2)) / 8.0
    mask = torch.arange(T2, device=cuda)[None, :] >= klens[:, None]
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    a = torch.softmax(s, -1).masked_fill(mask[:, None, None, :], 0.0)
    r = (a @ v.to(torch.bfloat16).float()).transpose(1, 2).reshape(B * T1, H * 64)
    assert rel_err(out, r) < 2e-2


# ------------------------------------------------------------------------------------------------
# greedy decoder: bit-exact against the reference's own outputs (golden) through the drop-in API
# ------------------------------------------------------------------------------------------------
def test_greedy_decoder_golden_bit_exact(lib, cuda):
    from ppasr_b200.decoders.ctc_greedy_decoder import greedy_decoder, greedy_decoder_batch, greedy_decoder_chunk
    z = np.load(os.path.join(GOLD, "greedy_golden.npz"))
    meta = json.load(open(os.path.join(GOLD, "greedy_golden.json"), encoding="utf-8"))
    for m in meta:
        if m["name"] == "__batch__":
            p = z["long_probs"]
            assert greedy_decoder_batch([p, p[:100]], gold_vocab(97)) == m["texts"]
            continue
        probs = z[m["name"] + "_probs"]
        score, text = greedy_decoder(probs, gold_vocab(m["V"]))
        assert text == m["text"], m["name"]
        assert repr(float(score)) == m["score"], (m["name"], score, m["score"])
        lp, li = None, None
        for c, s in enumerate(range(0, m["T"], 16)):
            sc, tx, lp, li = greedy_decoder_chunk(probs[s:s + 16], gold_vocab(m["V"]), lp, li)
            assert tx == m["chunks"][c]["text"]
            assert repr(float(sc)) == m["chunks"][c]["score"]


def test_greedy_full_size_matches_oracle(lib, cuda):
    """BASELINE configs[1] posterior size [32,248,4233]: ids / scores bit-exact against the NumPy restatement."""
    from oracle import decoders_oracle as DO
    from ppasr_b200.decoders.ctc_greedy_decoder import greedy_decode_ids
    torch.manual_seed(1)
    probs = torch.softmax(torch.randn(32, 248, 4233, device=cuda) * 3, -1)
    probs[:, :, 0] += 0.02
    ids, ol, sc, fi, fp = greedy_decode_ids(probs)
    pn = probs.cpu().numpy()
    vocab = [str(i) + "," for i in range(4233)]
    for b in range(0, 32, 5):
        _, coll, _ = DO.greedy_ids(pn[b])
        assert ids[b, :ol[b]].tolist() == coll
        score, _ = DO.greedy_decoder(pn[b], vocab)
        assert float(sc[b]) * 100.0 == score


# ------------------------------------------------------------------------------------------------
# whole hot path vs the oracle
# ------------------------------------------------------------------------------------------------
def _run_model(cuda, num_blocks, B, T, lens, vocab=4233, streaming=True, norm="layer_norm", model="conformer", **mkw):
    from oracle import decoders_oracle as DO
    from oracle.conformer_oracle import ConformerConf, ConformerOracle
    from oracle.squeezeformer_oracle import SqueezeformerConf, SqueezeformerOracle
    from ppasr_b200.engine import ConformerEngine, out_frames
    from ppasr_b200.weights import (ConformerConfig, SqueezeformerConfig, init_conformer_weights,
                                    init_squeezeformer_weights, synthetic_fbank)
    if model == "efficient_conformer":
        from oracle.efficient_conformer_oracle import EfficientConformerConf, EfficientConformerOracle
        from ppasr_b200.weights import EfficientConformerConfig, init_efficient_conformer_weights
        cfg = EfficientConformerConfig(num_blocks=num_blocks, vocab_size=vocab, streaming=streaming, cnn_module_norm=norm, **mkw)
        w = init_efficient_conformer_weights(cfg)
        oracle = EfficientConformerOracle(EfficientConformerConf(**cfg.to_dict()), w)
    elif model == "squeezeformer":
        cfg = SqueezeformerConfig(num_blocks=num_blocks, vocab_size=vocab, streaming=streaming, cnn_norm_type=norm, **mkw)
        w = init_squeezeformer_weights(cfg)
        oracle = SqueezeformerOracle(SqueezeformerConf(**cfg.to_dict()), w)
    else:
        cfg = ConformerConfig(num_blocks=num_blocks, vocab_size=vocab, streaming=streaming, cnn_module_norm=norm)
        w = init_conformer_weights(cfg)
        oracle = ConformerOracle(ConformerConf(**cfg.to_dict()), w)
    feats = synthetic_fbank(B, T)
    for b in range(B):
        feats[b, lens[b]:] = 0
    eng = ConformerEngine(cfg, w)
    eng.encode(torch.from_numpy(feats).to(cuda), lens)
    logits = eng.ctc_logits().cpu()
    probs = eng.ctc_probs().cpu()
    ids, ol, sc, fi, fp = eng.ctc_greedy(to_host=True, with_frames=True)
    ref_logits = oracle.get_encoder_out(torch.from_numpy(feats), torch.tensor(lens), return_logits=True)
    Tp = out_frames(T)
    vl = [min(Tp, (l + 3) // 4) for l in lens]
    if model == "efficient_conformer" and cfg.stride_layer_idx is not None:
        Tp = (Tp + 1) // 2                      # efficient_conformer/encoder.py:255-260
        vl = [(v + 1) // 2 for v in vl]
    assert tuple(logits.shape) == tuple(ref_logits.shape) == (B, Tp, vocab)
    scale = ref_logits.abs().max().item()
    worst = max((logits[b, :vl[b]] - ref_logits[b, :vl[b]]).abs().max().item() for b in range(B) if vl[b] > 0) / scale
    assert worst < 1e-2, f"logits rel err {worst}"
    # greedy ids vs the fp32 oracle. Bit-exact is what the decode kernels deliver ON A GIVEN posterior (checked below and
    # against the reference's own outputs in test_greedy_decoder_golden_bit_exact); end to end the bf16 encoder may flip an
    # arg-max only where the oracle's top-2 margin is below twice the logit error. So: every frame whose margin exceeds
    # 3 x the MEASURED max |logit error| must agree, and the unfiltered agreement rate is reported and bounded.
    top2 = ref_logits.topk(2, -1).values
    margin = top2[..., 0] - top2[..., 1]
    big = margin > 3.0 * worst * scale
    ref_ids = ref_logits.argmax(-1)
    n_frames = n_agree = n_small = 0
    for b in range(B):
        same = torch.from_numpy(fi)[b, :vl[b]] == ref_ids[b, :vl[b]]
        assert bool((same | ~big[b, :vl[b]]).all())
        n_frames += int(vl[b])
        n_agree += int(same.sum())
        n_small += int((~big[b, :vl[b]]).sum())
    rate = n_agree / max(1, n_frames)
    print(f"[greedy ids] model={model} L={num_blocks} B={B} T={T}: logits rel err {worst:.3e}; unfiltered frame agreement "
          f"{n_agree}/{n_frames} = {rate:.5f}; frames inside the 3x-error margin: {n_small}")
    assert rate >= 0.95, f"unfiltered greedy agreement {rate}"
    # fused head == reference greedy on the materialised posterior of the same engine (bit-exact ids)
    pn = probs.numpy()
    for b in range(B):
        _, coll, _ = DO.greedy_ids(pn[b])
        assert ids[b, :ol[b]].tolist() == coll
    assert torch.allclose(probs.sum(-1), torch.ones(B, Tp), atol=1e-4)
    eng.close()
    return worst


@pytest.mark.parametrize("kw", [
    dict(num_blocks=1, B=2, T=131, lens=[131, 90], vocab=97),
    dict(num_blocks=2, B=3, T=400, lens=[400, 333, 250]),
    dict(num_blocks=2, B=2, T=300, lens=[300, 200], streaming=False),
    dict(num_blocks=2, B=2, T=300, lens=[300, 200], streaming=False, norm="batch_norm"),
    dict(num_blocks=1, B=1, T=7, lens=[7], vocab=50),           # minimum length: one output frame
    dict(num_blocks=1, B=3, T=523, lens=[523, 3, 260], vocab=50),  # T' = 130 crosses the 128-row tile, tiny len
])
def test_model_matches_oracle_small(lib, cuda, kw):
    _run_model(cuda, **kw)


def test_model_matches_oracle_conformer_12_layers(lib, cuda):
    """conformer.yml sizes (12 blocks, d256, ff2048, V=4233), 4 x 10 s, ragged."""
    _run_model(cuda, 12, 4, 998, [998, 998, 900, 500])


@pytest.mark.parametrize("kw", [
    dict(num_blocks=1, B=2, T=131, lens=[131, 90], vocab=97, reduce_idx=None, recover_idx=None),
    dict(num_blocks=3, B=3, T=203, lens=[203, 150, 99], vocab=300, reduce_idx=1, recover_idx=2),          # even T' = 50
    dict(num_blocks=3, B=3, T=207, lens=[207, 150, 5], vocab=300, reduce_idx=1, recover_idx=2, streaming=False),  # odd T'
    dict(num_blocks=4, B=2, T=1051, lens=[1051, 700], vocab=300, reduce_idx=1, recover_idx=3),           # T' = 262 > 2 tiles
    dict(num_blocks=2, B=2, T=300, lens=[300, 200], vocab=300, reduce_idx=None, recover_idx=None, norm="batch_norm",
         streaming=False),
])
def test_squeezeformer_matches_oracle_small(lib, cuda, kw):
    """squeezeformer/encoder.py:172-236 incl. time reduction ('stream' k1 for streaming, conv1d k5 otherwise) + recover."""
    _run_model(cuda, model="squeezeformer", **kw)


def test_squeezeformer_matches_oracle_12_layers(lib, cuda):
    """configs/squeezeformer.yml sizes (12 blocks, reduce 5 / recover 11, k31 causal, V=4233), 4 x 10 s, ragged."""
    _run_model(cuda, 12, 4, 998, [998, 998, 900, 500], model="squeezeformer")


def test_squeezeformer_inference_predictor(lib, cuda):
    """InferencePredictor(use_model='squeezeformer').predict / predict_decode (the chunk API has its own test)."""
    from ppasr_b200.infer_utils.inference_predictor import InferencePredictor
    from ppasr_b200.weights import SqueezeformerConfig, init_squeezeformer_weights, make_vocab, synthetic_fbank
    from oracle import decoders_oracle as DO
    cfg = SqueezeformerConfig(num_blocks=3, vocab_size=200, reduce_idx=1, recover_idx=2)
    w = init_squeezeformer_weights(cfg)
    pred = InferencePredictor({"encoder_conf": cfg.to_dict(), "preprocess_conf": {"n_mels": 80}}, "squeezeformer",
                              streaming=True, weights=w)
    x = synthetic_fbank(2, 260)
    probs = pred.predict(x, np.array([260, 260]))
    assert probs.shape == (2, 64, 200) and np.allclose(probs.sum(-1), 1.0, atol=1e-4)
    vocab = make_vocab(200)
    got = pred.predict_decode(x, None, vocabulary=vocab)
    for b in range(2):
        score, text = DO.greedy_decoder(probs[b], vocab)
        assert got[b][1] == text and abs(got[b][0] - score) < 1e-3


@pytest.mark.parametrize("kw", [
    dict(num_blocks=1, B=2, T=131, lens=[131, 90], vocab=97, group_layer_idx=(0,), stride_layer_idx=None),  # T'=32 -> pad 1
    dict(num_blocks=1, B=2, T=135, lens=[135, 90], vocab=97, group_layer_idx=(), stride_layer_idx=0),       # odd T'=33
    dict(num_blocks=2, B=3, T=203, lens=[203, 150, 99], vocab=300, group_layer_idx=(0, 1), stride_layer_idx=1),
    dict(num_blocks=2, B=3, T=207, lens=[207, 150, 5], vocab=300, group_layer_idx=(0, 1), stride_layer_idx=1,
         streaming=False),
    dict(num_blocks=3, B=2, T=1051, lens=[1051, 700], vocab=300, group_layer_idx=(0, 1), stride_layer_idx=1),  # 88 groups
    dict(num_blocks=2, B=2, T=300, lens=[300, 200], vocab=300, group_layer_idx=(0,), stride_layer_idx=0,
         streaming=False, norm="batch_norm"),
])
def test_efficient_conformer_matches_oracle_small(lib, cuda, kw):
    """efficient_conformer/encoder.py:212-264: grouped attention (attention.py:128-193), stride block with AvgPool residual
    (encoder.py:455-548), kernel 15 -> 7, output at ceil(T'/2)."""
    _run_model(cuda, model="efficient_conformer", **kw)


def test_efficient_conformer_matches_oracle_12_layers(lib, cuda):
    """configs/efficient_conformer.yml sizes (12 blocks, grouped 0-3, stride block 3, V=4233), 4 x 5 s, ragged."""
    _run_model(cuda, 12, 4, 498, [498, 498, 400, 250], model="efficient_conformer")


def test_full_size_properties(lib, cuda):
    """BASELINE configs[1] shape [32,998,80]: run-to-run determinism and batch independence, bit-exact."""
    from ppasr_b200.engine import ConformerEngine
    from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank
    cfg = ConformerConfig()
    eng = ConformerEngine(cfg, init_conformer_weights(cfg))
    feats = torch.from_numpy(synthetic_fbank(32, 998)).to(cuda)
    eng.encode(feats)
    ids1, ol1, sc1, fi1, fp1 = eng.ctc_greedy(to_host=True, with_frames=True)
    eng.encode(feats)
    ids2, ol2, sc2, fi2, fp2 = eng.ctc_greedy(to_host=True, with_frames=True)
    assert np.array_equal(fi1, fi2) and np.array_equal(fp1, fp2) and np.array_equal(sc1, sc2)
    eng.encode(feats[5:6].contiguous())
    ids3, ol3, sc3, fi3, fp3 = eng.ctc_greedy(to_host=True, with_frames=True)
    assert np.array_equal(fi3[0], fi1[5]) and np.array_equal(fp3[0], fp1[5])
    assert ids3[0, :ol3[0]].tolist() == ids1[5, :ol1[5]].tolist() and sc3[0] == sc1[5]
    assert (ol1 > 0).all()
    eng.close()


def test_conformer_non_streaming_12_layers_30s(lib, cuda):
    """BASELINE configs[2] geometry: conformer.yml non-streaming (symmetric conv, full attention), 12 blocks, 30 s utterances
    (T = 2998 -> T' = 748 keys per query: the attention kernel's multi-block key loop at full depth), ragged, vs the oracle."""
    _run_model(cuda, 12, 2, 2998, [2998, 2501], streaming=False)


def test_full_size_batch_matches_oracle_on_a_subset(lib, cuda):
    """BASELINE configs[1] at its real size, B = 32 x 998 frames, 12 blocks: the logits of utterances 0, 13 and 31 of the
    batched run against the oracle run on each of them alone (utterances are independent; all 32 are full length)."""
    from oracle.conformer_oracle import ConformerConf, ConformerOracle
    from ppasr_b200.engine import ConformerEngine
    from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank
    cfg = ConformerConfig()
    w = init_conformer_weights(cfg)
    feats = synthetic_fbank(32, 998)
    eng = ConformerEngine(cfg, w)
    eng.encode(torch.from_numpy(feats).to(cuda))
    logits = eng.ctc_logits().float().cpu()
    eng.close()
    orc = ConformerOracle(ConformerConf(**cfg.to_dict()), w)
    n_frames = n_agree = 0
    for b in (0, 13, 31):
        ref = orc.get_encoder_out(torch.from_numpy(feats[b:b + 1]), torch.tensor([998]), return_logits=True)[0]
        scale = ref.abs().max().item()
        err = (logits[b] - ref).abs().max().item() / scale
        assert err < 1e-2, f"utterance {b}: logits rel err {err}"
        top2 = ref.topk(2, -1).values
        big = (top2[:, 0] - top2[:, 1]) > 3.0 * err * scale
        same = logits[b].argmax(-1) == ref.argmax(-1)
        assert bool((same | ~big).all())
        n_frames += ref.shape[0]
        n_agree += int(same.sum())
    print(f"[C2 full size] unfiltered arg-max agreement {n_agree}/{n_frames}")
    assert n_agree >= 0.95 * n_frames


def test_cuda_graph_replay_matches_direct_run(lib, cuda):
    """ppasr_b200_graph_begin / _end / _launch: the captured step (encode + fused CTC head + greedy, with the programmatic
    dependent launch edges and the 2-CTA cluster launches) replayed on NEW feature values gives bit-identical ids / lens / scores
    to running the same calls directly."""
    from ppasr_b200.engine import ConformerEngine
    from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank
    cfg = ConformerConfig(num_blocks=3, vocab_size=300)
    eng = ConformerEngine(cfg, init_conformer_weights(cfg))
    B, T = 4, 363
    lens = [363, 200, 363, 90]
    feats = torch.from_numpy(synthetic_fbank(B, T, seed=1)).to(cuda)
    other = torch.from_numpy(synthetic_fbank(B, T, seed=2)).to(cuda)
    gs = torch.cuda.Stream()
    Tp = eng.encode(feats, lens, stream=gs).Tp
    ids = torch.zeros((B, Tp), dtype=torch.int32, device=cuda)
    ol = torch.zeros((B,), dtype=torch.int32, device=cuda)
    sc = torch.zeros((B,), dtype=torch.float32, device=cuda)
    eng.ctc_greedy_into(ids, ol, sc, trim_to_lens=True, stream=gs)
    gs.synchronize()
    eng.graph_begin(gs)
    eng.encode(feats, lens, stream=gs)
    eng.ctc_greedy_into(ids, ol, sc, trim_to_lens=True, stream=gs)
    nk = eng.graph_end(gs)
    assert nk > 10
    feats.copy_(other)            # the graph re-reads the captured pointer: new values, same buffer
    torch.cuda.synchronize()
    n0 = lib.ppasr_b200_launch_count()
    eng.graph_launch(gs)
    gs.synchronize()
    assert lib.ppasr_b200_launch_count() - n0 == nk
    got = (ids.clone(), ol.clone(), sc.clone())
    eng.encode(other, lens)
    rids, rol, rsc = eng.ctc_greedy(to_host=False, trim_to_lens=True)
    torch.cuda.synchronize()
    assert torch.equal(got[1], rol) and torch.equal(got[2], rsc)
    for b in range(B):
        assert torch.equal(got[0][b, :int(rol[b])], rids[b, :int(rol[b])])
    eng.close()


def test_inference_predictor_api(lib, cuda):
    """Drop-in surface: predict -> probs [B,T',V] host; predict_decode == reference greedy on those probs."""
    from oracle import decoders_oracle as DO
    from ppasr_b200.infer_utils.inference_predictor import InferencePredictor
    from ppasr_b200.weights import ConformerConfig, init_conformer_weights, make_vocab, synthetic_fbank
    cfg = ConformerConfig(num_blocks=2, vocab_size=300)
    w = init_conformer_weights(cfg)
    pred = InferencePredictor({"encoder_conf": cfg.to_dict(), "preprocess_conf": {"n_mels": 80}}, "conformer",
                              streaming=True, weights=w)
    feats = synthetic_fbank(3, 211)
    lens = np.array([211, 211, 211], dtype=np.int64)
    probs = pred.predict(feats, lens)
    assert probs.shape == (3, 52, 300) and probs.dtype == np.float32
    vocab = make_vocab(300)
    res = pred.predict_decode(feats, lens, vocabulary=vocab)
    ref = [DO.greedy_decoder(probs[b], vocab) for b in range(3)]
    assert [r[1] for r in res] == [r[1] for r in ref]
    with pytest.raises(Exception):
        InferencePredictor({}, "deepspeech2", weights=w)
    with pytest.raises(Exception):
        InferencePredictor({}, "conformer", model_dir="/nonexistent/dir")


# ------------------------------------------------------------------------------------------------
# chunk streaming (forward_chunk / predict_chunk_conformer / predict_stream)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("required", [-16, 32, 0])
def test_chunk_streaming_matches_oracle(lib, cuda, required):
    from oracle.conformer_oracle import ConformerConf, ConformerOracle, stream_windows
    from ppasr_b200.infer_utils.inference_predictor import InferencePredictor
    from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank
    cfg = ConformerConfig(num_blocks=2, vocab_size=211)
    w = init_conformer_weights(cfg)
    T = 215  # 3 full 67-frame windows + a short tail consumed at is_end
    feats = synthetic_fbank(1, T)
    orc = ConformerOracle(ConformerConf(**cfg.to_dict()), w)
    pred = InferencePredictor({"encoder_conf": cfg.to_dict(), "preprocess_conf": {"n_mels": 80}}, "conformer",
                              streaming=True, weights=w)
    att = torch.zeros(0, 0, 0, 0)
    cnn = torch.zeros(0, 0, 0, 0)
    off = 0
    for (s, e) in stream_windows(T, is_end=True):
        ref, att, cnn = orc.get_encoder_out_chunk(torch.from_numpy(feats[:, s:e]), off, required, att, cnn)
        off += ref.shape[1]
        got = pred.predict_chunk_conformer(feats[:, s:e], required)
        assert got.shape == tuple(ref.shape)
        assert np.abs(got - ref.numpy()).max() < 3e-2, (s, e)
        assert int(pred.offset[0]) == off
    a, c = pred.att_cache, pred.cnn_cache
    assert a.shape == tuple(att.shape) and c.shape == tuple(cnn.shape)
    if att.numel():
        assert rel_err(torch.from_numpy(a), att) < 2e-2
    assert rel_err(torch.from_numpy(c), cnn) < 2e-2
    pred.reset_stream()
    assert pred.att_cache.shape == (0, 0, 0, 0) and int(pred.offset[0]) == 0
    # a second stream after reset reproduces the first chunk bit-for-bit
    s, e = stream_windows(T, is_end=True)[0]
    g1 = pred.predict_chunk_conformer(feats[:, s:e], required)
    pred.reset_stream()
    g2 = pred.predict_chunk_conformer(feats[:, s:e], required)
    assert np.array_equal(g1, g2)


def test_predict_stream_window_logic(lib, cuda):
    """PPASRPredictor.predict_stream_features: arbitrary feed sizes give the same transcript as feeding the
    reference's 67/64 windows by hand; None while fewer than a window is buffered (predict.py:287-288)."""
    from oracle import decoders_oracle as DO
    from oracle.conformer_oracle import stream_windows
    from ppasr_b200.predict import PPASRPredictor
    from ppasr_b200.weights import ConformerConfig, init_conformer_weights, make_vocab, synthetic_fbank
    cfg = ConformerConfig(num_blocks=2, vocab_size=150)
    w = init_conformer_weights(cfg)
    vocab = make_vocab(150)
    configs = {"use_model": "conformer", "streaming": True, "decoder": "ctc_greedy", "encoder_conf": cfg.to_dict(),
               "preprocess_conf": {"feature_method": "fbank", "n_mels": 80}}
    p = PPASRPredictor(configs, vocab_list=vocab, weights=w)
    T = 300
    feats = synthetic_fbank(1, T)
    assert p.predict_stream_features(feats[:, :30]) is None
    res = None
    pos = 30
    for n in (50, 10, 130, 80):
        r = p.predict_stream_features(feats[:, pos:pos + n], is_end=(pos + n >= T))
        pos += n
        res = r if r is not None else res
    assert pos == T and res is not None
    # by hand through the InferencePredictor with the reference windows
    p.reset_stream()
    chunks = [p.predictor.predict_chunk_conformer(feats[:, s:e], -16)[0] for (s, e) in stream_windows(T, is_end=True)]
    score, text = DO.greedy_decoder(np.concatenate(chunks, 0), vocab)
    assert res["text"] == text
    assert abs(res["score"] - score) < 1e-3
    # offline predict on the same features returns a dict with the same keys
    out = p.predict_features(feats[0])
    assert set(out) == {"text", "score"}


# ------------------------------------------------------------------------------------------------
# CTC prefix beam search (no scorer) vs the oracle restatement (parity unpinned: see oracle header)
# ------------------------------------------------------------------------------------------------
def _peaky_probs(rng, T, V, temp, blank_boost=0.0):
    logits = rng.randn(T, V).astype(np.float32) * temp
    logits[:, 0] += blank_boost
    e = np.exp(logits - logits.max(-1, keepdims=True))
    return (e / e.sum(-1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("T,V,beam,cp,topn,temp", [
    (12, 5, 100, 1.0, 5, 1.0), (40, 30, 10, 0.99, 40, 3.0), (60, 97, 20, 0.99, 40, 4.0), (248, 300, 10, 1.0, 40, 5.0),
    (30, 50, 1, 1.0, 50, 6.0), (25, 10, 16, 0.9, 3, 1.5),
])
def test_beam_search_matches_oracle(lib, cuda, T, V, beam, cp, topn, temp):
    from oracle import decoders_oracle as DO
    from ppasr_b200.decoders.beam_search_decoder import BeamSearchDecoder
    rng = np.random.RandomState(T * 31 + V)
    vocab = [f"<{i}>" for i in range(V)]
    dec = BeamSearchDecoder(beam_size=beam, cutoff_prob=cp, cutoff_top_n=topn, vocab_list=vocab)
    probs = np.stack([_peaky_probs(rng, T, V, temp, bb) for bb in (0.0, 2.0, 0.5)])
    got = dec.decode_ids_batch(probs)
    for b in range(probs.shape[0]):
        ref = DO.ctc_beam_search_decoding(probs[b], vocab, beam, cp, topn)
        assert got[b][0][1] == ref[0][1], f"best hypothesis differs (utt {b})"
        assert abs(got[b][0][0] - ref[0][0]) < 1e-3 * max(1.0, abs(ref[0][0]))
        # the whole beam agrees as a set of (text -> score), up to fp32 near-ties at the beam edge
        rd = dict((t, s) for s, t in ref)
        hits = sum(1 for s, t in got[b] if t in rd and abs(rd[t] - s) < 1e-3 * max(1.0, abs(s)))
        assert hits >= max(1, int(0.8 * min(len(ref), len(got[b]))))
    # single-utterance API and batch API
    s1, t1 = dec.decode_beam_search_offline(probs[0])
    assert t1 == got[0][0][1]
    assert dec.decode_batch_beam_search_offline([probs[0], probs[1][: T // 2]])[0] == t1


@pytest.mark.parametrize("T,V,beam,cp,topn,temp", [(60, 120, 300, 0.99, 40, 1.5), (40, 60, 200, 1.0, 40, 1.0)])
def test_beam_search_shipped_beam_size_matches_oracle(lib, cuda, T, V, beam, cp, topn, temp):
    """configs/conformer.yml:80-90 ships beam_size 300 / cutoff_top_n 40: the wide-beam path (512 threads per utterance,
    shared-memory bitonic selection over up to beam * top_n candidates) against the oracle restatement."""
    from oracle import decoders_oracle as DO
    from ppasr_b200.decoders.beam_search_decoder import BeamSearchDecoder
    rng = np.random.RandomState(T + V + beam)
    vocab = [f"<{i}>" for i in range(V)]
    dec = BeamSearchDecoder(beam_size=beam, cutoff_prob=cp, cutoff_top_n=topn, vocab_list=vocab)
    probs = np.stack([_peaky_probs(rng, T, V, temp, bb) for bb in (0.0, 1.5)])
    got = dec.decode_ids_batch(probs)
    for b in range(probs.shape[0]):
        ref = DO.ctc_beam_search_decoding(probs[b], vocab, beam, cp, topn)
        assert got[b][0][1] == ref[0][1], f"best hypothesis differs (utt {b})"
        assert abs(got[b][0][0] - ref[0][0]) < 1e-3 * max(1.0, abs(ref[0][0]))
        assert len(got[b]) == len(ref)
        rd = dict((t, s) for s, t in ref)
        hits = sum(1 for s, t in got[b] if t in rd and abs(rd[t] - s) < 1e-3 * max(1.0, abs(s)))
        assert hits >= int(0.9 * len(ref)), f"only {hits} of {len(ref)} beam entries agree"
        # best first
        sc = [s for s, _ in got[b]]
        assert all(sc[i] <= sc[i + 1] + 1e-6 for i in range(len(sc) - 1))


@pytest.mark.parametrize("V", [4233, 5003, 300])
def test_ctc_prune_matches_oracle_rows(lib, cuda, V):
    """decoder_utils.cpp get_pruned_log_probs on peaked, flat and tie-heavy rows (the tie rows take the fallback path). The
    vocabulary sizes cover both kernels: CTA per row (V >= 512; 4233 = two load batches, 5003 = a third partial one) and warp
    per row (V = 300)."""
    from oracle import decoders_oracle as DO
    from ppasr_b200 import _lib as L
    rng = np.random.RandomState(5)
    rows = []
    for temp in (8.0, 3.0, 1.0, 0.2):
        lg = rng.randn(6, V).astype(np.float32) * temp
        e = np.exp(lg - lg.max(-1, keepdims=True))
        rows.append((e / e.sum(-1, keepdims=True)).astype(np.float32))
    rows.append(np.full((2, V), 1.0 / V, np.float32))                    # uniform: every element ties
    tie = np.zeros((2, V), np.float32); tie[:, 100:170] = 1.0 / 70       # 70 equal candidates > 64
    rows.append(tie)
    probs = torch.from_numpy(np.concatenate(rows)).to(cuda).contiguous()
    R = probs.shape[0]
    for cp, topn in ((0.99, 40), (1.0, 40), (0.5, 64), (0.999, 5)):
        ws = torch.zeros(lib.ppasr_b200_beam_workspace_bytes(1, R), dtype=torch.uint8, device=cuda)
        L.check(lib.ppasr_b200_op_ctc_prune(L.ptr(probs), R, V, cp, topn, L.ptr(ws), L.stream_ptr()))
        torch.cuda.synchronize()
        w = ws.cpu().numpy()
        cnt = w[:R * 4].view(np.int32)
        off = ((R + 63) // 64 * 64) * 4
        cid = w[off:off + R * 64 * 4].view(np.int32).reshape(R, 64)
        clp = w[off + R * 64 * 4:off + 2 * R * 64 * 4].view(np.float32).reshape(R, 64)
        pn = probs.cpu().numpy()
        for r in range(R):
            ref = DO.get_pruned_log_probs(pn[r], cp, topn)
            assert cnt[r] == len(ref), (r, cp, topn, cnt[r], len(ref))
            assert cid[r, :cnt[r]].tolist() == [i for i, _ in ref], (r, cp, topn)
            np.testing.assert_allclose(clp[r, :cnt[r]], [l for _, l in ref], rtol=1e-5, atol=1e-5)


def test_beam_search_stream_overflow_raises(lib, cuda):
    """A stream fed past max_frames must fail loudly instead of returning truncated transcripts."""
    from ppasr_b200 import _lib as L
    from ppasr_b200.decoders.beam_search_decoder import BeamSearchDecoder
    rng = np.random.RandomState(2)
    V = 30
    vocab = [f"<{i}>" for i in range(V)]
    dec = BeamSearchDecoder(beam_size=8, vocab_list=vocab, max_frames=40)
    probs = _peaky_probs(rng, 64, V, 3.0, 0.5)
    dec.decode_chunk(probs[None, :32], np.array([32]))
    with pytest.raises(L.PPASRB200Error):
        dec.decode_chunk(probs[None, 32:64], np.array([32]))
    # the C-ABI itself refuses as well: the state is flagged and result() reports -2
    st = dec._alloc_state(1, 16)
    dec._advance(st, dec._to_cuda(probs[None, :32]), None, 16)
    with pytest.raises(L.PPASRB200Error):
        dec._results(st, 1, 16, 16, nbest=1)
    dec.reset_decoder()
    assert dec.decode_chunk(probs[None, :32], np.array([32]))[1] is not None


def test_beam_search_streaming_equals_offline(lib, cuda):
    from ppasr_b200.decoders.beam_search_decoder import BeamSearchDecoder
    rng = np.random.RandomState(9)
    V, T = 80, 96
    vocab = [f"<{i}>" for i in range(V)]
    probs = _peaky_probs(rng, T, V, 4.0, 1.0)
    dec = BeamSearchDecoder(beam_size=10, cutoff_prob=0.99, cutoff_top_n=40, vocab_list=vocab, max_frames=200)
    off = dec.decode_beam_search_offline(probs)
    res = None
    for s in range(0, T, 16):
        res = dec.decode_chunk(probs[None, s:s + 16], np.array([16]))
    assert res[1] == off[1] and abs(res[0] - off[0]) < 1e-4
    dec.reset_decoder()
    r2 = dec.decode_chunk(probs[None, :16], np.array([16]))
    d2 = BeamSearchDecoder(beam_size=10, cutoff_prob=0.99, cutoff_top_n=40, vocab_list=vocab, max_frames=200)
    assert r2 == d2.decode_chunk(probs[None, :16], np.array([16]))


def test_beam_search_full_size_vs_greedy(lib, cuda):
    """BASELINE configs[2]-like posterior [8,748,4233], beam 10: on peaked posteriors the best beam equals greedy."""
    from ppasr_b200.decoders.beam_search_decoder import BeamSearchDecoder
    from ppasr_b200.decoders.ctc_greedy_decoder import greedy_decode_ids
    torch.manual_seed(3)
    V = 4233
    probs = torch.softmax(torch.randn(8, 748, V, device=cuda) * 8, -1)
    vocab = [f"<{i}>" for i in range(V)]
    dec = BeamSearchDecoder(beam_size=10, cutoff_prob=0.99, cutoff_top_n=40, vocab_list=vocab)
    res = dec.decode_ids_batch(probs)
    ids, ol, _, _, _ = greedy_decode_ids(probs)
    for b in range(8):
        assert res[b][0][1] == "".join(vocab[i] for i in ids[b, :ol[b]])


def test_decode_pipeline_matches_sync_api(lib, cuda):
    """Double-buffered public API (two engines / streams) returns exactly what predict_decode returns -- bit for bit when both
    run the same fused_ffn variant (the pipeline switches to the one-CTA-per-tile variant, which differs from the cluster
    variant in fp32 summation order)."""
    from ppasr_b200.infer_utils.inference_predictor import InferencePredictor
    from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank
    cfg = ConformerConfig(num_blocks=2, vocab_size=300)
    w = init_conformer_weights(cfg)
    pred = InferencePredictor({"encoder_conf": cfg.to_dict(), "preprocess_conf": {"n_mels": 80}}, "conformer",
                              streaming=True, weights=w)
    batches = [torch.from_numpy(synthetic_fbank(4, 300, seed=s)).pin_memory() for s in (1, 2, 3)]
    lib.ppasr_b200_set_ffn_split(2)   # the pipeline's variant
    try:
        ref = [pred.predict_decode(b.numpy()) for b in batches]
    finally:
        lib.ppasr_b200_set_ffn_split(1)
    pipe = pred.pipeline(depth=2)
    tickets = []
    out = []
    for b in batches:
        tickets.append(pipe.submit(b))
        if len(tickets) == 2:
            out.append(pipe.result(tickets.pop(0)))
    while tickets:
        out.append(pipe.result(tickets.pop(0)))
    pipe.close()
    for (ids, ol, sc), (rids, rol, rsc) in zip(out, ref):
        assert np.array_equal(ol, rol) and sc == rsc
        for b in range(4):
            assert np.array_equal(ids[b, :ol[b]], rids[b, :rol[b]])


@pytest.mark.parametrize("B,T,lens,n_mels", [(3, 523, [523, 333, 260], 80), (5, 67, [67, 67, 50, 30, 67], 80), (1, 998, [998], 80),
                                             (3, 300, [300, 211, 64], 40), (2, 131, [131, 99], 64)])
def test_fused_conv_front_bit_identical(lib, cuda, B, T, lens, n_mels):
    """The fused front ends == the two-kernel paths they mirror, bit for bit: conv_front.cu (CUDA-core conv1 inside the conv2
    GEMM's A producer) vs conv1_subsample + conv2 GEMM (same fp32 FMA order), and conv_front_tc (split-tf32 conv1 GEMM as the
    producer, conv1_tc.cu) vs conv1_tc + conv2 GEMM (same MMAs in the same order)."""
    from ppasr_b200.engine import ConformerEngine
    from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank
    cfg = ConformerConfig(num_blocks=1, vocab_size=300, input_dim=n_mels)
    eng = ConformerEngine(cfg, init_conformer_weights(cfg))
    feats = synthetic_fbank(B, T, n_mels=n_mels)
    for b in range(B):
        feats[b, lens[b]:] = 0
    fd = torch.from_numpy(feats).cuda()
    # conv_front.cu's patch geometry is built for 80 mel bins; the tensor-core front end takes any feature width <= 96
    for conv1_tc, fused in (((0, 1), (1, 2)) if n_mels == 80 else ((1, 2),)):
        outs = []
        eng.set_option("conv1_tc", conv1_tc)
        for f in (0, fused):
            eng.set_option("fused_conv", f)
            eng.encode(fd, lens)
            outs.append(eng.ctc_logits().float().cpu())
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]), f"fused_conv={fused}: max diff {(outs[0] - outs[1]).abs().max().item():.3e}"
    eng.close()


@pytest.mark.parametrize("B,T,lens,n_mels", [(3, 523, [523, 333, 260], 80), (5, 67, [67, 67, 50, 30, 67], 80), (2, 998, [998, 700], 80),
                                             (32, 200, [200] * 32, 80), (3, 300, [300, 211, 64], 40)])
def test_conv1_phase_images_match_fp32_conv2d(lib, cuda, B, T, lens, n_mels):
    """GlobalCMVN + Conv2d(1, 256, 3, 2) + ReLU (cmvn.py:29-32, subsampling.py:84-85) as written into the stride-phase images:
    the tensor-core kernel (split-tf32, conv1_tc.cu, default) and the CUDA-core kernel against torch fp32 conv2d. Both carry
    fp32-accurate sums and ONE bf16 rounding, so each must sit within one bf16 ulp of the fp32 value, and the two kernels must
    agree bit for bit on all but a few ties."""
    import torch.nn.functional as F
    from ppasr_b200.engine import ConformerEngine
    from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank
    cfg = ConformerConfig(num_blocks=1, vocab_size=300, input_dim=n_mels)
    w = init_conformer_weights(cfg)
    eng = ConformerEngine(cfg, w)
    feats = synthetic_fbank(B, T, n_mels=n_mels)
    for b in range(B):
        feats[b, lens[b]:] = 0
    fd = torch.from_numpy(feats).cuda()
    imgs = []
    eng.set_option("fused_conv", 0)  # the default front end (conv_front_tc) never materialises the phase images
    for tc in (0, 1):
        eng.set_option("conv1_tc", tc)
        eng.encode(fd, lens)
        imgs.append(eng.debug_phase().float().cpu())
    torch.cuda.synchronize()
    eng.close()
    mean = torch.from_numpy(np.asarray(w["encoder.global_cmvn.mean"], dtype=np.float32))
    istd = torch.from_numpy(np.asarray(w["encoder.global_cmvn.istd"], dtype=np.float32))
    xn = (torch.from_numpy(feats) - mean) * istd
    c1 = F.relu(F.conv2d(xn.double().unsqueeze(1), torch.from_numpy(np.asarray(w["encoder.embed.conv.0.weight"])).double(),
                         torch.from_numpy(np.asarray(w["encoder.embed.conv.0.bias"])).double(), stride=2))  # [B, 256, T1, F1]
    _, _, T1, F1 = c1.shape
    Th, FH = imgs[0].shape[2], imgs[0].shape[3]
    ref = torch.zeros((4, B, Th, FH, 256), dtype=torch.float64)
    for pt in range(2):
        for pf in range(2):
            sub = c1[:, :, pt::2, pf::2].permute(0, 2, 3, 1)  # [B, th, f2, C]
            ref[2 * pt + pf, :, :sub.shape[1], :sub.shape[2]] = sub
    for name, img in zip(("cuda-core", "tensor-core"), imgs):
        err = (img.double() - ref).abs()
        tol = ref.abs() * 2.0 ** -8 + 1e-5  # half a bf16 ulp (round to nearest) of the exact value
        bad = (err > tol * 1.001).sum().item()
        assert bad == 0, f"{name}: {bad} elements beyond one bf16 rounding of the fp64 conv (max err {err.max().item():.3e})"
    differ = (imgs[0] != imgs[1]).float().mean().item()
    assert differ < 1e-3, f"tensor-core and CUDA-core conv1 differ on {differ:.2e} of the elements"


# ------------------------------------------------------------------------------------------------
# DeepSpeech2 (deepspeech2/encoder.py:61-104): conv -> {bi-}LSTM / GRU + LayerNorm x N -> CTC
# ------------------------------------------------------------------------------------------------
def _run_ds2(cuda, nl, H, B, T, lens, streaming=False, gru=False, vocab=300):
    from oracle.deepspeech2_oracle import DeepSpeech2Conf, DeepSpeech2Oracle
    from oracle import decoders_oracle as DO
    from ppasr_b200.engine import ConformerEngine
    from ppasr_b200.weights import DeepSpeech2Config, init_deepspeech2_weights, synthetic_fbank
    cfg = DeepSpeech2Config(num_rnn_layers=nl, rnn_size=H, vocab_size=vocab, streaming=streaming, use_gru=gru)
    w = init_deepspeech2_weights(cfg)
    feats = synthetic_fbank(B, T)
    for b in range(B):
        feats[b, lens[b]:] = 0
    eng = ConformerEngine(cfg, w)
    eng.encode(torch.from_numpy(feats).to(cuda), lens)
    logits = eng.ctc_logits().float().cpu()
    probs = eng.ctc_probs().cpu().numpy()
    ids, ol, sc = eng.ctc_greedy(to_host=True)
    ref = DeepSpeech2Oracle(DeepSpeech2Conf(**cfg.to_dict()), w).get_encoder_out(
        torch.from_numpy(feats), torch.tensor(lens), return_logits=True)
    rel = ((logits - ref).abs().max() / ref.abs().max()).item()
    assert rel < 1e-2, f"logits rel err {rel}"
    for b in range(B):  # fused head == reference greedy on the materialised posterior (bit-exact ids)
        _, coll, _ = DO.greedy_ids(probs[b])
        assert ids[b, :ol[b]].tolist() == coll
    eng.close()


@pytest.mark.parametrize("kw", [
    dict(nl=1, H=128, B=2, T=131, lens=[131, 90]),                       # bidirectional LSTM, ragged
    dict(nl=2, H=256, B=3, T=203, lens=[203, 150, 99], streaming=True),  # forward LSTM
    dict(nl=2, H=256, B=3, T=203, lens=[203, 150, 99], gru=True),        # bidirectional GRU
    dict(nl=2, H=128, B=11, T=131, lens=[131] * 5 + [100] * 6),          # batch > one shared-memory tile of 8
    dict(nl=1, H=64, B=1, T=7, lens=[7]),                                # one output frame
])
def test_deepspeech2_matches_oracle_small(lib, cuda, kw):
    _run_ds2(cuda, **kw)


def test_deepspeech2_matches_oracle_full_size(lib, cuda):
    """configs/deepspeech2.yml (5 x bi-LSTM 1024, V=4233) on BASELINE configs[0]: one 5 s utterance."""
    _run_ds2(cuda, 5, 1024, 1, 498, [498], vocab=4233)


def test_deepspeech2_chunk_streaming(lib, cuda):
    """predict_chunk_deepspeech (inference_predictor.py:147-182): chunk-by-chunk with carried states == the oracle's
    get_encoder_out_chunk chain; returned state arrays have the reference shape."""
    from oracle.deepspeech2_oracle import DeepSpeech2Conf, DeepSpeech2Oracle
    from ppasr_b200.infer_utils.inference_predictor import InferencePredictor
    from ppasr_b200.weights import DeepSpeech2Config, init_deepspeech2_weights, synthetic_fbank
    cfg = DeepSpeech2Config(num_rnn_layers=2, rnn_size=256, vocab_size=120, streaming=True)
    w = init_deepspeech2_weights(cfg)
    pred = InferencePredictor({"encoder_conf": cfg.to_dict(), "preprocess_conf": {"n_mels": 80}}, "deepspeech2",
                              streaming=True, weights=w)
    o = DeepSpeech2Oracle(DeepSpeech2Conf(**cfg.to_dict()), w)
    x = synthetic_fbank(2, 67 + 64 * 2)
    h = c = None
    for s in range(0, x.shape[1] - 66, 64):
        ch = x[:, s:s + 67]
        probs, lens = pred.predict_chunk_deepspeech(ch)
        ref, rl, h, c = o.get_encoder_out_chunk(torch.from_numpy(ch), torch.tensor([67, 67]), h, c, return_logits=True)
        assert probs.shape == tuple(ref.shape) and lens.tolist() == rl.tolist()
        assert np.allclose(probs.sum(-1), 1.0, atol=1e-4)
        lg = pred.engine.ctc_logits().float().cpu()
        assert ((lg - ref).abs().max() / ref.abs().max()).item() < 1e-2
        assert pred.output_state_h.shape == (2, 2, 256)
        assert np.abs(pred.output_state_h - h.numpy()).max() < 2e-2 and np.abs(pred.output_state_c - c.numpy()).max() < 5e-2
    pred.reset_stream()
    with pytest.raises(Exception):
        pred.predict_chunk_conformer(x[:1, :67], -1)


# ------------------------------------------------------------------------------------------------
# beam search with the external n-gram scorer (swig_wrapper.py:4-19, beam_search_decoder.py:29-40)
# ------------------------------------------------------------------------------------------------
def _toy_lm(V, order, seed=0):
    from ppasr_b200.decoders.ngram_lm import NGramLM
    rng = np.random.RandomState(seed)
    vocab = ["<blank>", "<unk>"] + [chr(0x4E00 + i) for i in range(V - 3)] + ["<eos>"]
    sents = [[vocab[2 + int(z) % (V - 3)] for z in rng.zipf(1.5, size=rng.randint(3, 12))] for _ in range(400)]
    return vocab, NGramLM.from_counts(sents, order=order)


@pytest.mark.parametrize("T,V,beam,order,alpha,beta,temp", [
    (30, 20, 10, 4, 2.2, 4.3, 3.0), (60, 40, 20, 3, 1.0, 0.5, 4.0), (40, 30, 5, 2, 0.5, 0.0, 2.0), (25, 12, 8, 4, 2.2, 4.3, 1.0),
    (40, 20, 10, 5, 2.2, 4.3, 2.0), (30, 15, 16, 5, 1.0, 1.0, 1.0),   # order 5: the order the reference trains its LMs with
])
def test_beam_search_with_ngram_scorer_matches_oracle(lib, cuda, tmp_path, T, V, beam, order, alpha, beta, temp):
    from oracle import decoders_oracle as DO
    from ppasr_b200.decoders.beam_search_decoder import BeamSearchDecoder
    from ppasr_b200.decoders.ngram_lm import Scorer
    vocab, lm = _toy_lm(V, order)
    arpa = str(tmp_path / "toy.arpa")
    lm.write_arpa(arpa)
    dec = BeamSearchDecoder(alpha=alpha, beta=beta, beam_size=beam, cutoff_prob=0.99, cutoff_top_n=40, vocab_list=vocab,
                            language_model_path=arpa)
    sc = Scorer(alpha, beta, arpa, vocab)
    rng = np.random.RandomState(T * 7 + V)
    probs = np.stack([_peaky_probs(rng, T, V, temp, bb) for bb in (0.0, 1.0)])
    got = dec.decode_ids_batch(probs)
    for b in range(probs.shape[0]):
        ref = DO.ctc_beam_search_decoding(probs[b], vocab, beam, 0.99, 40, ext_scorer=sc)
        assert got[b][0][1] == ref[0][1], f"best hypothesis differs (utt {b})"
        assert abs(got[b][0][0] - ref[0][0]) < 2e-3 * max(1.0, abs(ref[0][0]))
        # the whole beam (as a set of strings) agrees as well when the margins are comfortable
        assert len(got[b]) == len(ref)


def test_beam_search_scorer_changes_result_and_rejects_bad_models(lib, cuda, tmp_path):
    from ppasr_b200.decoders.beam_search_decoder import BeamSearchDecoder
    vocab, lm = _toy_lm(20, 3)
    arpa = str(tmp_path / "toy.arpa")
    lm.write_arpa(arpa)
    rng = np.random.RandomState(3)
    probs = _peaky_probs(rng, 40, 20, 1.0, 0.0)[None]
    plain = BeamSearchDecoder(beam_size=10, vocab_list=vocab).decode_ids_batch(probs)[0][0]
    with_lm = BeamSearchDecoder(alpha=3.0, beta=1.0, beam_size=10, vocab_list=vocab, language_model_path=arpa).decode_ids_batch(probs)[0][0]
    assert plain[1] != with_lm[1]  # flat posteriors: the LM decides
    bad = tmp_path / "model.klm"
    bad.write_bytes(b"mmap lm http://kheafield.com/code format version 5\\n\\x00")
    with pytest.raises(Exception):
        BeamSearchDecoder(beam_size=10, vocab_list=vocab, language_model_path=str(bad))


# ------------------------------------------------------------------------------------------------
# GPU fbank front end (csrc/fbank.cu) vs the oracle restatement of AudioFeaturizer.featurize
# ------------------------------------------------------------------------------------------------
def _wave(seconds, seed, amp=0.2):
    rng = np.random.RandomState(seed)
    t = np.arange(int(16000 * seconds)) / 16000.0
    x = 0.1 * rng.randn(t.size) + 0.3 * np.sin(2 * np.pi * (200 + 300 * seed) * t) * (1 + 0.5 * np.sin(2 * np.pi * 4 * t))
    return (x * amp).astype(np.float32)


@pytest.mark.parametrize("db", [True, False])
def test_gpu_fbank_matches_oracle(lib, cuda, db):
    from oracle import fbank_oracle as FO
    from ppasr_b200.featurizer import GpuFbank
    fb = GpuFbank(use_dB_normalization=db)
    waves = [_wave(2.0, 0), _wave(2.0, 1, 0.01), _wave(2.0, 2, 0.9)]
    lens = [32000, 20000, 399]
    batch = np.stack(waves)
    for b in range(3):
        batch[b, lens[b]:] = 0
    out, counts = fb.featurize_batch(batch, lens)
    out = out.cpu().numpy()
    assert out.shape == (3, 198, 80) and counts == [198, 123, 0]
    for b in range(3):
        ref = FO.featurize(batch[b, :lens[b]], use_db_normalization=db) if lens[b] >= 400 else np.zeros((0, 80), np.float32)
        got = out[b, :counts[b]]
        # fp32 FFT orderings differ (pocketfft vs radix-2): compare energies relatively, logs absolutely
        if counts[b]:
            assert np.abs(got - ref).max() < 5e-3, np.abs(got - ref).max()
        assert np.all(out[b, counts[b]:] == 0)
    assert fb.featurize(waves[0]).shape == (198, 80)


def test_predict_batch_from_waveforms(lib, cuda):
    """Waveform batch -> text entirely on the GPU == fbank on the host (reference path) -> predict."""
    from ppasr_b200.predict import PPASRPredictor
    from ppasr_b200.weights import ConformerConfig, init_conformer_weights, make_vocab
    cfg = ConformerConfig(num_blocks=2, vocab_size=200)
    vocab = make_vocab(200)
    p = PPASRPredictor({"use_model": "conformer", "streaming": True, "decoder": "ctc_greedy", "encoder_conf": cfg.to_dict(),
                        "preprocess_conf": {"feature_method": "fbank", "n_mels": 80}}, vocab_list=vocab,
                       weights=init_conformer_weights(cfg))
    waves = np.stack([_wave(3.0, 0), _wave(3.0, 1)])
    got = p.predict_batch(waves)

    def edit_distance(a, b):
        prev = list(range(len(b) + 1))
        for i, ca in enumerate(a, 1):
            cur = [i]
            for j, cb in enumerate(b, 1):
                cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
            prev = cur
        return prev[-1]

    for b in range(2):
        ref = p.predict(waves[b])
        # The two feature paths differ by ~1e-3 (fp32 FFT orderings, see test_gpu_fbank_matches_oracle) and a random-init model
        # has near-tied posteriors, so a few frames may flip between two symbols: bound the edit distance, not equality.
        assert edit_distance(got[b]['text'], ref['text']) <= max(2, len(ref['text']) // 10), (got[b]['text'], ref['text'])
        assert abs(got[b]['score'] - ref['score']) < 0.5


@pytest.mark.parametrize("nb,reduce_idx,recover_idx,required", [(4, 1, 3, -1), (4, 1, 3, 32), (3, None, None, -1), (12, 5, 11, -1)])
def test_squeezeformer_chunk_streaming_matches_oracle(lib, cuda, nb, reduce_idx, recover_idx, required):
    """squeezeformer/encoder.py:260-383 forward_chunk: 67-frame windows with stride 64, device-resident caches (the half-rate
    blocks keep theirs at half rate); compared chunk by chunk on the logits with the oracle's forward_chunk chain (which equals
    the chunk-masked offline forward, tests/test_oracle_cpu.py)."""
    from oracle.squeezeformer_oracle import SqueezeformerConf, SqueezeformerOracle
    from ppasr_b200.infer_utils.inference_predictor import InferencePredictor
    from ppasr_b200.weights import SqueezeformerConfig, init_squeezeformer_weights, synthetic_fbank
    cfg = SqueezeformerConfig(num_blocks=nb, vocab_size=120, reduce_idx=reduce_idx, recover_idx=recover_idx)
    w = init_squeezeformer_weights(cfg)
    orc = SqueezeformerOracle(SqueezeformerConf(**cfg.to_dict()), w)
    pred = InferencePredictor({"encoder_conf": cfg.to_dict(), "preprocess_conf": {"n_mels": 80}}, "squeezeformer",
                              streaming=True, weights=w)
    feats = synthetic_fbank(1, 67 + 64 * 3)
    att = torch.zeros(0, 0, 0, 0)
    cnn = torch.zeros(0, 0, 0, 0)
    off = 0
    for s in range(0, feats.shape[1] - 66, 64):
        ch = feats[:, s:s + 67]
        ref, att, cnn = orc.get_encoder_out_chunk(torch.from_numpy(ch), off, required, att, cnn, return_logits=True)
        off += ref.shape[1]
        probs = pred.predict_chunk_conformer(ch, required)
        assert probs.shape == tuple(ref.shape) and np.allclose(probs.sum(-1), 1.0, atol=1e-4)
        lg = pred.engine.ctc_logits().float().cpu()
        assert ((lg - ref).abs().max() / ref.abs().max()).item() < 1e-2, s
        assert int(pred.offset[0]) == off
    pred.reset_stream()


@pytest.mark.parametrize("nb,group_idx,stride_idx,batch", [(4, (0, 1), 1, 1), (4, (0, 1, 2), 2, 3), (3, (), None, 2), (12, (0, 1, 2, 3), 3, 2)])
def test_efficient_conformer_chunk_streaming_matches_oracle(lib, cuda, nb, group_idx, stride_idx, batch):
    """efficient_conformer/encoder.py:266-394 forward_chunk on the device (append-only grouped K / V^T caches, per-chunk padded
    positional operand, stride block over [cache | chunk], half-rate blocks behind it): 67-frame windows with stride 64 and a
    short last window, `batch` lock-step streams, chunk by chunk on the logits against the oracle's forward_chunk chain (which
    is pinned to the reference code's own streaming outputs, tests/test_encoder_golden_cpu.py)."""
    from oracle.efficient_conformer_oracle import EfficientConformerConf, EfficientConformerOracle
    from ppasr_b200.infer_utils.inference_predictor import InferencePredictor
    from ppasr_b200.weights import EfficientConformerConfig, init_efficient_conformer_weights, synthetic_fbank
    cfg = EfficientConformerConfig(num_blocks=nb, vocab_size=120, group_layer_idx=group_idx, stride_layer_idx=stride_idx)
    w = init_efficient_conformer_weights(cfg)
    orc = EfficientConformerOracle(EfficientConformerConf(**cfg.to_dict()), w)
    pred = InferencePredictor({"encoder_conf": cfg.to_dict(), "preprocess_conf": {"n_mels": 80}}, "efficient_conformer",
                              streaming=True, weights=w)
    feats = synthetic_fbank(batch, 67 + 64 * 3 + 24)
    states = [(torch.zeros(0, 0, 0, 0), torch.zeros(0, 0, 0, 0), 0) for _ in range(batch)]
    total = 0
    for s in range(0, feats.shape[1] - 6, 64):
        ch = feats[:, s:s + 67]
        refs = []
        for b in range(batch):   # the oracle streams one utterance at a time like the reference
            att, cnn, off = states[b]
            ref, att, cnn = orc.get_encoder_out_chunk(torch.from_numpy(ch[b:b + 1]), off, -16, att, cnn, return_logits=True)
            states[b] = (att, cnn, off + ref.shape[1])
            refs.append(ref)
        ref = torch.cat(refs, 0)
        total += ref.shape[1]
        probs = pred.predict_chunk_conformer(ch, -16)
        assert probs.shape == tuple(ref.shape) and np.allclose(probs.sum(-1), 1.0, atol=1e-4)
        lg = pred.engine.ctc_logits().float().cpu()
        assert ((lg - ref).abs().max() / ref.abs().max()).item() < 1e-2, s
        assert int(pred.offset[0]) == total
    # bounded histories are not what PPASRPredictor asks for and are refused; so is a stream past the grouped capacity
    with pytest.raises(Exception):
        pred.predict_chunk_conformer(feats[:, :67], 32)
    pred.reset_stream()
    probs = pred.predict_chunk_conformer(feats[:, :67], -16)
    assert probs.shape[1] == (16 if stride_idx is None else 8)
    pred.reset_stream()


@pytest.mark.parametrize("model", ["conformer", "squeezeformer"])
def test_fused_dwconv_bit_identical(lib, cuda, model):
    """The depthwise conv + norm + swish stage computed inside the chained FFN kernel (option fused_dwconv, default) equals the
    stand-alone dwconv kernel bit for bit (same FMA order, same bf16 roundings), incl. utterance boundaries inside a row tile
    and padded frames."""
    from ppasr_b200.engine import ConformerEngine
    from ppasr_b200 import weights as W
    if model == "conformer":
        cfg = W.ConformerConfig(num_blocks=2, vocab_size=300)
        w = W.init_conformer_weights(cfg)
    else:
        cfg = W.SqueezeformerConfig(num_blocks=3, vocab_size=300, reduce_idx=1, recover_idx=2)
        w = W.init_squeezeformer_weights(cfg)
    eng = ConformerEngine(cfg, w)
    B, T, lens = 5, 363, [363, 200, 363, 90, 300]   # T' = 90: row tiles of 128 straddle utterances
    feats = W.synthetic_fbank(B, T)
    for b in range(B):
        feats[b, lens[b]:] = 0
    fd = torch.from_numpy(feats).cuda()
    outs = []
    lib.ppasr_b200_set_ffn_split(0)   # fused_dwconv lives in the round-1 fused_ffn kernel: compare within that variant
    try:
        for fused in (0, 1):
            eng.set_option("fused_dwconv", fused)
            eng.encode(fd, lens)
            outs.append(eng.ctc_logits().float().cpu())
        torch.cuda.synchronize()
    finally:
        lib.ppasr_b200_set_ffn_split(1)
    eng.close()
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("use_model", ["deepspeech2", "squeezeformer", "efficient_conformer"])
def test_predict_stream_other_models(lib, cuda, use_model):
    """PPASRPredictor.predict_stream_features dispatches to predict_chunk_deepspeech / predict_chunk_conformer by model
    (predict.py:302-311); feeding arbitrary slices equals feeding the reference's 67/64 windows by hand."""
    from oracle import decoders_oracle as DO
    from oracle.conformer_oracle import stream_windows
    from ppasr_b200.predict import PPASRPredictor
    from ppasr_b200 import weights as W
    if use_model == "deepspeech2":
        cfg = W.DeepSpeech2Config(num_rnn_layers=2, rnn_size=256, vocab_size=150, streaming=True)
        w = W.init_deepspeech2_weights(cfg)
    elif use_model == "efficient_conformer":   # forward_chunk on the device: grouped blocks 0-1, stride block 1, odd last chunk
        cfg = W.EfficientConformerConfig(num_blocks=4, vocab_size=150, group_layer_idx=(0, 1), stride_layer_idx=1)
        w = W.init_efficient_conformer_weights(cfg)
    else:
        cfg = W.SqueezeformerConfig(num_blocks=3, vocab_size=150, reduce_idx=1, recover_idx=2)
        w = W.init_squeezeformer_weights(cfg)
    vocab = W.make_vocab(150)
    p = PPASRPredictor({"use_model": use_model, "streaming": True, "decoder": "ctc_greedy", "encoder_conf": cfg.to_dict(),
                        "preprocess_conf": {"feature_method": "fbank", "n_mels": 80}}, vocab_list=vocab, weights=w)
    T = 67 + 64 * 3 + 21   # the tail window at is_end is short (odd number of output frames)
    feats = W.synthetic_fbank(1, T)
    res, pos = None, 0
    for n in (40, 90, 100, 50):
        r = p.predict_stream_features(feats[:, pos:pos + n], is_end=(pos + n >= T))
        pos += n
        res = r if r is not None else res
    assert pos == T and res is not None
    p.reset_stream()
    chunks = []
    for (s, e) in stream_windows(T, is_end=True):
        if use_model == "deepspeech2":
            chunks.append(p.predictor.predict_chunk_deepspeech(feats[:, s:e])[0][0])
        else:
            chunks.append(p.predictor.predict_chunk_conformer(feats[:, s:e], -16)[0])
    score, text = DO.greedy_decoder(np.concatenate(chunks, 0), vocab)
    assert res["text"] == text and abs(res["score"] - score) < 1e-3
    p.reset_stream()


def test_stream_scheduler_ragged_sessions(lib, cuda):
    """Three sessions opened at different times, fed in uneven pieces and stepped in shared batches give exactly the transcript
    each of them gets alone through PPASRPredictor.predict_stream_features (same windows, same chunk-wise greedy decoder); the
    chunk posteriors of a session stepped in a ragged batch equal those of a solo stream up to batch-shape effects."""
    from ppasr_b200.infer_utils.stream_scheduler import StreamScheduler
    from ppasr_b200.predict import PPASRPredictor
    from ppasr_b200 import weights as W
    cfg = W.ConformerConfig(num_blocks=2, vocab_size=150)
    w = W.init_conformer_weights(cfg)
    vocab = W.make_vocab(150)
    configs = {"use_model": "conformer", "streaming": True, "decoder": "ctc_greedy", "encoder_conf": cfg.to_dict(),
               "preprocess_conf": {"feature_method": "fbank", "n_mels": 80}}
    p = PPASRPredictor(configs, vocab_list=vocab, weights=w)
    lens = [67 + 64 * 3 + 10, 67 + 64 * 1 + 30, 67 + 64 * 2]
    feats = [W.synthetic_fbank(1, n, seed=40 + i)[0] for i, n in enumerate(lens)]
    # solo references
    solo = []
    for f in feats:
        p.reset_stream()
        r = None
        for s in range(0, f.shape[0], 50):
            q = p.predict_stream_features(f[None, s:s + 50], is_end=(s + 50 >= f.shape[0]))
            r = q if q is not None else r
        solo.append(r)
    p.reset_stream()
    # scheduler: sessions start at rounds 0, 1, 3; 50 frames per round each
    sch = StreamScheduler(p.predictor, vocab, max_sessions=4)
    start = [0, 1, 3]
    sid = [None] * 3
    pos = [0] * 3
    last = {}
    rnd = 0
    while any(pos[i] < lens[i] for i in range(3)) or sch.pending():
        for i in range(3):
            if rnd >= start[i] and pos[i] < lens[i]:
                if sid[i] is None:
                    sid[i] = sch.open()
                n = min(50, lens[i] - pos[i])
                sch.feed(sid[i], feats[i][pos[i]:pos[i] + n], is_end=(pos[i] + n >= lens[i]))
                pos[i] += n
        last.update(sch.step())
        rnd += 1
        assert rnd < 100
    for i in range(3):
        assert last[sid[i]]["text"] == solo[i]["text"], i
        assert abs(last[sid[i]]["score"] - solo[i]["score"]) < 0.5
        assert sch.close(sid[i])["text"] == solo[i]["text"]
    assert len(sch._free) == 4


def test_predict_long_batched_regions(lib, cuda):
    """predict_long (predict.py:190-229): the speech regions of one recording as ragged GPU batches vs one predict() per
    region. The two routes differ only in the fbank implementation (GPU kernel vs torchaudio) and in batch padding, so the
    composed result must have the same structure and nearly the same score; every region on its own must decode identically
    wherever predict_batch on equal-length input already does (test_predict_batch_from_waveforms)."""
    from ppasr_b200.predict import PPASRPredictor
    from ppasr_b200.weights import ConformerConfig, init_conformer_weights, make_vocab
    cfg = ConformerConfig(num_blocks=2, vocab_size=200)
    p = PPASRPredictor({"use_model": "conformer", "streaming": True, "decoder": "ctc_greedy", "encoder_conf": cfg.to_dict(),
                        "preprocess_conf": {"feature_method": "fbank", "n_mels": 80}}, vocab_list=make_vocab(200),
                       weights=init_conformer_weights(cfg))
    audio = np.concatenate([_wave(2.0, 0), _wave(1.0, 1, 0.02), _wave(3.0, 2, 0.6)])
    stamps = [{"start": 800, "end": 30000}, {"start": 33000, "end": 47000}, {"start": 50000, "end": 95000},
              {"start": 95000, "end": 95200}]
    seq = p.predict_long(audio, speech_timestamps=stamps, batched=False)
    bat = p.predict_long(audio, speech_timestamps=stamps, batched=True)
    two = p.predict_long(audio, speech_timestamps=stamps, batched=True, max_batch_samples=2 * 45000)
    assert isinstance(bat["text"], str) and len(seq["text"]) > 0
    assert bat["text"].count("，") == seq["text"].count("，") == two["text"].count("，")
    assert abs(len(bat["text"]) - len(seq["text"])) <= 3
    assert abs(bat["score"] - seq["score"]) < 2.0 and abs(two["score"] - bat["score"]) < 2.0
    with pytest.raises(Exception, match="speech_timestamps"):
        p.predict_long(audio)
