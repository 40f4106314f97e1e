"""CPU tests of the oracle (test infrastructure) against the reference's known answers and golden vectors."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import conformer_oracle as CO
from oracle import decoders_oracle as DO
from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold_vocab(V):
    v = ["<blank>", "<unk>"] + [chr(0x4E00 + i) for i in range(V - 4)] + ["<space>", "<eos>"]
    return v[:V]


# ---- docstring known answers of ppasr/model_utils/utils/mask.py ------------------------------------
def test_make_pad_mask_kat():
    # mask.py:30-35
    m = CO.make_pad_mask(torch.tensor([5, 3, 2]))
    assert m.int().tolist() == [[0, 0, 0, 0, 0], [0, 0, 0, 1, 1], [0, 0, 1, 1, 1]]
    assert CO.make_non_pad_mask(torch.tensor([5, 3, 2])).int().tolist() == [[1, 1, 1, 1, 1], [1, 1, 1, 0, 0],
                                                                          [1, 1, 0, 0, 0]]


def test_subsequent_chunk_mask_kat():
    # mask.py:107-112
    assert CO.subsequent_chunk_mask(4, 2).int().tolist() == [[1, 1, 0, 0], [1, 1, 0, 0], [1, 1, 1, 1], [1, 1, 1, 1]]
    m = CO.subsequent_chunk_mask(6, 2, 1).int().tolist()
    assert m[4] == [0, 0, 1, 1, 1, 1] and m[0] == [1, 1, 0, 0, 0, 0]


def test_subsampled_mask_rule():
    """mask[:, :, :-2:2][:, :, :-2:2] keeps output frame j iff 4*j < len (conformer/subsampling.py:115)."""
    for T in (7, 8, 67, 131, 998):
        for L in (1, 3, 4, 5, T - 1, T):
            m = CO.make_non_pad_mask(torch.tensor([T, L]))[:, None, :]
            sub = m[:, :, :-2:2][:, :, :-2:2][1, 0]
            Tp = ((T - 1) // 2 - 1) // 2
            assert sub.shape[0] == Tp
            assert sub.tolist() == [4 * j < L for j in range(Tp)]


# ---- greedy decoder restatement pinned on the reference's own outputs ------------------------------
def _gold():
    z = np.load(os.path.join(GOLD, "greedy_golden.npz"))
    meta = json.load(open(os.path.join(GOLD, "greedy_golden.json"), encoding="utf-8"))
    return z, meta


def test_greedy_oracle_matches_reference_golden():
    z, meta = _gold()
    for m in meta:
        if m["name"] == "__batch__":
            continue
        probs = z[m["name"] + "_probs"]
        score, text = DO.greedy_decoder(probs, gold_vocab(m["V"]))
        assert text == m["text"], m["name"]
        assert repr(float(score)) == m["score"], m["name"]


def test_greedy_chunk_oracle_matches_reference_golden():
    z, meta = _gold()
    for m in meta:
        if m["name"] == "__batch__":
            continue
        probs = z[m["name"] + "_probs"]
        lp, li = None, None
        for c, s in enumerate(range(0, m["T"], 16)):
            sc, tx, lp, li = DO.greedy_decoder_chunk(probs[s:s + 16], gold_vocab(m["V"]), lp, li)
            assert tx == m["chunks"][c]["text"]
            assert repr(float(sc)) == m["chunks"][c]["score"]


def test_greedy_batch_oracle_matches_reference_golden():
    z, meta = _gold()
    m = [x for x in meta if x["name"] == "__batch__"][0]
    p = z["long_probs"]
    assert DO.greedy_decoder_batch([p, p[:100]], gold_vocab(97)) == m["texts"]


# ---- encoder oracle self-consistency (the pin against the reference code is tests/test_encoder_golden_cpu.py) ----
@pytest.fixture(scope="module")
def small():
    cfg = ConformerConfig(num_blocks=2, vocab_size=61)
    w = init_conformer_weights(cfg)
    return cfg, w


def test_chunked_equals_chunk_masked_offline(small):
    """forward(decoding_chunk_size=16, left=-1) == concat of forward_chunk over 67/64 windows."""
    cfg, w = small
    o = CO.ConformerOracle(CO.ConformerConf(**cfg.to_dict()), w)
    T = 195
    x = torch.from_numpy(synthetic_fbank(1, T))
    enc, _ = o.encoder_forward(x, torch.tensor([T]), decoding_chunk_size=16, num_decoding_left_chunks=-1)
    att = torch.zeros(0, 0, 0, 0)
    cnn = torch.zeros(0, 0, 0, 0)
    off, outs = 0, []
    for (s, e) in CO.stream_windows(T, is_end=True):
        y, att, cnn = o.encoder_forward_chunk(x[:, s:e], off, -16, att, cnn)
        off += y.shape[1]
        outs.append(y)
    yc = torch.cat(outs, 1)
    assert yc.shape[1] == enc.shape[1]
    assert (enc - yc).abs().max() < 2e-5
    assert att.shape == (cfg.num_blocks, cfg.attention_heads, yc.shape[1], 128)
    assert cnn.shape == (cfg.num_blocks, 1, cfg.output_size, cfg.cnn_module_kernel - 1)


def test_batch_invariance(small):
    """An utterance that fills the batch tensor is independent of its batch mates. (A *padded* utterance is
    not equal to its stand-alone run in the reference: the subsampled mask keeps frame j iff 4*j < len, i.e.
    one or two frames more than the stand-alone T', and those frames are attended to -- replicated as is.)"""
    cfg, w = small
    o = CO.ConformerOracle(CO.ConformerConf(**cfg.to_dict()), w)
    x = torch.from_numpy(synthetic_fbank(2, 131))
    x[1, 99:] = 0
    pb = o.get_encoder_out(x, torch.tensor([131, 99]))
    p0 = o.get_encoder_out(x[:1], torch.tensor([131]))
    assert (pb[0] - p0[0]).abs().max() < 1e-5
    # same padded utterance next to a different neighbour: unchanged
    x2 = x.clone()
    x2[0] = torch.from_numpy(synthetic_fbank(1, 131, seed=77))[0]
    pb2 = o.get_encoder_out(x2, torch.tensor([131, 99]))
    assert (pb[1] - pb2[1]).abs().max() < 1e-5


def test_fp64_noise_floor(small):
    cfg, w = small
    x = torch.from_numpy(synthetic_fbank(1, 131))
    l32 = CO.ConformerOracle(CO.ConformerConf(**cfg.to_dict()), w).get_encoder_out(x, torch.tensor([131]), True)
    l64 = CO.ConformerOracle(CO.ConformerConf(**cfg.to_dict()), w, torch.float64).get_encoder_out(x, torch.tensor([131]), True)
    rel = (l32.double() - l64).abs().max() / l64.abs().max()
    assert rel < 1e-4


def test_batchnorm_variant_runs():
    cfg = ConformerConfig(num_blocks=1, vocab_size=31, streaming=False, cnn_module_norm="batch_norm")
    w = init_conformer_weights(cfg)
    o = CO.ConformerOracle(CO.ConformerConf(**cfg.to_dict()), w)
    p = o.get_encoder_out(torch.from_numpy(synthetic_fbank(1, 67)), torch.tensor([67]))
    assert p.shape == (1, 16, 31) and torch.allclose(p.sum(-1), torch.ones(1, 16), atol=1e-5)


# ---- beam search restatement: properties (parity unpinned, see oracle/decoders_oracle.py) ----------
def test_beam_search_equals_bruteforce_on_tiny():
    rng = np.random.RandomState(3)
    for _ in range(5):
        T, V = 4, 3
        p = rng.dirichlet(np.ones(V), size=T).astype(np.float32)
        res = DO.ctc_beam_search_ids(p, beam_size=100, cutoff_prob=1.0, cutoff_top_n=V)
        brute = DO.ctc_prefix_total_logprob_bruteforce(p)
        got = {tuple(ids): s for s, ids in res}
        for lab, lp in brute.items():
            assert lab in got
            assert abs(got[lab] - lp) < 1e-4
        best = max(brute.items(), key=lambda kv: kv[1])[0]
        assert tuple(res[0][1]) == best


def test_beam1_equals_greedy_on_peaked():
    rng = np.random.RandomState(5)
    T, V = 30, 12
    logits = rng.randn(T, V) * 8
    p = np.exp(logits - logits.max(-1, keepdims=True))
    p = (p / p.sum(-1, keepdims=True)).astype(np.float32)
    _, coll, _ = DO.greedy_ids(p)
    res = DO.ctc_beam_search_ids(p, beam_size=1, cutoff_prob=1.0, cutoff_top_n=V)
    assert res[0][1] == coll


def test_pruned_log_probs():
    p = np.array([0.05, 0.6, 0.25, 0.1], dtype=np.float32)
    r = DO.get_pruned_log_probs(p, 0.8, 40)
    assert [i for i, _ in r] == [1, 2]
    r = DO.get_pruned_log_probs(p, 1.0, 3)
    assert [i for i, _ in r] == [1, 2, 3]
    assert abs(r[0][1] - math.log(0.6 + DO.NUM_FLT_MIN)) < 1e-7


# ------------------------------------------------------------------------------------------------
# Squeezeformer oracle self-consistency (pin against the reference code: tests/test_encoder_golden_cpu.py)
# ------------------------------------------------------------------------------------------------
def _squeeze(streaming=True, nb=3, **kw):
    from oracle.squeezeformer_oracle import SqueezeformerConf, SqueezeformerOracle
    from ppasr_b200.weights import SqueezeformerConfig, init_squeezeformer_weights
    cfg = SqueezeformerConfig(num_blocks=nb, vocab_size=60, streaming=streaming, **kw)
    w = init_squeezeformer_weights(cfg)
    return cfg, w, SqueezeformerOracle(SqueezeformerConf(**cfg.to_dict()), w)


@pytest.mark.parametrize("streaming", [True, False])
def test_squeezeformer_oracle_shapes_and_batch_invariance(streaming):
    from ppasr_b200.weights import synthetic_fbank
    cfg, w, o = _squeeze(streaming, reduce_idx=1, recover_idx=2)
    x = torch.from_numpy(synthetic_fbank(3, 203))
    lens = torch.tensor([203, 150, 99])
    y = o.get_encoder_out(x, lens, return_logits=True)
    assert y.shape == (3, 50, 60)
    y0 = o.get_encoder_out(x[:1], lens[:1], return_logits=True)
    assert (y0[0] - y[0]).abs().max() < 1e-3 * y.abs().max()
    # time-reduction kernel size follows the streaming switch (squeezeformer/model.py:35-41)
    assert w["encoder.time_reduction_layer.dw_conv.weight"].shape[-1] == (1 if streaming else 5)


def test_squeezeformer_oracle_fp64_noise_floor():
    from oracle.squeezeformer_oracle import SqueezeformerConf, SqueezeformerOracle
    from ppasr_b200.weights import synthetic_fbank
    cfg, w, o = _squeeze(True, reduce_idx=1, recover_idx=2)
    o64 = SqueezeformerOracle(SqueezeformerConf(**cfg.to_dict()), w, dtype=torch.float64)
    x = torch.from_numpy(synthetic_fbank(2, 131))
    lens = torch.tensor([131, 90])
    a = o.get_encoder_out(x, lens, return_logits=True)
    b = o64.get_encoder_out(x, lens, return_logits=True)
    assert ((a - b.float()).abs().max() / b.abs().max()).item() < 1e-4


def test_squeezeformer_oracle_identity_ada_and_no_reduction_is_plain_postnorm():
    """With ada_scale = 1 / ada_bias = 0 the adaptive-scale code path must be a no-op (positionwise.py:62-63)."""
    from oracle.squeezeformer_oracle import SqueezeformerConf, SqueezeformerOracle
    from ppasr_b200.weights import synthetic_fbank
    cfg, w, o = _squeeze(True, nb=2, reduce_idx=None, recover_idx=None)
    w2 = {k: (np.ones_like(v) if k.endswith("ada_scale") else np.zeros_like(v) if k.endswith("ada_bias") else v)
          for k, v in w.items()}
    conf_off = SqueezeformerConf(**{**cfg.to_dict(), "adaptive_scale": False})
    x = torch.from_numpy(synthetic_fbank(1, 131))
    lens = torch.tensor([131])
    a = SqueezeformerOracle(SqueezeformerConf(**cfg.to_dict()), w2).get_encoder_out(x, lens, return_logits=True)
    b = SqueezeformerOracle(conf_off, w).get_encoder_out(x, lens, return_logits=True)
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------
# DeepSpeech2 oracle: the restated paddle.nn.LSTM / GRU equations == torch.nn.LSTM / GRU on packed sequences
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("streaming,gru", [(False, False), (True, False), (False, True)])
def test_deepspeech2_oracle_rnn_matches_torch_nn(streaming, gru):
    from oracle.deepspeech2_oracle import DeepSpeech2Conf, DeepSpeech2Oracle
    from ppasr_b200.weights import DeepSpeech2Config, init_deepspeech2_weights
    cfg = DeepSpeech2Config(num_rnn_layers=1, rnn_size=64, vocab_size=30, streaming=streaming, use_gru=gru)
    w = init_deepspeech2_weights(cfg)
    o = DeepSpeech2Oracle(DeepSpeech2Conf(**cfg.to_dict()), w)
    x = torch.from_numpy(synthetic_fbank(3, 131))
    lens = torch.tensor([131, 100, 60])
    xs = (x - o.w["encoder.global_cmvn.mean"]) * o.w["encoder.global_cmvn.istd"]
    c1, l1 = o.conv(xs, lens)
    assert c1.shape == (3, 32, 32 * 19) and l1.tolist() == [32, 24, 14]
    nd = cfg.num_directions
    mod = (torch.nn.GRU if gru else torch.nn.LSTM)(c1.shape[-1], 64, batch_first=True, bidirectional=nd == 2)
    sd = {}
    for suf in ([""] if nd == 1 else ["", "_reverse"]):
        for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            sd[f"{n}_l0{suf}"] = o.w[f"encoder.rnn.0.{n}_l0{suf}"]
    mod.load_state_dict(sd)
    pk = torch.nn.utils.rnn.pack_padded_sequence(c1, l1, batch_first=True, enforce_sorted=False)
    with torch.no_grad():
        out, st = mod(pk)
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=c1.shape[1])
    outs, hs = [], []
    for d in range(nd):
        oo, hh, _ = o.run_direction("encoder.rnn.0", "_reverse" if d else "", c1, l1, None, None, d == 1)
        outs.append(oo)
        hs.append(hh)
    assert (torch.cat(outs, -1) - out).abs().max() < 1e-5
    h_n = st if gru else st[0]
    assert (torch.stack(hs, 0) - h_n).abs().max() < 1e-5


def test_deepspeech2_oracle_chunked_equals_offline_forward():
    """Forward (streaming) model: feeding the subsampled sequence in two chunks with carried states == one pass."""
    from oracle.deepspeech2_oracle import DeepSpeech2Conf, DeepSpeech2Oracle
    from ppasr_b200.weights import DeepSpeech2Config, init_deepspeech2_weights
    cfg = DeepSpeech2Config(num_rnn_layers=2, rnn_size=64, vocab_size=30, streaming=True)
    w = init_deepspeech2_weights(cfg)
    o = DeepSpeech2Oracle(DeepSpeech2Conf(**cfg.to_dict()), w)
    x = torch.from_numpy(synthetic_fbank(1, 67 + 64))
    full = o.get_encoder_out(x, torch.tensor([131]), return_logits=True)
    a, _, h, c = o.get_encoder_out_chunk(x[:, :67], torch.tensor([67]), None, None, return_logits=True)
    b, _, _, _ = o.get_encoder_out_chunk(x[:, 64:131], torch.tensor([67]), h, c, return_logits=True)
    assert (torch.cat([a, b], 1) - full).abs().max() < 1e-4


# ------------------------------------------------------------------------------------------------
# Efficient-Conformer oracle self-consistency (pin against the reference code: tests/test_encoder_golden_cpu.py)
# ------------------------------------------------------------------------------------------------
def test_efficient_conformer_oracle_shapes_and_batch_invariance():
    from oracle.efficient_conformer_oracle import EfficientConformerConf, EfficientConformerOracle
    from ppasr_b200.weights import EfficientConformerConfig, init_efficient_conformer_weights
    cfg = EfficientConformerConfig(num_blocks=5, vocab_size=60)
    w = init_efficient_conformer_weights(cfg)
    assert w["encoder.encoders.3.conv_module.depthwise_conv.weight"].shape[-1] == 15
    assert w["encoder.encoders.4.conv_module.depthwise_conv.weight"].shape[-1] == 7      # encoder.py:123-128
    assert w["encoder.encoders.0.self_attn.pos_bias_u"].shape == (4, 192)                # attention.py:35-38
    assert w["encoder.encoders.4.self_attn.pos_bias_u"].shape == (4, 64)
    o = EfficientConformerOracle(EfficientConformerConf(**cfg.to_dict()), w)
    x = torch.from_numpy(synthetic_fbank(3, 207))
    lens = torch.tensor([207, 150, 99])
    y = o.get_encoder_out(x, lens, return_logits=True)
    assert y.shape == (3, 26, 60)  # T' = 51 -> ceil(51 / 2)
    y0 = o.get_encoder_out(x[:1], lens[:1], return_logits=True)
    assert (y0[0] - y[0]).abs().max() < 1e-3 * y.abs().max()


def test_efficient_conformer_grouped_attention_group1_equals_plain():
    """With group_size 1 the grouped attention must reduce to the plain rel-pos attention plus the linear_pos bias."""
    from oracle.efficient_conformer_oracle import EfficientConformerConf, EfficientConformerOracle
    from ppasr_b200.weights import EfficientConformerConfig, init_efficient_conformer_weights
    cfg = EfficientConformerConfig(num_blocks=1, vocab_size=30, group_layer_idx=(), stride_layer_idx=None)
    w = init_efficient_conformer_weights(cfg)
    w["encoder.encoders.0.self_attn.linear_pos.bias"] = np.zeros(256, dtype=np.float32)
    conf = EfficientConformerConf(**cfg.to_dict())
    o = EfficientConformerOracle(conf, w)
    x = torch.randn(2, 20, 256)
    pos = o.position_encoding(0, 20)
    mask = torch.ones(2, 20, 20, dtype=torch.bool)
    a = o.rel_mha("encoder.encoders.0.self_attn", x, mask, pos, None)[0]
    conf.group_size = 1
    b = o.grouped_mha("encoder.encoders.0.self_attn", x, mask, pos)
    assert (a - b).abs().max() < 1e-5


# ------------------------------------------------------------------------------------------------
# n-gram scorer (host side of the GPU scorer) and the oracle beam search with it
# ------------------------------------------------------------------------------------------------
def _toy_lm(V=20, order=4, seed=0):
    from ppasr_b200.decoders.ngram_lm import NGramLM
    rng = np.random.RandomState(seed)
    vocab = ["<blank>", "<unk>"] + [chr(0x4E00 + i) for i in range(V - 3)] + ["<eos>"]
    sents = [[vocab[2 + int(z) % (V - 3)] for z in rng.zipf(1.5, size=rng.randint(3, 12))] for _ in range(300)]
    return vocab, sents, NGramLM.from_counts(sents, order=order)


def test_ngram_lm_backoff_semantics_and_arpa_roundtrip(tmp_path):
    from ppasr_b200.decoders.ngram_lm import NGramLM, OOV_SCORE, LOG10_E, lm_hash64, lm_hash64
    vocab, sents, lm = _toy_lm()
    assert lm.is_character_based() and lm.order == 4
    # P(. | seen context) sums to ~1 (discounted mass redistributed through the back-off weights)
    ctx = tuple(sents[0][:3])
    tot = sum(10 ** lm._score(ctx, w) for w in lm.vocab if w != "<s>")
    assert abs(tot - 1.0) < 0.02
    # unseen context backs off to shorter histories; manual recursion == log_cond_prob
    words = ["<s>", "<s>", vocab[5], vocab[9]]
    manual = lm._score(tuple(words[:-1])[-3:], words[-1]) / LOG10_E
    assert abs(lm.log_cond_prob(words) - manual) < 1e-9
    assert lm.log_cond_prob(["<s>", "zz"]) == OOV_SCORE
    p = str(tmp_path / "t.arpa")
    lm.write_arpa(p)
    lm2 = NGramLM.from_arpa(p)
    assert len(lm2.ngrams) == len(lm.ngrams) and lm2.order == 4
    for s in sents[:5]:
        assert abs(lm2.sent_log_prob(s) - lm.sent_log_prob(s)) < 1e-3
    # device hash table: every n-gram whose tokens are in the vocabulary is retrievable
    keys, vals, in_lm = lm.device_tables(vocab)
    tok = {"<s>": 1, **{w: i + 2 for i, w in enumerate(vocab)}}
    mask = len(keys) - 1
    for ng, (lp, bo) in list(lm.ngrams.items())[:200]:
        if any(w not in tok for w in ng):
            continue
        key = 0
        for w in ng:
            key = (key << 16) | tok[w]
        slot = lm_hash64(key) & mask
        while int(keys[slot]) != key:
            assert int(keys[slot]) != 0
            slot = (slot + 1) & mask
        assert abs(vals[slot][0] - lp / LOG10_E) < 1e-4
    # the home-slot hash mixes every token field: no long probe runs (a multiplicative hash taken from the middle of the
    # product ignored the oldest tokens and produced runs of thousands of slots on a 4-gram table)
    occ = (keys != 0).astype(np.int8)
    edges = np.diff(np.concatenate([[0], occ, [0]]))
    runs = np.flatnonzero(edges == -1) - np.flatnonzero(edges == 1)
    assert occ.mean() <= 0.5 and runs.max() <= 64


def test_ngram_lm_order5_device_table():
    """Order 5 (what the reference's LM recipe trains, docs/beam_search.md): the device table splits a 5-gram key into the four
    most recent tokens (64 bits) + the oldest token; every n-gram is retrievable and 5-grams do not alias their 4-gram suffix."""
    from ppasr_b200.decoders.ngram_lm import NGramLM, LOG10_E, lm_hash64_2
    rng = np.random.RandomState(3)
    vocab = ["<blank>", "<unk>"] + [chr(0x4E00 + i) for i in range(30)] + ["<eos>"]
    sents = [[vocab[2 + int(z) % 30] for z in rng.zipf(1.4, size=rng.randint(4, 14))] for _ in range(300)]
    lm = NGramLM.from_counts(sents, order=5)
    assert lm.order == 5 and any(len(k) == 5 for k in lm.ngrams)
    keys, keys_hi, vals, in_lm = lm.device_tables(vocab, with_hi=True)
    tok = {"<s>": 1, **{w: i + 2 for i, w in enumerate(vocab)}}
    mask = len(keys) - 1
    n5 = 0
    for ng, (lp, bo) in lm.ngrams.items():
        if any(w not in tok for w in ng):
            continue
        ids = [tok[w] for w in ng]
        lo = 0
        for i in ids[-4:]:
            lo = (lo << 16) | i
        hi = ids[0] if len(ids) == 5 else 0
        slot = lm_hash64_2(lo, hi) & mask
        while not (int(keys[slot]) == lo and int(keys_hi[slot]) == hi):
            assert int(keys[slot]) != 0, ng
            slot = (slot + 1) & mask
        assert abs(vals[slot][0] - lp / LOG10_E) < 1e-4 and abs(vals[slot][1] - bo / LOG10_E) < 1e-4
        n5 += len(ids) == 5
    assert n5 > 50


def test_oracle_beam_search_with_scorer():
    from ppasr_b200.decoders.ngram_lm import Scorer
    vocab, sents, lm = _toy_lm()
    rng = np.random.RandomState(1)
    x = rng.randn(25, 20) * 2.0
    p = np.exp(x - x.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    # alpha = beta = 0: the scorer only adds the min_cutoff pruning; with a beam that never fills it is a no-op
    sc0 = Scorer(0.0, 0.0, None, vocab, lm=lm)
    ps = p[:5, :6] / p[:5, :6].sum(1, keepdims=True)
    a0 = DO.ctc_beam_search_decoding(ps, vocab[:6], 10000, 1.0, 40)
    b0 = DO.ctc_beam_search_decoding(ps, vocab[:6], 10000, 1.0, 40, ext_scorer=sc0)
    assert a0[0][1] == b0[0][1] and abs(a0[0][0] - b0[0][0]) < 1e-4
    a = DO.ctc_beam_search_decoding(p, vocab, 20, 0.99, 40)
    # a strong LM pulls the result towards likely character sequences: LM log-prob per char of the output improves
    sc = Scorer(3.0, 0.0, None, vocab, lm=lm)
    c = DO.ctc_beam_search_decoding(p, vocab, 20, 0.99, 40, ext_scorer=sc)
    la = lm.sent_log_prob(list(a[0][1])) / max(1, len(a[0][1]))
    lc = lm.sent_log_prob(list(c[0][1])) / max(1, len(c[0][1]))
    assert lc > la


# ------------------------------------------------------------------------------------------------
# fbank front end oracle == torchaudio's Kaldi port (the reference uses paddleaudio's port of the same code)
# ------------------------------------------------------------------------------------------------
def _test_wave(seconds=1.3, seed=0, amp=0.2):
    rng = np.random.RandomState(seed)
    t = np.arange(int(16000 * seconds)) / 16000.0
    x = 0.1 * rng.randn(t.size) + 0.3 * np.sin(2 * np.pi * 440 * t) * (1 + 0.5 * np.sin(2 * np.pi * 4 * t))
    return (x * amp).astype(np.float32)


def test_fbank_oracle_matches_torchaudio_kaldi():
    import torchaudio
    from oracle import fbank_oracle as FO
    for seed, amp in ((0, 0.2), (1, 0.01), (2, 0.9)):
        x = _test_wave(seed=seed, amp=amp)
        xs = FO.to_int16_scale(FO.db_normalize(x))
        a = FO.kaldi_fbank(xs)
        b = torchaudio.compliance.kaldi.fbank(torch.from_numpy(xs)[None], num_mel_bins=80, frame_length=25, frame_shift=10,
                                              dither=0.0, sample_frequency=16000.0).numpy()
        assert a.shape == b.shape == (1 + (x.size - 400) // 160, 80)
        assert np.abs(a - b).max() < 2e-3
    assert FO.kaldi_fbank(np.zeros(399, dtype=np.float32)).shape == (0, 80)
    # AudioFeaturizer of the product (host path) is the same computation
    from ppasr_b200.predict import AudioFeaturizer
    x = _test_wave(seed=3)
    assert np.abs(AudioFeaturizer().featurize(x) - FO.featurize(x)).max() < 2e-3


def test_squeezeformer_oracle_chunked_equals_chunk_masked_offline():
    """forward_chunk over 67-frame windows == offline forward with decoding_chunk_size=16 (only true if the cache rate rules of
    the time-reduced blocks, the positional offsets and the causal conv caches are wired as in the reference)."""
    cfg, w, o = _squeeze(True, nb=4, reduce_idx=1, recover_idx=3)
    from ppasr_b200.weights import synthetic_fbank
    x = torch.from_numpy(synthetic_fbank(1, 67 + 64 * 2))
    att = torch.zeros(0, 0, 0, 0)
    cnn = torch.zeros(0, 0, 0, 0)
    off, outs = 0, []
    for s in range(0, x.shape[1] - 66, 64):
        y, att, cnn = o.get_encoder_out_chunk(x[:, s:s + 67], off, -1, att, cnn, return_logits=True)
        off += y.shape[1]
        outs.append(y)
    chunked = torch.cat(outs, 1)
    enc, _ = o.encoder_forward(x, torch.tensor([x.shape[1]]), decoding_chunk_size=16, num_decoding_left_chunks=-1)
    full = o.ctc_logits(enc)
    assert (chunked - full[:, :chunked.shape[1]]).abs().max() < 1e-3 * full.abs().max()
    assert att.shape == (4, 4, 48, 128) and cnn.shape == (4, 1, 256, 30)


def test_efficient_conformer_oracle_chunk_path():
    """forward_chunk restatement (efficient_conformer/encoder.py:266-394). Without grouped attention the chunked run equals the
    chunk-masked offline forward (stride block, kernel 15 -> 7, half-rate cache rules and positions are consistent); WITH the
    grouped attention of the shipped config it does not -- the reference re-forms the key groups from the first cached frame
    for every chunk (attention.py:153-160) and strides the chunk mask by 3 offline (:63-64), so its own streaming and offline
    paths disagree. Both facts are pinned here so a GPU implementation knows what it has to reproduce."""
    from oracle.efficient_conformer_oracle import EfficientConformerConf, EfficientConformerOracle
    from ppasr_b200.weights import EfficientConformerConfig, init_efficient_conformer_weights
    x = torch.from_numpy(synthetic_fbank(1, 67 + 64 * 2))
    gaps = {}
    for name, groups in (("plain", ()), ("grouped", (0, 1))):
        cfg = EfficientConformerConfig(num_blocks=3, vocab_size=40, group_layer_idx=groups, stride_layer_idx=1)
        w = init_efficient_conformer_weights(cfg)
        o = EfficientConformerOracle(EfficientConformerConf(**cfg.to_dict()), w)
        att = torch.zeros(0, 0, 0, 0)
        cnn = torch.zeros(0, 0, 0, 0)
        off, outs = 0, []
        for s in range(0, x.shape[1] - 66, 64):
            y, att, cnn = o.get_encoder_out_chunk(x[:, s:s + 67], off, -1, att, cnn, return_logits=True)
            assert y.shape[1] == 8  # 16 full-rate frames -> 8 output frames per chunk
            off += y.shape[1]
            outs.append(y)
        assert att.shape == (3, 4, 48, 128) and cnn.shape == (3, 1, 256, 14)
        chunked = torch.cat(outs, 1)
        enc, _ = o.encoder_forward(x, torch.tensor([x.shape[1]]), decoding_chunk_size=16, num_decoding_left_chunks=-1)
        full = o.ctc_logits(enc)
        gaps[name] = ((chunked - full[:, :chunked.shape[1]]).abs().max() / full.abs().max()).item()
    assert gaps["plain"] < 1e-4
    assert gaps["grouped"] > 1e-3


def test_audio_normalisation_oracle_matches_reference_golden():
    """db_normalize / to_int16_scale of oracle/fbank_oracle.py against outputs of the reference's own AudioSegment.normalize /
    to('int16') (tests/golden/make_audio_golden.py). Bit-exact under NumPy >= 2 (the golden was made with its promotion rules:
    the gain is a float32 quantity); one int16 LSB otherwise."""
    from oracle import fbank_oracle as FO
    g = np.load(os.path.join(GOLD, "audio_golden.npz"))
    exact = int(np.__version__.split(".")[0]) >= 2
    for k in ("quiet", "loud_clipping", "tone", "int16_input", "silence"):
        x = g[k + "/float32"]
        n = FO.db_normalize(x)
        i16 = FO.to_int16_scale(n)
        if exact:
            assert np.array_equal(n, g[k + "/normalized"]), k
            assert np.array_equal(i16, g[k + "/int16"].astype(np.float32)), k
        else:
            assert np.abs(i16 - g[k + "/int16"].astype(np.float32)).max() <= 1, k
    # the product's host featurizer applies the same normalisation (then torchaudio's Kaldi port)
    from ppasr_b200.predict import AudioFeaturizer
    x = g["tone/float32"]
    assert np.abs(AudioFeaturizer().featurize(x) - FO.kaldi_fbank(g["tone/int16"].astype(np.float32))).max() < 2e-3


def test_collapse_frames_matches_reference_golden():
    """ppasr_b200.decoders.ctc_greedy_decoder.collapse_frames (the host half of the ragged-batch decode: per-frame arg-max
    ids / probabilities -> text, score) against the reference decoder's own outputs, bit for bit, incl. a padded row cut by
    n_frames."""
    from ppasr_b200.decoders.ctc_greedy_decoder import collapse_frames
    z, meta = _gold()
    for m in meta:
        if m["name"] == "__batch__":
            continue
        probs = z[m["name"] + "_probs"]
        ids, mx = probs.argmax(-1), probs.max(-1)
        score, text = collapse_frames(ids, mx, gold_vocab(m["V"]))
        assert text == m["text"] and repr(float(score)) == m["score"], m["name"]
        pad_ids = np.concatenate([ids, np.full(5, 3, ids.dtype)])
        pad_mx = np.concatenate([mx, np.full(5, 0.9, mx.dtype)])
        score2, text2 = collapse_frames(pad_ids, pad_mx, gold_vocab(m["V"]), n_frames=len(ids))
        assert (text2, repr(float(score2))) == (m["text"], m["score"])
    assert collapse_frames(np.zeros(4, np.int64), np.ones(4, np.float32), ["<blank>", "a"]) == (0, "")
    assert collapse_frames(np.array([1, 1, 0, 1]), np.ones(4, np.float32), ["<blank>", "a"], n_frames=0) == (0, "")


def test_split_tf32_conv1_arithmetic_is_fp32_accurate():
    """Numerics of the tensor-core conv1 (csrc/conv1_tc.cu): x = x_hi + x_lo, w = w_hi + w_lo with the halves rounded to tf32
    (10 explicit mantissa bits, round to nearest, ties away: cvt.rna.tf32.f32); the GEMM carries x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
    and the bias as b_hi + b_lo. Restated in NumPy: the dropped term and the roundings leave a relative error below 2^-20 of
    sum |x||w| -- two hundred times finer than the bf16 rounding applied to the result (2^-9) -- so the bf16 outputs agree with an
    exact convolution on all but the rare ties. GPU side: tests/test_gpu_parity.py::test_conv1_phase_images_match_fp32_conv2d."""
    def tf32_rna(a):
        a = np.asarray(a, dtype=np.float32)
        bits = a.view(np.uint32).astype(np.uint64)
        bits = (bits + 0x1000) & 0xFFFFE000          # add half an ulp of the 13 dropped bits (ties away), truncate
        return bits.astype(np.uint32).view(np.float32)

    rng = np.random.RandomState(7)
    x = rng.normal(0, 1, (4096, 9)).astype(np.float32)           # normalised fbank windows
    w = (rng.normal(0, 1, (256, 9)) / 3).astype(np.float32)      # conv1 weights
    b = rng.normal(0, 0.1, 256).astype(np.float32)
    xh = tf32_rna(x); xl = tf32_rna(x - xh)
    wh = tf32_rna(w); wl = tf32_rna(w - wh)
    bh = tf32_rna(b); bl = tf32_rna(b - bh)
    assert np.all((xh.view(np.uint32) & 0x1FFF) == 0) and np.all((wl.view(np.uint32) & 0x1FFF) == 0)
    f = np.float64
    got = xh.astype(f) @ wh.astype(f).T + xl.astype(f) @ wh.astype(f).T + xh.astype(f) @ wl.astype(f).T + bh.astype(f) + bl.astype(f)
    ref = x.astype(f) @ w.astype(f).T + b.astype(f)
    scale = np.abs(x).astype(f) @ np.abs(w).astype(f).T + np.abs(b)
    assert np.max(np.abs(got - ref) / scale) < 2.0 ** -20

    def bf16_rn(a):
        bits = np.asarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
        bits = (bits + 0x7FFF + ((bits >> 16) & 1)) & 0xFFFF0000
        return bits.astype(np.uint32).view(np.float32)
    same = bf16_rn(np.maximum(got, 0).astype(np.float32)) == bf16_rn(np.maximum(ref, 0).astype(np.float32))
    assert same.mean() > 0.999
