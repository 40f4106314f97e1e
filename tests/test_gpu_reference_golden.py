"""The CUDA hot path against outputs of the reference's OWN model code -- no oracle in between (pytest -m gpu).

tests/golden/encoder_golden_*_gpu_*.npz hold what the unmodified reference encoders (ConformerEncoder, SqueezeformerEncoder,
EfficientConformerEncoder, CRNNEncoder + CTC head; imported from the reference tree on the torch-backed paddle stand-in by
tests/golden/make_encoder_golden.py) produce for seeded weights at the engine's widths (d_model 256, 4 heads, FFN 2048 / LSTM
256). Here the same weights (regenerated from the seed) and inputs go through the C-ABI engine. Tolerance as everywhere for the
bf16 tensor-core path: max |logit diff| over valid frames < 1e-2 of max |logit| (north_star: 1e-2 rel for bf16); caches 2e-2.
"""
import ast
import glob
import os

import numpy as np
import pytest
import torch

# Hard assertions: seen green on a B200 at the end of round 1 (GPUTEST_r01: all XPASS), the non-strict xfail marks are gone.
pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FILES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "encoder_golden_*_gpu_*.npz")))
FAMILIES = ("efficient_conformer", "squeezeformer", "conformer", "deepspeech2")


def _family(fname):
    stem = fname[len("encoder_golden_"):]
    return next(f for f in FAMILIES if stem.startswith(f))


def _config(g, family):
    from ppasr_b200 import weights as W
    cls, init = {"conformer": (W.ConformerConfig, W.init_conformer_weights),
                 "squeezeformer": (W.SqueezeformerConfig, W.init_squeezeformer_weights),
                 "efficient_conformer": (W.EfficientConformerConfig, W.init_efficient_conformer_weights),
                 "deepspeech2": (W.DeepSpeech2Config, W.init_deepspeech2_weights)}[family]
    cfg = cls(**ast.literal_eval(str(g["cfg"])))
    return cfg, init(cfg, seed=int(g["seed"]))


def _valid_frames(g, family, cfg):
    n = ((g["lens"] - 1) // 2 - 1) // 2
    if family == "efficient_conformer" and cfg.stride_layer_idx is not None:
        n = (n + 1) // 2  # efficient_conformer/encoder.py:255-260
    if family == "deepspeech2":
        assert n.tolist() == g["offline_lens"].tolist()
    return [int(v) for v in n]


def _windows(num_frames, window=67, stride=64, context=7):
    """predict.py:281-297 with is_end on the last call (the windows the fixtures were recorded with)."""
    return [(cur, min(cur + window, num_frames)) for cur in range(0, num_frames - context + 1, stride)]


def test_fixtures_present():
    assert len(FILES) == 10 and {_family(f) for f in FILES} == set(FAMILIES)


@pytest.mark.parametrize("fname", FILES)
def test_offline_logits_match_reference_code(lib, cuda, fname):
    from ppasr_b200.engine import ConformerEngine
    g = np.load(os.path.join(GOLDEN, fname))
    family = _family(fname)
    cfg, w = _config(g, family)
    eng = ConformerEngine(cfg, w)
    eng.encode(torch.from_numpy(g["feats"]).to(cuda), [int(v) for v in g["lens"]])
    logits = eng.ctc_logits().float().cpu().numpy()
    probs = eng.ctc_probs().float().cpu().numpy()
    eng.close()
    ref, ref_p = g["offline_logits"], g["offline_probs"]
    assert logits.shape == ref.shape
    scale = float(np.abs(ref).max())
    vl = _valid_frames(g, family, cfg)
    assert max(vl) == ref.shape[1]
    worst = max(float(np.abs(logits[b, :n] - ref[b, :n]).max()) for b, n in enumerate(vl) if n > 0) / scale
    assert worst < 1e-2, f"{fname}: logits rel err {worst}"
    # arg-max path: every frame whose reference top-2 margin exceeds 3 x the MEASURED max |logit error| must agree (an arg-max
    # can only flip below twice the error); the unfiltered agreement rate is reported and bounded
    top2 = np.sort(ref, -1)[..., -2:]
    big = (top2[..., 1] - top2[..., 0]) > 3.0 * worst * scale
    n_checked = n_frames = n_agree = 0
    for b, n in enumerate(vl):
        same = logits[b, :n].argmax(-1) == ref[b, :n].argmax(-1)
        assert bool((same | ~big[b, :n]).all())
        n_checked += int(big[b, :n].sum())
        n_frames += int(n)
        n_agree += int(same.sum())
    print(f"[{fname}] logits rel err {worst:.3e}; unfiltered arg-max agreement {n_agree}/{n_frames}")
    assert n_agree >= 0.95 * n_frames
    for b, n in enumerate(vl):
        assert np.allclose(probs[b, :n].sum(-1), 1.0, atol=1e-4)
        # probabilities: only where the reference is decided (a 1e-2-of-scale logit error moves an undecided softmax a lot)
        sure = ref_p[b, :n].max(-1) > 0.99
        assert float(np.abs(probs[b, :n][sure] - ref_p[b, :n][sure]).max()) < 5e-2 if sure.any() else True
    assert n_checked > 0


@pytest.mark.parametrize("fname", [f for f in FILES if "stream" in f])
def test_chunk_chain_matches_reference_code(lib, cuda, fname):
    from ppasr_b200.infer_utils.inference_predictor import InferencePredictor
    g = np.load(os.path.join(GOLDEN, fname))
    family = _family(fname)
    cfg, w = _config(g, family)
    pred = InferencePredictor({"encoder_conf": cfg.to_dict(), "preprocess_conf": {"n_mels": 80}}, family, streaming=True,
                              weights=w)
    cf = g["chunk_feats"]
    ref = g["chunk_logits"]
    scale = float(np.abs(ref).max())
    outs = []
    if family == "deepspeech2":
        assert cf.ndim == 3
        for (a, b) in _windows(cf.shape[1]):
            probs, lens = pred.predict_chunk_deepspeech(np.ascontiguousarray(cf[:, a:b]))
            lg = pred.engine.ctc_logits().float().cpu().numpy()
            assert lg.shape[0] == cf.shape[0] and lens.tolist() == [lg.shape[1]] * cf.shape[0]
            outs.append(lg)
        got = np.concatenate(outs, 1)
        assert np.abs(pred.output_state_h - g["chunk_state_h"]).max() < 2e-2
        if "chunk_state_c" in g.files:
            assert np.abs(pred.output_state_c - g["chunk_state_c"]).max() < 5e-2
    else:
        assert cf.ndim == 2
        off = 0
        for (a, b) in _windows(cf.shape[0]):
            probs = pred.predict_chunk_conformer(np.ascontiguousarray(cf[None, a:b]), -16)
            lg = pred.engine.ctc_logits().float().cpu().numpy()
            assert lg.shape == probs.shape and lg.shape[0] == 1
            off += lg.shape[1]
            assert int(pred.offset[0]) == off
            outs.append(lg[0])
        got = np.concatenate(outs, 0)
        if family == "conformer":  # exported in the reference layout (inference_predictor.py:204-207)
            att, cnn = pred.att_cache, pred.cnn_cache
            ra, rc = g["chunk_att_cache"], g["chunk_cnn_cache"]
            assert att.shape == ra.shape and cnn.shape == rc.shape
            assert float(np.abs(att - ra).max()) / float(np.abs(ra).max()) < 2e-2
            assert float(np.abs(cnn - rc).max()) / float(np.abs(rc).max()) < 2e-2
    assert got.shape == ref.shape
    worst = float(np.abs(got - ref).max()) / scale
    assert worst < 1e-2, f"{fname}: chunk logits rel err {worst}"
    pred.reset_stream()


def test_model_level_api_matches_reference_code(lib, cuda):
    """ppasr_b200.model_utils.ConformerModel driven exactly like the reference model (set_state_dict, get_encoder_out,
    functional get_encoder_out_chunk loop with the caches handed back) against the reference code's own probabilities."""
    from ppasr_b200.model_utils import ConformerModel
    g = np.load(os.path.join(GOLDEN, "encoder_golden_conformer_gpu_stream.npz"))
    cfg, w = _config(g, "conformer")
    m = ConformerModel(input_dim=80, vocab_size=cfg.vocab_size, streaming=True,
                       encoder_conf={k: v for k, v in cfg.to_dict().items() if k in (
                           "output_size", "attention_heads", "linear_units", "num_blocks", "cnn_module_kernel", "cnn_module_norm",
                           "max_len")})
    m.set_state_dict(w)
    probs = m.eval().get_encoder_out(g["feats"], g["lens"]).numpy()
    ref = g["offline_probs"]
    assert probs.shape == ref.shape
    for b, n in enumerate(_valid_frames(g, "conformer", cfg)):
        sure = ref[b, :n].max(-1) > 0.99
        assert sure.any() and (probs[b, :n][sure].argmax(-1) == ref[b, :n][sure].argmax(-1)).all()
        assert float(np.abs(probs[b, :n][sure] - ref[b, :n][sure]).max()) < 5e-2
    cf = g["chunk_feats"]
    ref_l = g["chunk_logits"]
    ref_p = np.exp(ref_l - ref_l.max(-1, keepdims=True))
    ref_p /= ref_p.sum(-1, keepdims=True)
    att = np.zeros([0, 0, 0, 0], np.float32)
    cnn = np.zeros([0, 0, 0, 0], np.float32)
    offset, outs = 0, []
    for (a, b) in _windows(cf.shape[0]):
        pr, att, cnn = m.get_encoder_out_chunk(cf[None, a:b], np.array([offset], np.int32), np.array([-16], np.int32), att, cnn)
        offset += pr.shape[1]
        outs.append(pr.numpy()[0])
    got = np.concatenate(outs, 0)
    assert got.shape == ref_p.shape and att.shape == g["chunk_att_cache"].shape and cnn.shape == g["chunk_cnn_cache"].shape
    sure = ref_p.max(-1) > 0.99
    assert sure.any() and (got[sure].argmax(-1) == ref_p[sure].argmax(-1)).all()
    assert float(np.abs(got[sure] - ref_p[sure]).max()) < 5e-2
    m.close()
