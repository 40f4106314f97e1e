"""Generate tests/golden/stream_golden.json by running the REFERENCE's own PPASRPredictor.predict_stream / predict
(/root/reference/ppasr/predict.py:163-187,232-347), AudioSegment, AudioFeaturizer.featurize and greedy_decoder_chunk, unmodified.

What is replaced, and why: `paddle` -> tests/golden/paddle_shim (PaddlePaddle is not installable here);
`paddleaudio.compliance.kaldi.fbank` -> torchaudio.compliance.kaldi.fbank (both are ports of Kaldi's compute-fbank-feats; the
argument names are mapped); the exported-model runner `self.predictor` -> tests/golden/stream_fake.FakePredictor (no weights
exist); av / zhconv / resampy / soundfile / termcolor -> empty modules (file decoding, resampling, coloured logs: not on this
path). PPASRPredictor.__init__ (model download, paddle.inference) is bypassed with object.__new__ and the attributes it would
set are assigned by hand. Everything else that runs -- sample conversion, in-place dB normalisation of the remaining audio,
frame bookkeeping, the 67/64/3 window logic, None returns, greedy chunk decoding, reset -- is the reference's code.

Run in the build container only:  python tests/golden/make_stream_golden.py
"""
import json
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "paddle_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

if not hasattr(np, "sctypes"):  # removed in NumPy 2; audio.py:542,567
    np.sctypes = {"float": [np.float16, np.float32, np.float64], "int": [np.int8, np.int16, np.int32, np.int64]}
for name in ("av", "zhconv", "resampy", "soundfile", "termcolor", "paddleaudio", "paddleaudio.compliance",
             "paddleaudio.compliance.kaldi", "paddle.inference"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["zhconv"].convert = lambda s, *_: s
sys.modules["termcolor"].colored = lambda s, *a, **k: s

import paddle  # noqa: E402  (shim)
import torch  # noqa: E402
import torchaudio  # noqa: E402

paddle.inference = sys.modules["paddle.inference"]


def _fbank(waveform, n_mels, frame_length, frame_shift, dither, sr):
    m = torchaudio.compliance.kaldi.fbank(waveform.as_subclass(torch.Tensor), num_mel_bins=n_mels, frame_length=frame_length,
                                          frame_shift=frame_shift, dither=dither, sample_frequency=float(sr))
    return paddle.Tensor(m)


sys.modules["paddleaudio.compliance.kaldi"].fbank = _fbank
sys.modules["paddleaudio.compliance.kaldi"].mfcc = None

from ppasr.predict import PPASRPredictor  # noqa: E402  (the reference)
from ppasr.data_utils.featurizer.audio_featurizer import AudioFeaturizer  # noqa: E402
from ppasr.utils.utils import dict_to_object  # noqa: E402

from stream_fake import VOCAB, FakePredictor, audio_script  # noqa: E402


def make_reference_predictor():
    p = object.__new__(PPASRPredictor)
    p.configs = dict_to_object({"use_model": "conformer", "streaming": True, "decoder": "ctc_greedy",
                                "preprocess_conf": {"feature_method": "fbank", "n_mels": 80, "use_dB_normalization": True,
                                                    "target_dB": -20, "sample_rate": 16000}})
    p.running = False
    p.inv_normalizer = None
    p.pun_predictor = None
    p.vad_predictor = None
    p._text_featurizer = types.SimpleNamespace(vocab_list=VOCAB, vocab_size=len(VOCAB))
    p._audio_featurizer = AudioFeaturizer(feature_method="fbank", n_mels=80, use_dB_normalization=True, target_dB=-20)
    p.remained_wav = None
    p.cached_feat = None
    p.greedy_last_max_prob_list = None
    p.greedy_last_max_index_list = None
    p.predictor = FakePredictor()
    return p


def snapshot(p, res, seen):
    wins = p.predictor.windows[seen:]
    return {
        "result": None if res is None else {"text": res["text"], "score": float(res["score"])},
        "windows": [{"frames": n, "mel_sums": [float(v) for v in s]} for n, s in wins],
        "remained_samples": None if p.remained_wav is None else int(len(p.remained_wav.samples)),
        "remained_abs_sum": None if p.remained_wav is None else float(np.abs(p.remained_wav.samples.astype(np.float64)).sum()),
        "cached_frames": None if p.cached_feat is None else int(p.cached_feat.shape[1]),
    }


def main():
    out = {"vocab": VOCAB, "stream": [], "offline": []}
    p = make_reference_predictor()
    for rnd in range(2):  # second round after reset_stream(): state must be clean
        for kind, payload, is_end in audio_script(seed=11 + rnd):
            seen = len(p.predictor.windows)
            res = p.predict_stream(payload, is_end=is_end)
            out["stream"].append(dict(round=rnd, kind=kind, is_end=is_end, **snapshot(p, res, seen)))
        p.reset_stream()
        assert p.remained_wav is None and p.cached_feat is None
    out["resets"] = p.predictor.resets
    # offline predict() on ndarray inputs (predict.py:163-187), fresh predictor (the fake's output depends on its call count)
    p = make_reference_predictor()
    for kind, payload, _ in audio_script(seed=21):
        if kind == "bytes" or len(payload) < 1000:
            continue
        seen = len(p.predictor.windows)
        res = p.predict(audio_data=payload)
        out["offline"].append(dict(kind=kind, **snapshot(p, res, seen)))
    # predict_long (predict.py:190-229) with a fixed segmentation instead of the silero VAD model
    from stream_fake import long_audio
    audio, stamps = long_audio()
    p = make_reference_predictor()
    p.vad_predictor = types.SimpleNamespace(get_speech_timestamps=lambda samples, sr: stamps)
    res = p.predict_long(audio_data=audio)
    out["long"] = {"result": {"text": res["text"], "score": float(res["score"])}, "stamps": stamps,
                   "windows": [{"frames": n, "mel_sums": [float(v) for v in sm]} for n, sm in p.predictor.windows]}
    with open(os.path.join(HERE, "stream_golden.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False)
    n_none = sum(1 for s in out["stream"] if s["result"] is None)
    print("stream calls", len(out["stream"]), "None results", n_none, "windows", sum(len(s["windows"]) for s in out["stream"]),
          "offline", len(out["offline"]))
    for s in out["stream"][:9]:
        print(s["kind"], s["is_end"], s["result"], [w["frames"] for w in s["windows"]], s["remained_samples"], s["cached_frames"])


if __name__ == "__main__":
    main()
