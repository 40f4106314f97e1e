"""Host-side streaming logic of ppasr_b200.PPASRPredictor against the reference's own PPASRPredictor.

tests/golden/stream_golden.json was recorded by tests/golden/make_stream_golden.py from the UNMODIFIED reference
predict_stream / predict / reset_stream (ppasr/predict.py:163-187,232-347), AudioSegment and AudioFeaturizer.featurize, driven
with tests/golden/stream_fake.py (a deterministic stand-in for the exported-model runner that logs every feature window it gets).
Here the same script drives ppasr_b200.PPASRPredictor with the same stand-in. Must agree call by call: None vs result, the
number and length of the decoding windows, the window contents (per-mel sums, 2e-3 relative: the features come from the same
torchaudio Kaldi fbank in both), the samples left in `remained_wav` INCLUDING their in-place dB scaling, the cached frame count,
the decoded text (exact) and score. The product greedy decoders need the GPU, so the oracle's (pinned bit-for-bit against the
reference's in tests/test_oracle_cpu.py) are patched in.
"""
import json
import os
import sys

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
from stream_fake import VOCAB, FakePredictor, audio_script, long_audio  # noqa: E402

from oracle import decoders_oracle as DO  # noqa: E402
from ppasr_b200 import predict as P  # noqa: E402


@pytest.fixture
def predictor(monkeypatch):
    monkeypatch.setattr(P, "greedy_decoder_chunk", DO.greedy_decoder_chunk)
    monkeypatch.setattr(P, "greedy_decoder", DO.greedy_decoder)
    p = object.__new__(P.PPASRPredictor)  # __init__ builds the CUDA engine; the attributes it sets are assigned by hand
    p.configs = {}
    p.use_model, p.streaming, p.decoder = "conformer", True, "ctc_greedy"
    p._audio_featurizer = P.AudioFeaturizer(feature_method="fbank", n_mels=80, use_dB_normalization=True, target_dB=-20)
    p.vocab_list = VOCAB
    p.running = False
    p.vad_predictor = None
    p.remained_wav = p.cached_feat = p.greedy_last_max_prob_list = p.greedy_last_max_index_list = None
    p.predictor = FakePredictor()
    return p


def _check(p, res, seen, g):
    if g["result"] is None:
        assert res is None
    else:
        assert res is not None and res["text"] == g["result"]["text"]
        assert abs(float(res["score"]) - g["result"]["score"]) < 1e-3
    wins = p.predictor.windows[seen:]
    assert [n for n, _ in wins] == [w["frames"] for w in g["windows"]]
    for (_, s), w in zip(wins, g["windows"]):
        np.testing.assert_allclose(s, np.array(w["mel_sums"]), rtol=2e-3, atol=1e-2)


def test_predict_stream_matches_reference_call_by_call(predictor):
    g = json.load(open(os.path.join(GOLDEN, "stream_golden.json"), encoding="utf-8"))
    assert g["vocab"] == VOCAB
    it = iter(g["stream"])
    n_none = n_res = 0
    for rnd in range(2):
        for kind, payload, is_end in audio_script(seed=11 + rnd):
            rec = next(it)
            assert (rec["round"], rec["kind"], rec["is_end"]) == (rnd, kind, is_end)
            seen = len(predictor.predictor.windows)
            res = predictor.predict_stream(payload, is_end=is_end)
            _check(predictor, res, seen, rec)
            assert len(predictor.remained_wav) == rec["remained_samples"]
            got = float(np.abs(predictor.remained_wav.astype(np.float64)).sum())
            assert abs(got - rec["remained_abs_sum"]) <= 1e-5 * max(1.0, rec["remained_abs_sum"])  # in-place dB scaling carried over
            assert predictor.cached_feat.shape[1] == rec["cached_frames"]
            n_none += res is None
            n_res += res is not None
        predictor.reset_stream()
        assert predictor.remained_wav is None and predictor.cached_feat is None
        assert predictor.greedy_last_max_index_list is None and predictor.greedy_last_max_prob_list is None
    assert predictor.predictor.resets == g["resets"] == 2
    assert n_none >= 4 and n_res >= 4  # both branches exercised


def test_predict_offline_matches_reference(predictor):
    g = json.load(open(os.path.join(GOLDEN, "stream_golden.json"), encoding="utf-8"))
    it = iter(g["offline"])
    n = 0
    for kind, payload, _ in audio_script(seed=21):
        if kind == "bytes" or len(payload) < 1000:
            continue
        rec = next(it)
        assert rec["kind"] == kind
        keep = payload.copy()
        seen = len(predictor.predictor.windows)
        res = predictor.predict(audio_data=payload)
        _check(predictor, res, seen, rec)
        assert np.array_equal(keep, payload)  # the caller's buffer is never scaled in place
        n += 1
    assert n == len(g["offline"]) >= 3


def test_unsupported_inputs_raise_like_reference(predictor):
    with pytest.raises(Exception, match="不支持该数据类型"):
        predictor.predict_stream([0.0] * 1000)
    with pytest.raises(Exception, match="不支持该数据类型"):
        predictor.predict(audio_data=[0.0] * 1000)
    predictor.streaming = False
    with pytest.raises(Exception, match="不支持改该模型流式识别"):
        predictor.predict_stream(np.zeros(1000, np.float32))


def test_predict_long_matches_reference(predictor):
    """predict.py:190-229 with a fixed segmentation in place of the silero VAD: region-by-region recognition, '，' joining,
    mean score rounded to 2 places -- first one predict() per region like the reference, then the ragged-batch route (with
    predict_batch, which needs the GPU, replaced by a per-row loop so that the grouping / padding / ordering logic is what
    is tested)."""
    g = json.load(open(os.path.join(GOLDEN, "stream_golden.json"), encoding="utf-8"))["long"]
    audio, stamps = long_audio()
    assert stamps == g["stamps"]
    with pytest.raises(Exception, match="speech_timestamps"):
        predictor.predict_long(audio)
    res = predictor.predict_long(audio, speech_timestamps=stamps, batched=False)
    assert res["text"] == g["result"]["text"] and abs(res["score"] - g["result"]["score"]) < 1e-6
    wins = predictor.predictor.windows
    assert [n for n, _ in wins] == [w["frames"] for w in g["windows"]]
    for (_, sm), w in zip(wins, g["windows"]):
        np.testing.assert_allclose(sm, np.array(w["mel_sums"]), rtol=2e-3, atol=1e-2)

    # the VAD hook: any object with the reference's get_speech_timestamps(samples, sample_rate)
    class Vad:
        def get_speech_timestamps(self, samples, sr):
            assert samples.dtype == np.float32 and sr == 16000
            return stamps

    calls = []

    def fake_predict_batch(batch, n_samples=None):
        calls.append((batch.shape, [int(v) for v in n_samples]))
        assert batch.dtype == np.float32 and all((batch[r, n:] == 0).all() for r, n in enumerate(n_samples))
        return [predictor.predict(audio_data=batch[r, :n]) for r, n in enumerate(n_samples)]

    predictor.predict_batch = fake_predict_batch
    predictor.vad_predictor = Vad()
    predictor.predictor = FakePredictor()
    res2 = predictor.predict_long(audio, max_batch_samples=2 * 71000)   # forces 3 batches: [0,1] [2,3] [4]
    assert res2 == res
    assert [c[1] for c in calls] == [[28400, 3000], [71000, 2000], [41500]]
    assert [c[0] for c in calls] == [(2, 28400), (2, 71000), (1, 41500)]
    # regions shorter than 7 fbank frames give no text and do not reach the GPU; no region at all -> empty result
    calls.clear()
    r3 = predictor.predict_long(audio, speech_timestamps=[{"start": 0, "end": 300}, {"start": 1600, "end": 30000}])
    assert [c[1] for c in calls] == [[28400]] and r3["text"] != ""
    assert predictor.predict_long(audio, speech_timestamps=[]) == {"text": "", "score": 0}


def test_predict_accepts_wav_files_like_reference_load_audio(predictor, tmp_path):
    """predict.py:142-161 _load_audio: path / open file / bytes of a complete file / ndarray. WAV PCM is decoded with the
    standard library; the result must equal feeding the same int16 samples as an array (soundfile.read(dtype='float32')
    scales int16 by 1/32768, exactly what AudioSegment does with an int16 array)."""
    import wave
    x = next(p for k, p, _ in audio_script(seed=21) if k == "int16" and len(p) > 1000)
    path = str(tmp_path / "a.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(1), w.setsampwidth(2), w.setframerate(16000)
        w.writeframes(x.tobytes())

    def run(arg):
        predictor.predictor = FakePredictor()
        return predictor.predict(audio_data=arg)

    ref = run(x)
    assert run(path) == ref
    assert run(open(path, "rb").read()) == ref
    with open(path, "rb") as f:
        assert run(f) == ref
    def run_long(arg):
        predictor.predictor = FakePredictor()   # the stand-in's output depends on its call count
        return predictor.predict_long(arg, speech_timestamps=[{"start": 0, "end": len(x)}], batched=False)

    assert run_long(path) == run_long(x) and run_long(x)["text"] == ref["text"]
    # stereo: channels are averaged (audio.py:31-32)
    st = str(tmp_path / "st.wav")
    with wave.open(st, "wb") as w:
        w.setnchannels(2), w.setsampwidth(2), w.setframerate(16000)
        w.writeframes(np.stack([x, x], 1).tobytes())
    assert run(st) == ref
    with pytest.raises(Exception, match="RIFF/WAVE"):
        run(b"ID3\x03 not a wav file at all")
    with pytest.raises(AssertionError, match="文件不存在"):
        run(str(tmp_path / "missing.wav"))
    with wave.open(st, "wb") as w:
        w.setnchannels(1), w.setsampwidth(2), w.setframerate(8000)
        w.writeframes(x.tobytes())
    with pytest.raises(Exception, match="resampling"):
        run(st)


def test_stream_scheduler_feed_audio_matches_reference_predict_stream(monkeypatch):
    """StreamScheduler.feed_audio (many sessions on one engine) must cut each session's audio into exactly the feature windows
    the reference's single-stream predict_stream produced (stream_golden.json: per-window frame count and per-mel sums, incl.
    the in-place dB normalisation carried between calls) -- here with the two recorded rounds running as two INTERLEAVED
    sessions, and the transcripts must equal the reference's final results. The engine is a stand-in built on the same
    FakePredictor (its output depends on how many windows it has seen, so each session gets its own)."""
    from ppasr_b200.infer_utils import stream_scheduler as SS
    g = json.load(open(os.path.join(GOLDEN, "stream_golden.json"), encoding="utf-8"))
    fakes = {}

    class Engine:
        def sessions_init(self, n):
            pass

        def sessions_reset(self, slot):
            fakes[slot] = FakePredictor()

        def sessions_step(self, batch, slots, required):
            assert required == -16 and len(set(slots)) == len(slots)
            self.out = [fakes[s].predict_chunk_conformer(batch[b:b + 1], required)[0] for b, s in enumerate(slots)]

        def ctc_probs(self, to_host=True):
            return np.stack(self.out)

    class Pred:
        use_model, streaming = "conformer", True
        configs = {"preprocess_conf": {"feature_method": "fbank", "n_mels": 80, "use_dB_normalization": True, "target_dB": -20}}

        class model_config:
            input_dim = 80

        engine = Engine()

    monkeypatch.setattr(SS, "greedy_decoder_chunk", DO.greedy_decoder_chunk)
    sch = SS.StreamScheduler(Pred(), VOCAB, max_sessions=4, required_cache_size=-16)
    scripts = {rnd: audio_script(seed=11 + rnd) for rnd in range(2)}
    sids = {rnd: sch.open() for rnd in range(2)}
    # the reference run used ONE stand-in for both rounds: round 1 started after the windows of round 0
    n0 = sum(len(r["windows"]) for r in g["stream"] if r["round"] == 0)
    fakes[sch._sessions[sids[1]].slot].windows = [(0, None)] * n0
    last = {}
    for i in range(len(scripts[0])):
        for rnd in range(2):
            kind, payload, is_end = scripts[rnd][i]
            sch.feed_audio(sids[rnd], payload, is_end=is_end)
        while sch.pending():
            last.update(sch.step())
    for rnd in range(2):
        recs = [r for r in g["stream"] if r["round"] == rnd]
        ref_wins = [w for r in recs for w in r["windows"]]
        slot = sch._sessions[sids[rnd]].slot
        wins = fakes[slot].windows[n0 if rnd == 1 else 0:]
        assert [n for n, _ in wins] == [w["frames"] for w in ref_wins]
        for (_, sm), w in zip(wins, ref_wins):
            np.testing.assert_allclose(sm, np.array(w["mel_sums"]), rtol=2e-3, atol=1e-2)
        final = recs[-1]["result"]
        assert last[sids[rnd]]["text"] == final["text"] and abs(last[sids[rnd]]["score"] - final["score"]) < 1e-3
        assert len(sch._sessions[sids[rnd]].wav) == recs[-1]["remained_samples"]
        assert sch.close(sids[rnd])["text"] == final["text"]
    with pytest.raises(Exception, match="不支持该数据类型"):
        sch.feed_audio(sch.open(), [0.0] * 100)
