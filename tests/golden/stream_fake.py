"""A deterministic stand-in for InferencePredictor used by make_stream_golden.py (driving the REFERENCE PPASRPredictor) and by
tests/test_stream_golden_cpu.py (driving ppasr_b200.PPASRPredictor): CTC 'probabilities' that are a fixed function of the
feature window, plus a log of every window it was handed. No model, no GPU -- it isolates the host-side streaming logic."""
import numpy as np

VOCAB = ["<blank>", "<unk>"] + [chr(0x4E00 + i) for i in range(27)] + ["<eos>"]


class FakePredictor:
    def __init__(self, n_mels=80, vocab=len(VOCAB), seed=5):
        rng = np.random.RandomState(seed)
        self.proj = rng.randn(n_mels, vocab).astype(np.float64)
        self.windows = []      # (frames, per-mel sums) of every x_chunk seen
        self.resets = 0

    def _probs(self, x):
        x = np.asarray(x, dtype=np.float64)[0]
        n = (x.shape[0] - 7) // 4 + 1
        fr = np.stack([x[4 * i:4 * i + 7].mean(0) for i in range(n)], 0)
        fr = np.round(fr, 2)   # coarse, so that 1e-4-level feature noise cannot flip an argmax
        lg = (fr - fr.mean(1, keepdims=True)) @ self.proj * 0.05
        lg += 3.0 * np.random.RandomState(1000 + len(self.windows)).randn(n, lg.shape[1])   # varies with the window count
        lg[:, 0] += 2.0 * np.sin(np.arange(n) * 1.3)   # blanks come and go
        p = np.exp(lg - lg.max(1, keepdims=True))
        return (p / p.sum(1, keepdims=True)).astype(np.float32)[None]

    def predict_chunk_conformer(self, x_chunk, required_cache_size=-1):
        assert required_cache_size == -16
        self.windows.append((int(x_chunk.shape[1]), np.asarray(x_chunk, np.float64)[0].sum(0)))
        return self._probs(x_chunk)

    def predict(self, speech, speech_lens):
        self.windows.append((int(speech.shape[1]), np.asarray(speech, np.float64)[0].sum(0)))
        return self._probs(speech)

    def reset_stream(self):
        self.resets += 1


def audio_script(seed=11):
    """The sequence of predict_stream calls: (kind, payload, is_end). Chunk sizes straddle the 400-sample fbank window, the
    67-frame decoding window and the 64-frame stride; level changes between chunks make the per-call dB normalisation matter."""
    rng = np.random.RandomState(seed)
    calls = []
    sizes = [8000, 300, 12345, 160, 33000, 161, 16000, 4801, 20000]  # every call leaves >= 400 samples (kaldi fbank asserts)
    for i, n in enumerate(sizes):
        level = 0.02 * (1 + (i % 3) * 4)
        t = np.arange(n) / 16000.0
        x = (level * (np.sin(2 * np.pi * (200 + 90 * i) * t) + 0.5 * rng.randn(n))).astype(np.float32)
        if i % 3 == 1:
            calls.append(("int16", np.clip(x * 32768, -32768, 32767).astype(np.int16), False))
        elif i % 3 == 2:
            calls.append(("bytes", np.clip(x * 32768, -32768, 32767).astype(np.int16).tobytes(), False))
        else:
            calls.append(("float32", x, False))
    k, p, _ = calls[-1]
    calls[-1] = (k, p, True)
    return calls


def long_audio(seed=31):
    """A 12 s recording with level changes and a fixed 'VAD' segmentation (samples) for predict_long."""
    rng = np.random.RandomState(seed)
    n = 12 * 16000
    t = np.arange(n) / 16000.0
    env = 0.02 + 0.2 * (np.sin(2 * np.pi * 0.4 * t) > 0)
    x = (env * (np.sin(2 * np.pi * 310 * t) + 0.7 * rng.randn(n))).astype(np.float32)
    stamps = [{"start": 1600, "end": 30000}, {"start": 41000, "end": 44000}, {"start": 60000, "end": 131000},
              {"start": 140000, "end": 142000}, {"start": 150000, "end": 191500}]
    return x, stamps
