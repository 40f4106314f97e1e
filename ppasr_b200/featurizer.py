"""GPU fbank front end: batched drop-in for AudioFeaturizer.featurize (feature_method='fbank') feeding the encoder without a
host round trip (ppasr/data_utils/featurizer/audio_featurizer.py:37-69,120-138; kernel: csrc/fbank.cu)."""
import ctypes

import numpy as np

from . import _lib as L


class GpuFbank:
    def __init__(self, n_mels=80, sample_rate=16000, use_dB_normalization=True, target_dB=-20, device=0):
        import torch
        if not torch.cuda.is_available():
            raise L.PPASRB200Error("ppasr_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if sample_rate != 16000:
            raise Exception("the GPU fbank front end handles 16 kHz audio (resampling is outside the hot path)")
        self.torch = torch
        self.lib = L.load()
        self.n_mels = int(n_mels)
        self.sample_rate = int(sample_rate)
        self.db = bool(use_dB_normalization)
        self.target_db = float(target_dB)
        self.device = torch.device("cuda", device)

    def num_frames(self, n_samples):
        return int(self.lib.ppasr_b200_fbank_frames(int(n_samples)))

    def featurize_batch(self, audio, n_samples=None, stream=None):
        """audio: float32 [B, N] in [-1, 1] (NumPy / torch, host or CUDA); n_samples: optional per-utterance valid samples.
        Returns a CUDA tensor [B, T, n_mels] (T = frames of N) -- rows of shorter utterances are zero beyond their own frames --
        and the per-utterance frame counts (list)."""
        torch = self.torch
        a = audio if isinstance(audio, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
        a = a.to(self.device, dtype=torch.float32).contiguous()
        B, N = a.shape
        T = self.num_frames(N)
        out = torch.empty((B, T, self.n_mels), dtype=torch.float32, device=self.device)
        ns = None
        counts = [T] * B
        if n_samples is not None:
            ns_np = np.minimum(np.asarray(n_samples, dtype=np.int32), N)
            ns = torch.from_numpy(ns_np).to(self.device)
            counts = [self.num_frames(int(v)) for v in ns_np]
        gain = torch.empty(B, dtype=torch.float32, device=self.device) if self.db else None
        L.check(self.lib.ppasr_b200_fbank(L.ptr(a), B, N, N, L.ptr(ns), self.n_mels, self.sample_rate, int(self.db),
                                          ctypes.c_float(self.target_db), L.ptr(gain), L.ptr(out), T, L.stream_ptr(stream)))
        return out, counts

    def featurize(self, samples):
        """Single utterance, like AudioFeaturizer.featurize: float32 [N] -> NumPy [T, n_mels]."""
        out, _ = self.featurize_batch(np.asarray(samples, dtype=np.float32)[None])
        return out[0].cpu().numpy()
