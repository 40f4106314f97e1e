"""CPU oracle for the fbank front end (TEST INFRASTRUCTURE ONLY -- see oracle/conformer_oracle.py header).

NumPy restatement of AudioFeaturizer.featurize for feature_method='fbank'
(ppasr/data_utils/featurizer/audio_featurizer.py:37-69,120-138; paths relative to /root/reference):
  1. AudioSegment.normalize(target_db=-20): gain = target_db - rms_db, samples *= 10^(gain/20)   (data_utils/audio.py:287-304)
  2. AudioSegment.to('int16'): samples * 2^15, clipped, truncated to int16                        (audio.py:549-574)
  3. paddleaudio.compliance.kaldi.fbank(waveform, n_mels, frame_length=25, frame_shift=10, dither=0, sr) -- third-party
     (paddleaudio >= 1.0.1, requirements.txt:14; not installable here), a port of Kaldi's compute-fbank-feats with the
     defaults: snip_edges, remove_dc_offset, preemphasis 0.97, povey window, FFT size 512, power spectrum, mel filters from
     20 Hz to Nyquist on the scale 1127 ln(1 + f/700), log(max(e, FLT_EPSILON)).
PARITY: steps 1-2 are PINNED bit-for-bit against the reference's own AudioSegment code (tests/golden/audio_golden.npz, made by
tests/golden/make_audio_golden.py); step 3 is pinned against torchaudio.compliance.kaldi.fbank (the same Kaldi port,
importable here) in tests/test_oracle_cpu.py.
"""
import numpy as np

EPS = np.float32(1.1920928955078125e-07)


def db_normalize(samples, target_db=-20, max_gain_db=300.0):
    """AudioSegment.normalize / rms_db / gain_db (audio.py:287-304,519-529,256-264), expression by expression: the mean square
    and the gain are float32 quantities (NumPy >= 2 promotion rules), an all-zero segment has rms_db = 0."""
    samples = np.asarray(samples, dtype=np.float32).copy()
    mean_square = np.mean(samples ** 2)
    if mean_square == 0:
        mean_square = 1
    rms_db = 10 * np.log10(mean_square)
    gain = target_db - rms_db
    if gain > max_gain_db:
        raise ValueError("gain exceeds max_gain_db")
    samples *= 10. ** (min(max_gain_db, target_db - rms_db) / 20.)
    return samples


def to_int16_scale(samples):
    """AudioSegment.to('int16') (audio.py:244-254,549-574): scale by 2^15, clip to the int16 range, truncate; returned as
    float32 (what the fbank routine is fed with, audio_featurizer.py:128)."""
    out = np.asarray(samples, dtype=np.float32).copy()
    out *= (2 ** 15 / 1.)
    out[out > 32767] = 32767
    out[out < -32768] = -32768
    return out.astype(np.int16).astype(np.float32)


def mel_banks(n_mels, nfft=512, sr=16000, low=20.0):
    def mel(f):
        return np.float32(1127.0) * np.log(np.float32(1.0) + np.asarray(f, dtype=np.float32) / np.float32(700.0))
    nyq = 0.5 * sr
    bw = np.float32(sr / nfft)
    mlo, mhi = mel(low), mel(nyq)
    delta = (mhi - mlo) / np.float32(n_mels + 1)
    b = np.arange(n_mels, dtype=np.float32)[:, None]
    left, center, right = mlo + b * delta, mlo + (b + 1) * delta, mlo + (b + 2) * delta
    mk = mel(bw * np.arange(nfft // 2, dtype=np.float32))[None, :]
    up = (mk - left) / (center - left)
    down = (right - mk) / (right - center)
    w = np.maximum(0.0, np.minimum(up, down)).astype(np.float32)
    return np.pad(w, ((0, 0), (0, 1)))  # zero column for the Nyquist bin


def kaldi_fbank(wave, n_mels=80, sr=16000):
    """wave: float32 on the int16 scale, [N]. Returns [T, n_mels] float32."""
    wave = np.asarray(wave, dtype=np.float32)
    win, shift, nfft = 400, 160, 512
    if wave.shape[0] < win:
        return np.zeros((0, n_mels), dtype=np.float32)
    T = 1 + (wave.shape[0] - win) // shift
    idx = np.arange(win)[None, :] + shift * np.arange(T)[:, None]
    fr = wave[idx]
    fr = fr - fr.mean(axis=1, keepdims=True)
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)
    fr = fr - np.float32(0.97) * prev
    window = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(win) / (win - 1))) ** 0.85
    fr = fr * window.astype(np.float32)
    fr = np.pad(fr, ((0, 0), (0, nfft - win)))
    power = (np.abs(np.fft.rfft(fr.astype(np.float32), axis=1)) ** 2).astype(np.float32)
    e = power @ mel_banks(n_mels, nfft, sr).T
    return np.log(np.maximum(e, EPS)).astype(np.float32)


def featurize(samples, n_mels=80, use_db_normalization=True, target_db=-20.0):
    """audio_featurizer.py:37-69 for float samples in [-1, 1] at 16 kHz."""
    if use_db_normalization:
        samples = db_normalize(samples, target_db)
    return kaldi_fbank(to_int16_scale(samples), n_mels)
