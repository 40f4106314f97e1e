"""Property tests (hypothesis) of the host-side pieces the GPU path relies on: sharding, the streaming window arithmetic,
the ragged-batch decode helper, the output-frame rule. Small example counts: the whole file runs in a few seconds."""
from itertools import groupby

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import decoders_oracle as DO
from oracle.conformer_oracle import stream_windows
from ppasr_b200.decoders.ctc_greedy_decoder import collapse_frames
from ppasr_b200.parallel import shard_range

FAST = settings(max_examples=60, deadline=None, derandomize=True)  # same examples on every run


@FAST
@given(st.integers(0, 5000), st.integers(1, 16))
def test_shard_range_partitions_the_batch(n, world):
    """parallel.shard_range: contiguous, disjoint, covering, sizes differ by at most one (sampler.py:89-102 slicing)."""
    blocks = [shard_range(n, world, r) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    sizes = [e - s for s, e in blocks]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@FAST
@given(st.integers(7, 3000))
def test_stream_windows_cover_every_output_frame_once(T):
    """predict.py:277-300: the 67/64 windows of a finished stream produce exactly the output frames of the offline
    subsampling, each once: sum over windows of out_frames(len) == out_frames(T) whenever the tail is long enough to be
    consumed, and never more."""
    def out(n):
        return max(((n - 1) // 2 - 1) // 2, 0)
    wins = stream_windows(T, is_end=True)
    assert wins and wins[0][0] == 0 and all(e - s <= 67 for s, e in wins)
    assert all(b[0] - a[0] == 64 for a, b in zip(wins, wins[1:]))
    total = sum(out(e - s) for s, e in wins)
    assert total <= out(T)
    # frames of a tail shorter than the 7-frame context are dropped by the reference too; otherwise nothing is lost
    last_end = wins[-1][1]
    assert last_end == T or T - (wins[-1][0] + 64) < 7
    if last_end == T:
        assert total == out(T)
    mid = stream_windows(T, is_end=False)
    assert mid == [w for w in wins if w[1] - w[0] == 67][:len(mid)]


@FAST
@given(st.lists(st.integers(0, 6), min_size=0, max_size=60), st.integers(0, 70), st.randoms(use_true_random=False))
def test_collapse_frames_is_the_reference_rule(ids, cut, rnd):
    """collapse_frames == groupby-collapse + blank removal + float32 running mean, for any id sequence and any cut."""
    probs = np.array([rnd.random() for _ in ids], dtype=np.float32)
    vocab = ["<blank>", "a", "b", "<space>", "d", "e", "f"]
    n = min(cut, len(ids))
    score, text = collapse_frames(np.array(ids, dtype=np.int64), probs, vocab, n_frames=n)
    kept_ids = [k for k, _ in groupby(ids[:n]) if k != 0]
    assert text == "".join(vocab[i] for i in kept_ids).replace("<space>", " ")
    nb = [np.float32(p) for i, p in zip(ids[:n], probs[:n]) if i != 0]
    assert score == (float(sum(nb) / len(nb)) * 100.0 if nb else 0)
    # and it is what the reference decoder returns for a one-hot-ish posterior with these arg-maxes
    if n > 0:
        post = np.full((n, len(vocab)), 1e-4, dtype=np.float32)
        for t in range(n):
            post[t, ids[t]] = max(probs[t], np.float32(1e-3))
        s2, t2 = DO.greedy_decoder(post, vocab)
        assert t2 == text


@FAST
@given(st.integers(1, 4000))
def test_ragged_length_trick_ends_the_mask_at_the_own_frame_count(frames):
    """predict_batch (ragged): a length of 4 T' - 3 makes the batch mask rule `4 j < len` (subsampling.py:115) keep exactly
    T' = ((frames - 1) // 2 - 1) // 2 frames."""
    t = max(((frames - 1) // 2 - 1) // 2, 0)
    length = 4 * t - 3 if t > 0 else 0
    kept = sum(1 for j in range(t + 3) if 4 * j < length)
    assert kept == t and (length + 3) // 4 == t


def test_hypothesis_is_bounded():
    assert FAST.max_examples <= 100


@settings(max_examples=25, deadline=None, derandomize=True)
@given(st.integers(0, 40000))
def test_fbank_frame_count_of_the_library_matches_kaldi(n):
    """ppasr_b200_fbank_frames (pure host arithmetic inside the C-ABI library) == the frame count of Kaldi's snip-edges
    framing (25 ms window, 10 ms shift: 1 + (n - 400) // 160, nothing below one window) == what torchaudio's Kaldi fbank (the
    CPU twin of the reference's paddleaudio call, audio_featurizer.py:120-138) actually returns."""
    import torch
    import torchaudio
    from ppasr_b200 import _lib as L
    got = int(L.load().ppasr_b200_fbank_frames(int(n)))
    assert got == (1 + (n - 400) // 160 if n >= 400 else 0)
    if n >= 400:
        m = torchaudio.compliance.kaldi.fbank(torch.zeros(1, n) + 1.0, num_mel_bins=80, frame_length=25, frame_shift=10,
                                              dither=0.0, sample_frequency=16000.0)
        assert m.shape == (got, 80)
