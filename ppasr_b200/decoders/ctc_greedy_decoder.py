"""Drop-in for ppasr/decoders/ctc_greedy_decoder.py: same names, arguments and return values, with the
frame arg-max + blank-collapse + score running in the sm_100a kernels behind
ppasr_b200_greedy_decode (include/ppasr_b200.h). Detokenisation (ids -> string) stays on the host,
exactly as ctc_greedy_decoder.py:27-31. There is no NumPy fallback: without the CUDA library these
functions raise.
"""
import numpy as np

from .. import _lib as L


def _as_cuda_probs(probs):
    import torch
    if isinstance(probs, torch.Tensor):
        t = probs
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(probs), dtype=np.float32))
    if not torch.cuda.is_available():
        raise L.PPASRB200Error("ppasr_b200 decoders need a CUDA device (no CPU fallback)")
    return t.to(device="cuda", dtype=torch.float32).contiguous()


def greedy_decode_ids(probs, frame_lens=None, blank_index=0):
    """probs: [B,T,V] or [T,V]. Returns (ids [B,T] int32, out_lens [B], mean_prob [B] fp32, frame_ids [B,T],
    frame_probs [B,T]) as NumPy arrays."""
    import torch
    t = _as_cuda_probs(probs)
    if t.dim() == 2:
        t = t.unsqueeze(0)
    B, T, V = t.shape
    dev = t.device
    ids = torch.empty((B, T), dtype=torch.int32, device=dev)
    ol = torch.empty((B,), dtype=torch.int32, device=dev)
    sc = torch.empty((B,), dtype=torch.float32, device=dev)
    ti = torch.empty((B * T,), dtype=torch.int32, device=dev)
    tm = torch.empty((B * T,), dtype=torch.float32, device=dev)
    fl = None
    if frame_lens is not None:
        fl = torch.as_tensor(np.asarray(frame_lens), dtype=torch.int32).to(dev)
    lib = L.load()
    L.check(lib.ppasr_b200_greedy_decode(L.ptr(t), B, T, V, L.ptr(fl), int(blank_index), L.ptr(ids), T, L.ptr(ol),
                                         L.ptr(sc), L.ptr(ti), L.ptr(tm), L.stream_ptr()))
    return (ids.cpu().numpy(), ol.cpu().numpy(), sc.cpu().numpy(), ti.view(B, T).cpu().numpy(),
            tm.view(B, T).cpu().numpy())


def _detok(ids, vocabulary):
    return "".join([vocabulary[i] for i in ids]).replace("<space>", " ")


def collapse_frames(frame_ids, frame_probs, vocabulary, n_frames=None, blank_index=0):
    """Host half of ctc_greedy_decoder.py:18-31 on per-frame arg-max ids / probabilities that the GPU already produced
    (ppasr_b200_ctc_greedy `frame_ids` / `frame_probs`): merge repeats, drop blanks, join; score = 100 x mean probability of the
    non-blank frames, summed in frame order in float32 like Python's sum() over np.float32. `n_frames` cuts a padded row."""
    from itertools import groupby
    ids = [int(i) for i in frame_ids[:n_frames]]
    kept = [np.float32(frame_probs[t]) for t, i in enumerate(ids) if i != blank_index]
    index_list = [i for i in (g[0] for g in groupby(ids)) if i != blank_index]
    text = "".join([vocabulary[i] for i in index_list])
    score = float(sum(kept) / len(kept)) * 100.0 if len(kept) > 0 else 0
    return score, text.replace("<space>", " ")


def greedy_decoder(probs_seq, vocabulary, blank_index=0):
    """ctc_greedy_decoder.py:6-31 -> (score, text)."""
    ids, ol, sc, _, _ = greedy_decode_ids(probs_seq, None, blank_index)
    n = int(ol[0])
    score = float(sc[0]) * 100.0 if n > 0 else 0
    return score, _detok(ids[0, :n].tolist(), vocabulary)


def greedy_decoder_batch(probs_split, vocabulary, blank_index=0):
    """ctc_greedy_decoder.py:34-49 -> [text]. Accepts a list of [T_i,V] arrays or one [B,T,V] tensor."""
    import torch
    if isinstance(probs_split, (list, tuple)):
        lens = [int(np.asarray(p.shape)[0]) for p in probs_split]
        T = max(lens)
        V = int(probs_split[0].shape[1])
        batch = torch.zeros((len(probs_split), T, V), dtype=torch.float32)
        for i, p in enumerate(probs_split):
            batch[i, :lens[i]] = torch.as_tensor(np.asarray(p) if not isinstance(p, torch.Tensor) else p.cpu())
        ids, ol, _, _, _ = greedy_decode_ids(batch, lens, blank_index)
    else:
        ids, ol, _, _, _ = greedy_decode_ids(probs_split, None, blank_index)
    return [_detok(ids[b, :int(ol[b])].tolist(), vocabulary) for b in range(ids.shape[0])]


def greedy_decoder_chunk(probs_seq, vocabulary, last_max_prob_list=None, last_max_index_list=None, blank_index=0):
    """ctc_greedy_decoder.py:52-89 -> (score, text, last_max_prob_list, last_max_index_list).
    Keeps the reference's (swapped) naming: `last_max_prob_list` holds ids, `last_max_index_list` probabilities."""
    from itertools import groupby
    if last_max_prob_list is None:
        last_max_prob_list = []
    if last_max_index_list is None:
        last_max_index_list = []
    _, _, _, fi, fp = greedy_decode_ids(probs_seq, None, blank_index)
    max_index_list = [np.int64(i) for i in fi[0]]
    max_prob_list = [np.float32(fp[0, i]) for i in range(len(max_index_list)) if max_index_list[i] != blank_index]
    last_max_prob_list.extend(max_index_list)
    last_max_index_list.extend(max_prob_list)
    index_list = [g[0] for g in groupby(last_max_prob_list)]
    index_list = [i for i in index_list if i != blank_index]
    text = "".join([vocabulary[i] for i in index_list])
    score = 0
    if len(last_max_index_list) > 0:
        score = float(sum(last_max_index_list) / len(last_max_index_list)) * 100.0
    return score, text.replace("<space>", " "), last_max_prob_list, last_max_index_list
