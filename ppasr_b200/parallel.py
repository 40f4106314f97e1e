"""Multi-GPU data parallelism for the hot path: utterances are independent, so the batch is cut into
contiguous per-rank blocks (the reference's own DP slicing, ppasr/data_utils/sampler.py:89-102), every
rank runs the same single-GPU engine on its block, and ONE all-gather of fixed-shape records
(ids int32 [B_local, Lmax] | out_len | score bits) over NCCL / NVLink returns the decoded ids to every
rank. Strings are detokenised on the host from the ids. gloo is used for the CPU tests.
"""
import numpy as np


def shard_range(num_items: int, world_size: int, rank: int):
    """Contiguous block [start, end) of rank `rank`; the first (num_items % world_size) ranks get one extra."""
    base, rem = divmod(num_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_records(ids, out_lens, scores, max_local: int, lmax: int):
    """[max_local, lmax + 2] int32 records: ids row (zero padded), out_len, score (fp32 bit pattern)."""
    import torch
    rec = torch.zeros((max_local, lmax + 2), dtype=torch.int32, device=ids.device)
    n = ids.shape[0]
    rec[:n, :ids.shape[1]] = ids[:, :lmax]
    rec[:n, lmax] = out_lens
    rec[:n, lmax + 1] = scores.view(torch.int32)
    return rec


def all_gather_results(ids, out_lens, scores, num_items: int, lmax: int, group=None):
    """Single collective: every rank contributes its block and receives all `num_items` results.
    ids: int32 [B_local, >=lmax] device tensor; out_lens int32 [B_local]; scores fp32 [B_local].
    Returns (ids [num_items, lmax], out_lens [num_items], scores [num_items]) on the calling device."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ids[:, :lmax], out_lens, scores
    max_local = (num_items + world - 1) // world
    rec = pack_records(ids, out_lens, scores, max_local, lmax)
    out = torch.empty((world * max_local, lmax + 2), dtype=torch.int32, device=rec.device)
    dist.all_gather_into_tensor(out, rec, group=group)
    rows = []
    for r in range(world):
        s, e = shard_range(num_items, world, r)
        rows.append(out[r * max_local: r * max_local + (e - s)])
    allr = torch.cat(rows, 0)
    return allr[:, :lmax].contiguous(), allr[:, lmax].contiguous(), allr[:, lmax + 1].contiguous().view(torch.float32)


def detokenize(ids, out_lens, vocabulary):
    """Host-side string assembly (ppasr/decoders/ctc_greedy_decoder.py:27-31)."""
    ids = np.asarray(ids)
    out_lens = np.asarray(out_lens)
    key = id(vocabulary)
    tab = _VOCAB_ARRAYS.get(key)
    if tab is None or len(tab) != len(vocabulary) or tab[0] != vocabulary[0] or tab[-1] != vocabulary[-1]:
        tab = np.asarray(vocabulary, dtype=object)  # one object-array gather per utterance instead of a Python loop per token
        _VOCAB_ARRAYS.clear()
        _VOCAB_ARRAYS[key] = tab
    return ["".join(tab[ids[b, :int(out_lens[b])]].tolist()).replace("<space>", " ") for b in range(ids.shape[0])]


_VOCAB_ARRAYS = {}
