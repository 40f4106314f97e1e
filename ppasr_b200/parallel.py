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


def pack_records(ids, out_lens, scores, max_local: int, lmax: int, out=None):
    """[max_local, lmax + 2] int32 records: ids row (zero padded), out_len, score (fp32 bit pattern). `out` (optional): a
    zero-initialised record tensor of that shape to fill in place (the serving loop keeps one per pipeline slot)."""
    import torch
    rec = out if out is not None else torch.zeros((max_local, lmax + 2), dtype=torch.int32, device=ids.device)
    n = ids.shape[0]
    rec[:n, :min(ids.shape[1], lmax)] = ids[:, :lmax]
    rec[:n, lmax] = out_lens
    rec[:n, lmax + 1] = scores.view(torch.int32)
    return rec


def all_gather_records(ids, out_lens, scores, num_items: int, lmax: int, group=None, rec=None, out=None):
    """The collective itself: every rank contributes [max_local, lmax + 2] int32 records (pack_records) and receives the
    [world * max_local, lmax + 2] concatenation on its device -- pack (3 small copies) + ONE all-gather, nothing else on the
    stream. `rec` / `out` are optional preallocated buffers. Split it on the host with `unpack_records`."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    max_local = (num_items + world - 1) // world
    rec = pack_records(ids, out_lens, scores, max_local, lmax, out=rec)
    if world == 1:
        return rec
    if out is None:
        out = torch.empty((world * max_local, lmax + 2), dtype=torch.int32, device=rec.device)
    dist.all_gather_into_tensor(out, rec, group=group)
    return out


def unpack_records(records, num_items: int, world: int, lmax: int):
    """(ids [num_items, lmax], out_lens [num_items], scores fp32 [num_items]) views / selections of a gathered record array
    (NumPy array or torch tensor); drops the padding rows of the ranks that hold one item less."""
    max_local = (num_items + world - 1) // world
    if num_items != world * max_local:
        keep = []
        for r in range(world):
            s, e = shard_range(num_items, world, r)
            keep.extend(range(r * max_local, r * max_local + (e - s)))
        records = records[keep]
    sc = records[:, lmax + 1]
    sc = sc.view(np.float32) if isinstance(sc, np.ndarray) else sc.contiguous().view(_torch_float32())
    return records[:, :lmax], records[:, lmax], sc


def _torch_float32():
    import torch
    return torch.float32


def all_gather_results(ids, out_lens, scores, num_items: int, lmax: int, group=None):
    """Single collective: every rank contributes its block and receives all `num_items` results.
    ids: int32 [B_local, >=lmax] device tensor; out_lens int32 [B_local]; scores fp32 [B_local].
    Returns (ids [num_items, lmax], out_lens [num_items], scores [num_items]) on the calling device."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ids[:, :lmax], out_lens, scores
    out = all_gather_records(ids, out_lens, scores, num_items, lmax, group=group)
    gi, gl, gs = unpack_records(out, num_items, world, lmax)
    return gi.contiguous(), gl.contiguous(), gs


def detokenize(ids, out_lens, vocabulary):
    """Host-side string assembly (ppasr/decoders/ctc_greedy_decoder.py:27-31): "".join(vocabulary[i]) with "<space>" -> " ".

    Vectorised over the whole batch: tokens that are a single character (every entry of a Chinese character vocabulary, and
    "<space>" after its replacement) are gathered as UTF-32 code points in ONE NumPy take over all utterances and decoded in
    one call; the per-utterance strings are slices of that. An utterance containing any other multi-character token
    (<blank>, <unk>, <eos>, word pieces) takes the reference's join."""
    ids = np.asarray(ids)
    out_lens = np.asarray(out_lens).astype(np.int64)
    key = id(vocabulary)
    ent = _VOCAB_ARRAYS.get(key)
    if ent is None or len(ent[0]) != len(vocabulary) or ent[0][0] != vocabulary[0] or ent[0][-1] != vocabulary[-1]:
        tab = np.asarray(vocabulary, dtype=object)
        cp = np.zeros(len(vocabulary), dtype=np.uint32)   # code point of single-character tokens, 0 = needs the slow path
        for i, tok in enumerate(vocabulary):
            t = " " if tok == "<space>" else tok
            if len(t) == 1:
                cp[i] = ord(t)
        _VOCAB_ARRAYS.clear()
        ent = (tab, cp)
        _VOCAB_ARRAYS[key] = ent
    tab, cp = ent
    B = ids.shape[0]
    if B == 0:
        return []
    L = ids.shape[1]
    lens = np.minimum(out_lens, L)
    mask = np.arange(L)[None, :] < lens[:, None]
    flat = ids[mask]                                   # all tokens of all utterances, utterance-major
    pts = cp[flat]
    text = pts.astype("<u4").tobytes().decode("utf-32-le") if pts.size else ""
    ends = np.cumsum(lens)
    slow = np.zeros(B, dtype=bool)
    if pts.size and not pts.all():
        bad = np.flatnonzero(pts == 0)
        slow[np.searchsorted(ends, bad, side="right")] = True
    out = []
    start = 0
    for b in range(B):
        e = int(ends[b])
        if slow[b]:
            out.append("".join(tab[ids[b, :int(lens[b])]].tolist()).replace("<space>", " "))
        else:
            out.append(text[start:e])
        start = e
    return out


_VOCAB_ARRAYS = {}
