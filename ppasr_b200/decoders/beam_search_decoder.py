"""Drop-in for ppasr/decoders/beam_search_decoder.py::BeamSearchDecoder (and the swig_wrapper functions)
running the CTC prefix beam search on the B200 (ppasr_b200/csrc/beam.cu) instead of the
`paddlespeech_ctcdecoders` C++ library.

Same constructor arguments and methods (beam_search_decoder.py:9-96). Differences:
  * `language_model_path=None` (default here) decodes without an external scorer -- which the upstream C++
    supports via ext_scoring_func=None (swig_wrapper.py:35-41). With a path, the file must be an ARPA text n-gram model
    (order <= 5) that is character based (every LM word is one vocabulary token), as the reference's Chinese setups
    are; it is turned into a device hash table and scored inside the beam kernel (ppasr_b200/decoders/ngram_lm.py).
    KenLM *binary* files (.klm / trie) and word-based LMs (English, needs the dictionary FST) raise.
  * beam_size <= 512 and cutoff_top_n <= 64 in this build (the shipped configs use 300 / 40); larger values raise.
  * `num_processes` is accepted and ignored (utterances are decoded by one CTA each, all in parallel).
Returned scores follow the upstream convention: -log P(prefix) of the total CTC probability
(decoder_utils.cpp get_beam_search_result, `-approx_ctc`), best first.
"""
import ctypes

import numpy as np

from .. import _lib as L


class UnsupportedDecoderConfig(Exception):
    """The requested beam-search configuration is outside what the GPU decoder of this build implements."""


class BeamSearchDecoder:
    def __init__(self, alpha=2.2, beta=4.3, beam_size=300, cutoff_prob=0.99, cutoff_top_n=40, vocab_list=None,
                 num_processes=10, blank_id=0, language_model_path=None, max_frames=5000, ext_scorer=None):
        if beam_size > 512:
            raise UnsupportedDecoderConfig(f"beam_size {beam_size} > 512 is not supported by the GPU decoder in this build")
        if cutoff_top_n > 64:
            raise UnsupportedDecoderConfig(f"cutoff_top_n {cutoff_top_n} > 64 is not supported by the GPU decoder in this build")
        # ---- configuration checks first (they do not need a device) ----
        self.alpha = alpha
        self.beta = beta
        self.beam_size = int(beam_size)
        self.cutoff_prob = float(cutoff_prob)
        self.cutoff_top_n = int(cutoff_top_n)
        self.vocab_list = vocab_list
        self.num_processes = num_processes
        self.blank_id = int(blank_id)
        self.max_frames = int(max_frames)
        self._ext_scorer = ext_scorer
        if ext_scorer is not None:  # a ready-made scorer carries its own weights (swig_wrapper.py:35-41 ext_scoring_func)
            self.alpha, self.beta = ext_scorer.alpha, ext_scorer.beta
        if language_model_path is not None and ext_scorer is None:
            import os
            if not os.path.isfile(str(language_model_path)):
                # the reference downloads a missing LM (beam_search_decoder.py:19-25); there is no network here
                raise UnsupportedDecoderConfig("language model file not found: " + str(language_model_path))
            with open(language_model_path, "rb") as f:
                head = f.read(64)
            if not head.lstrip().startswith(b"\\data\\"):
                raise UnsupportedDecoderConfig("language model must be an ARPA text file (KenLM binary formats are not supported): "
                                + str(language_model_path))
        import torch
        if not torch.cuda.is_available():
            raise L.PPASRB200Error("ppasr_b200 decoders need a CUDA device (no CPU fallback)")
        self.torch = torch
        self.lib = L.load()
        if language_model_path is not None and ext_scorer is None:
            from .ngram_lm import Scorer
            self._ext_scorer = Scorer(alpha, beta, language_model_path, vocab_list)
        self._lm_dev = None
        if self._ext_scorer is not None:
            sc = self._ext_scorer
            if not sc.is_character_based():
                raise UnsupportedDecoderConfig("only character-based language models are supported by the GPU scorer")
            if sc.get_max_order() > 5:
                raise UnsupportedDecoderConfig("n-gram order > 5 is not supported by the GPU scorer")
            keys, keys_hi, vals, in_lm = sc.lm.device_tables(vocab_list, with_hi=True)
            self._lm_dev = (torch.from_numpy(keys.view(np.int64)).cuda(), torch.from_numpy(vals).cuda(),
                            torch.from_numpy(in_lm).cuda(), int(keys.shape[0]),
                            torch.from_numpy(keys_hi.view(np.int32)).cuda())
        self._stream_state = None  # persistent state of the streaming decoder (batch 1)
        self.reset_decoder()

    # ---- device helpers -------------------------------------------------------------------------
    def _alloc_state(self, B, max_frames):
        torch = self.torch
        n = self.lib.ppasr_b200_beam_state_bytes(B, max_frames, self.beam_size)
        st = torch.empty(n, dtype=torch.uint8, device="cuda")
        L.check(self.lib.ppasr_b200_beam_reset(L.ptr(st), B, max_frames, self.beam_size, L.stream_ptr()))
        return st

    def _advance(self, state, probs, frame_lens, max_frames):
        torch = self.torch
        B, T, V = probs.shape
        ws = torch.empty(self.lib.ppasr_b200_beam_workspace_bytes(B, T), dtype=torch.uint8, device="cuda")
        fl = None
        if frame_lens is not None:
            fl = torch.as_tensor(np.asarray(frame_lens), dtype=torch.int32).cuda()
        if self._lm_dev is not None:
            keys, vals, in_lm, cap, keys_hi = self._lm_dev
            sc = self._ext_scorer
            sc.reset_params(self.alpha, self.beta)  # beam_search_decoder.py:46-47,60-61: alpha / beta may be retuned between calls
            L.check(self.lib.ppasr_b200_beam_advance_lm(L.ptr(probs), B, T, V, L.ptr(fl), self.beam_size,
                                                        ctypes.c_float(self.cutoff_prob), self.cutoff_top_n, self.blank_id,
                                                        L.ptr(state), max_frames, L.ptr(ws), L.ptr(keys), L.ptr(keys_hi),
                                                        L.ptr(vals), L.ptr(in_lm), cap, sc.get_max_order(), ctypes.c_float(sc.alpha),
                                                        ctypes.c_float(sc.beta), L.stream_ptr()))
            return
        L.check(self.lib.ppasr_b200_beam_advance(L.ptr(probs), B, T, V, L.ptr(fl), self.beam_size,
                                                 ctypes.c_float(self.cutoff_prob), self.cutoff_top_n, self.blank_id,
                                                 L.ptr(state), max_frames, L.ptr(ws), L.stream_ptr()))

    def _results_device(self, state, B, max_frames, lmax, nb):
        """First `nb` beam entries as device tensors: ids int32 [B, nb, lmax], lens int32 [B, nb] (-1 none, -2 overflow),
        scores fp32 [B, nb] = log P(prefix), best first."""
        torch = self.torch
        ids = torch.zeros((B, nb, lmax), dtype=torch.int32, device="cuda")
        lens = torch.zeros((B, nb), dtype=torch.int32, device="cuda")
        sc = torch.zeros((B, nb), dtype=torch.float32, device="cuda")
        L.check(self.lib.ppasr_b200_beam_result_nbest(L.ptr(state), B, max_frames, self.beam_size, nb, L.ptr(ids), lmax,
                                                      L.ptr(lens), L.ptr(sc), L.stream_ptr()))
        return ids, lens, sc

    def _results(self, state, B, max_frames, lmax, nbest=None, approx=True):
        nb = self.beam_size if nbest is None else max(1, min(int(nbest), self.beam_size))
        lmax = max(1, int(lmax))
        ids, lens, sc = self._results_device(state, B, max_frames, lmax, nb)
        ids, lens, sc = ids.cpu().numpy(), lens.cpu().numpy(), sc.cpu().numpy()
        if (lens == -2).any():
            raise L.PPASRB200Error("beam-search state overflow: more frames were fed than the max_frames the decoder "
                                   "was created with (call reset_decoder(), or construct it with a larger max_frames)")
        out = []
        for b in range(B):
            res = []
            for k in range(nb):
                if lens[b, k] < 0:
                    continue
                toks = [self.vocab_list[i] for i in ids[b, k, :lens[b, k]]]
                score = float(sc[b, k])
                if self._ext_scorer is not None and approx:
                    # approx_ctc: take the word-insertion and LM terms out again (ctc_beam_search_decoder.cpp, end of decoding)
                    es = self._ext_scorer
                    score = score - len(toks) * es.beta - es.get_sent_log_prob(toks) * es.alpha
                res.append((-score, "".join(toks)))
            out.append(res)
        return out

    def decode_device(self, probs, frame_lens=None, nbest=1):
        """probs: CUDA fp32 [B,T,V] -> device tensors (ids [B,nbest,T], lens [B,nbest], log-prob scores [B,nbest]) without any
        host synchronisation (batch pipelines: bench.py, predict_batch)."""
        B, T, V = probs.shape
        st = self._alloc_state(B, T)
        self._advance(st, probs, frame_lens, T)
        return self._results_device(st, B, T, T, max(1, min(int(nbest), self.beam_size)))

    def _to_cuda(self, probs):
        torch = self.torch
        t = probs if isinstance(probs, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(probs, dtype=np.float32))
        return t.to(device="cuda", dtype=torch.float32).contiguous()

    def decode_ids_batch(self, probs, frame_lens=None, nbest=None, approx=True):
        """probs [B,T,V] -> per utterance list of (score, text), best first (all beam entries, or the first `nbest`).
        approx=False skips the host-side approx_ctc rescoring of the returned scores (texts and their order are unaffected)."""
        p = self._to_cuda(probs)
        B, T, V = p.shape
        st = self._alloc_state(B, T)
        self._advance(st, p, frame_lens, T)
        return self._results(st, B, T, T, nbest, approx)

    # ---- reference API --------------------------------------------------------------------------
    def decode_beam_search_offline(self, probs_split):
        """beam_search_decoder.py:45-56 -> (score, text) of the best hypothesis."""
        p = self._to_cuda(probs_split)
        return self.decode_ids_batch(p.unsqueeze(0), nbest=1)[0][0]

    def decode_batch_beam_search_offline(self, probs_split):
        """beam_search_decoder.py:59-73 -> [text]; probs_split: list of [T_i,V] arrays or a [B,T,V] tensor."""
        torch = self.torch
        if isinstance(probs_split, (list, tuple)):
            lens = [int(p.shape[0]) for p in probs_split]
            V = int(probs_split[0].shape[1])
            batch = torch.zeros((len(lens), max(lens), V), dtype=torch.float32)
            for i, p in enumerate(probs_split):
                batch[i, :lens[i]] = torch.as_tensor(np.asarray(p) if not isinstance(p, torch.Tensor) else p.cpu())
            res = self.decode_ids_batch(batch, lens, nbest=1, approx=False)  # texts only
        else:
            res = self.decode_ids_batch(probs_split, nbest=1, approx=False)
        return [r[0][1] for r in res]

    def decode_chunk(self, probs, logits_lens):
        """beam_search_decoder.py:75-91: feed one chunk [1,t,V] to the persistent decoder, return the current best."""
        p = self._to_cuda(probs)
        if p.dim() == 2:
            p = p.unsqueeze(0)
        lens = np.asarray(logits_lens).astype(np.int32)
        n_new = int(min(int(lens[0]), p.shape[1]))
        if self._stream_frames + n_new > self.max_frames:
            raise L.PPASRB200Error(f"decode_chunk: {self._stream_frames + n_new} frames exceed max_frames={self.max_frames} "
                                   "of this decoder (reset_decoder() between utterances, or construct with a larger max_frames)")
        self._advance(self._stream_state, p[:1], lens[:1], self.max_frames)
        self._stream_frames += n_new
        # a prefix is never longer than the number of frames consumed: copy back only that much
        return self._results(self._stream_state, 1, self.max_frames, self._stream_frames, nbest=1)[0][0]

    def reset_decoder(self):
        """beam_search_decoder.py:93-96."""
        self._stream_state = self._alloc_state(1, self.max_frames)
        self._stream_frames = 0


def ctc_beam_search_decoding(probs_seq, vocabulary, beam_size, cutoff_prob=1.0, cutoff_top_n=40, blank_id=0,
                             ext_scoring_func=None):
    """swig_wrapper.py:35-64 -> [(score, text)] best first."""
    d = BeamSearchDecoder(beam_size=beam_size, cutoff_prob=cutoff_prob, cutoff_top_n=cutoff_top_n, vocab_list=vocabulary,
                          blank_id=blank_id, max_frames=8, ext_scorer=ext_scoring_func)
    return d.decode_ids_batch(d._to_cuda(probs_seq).unsqueeze(0))[0]


def ctc_beam_search_decoding_batch(probs_split, vocabulary, beam_size, num_processes, cutoff_prob=1.0, cutoff_top_n=40,
                                   blank_id=0, ext_scoring_func=None):
    """swig_wrapper.py:67-103 -> [[(score, text)]]."""
    d = BeamSearchDecoder(beam_size=beam_size, cutoff_prob=cutoff_prob, cutoff_top_n=cutoff_top_n, vocab_list=vocabulary,
                          blank_id=blank_id, max_frames=8, ext_scorer=ext_scoring_func)
    import torch
    lens = [int(p.shape[0]) for p in probs_split]
    batch = torch.zeros((len(lens), max(lens), int(probs_split[0].shape[1])), dtype=torch.float32)
    for i, p in enumerate(probs_split):
        batch[i, :lens[i]] = torch.as_tensor(np.asarray(p))
    return d.decode_ids_batch(batch, lens)
