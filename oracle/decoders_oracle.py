"""CPU oracle for the CTC decoders (TEST INFRASTRUCTURE ONLY -- see oracle/conformer_oracle.py header).

greedy_*: restatement of ppasr/decoders/ctc_greedy_decoder.py:6-89 in NumPy. PARITY PINNED: the golden
vectors in tests/golden/greedy_*.npz were produced by importing the reference's own file
(tests/golden/make_greedy_golden.py) and this restatement is checked against them bit-for-bit
(ids, text and score).

ctc_beam_search_*: restatement of the algorithm behind ppasr/decoders/swig_wrapper.py:35-64 /
beam_search_decoder.py:45-73. That code lives in the third-party C++ package
`paddlespeech_ctcdecoders` (imported at swig_wrapper.py:1; NOT vendored under /root/reference and
NOT version-pinned anywhere: absent from requirements.txt / setup.py, installed `-U` from the
author's index per ppasr/predict.py:101-102). PARITY UNPINNED: the published algorithm
(PaddleSpeech third_party/ctc_decoders `ctc_beam_search_decoder.cpp`, `decoder_utils.cpp`,
`path_trie.cpp`; itself derived from Baidu DeepSpeech2's decoder) is restated from its documented
behaviour, anchored on the reference call sites (argument order/meaning, blank_id=0, probabilities
not log-probabilities as input, returned list of (score, text) sorted best-first). Checked by
properties instead: beam >= #prefixes on tiny problems == exact sum over all CTC alignments; beam=1
on peaked posteriors == greedy text.
"""
import math
from itertools import groupby

import numpy as np

NUM_FLT_INF = float("inf")
NUM_FLT_MIN = float(np.finfo(np.float32).tiny)  # std::numeric_limits<float>::min()


# ------------------------------------------------------------------------------------------------
# greedy  (ppasr/decoders/ctc_greedy_decoder.py)
# ------------------------------------------------------------------------------------------------
def greedy_ids(probs_seq, blank_index=0):
    """ctc_greedy_decoder.py:21-26 -> (argmax ids per frame, collapsed ids, non-blank max-prob list)."""
    probs_seq = np.asarray(probs_seq)
    max_index = probs_seq.argmax(axis=1)  # first maximum wins
    max_prob = [probs_seq[i][max_index[i]] for i in range(len(max_index)) if max_index[i] != blank_index]
    collapsed = [g[0] for g in groupby(list(max_index))]
    collapsed = [int(i) for i in collapsed if i != blank_index]
    return max_index, collapsed, max_prob


def greedy_decoder(probs_seq, vocabulary, blank_index=0):
    """ctc_greedy_decoder.py:6-31."""
    _, index_list, max_prob_list = greedy_ids(probs_seq, blank_index)
    text = "".join([vocabulary[i] for i in index_list])
    score = 0
    if len(max_prob_list) > 0:
        score = float(sum(max_prob_list) / len(max_prob_list)) * 100.0
    return score, text.replace("<space>", " ")


def greedy_decoder_batch(probs_split, vocabulary, blank_index=0):
    """ctc_greedy_decoder.py:34-49."""
    return [greedy_decoder(p, vocabulary, blank_index)[1] for p in probs_split]


def greedy_decoder_chunk(probs_seq, vocabulary, last_max_prob_list=None, last_max_index_list=None, blank_index=0):
    """ctc_greedy_decoder.py:52-89. NOTE the reference's swapped names: `last_max_prob_list`
    accumulates the argmax *ids*, `last_max_index_list` accumulates the *probabilities*."""
    if last_max_prob_list is None:
        last_max_prob_list = []
    if last_max_index_list is None:
        last_max_index_list = []
    probs_seq = np.asarray(probs_seq)
    max_index_list = list(probs_seq.argmax(axis=1))
    max_prob_list = [probs_seq[i][max_index_list[i]] for i in range(len(max_index_list))
                     if max_index_list[i] != blank_index]
    last_max_prob_list.extend(max_index_list)
    last_max_index_list.extend(max_prob_list)
    index_list = [g[0] for g in groupby(last_max_prob_list)]
    index_list = [i for i in index_list if i != blank_index]
    text = "".join([vocabulary[i] for i in index_list])
    score = 0
    if len(last_max_index_list) > 0:
        score = float(sum(last_max_index_list) / len(last_max_index_list)) * 100.0
    return score, text.replace("<space>", " "), last_max_prob_list, last_max_index_list


# ------------------------------------------------------------------------------------------------
# prefix beam search  (published algorithm of paddlespeech_ctcdecoders, no external scorer)
# ------------------------------------------------------------------------------------------------
def log_sum_exp(x, y):
    """decoder_utils.h log_sum_exp: -inf aware, max + log(exp(x-max) + exp(y-max))."""
    if x == -NUM_FLT_INF:
        return y
    if y == -NUM_FLT_INF:
        return x
    m = max(x, y)
    return math.log(math.exp(x - m) + math.exp(y - m)) + m


def get_pruned_log_probs(prob_step, cutoff_prob, cutoff_top_n):
    """decoder_utils.cpp get_pruned_log_probs: optional sort by prob desc, keep the smallest head whose
    cumulative probability >= cutoff_prob, capped at cutoff_top_n; returns [(index, log(p + FLT_MIN))]."""
    prob_step = np.asarray(prob_step)
    V = prob_step.shape[0]
    idx = list(range(V))
    cutoff_len = V
    if cutoff_prob < 1.0 or cutoff_top_n < cutoff_len:
        # std::sort with pair_comp_second_rev (descending by prob); ties: stable by index here
        idx = sorted(idx, key=lambda i: (-float(prob_step[i]), i))
        if cutoff_prob < 1.0:
            cum_prob = 0.0
            cutoff_len = 0
            for i in range(V):
                cum_prob += float(prob_step[idx[i]])
                cutoff_len += 1
                if cum_prob >= cutoff_prob or cutoff_len >= cutoff_top_n:
                    break
        else:
            cutoff_len = cutoff_top_n
        idx = idx[:cutoff_len]
    return [(i, math.log(float(prob_step[i]) + NUM_FLT_MIN)) for i in idx]


class _Node:
    """path_trie.h PathTrie without dictionary/FST support."""
    __slots__ = ("ch", "parent", "children", "exists", "b_prev", "nb_prev", "b_cur", "nb_cur", "score")

    def __init__(self, ch=-1, parent=None):
        self.ch = ch
        self.parent = parent
        self.children = {}
        self.exists = True
        self.b_prev = -NUM_FLT_INF
        self.nb_prev = -NUM_FLT_INF
        self.b_cur = -NUM_FLT_INF
        self.nb_cur = -NUM_FLT_INF
        self.score = -NUM_FLT_INF

    def get_path_trie(self, c):
        child = self.children.get(c)
        if child is not None:
            if not child.exists:
                child.exists = True
                child.b_prev = child.nb_prev = child.b_cur = child.nb_cur = -NUM_FLT_INF
            return child
        child = _Node(c, self)
        self.children[c] = child
        return child

    def iterate_to_vec(self, out):
        if self.exists:
            self.b_prev = self.b_cur
            self.nb_prev = self.nb_cur
            self.score = log_sum_exp(self.b_prev, self.nb_prev)
            self.b_cur = -NUM_FLT_INF
            self.nb_cur = -NUM_FLT_INF
            out.append(self)
        for c in sorted(self.children):
            self.children[c].iterate_to_vec(out)

    def remove(self):
        self.exists = False
        if not self.children and self.parent is not None:
            del self.parent.children[self.ch]
            if not self.parent.children and not self.parent.exists:
                self.parent.remove()

    def path(self):
        out = []
        n = self
        while n.parent is not None:
            out.append(n.ch)
            n = n.parent
        return out[::-1]


def _prefix_key(n):
    # prefix_compare: score descending, then character ascending
    return (-n.score, n.ch)


def make_ngram(prefix, vocabulary, max_order):
    """scorer.cpp make_ngram for a character-based LM: the last `max_order` tokens of the prefix (one token = one LM word),
    left-padded with <s> up to max_order items."""
    toks = []
    n = prefix
    while n.parent is not None and len(toks) < max_order:
        toks.append(vocabulary[n.ch])
        n = n.parent
    if len(toks) < max_order:
        toks += ["<s>"] * (max_order - len(toks))
    return toks[::-1]


def ctc_beam_search_ids(probs_seq, beam_size, cutoff_prob=1.0, cutoff_top_n=40, blank_id=0, ext_scorer=None, vocabulary=None):
    """ctc_beam_search_decoder.cpp ctc_beam_search_decoding. Returns [(score, [ids])] sorted best first; score is the log of
    the total CTC probability of the prefix (log_prob_b + log_prob_nb at the last frame), plus, with an external scorer
    (character based), alpha * ln P_LM(c | history) + beta for every emitted character.
    ext_scorer: object with .alpha, .beta, .get_max_order(), .get_log_cond_prob(words) (ppasr_b200.decoders.ngram_lm.Scorer
    mirrors paddlespeech_ctcdecoders.Scorer); with it the published min_cutoff pruning is applied as well."""
    probs_seq = np.asarray(probs_seq)
    root = _Node()
    root.score = root.b_prev = 0.0
    prefixes = [root]
    for t in range(probs_seq.shape[0]):
        log_prob_idx = get_pruned_log_probs(probs_seq[t], cutoff_prob, cutoff_top_n)
        min_cutoff = -NUM_FLT_INF
        full_beam = False
        if ext_scorer is not None:
            num_prefixes = min(len(prefixes), beam_size)
            prefixes[:num_prefixes] = sorted(prefixes[:num_prefixes], key=_prefix_key)
            pb = float(probs_seq[t][blank_id])
            min_cutoff = prefixes[num_prefixes - 1].score + (math.log(pb) if pb > 0 else -NUM_FLT_INF) - max(0.0, ext_scorer.beta)
            full_beam = num_prefixes == beam_size
        for c, log_prob_c in log_prob_idx:
            for i in range(min(len(prefixes), beam_size)):
                prefix = prefixes[i]
                if full_beam and log_prob_c + prefix.score < min_cutoff:
                    break
                if c == blank_id:
                    prefix.b_cur = log_sum_exp(prefix.b_cur, log_prob_c + prefix.score)
                    continue
                if c == prefix.ch:
                    prefix.nb_cur = log_sum_exp(prefix.nb_cur, log_prob_c + prefix.nb_prev)
                prefix_new = prefix.get_path_trie(c)
                log_p = -NUM_FLT_INF
                if c == prefix.ch and prefix.b_prev > -NUM_FLT_INF:
                    log_p = log_prob_c + prefix.b_prev
                elif c != prefix.ch:
                    log_p = log_prob_c + prefix.score
                if ext_scorer is not None and log_p > -NUM_FLT_INF:  # character-based LM: score every new character
                    ngram = make_ngram(prefix_new, vocabulary, ext_scorer.get_max_order())
                    log_p += ext_scorer.get_log_cond_prob(ngram) * ext_scorer.alpha + ext_scorer.beta
                prefix_new.nb_cur = log_sum_exp(prefix_new.nb_cur, log_p)
        prefixes = []
        root.iterate_to_vec(prefixes)
        if len(prefixes) >= beam_size:
            prefixes.sort(key=_prefix_key)  # nth_element + later sort: only the top-beam set matters
            for n in prefixes[beam_size:]:
                n.remove()
            prefixes = prefixes[:beam_size]
    prefixes.sort(key=_prefix_key)
    return [(n.score, n.path()) for n in prefixes[:beam_size]]


def ctc_beam_search_decoding(probs_seq, vocabulary, beam_size, cutoff_prob=1.0, cutoff_top_n=40, blank_id=0, ext_scorer=None):
    """Mirror of ppasr/decoders/swig_wrapper.py:35-64. The published decoder returns `-approx_ctc`, best hypothesis first;
    without a scorer approx_ctc is the prefix score itself, with one the word-insertion and LM terms are taken out again:
    approx_ctc = score - len * beta - alpha * get_sent_log_prob(words)  (ctc_beam_search_decoder.cpp, end of decoding)."""
    res = ctc_beam_search_ids(probs_seq, beam_size, cutoff_prob, cutoff_top_n, blank_id, ext_scorer, vocabulary)
    out = []
    for s, ids in res:
        approx = s
        if ext_scorer is not None:
            words = [vocabulary[i] for i in ids]
            approx = s - len(ids) * ext_scorer.beta - ext_scorer.get_sent_log_prob(words) * ext_scorer.alpha
        out.append((-approx, "".join(vocabulary[i] for i in ids)))
    return out


def ctc_prefix_total_logprob_bruteforce(probs_seq, blank_id=0):
    """Exact log P(prefix) for every label sequence by enumerating all V^T alignments (tiny cases only).
    Used as the property check of the beam search restatement."""
    probs_seq = np.asarray(probs_seq, dtype=np.float64)
    T, V = probs_seq.shape
    import itertools
    totals = {}
    for path in itertools.product(range(V), repeat=T):
        p = 1.0
        for t, c in enumerate(path):
            p *= probs_seq[t, c]
        lab = tuple(c for c, _ in groupby(path) if c != blank_id)
        totals[lab] = totals.get(lab, 0.0) + p
    return {k: math.log(v) if v > 0 else -NUM_FLT_INF for k, v in totals.items()}
