"""Back-off n-gram language model for the CTC beam-search scorer (ARPA text format).

Stands in for the KenLM model behind `paddlespeech_ctcdecoders.Scorer` (ppasr/decoders/swig_wrapper.py:4-19,
beam_search_decoder.py:29-40). KenLM itself is a third-party C++ dependency of that (also third-party) package; neither
is vendored under /root/reference. What is restated here is the published scorer behaviour (PaddleSpeech
third_party/ctc_decoders/scorer.cpp, recalled -- parity unpinned):
  * get_log_cond_prob(words): walk the words from the NULL context with the back-off recursion
        P(w | ctx) = p(ctx w) if the n-gram exists else backoff(ctx) * P(w | ctx[1:]);
    any out-of-vocabulary word -> OOV_SCORE = -1000; result converted from log10 to natural log;
  * is_character_based: every LM unigram other than <s>, </s>, <unk> is a single character;
  * make_ngram(prefix): the last `order` tokens of the prefix, left-padded with <s> to `order` items.
Only ARPA text files are read (the KenLM binary formats would need KenLM); `from_counts` builds a small model with
absolute discounting for tests and synthetic benchmarks. `device_tables()` exports the model as an open-addressing hash
table keyed by up to five 16-bit token ids (the reference trains its LMs with `-o 5`, docs/beam_search.md) for the GPU scorer
(csrc/beam.cu).
"""
import math
from collections import defaultdict

import numpy as np

OOV_SCORE = -1000.0
START_TOKEN, END_TOKEN, UNK_TOKEN = "<s>", "</s>", "<unk>"
LOG10_E = 0.4342944819032518  # NUM_FLT_LOGE in decoder_utils.h


class NGramLM:
    def __init__(self, ngrams, order):
        """ngrams: dict tuple(words) -> (log10 prob, log10 backoff)."""
        self.ngrams = ngrams
        self.order = order
        self.vocab = {k[0] for k in ngrams if len(k) == 1}

    # ---- construction -----------------------------------------------------------------------------------------
    @classmethod
    def from_arpa(cls, path):
        ngrams = {}
        order = 0
        cur = 0
        with open(path, "r", encoding="utf-8") as f:
            for line in f:
                line = line.rstrip("\n")
                if not line or line.startswith("ngram ") or line == "\\data\\":
                    continue
                if line.startswith("\\") and line.endswith("-grams:"):
                    cur = int(line[1:line.index("-")])
                    order = max(order, cur)
                    continue
                if line == "\\end\\":
                    break
                if cur == 0:
                    continue
                parts = line.split("\t") if "\t" in line else line.split()
                if "\t" in line:
                    prob = float(parts[0])
                    words = tuple(parts[1].split(" "))
                    bo = float(parts[2]) if len(parts) > 2 else 0.0
                else:
                    prob = float(parts[0])
                    words = tuple(parts[1:1 + cur])
                    bo = float(parts[1 + cur]) if len(parts) > 1 + cur else 0.0
                ngrams[words] = (prob, bo)
        return cls(ngrams, order)

    @classmethod
    def from_counts(cls, sentences, order=4, discount=0.7):
        """Small interpolated absolute-discounting model estimated from token sequences (lists of str)."""
        counts = [defaultdict(int) for _ in range(order + 1)]
        for s in sentences:
            toks = [START_TOKEN] + list(s) + [END_TOKEN]
            for n in range(1, order + 1):
                for i in range(len(toks) - n + 1):
                    counts[n][tuple(toks[i:i + n])] += 1
        vocab = sorted({k[0] for k in counts[1]} | {UNK_TOKEN})
        total = sum(counts[1].values())
        probs = {}  # tuple -> prob (interpolated)
        for w in vocab:
            c = counts[1].get((w,), 0)
            probs[(w,)] = (max(c - discount, 0.0) + discount * len(counts[1]) / len(vocab)) / total if total else 1.0 / len(vocab)
        ctx_types = [defaultdict(int) for _ in range(order + 1)]
        ctx_total = [defaultdict(int) for _ in range(order + 1)]
        for n in range(2, order + 1):
            for ng, c in counts[n].items():
                ctx_types[n][ng[:-1]] += 1
                ctx_total[n][ng[:-1]] += c
        bows = {}
        for n in range(2, order + 1):
            for ng, c in counts[n].items():
                ctx = ng[:-1]
                lam = discount * ctx_types[n][ctx] / ctx_total[n][ctx]
                lower = probs.get(ng[1:], probs.get(ng[-1:], 1e-9))
                probs[ng] = max(c - discount, 0.0) / ctx_total[n][ctx] + lam * lower
                bows[ctx] = lam
        ngrams = {}
        for ng, p in probs.items():
            bo = bows.get(ng, 1.0) if len(ng) < order else 1.0
            ngrams[ng] = (math.log10(max(p, 1e-30)), math.log10(max(bo, 1e-30)) if len(ng) < order else 0.0)
        return cls(ngrams, order)

    def write_arpa(self, path):
        by_n = defaultdict(list)
        for ng, v in self.ngrams.items():
            by_n[len(ng)].append((ng, v))
        with open(path, "w", encoding="utf-8") as f:
            f.write("\\data\\\n")
            for n in range(1, self.order + 1):
                f.write(f"ngram {n}={len(by_n[n])}\n")
            for n in range(1, self.order + 1):
                f.write(f"\n\\{n}-grams:\n")
                for ng, (p, bo) in sorted(by_n[n]):
                    f.write(f"{p:.6f}\t{' '.join(ng)}" + (f"\t{bo:.6f}\n" if n < self.order else "\n"))
            f.write("\n\\end\\\n")

    # ---- queries (log10, KenLM back-off semantics) --------------------------------------------------------------
    def _score(self, ctx, w):
        bo = 0.0
        for start in range(len(ctx) + 1):
            ng = ctx[start:] + (w,)
            hit = self.ngrams.get(ng)
            if hit is not None:
                return bo + hit[0]
            if start < len(ctx):
                c = self.ngrams.get(ctx[start:])
                if c is not None:
                    bo += c[1]
        return None  # w is OOV

    def log_cond_prob(self, words):
        """scorer.cpp get_log_cond_prob: natural-log P(words[-1] | words[:-1]) scoring the words in sequence from the NULL
        context; OOV anywhere -> OOV_SCORE."""
        cond = 0.0
        ctx = ()
        for w in words:
            if w not in self.vocab:
                return OOV_SCORE
            cond = self._score(ctx, w)
            ctx = (ctx + (w,))[-(self.order - 1):] if self.order > 1 else ()
        return cond / LOG10_E

    def sent_log_prob(self, words):
        """scorer.cpp get_sent_log_prob / get_log_prob."""
        words = list(words)
        if not words:
            sent = [START_TOKEN] * self.order
        else:
            sent = [START_TOKEN] * (self.order - 1) + words
        sent.append(END_TOKEN)
        score = 0.0
        for i in range(self.order - 1, len(sent)):
            score += self.log_cond_prob(sent[i - self.order + 1:i + 1])
        return score

    def is_character_based(self):
        return all(len(w) == 1 for w in self.vocab if w not in (START_TOKEN, END_TOKEN, UNK_TOKEN))

    # ---- GPU export -------------------------------------------------------------------------------------------------
    def device_tables(self, vocab_list, with_hi=False):
        """Hash table keyed by token ids: id 0 = empty slot, 1 = <s>, v + 2 = vocab_list[v]. Key = (lo, hi): lo = the four
        most recent ids packed 16 bits each (most recent token in the low bits), hi = the oldest id of a 5-gram (0 for shorter
        n-grams). Values (ln prob, ln backoff). Also returns in_lm[v] (unigram present). Returns (keys_lo, vals, in_lm) or,
        with_hi, (keys_lo, keys_hi, vals, in_lm)."""
        assert len(vocab_list) + 2 < 65536 and self.order <= 5
        tok = {START_TOKEN: 1}
        for v, w in enumerate(vocab_list):
            tok.setdefault(w, v + 2)
        items = []
        for ng, (p, bo) in self.ngrams.items():
            ids = [tok.get(w) for w in ng]
            if any(i is None for i in ids):
                continue
            key = 0
            for i in ids[-4:]:  # oldest first -> most recent ends in the low 16 bits
                key = (key << 16) | i
            items.append((key, ids[0] if len(ids) == 5 else 0, p / LOG10_E, bo / LOG10_E))
        cap = 1
        while cap < 2 * max(1, len(items)) + 2:
            cap *= 2
        keys = np.zeros(cap, dtype=np.uint64)
        keys_hi = np.zeros(cap, dtype=np.uint32)
        vals = np.zeros((cap, 2), dtype=np.float32)
        mask = cap - 1
        for key, hi, p, bo in items:
            slot = lm_hash64_2(key, hi) & mask
            while keys[slot] != 0:
                slot = (slot + 1) & mask
            keys[slot] = key
            keys_hi[slot] = hi
            vals[slot] = (p, bo)
        in_lm = np.array([1 if w in self.vocab else 0 for w in vocab_list], dtype=np.int32)
        if with_hi:
            return keys, keys_hi, vals, in_lm
        return keys, vals, in_lm


def lm_hash64(x):
    """Home slot hash of the device table (the same finaliser runs in csrc/beam.cu::lm_find). The keys are token ids packed
    16 bits each, so the hash must mix the HIGH fields into the low bits: a plain multiplicative hash taken from the middle
    of the product ignores the oldest tokens of a 4-gram and piles thousands of n-grams onto the same slots."""
    x &= 0xFFFFFFFFFFFFFFFF
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 33
    x = (x * 0xC4CEB9FE1A85EC53) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 33
    return x


def lm_hash64_2(lo, hi):
    """Home slot of the key (lo, hi); equals lm_hash64(lo) for n-grams of up to four tokens (hi = 0)."""
    return lm_hash64(lo ^ ((hi * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF))


class Scorer:
    """Mirror of paddlespeech_ctcdecoders.Scorer(alpha, beta, model_path, vocabulary) (swig_wrapper.py:4-19)."""

    def __init__(self, alpha, beta, model_path, vocabulary, lm=None):
        self.alpha = float(alpha)
        self.beta = float(beta)
        self.vocabulary = list(vocabulary)
        self.lm = lm if lm is not None else NGramLM.from_arpa(model_path)
        self.max_order = self.lm.order
        self.character_based = self.lm.is_character_based()

    def is_character_based(self):
        return self.character_based

    def get_max_order(self):
        return self.max_order

    def get_log_cond_prob(self, words):
        return self.lm.log_cond_prob(words)

    def get_sent_log_prob(self, words):
        return self.lm.sent_log_prob(words)

    def reset_params(self, alpha, beta):
        self.alpha, self.beta = float(alpha), float(beta)
