"""Generates tests/golden/greedy_golden.npz by running the REFERENCE's own greedy decoder
(/root/reference/ppasr/decoders/ctc_greedy_decoder.py -- pure NumPy, importable without Paddle).

Run once in the build container (where /root/reference exists):  python tests/golden/make_greedy_golden.py
The committed .npz is what the CPU and GPU tests compare against; /root/reference is never read at test time.
"""
import importlib.util
import json
import os

import numpy as np

REF = "/root/reference/ppasr/decoders/ctc_greedy_decoder.py"
spec = importlib.util.spec_from_file_location("ref_greedy", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

HERE = os.path.dirname(os.path.abspath(__file__))


def vocab(V):
    v = ["<blank>", "<unk>"] + [chr(0x4E00 + i) for i in range(V - 4)] + ["<space>", "<eos>"]
    return v[:V]


def softmax(x):
    e = np.exp(x - x.max(-1, keepdims=True))
    return (e / e.sum(-1, keepdims=True)).astype(np.float32)


def main():
    rng = np.random.RandomState(20260922)
    cases = {}
    meta = []
    specs = [
        ("peaked", 60, 40, 6.0), ("flat", 33, 17, 0.3), ("blanky", 80, 29, 4.0), ("single_frame", 1, 12, 3.0),
        ("all_blank", 25, 10, 5.0), ("repeats", 64, 8, 5.0), ("ties", 20, 16, 0.0), ("long", 248, 97, 5.0),
        ("space", 40, 9, 6.0),
    ]
    for name, T, V, temp in specs:
        logits = rng.randn(T, V).astype(np.float32) * temp
        if name == "blanky":
            logits[:, 0] += 6.0
        if name == "all_blank":
            logits[:, 0] += 100.0
        if name == "repeats":
            logits = np.repeat(logits[::4], 4, axis=0)[:T]
        probs = softmax(logits)
        if name == "ties":
            probs = np.full((T, V), 1.0 / V, dtype=np.float32)  # exact ties: first index wins
            probs[3, 5] = probs[3, 9] = 0.25
        if name == "space":
            probs[:, V - 2] += 0.4  # make '<space>' frequent
            probs = (probs / probs.sum(-1, keepdims=True)).astype(np.float32)
        v = vocab(V)
        score, text = ref.greedy_decoder(probs, v)
        cases[f"{name}_probs"] = probs
        m = {"name": name, "T": T, "V": V, "score": repr(float(score)), "text": text}
        # streaming variant: feed in chunks of 16 frames through the reference's greedy_decoder_chunk
        lp, li = None, None
        chunk_out = []
        for s in range(0, T, 16):
            sc, tx, lp, li = ref.greedy_decoder_chunk(probs[s:s + 16], v, lp, li)
            chunk_out.append({"score": repr(float(sc)), "text": tx})
        m["chunks"] = chunk_out
        meta.append(m)
    # batch API
    batch = [cases[f"{n}_probs"] for n in ("peaked", "long")]
    vb = vocab(97)
    meta.append({"name": "__batch__", "texts": ref.greedy_decoder_batch([cases["long_probs"], cases["long_probs"][:100]], vb)})
    np.savez_compressed(os.path.join(HERE, "greedy_golden.npz"), **cases)
    with open(os.path.join(HERE, "greedy_golden.json"), "w", encoding="utf-8") as f:
        json.dump(meta, f, ensure_ascii=False, indent=1)
    print("wrote", len(specs), "cases")


if __name__ == "__main__":
    main()
