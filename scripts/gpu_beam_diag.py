"""Where does the beam-search decode time go? Times prune / prefix kernels (with and without the n-gram scorer) on model-like and
synthetic posteriors at the C3 and C5 shapes, and the Python wrapper around them."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from ppasr_b200 import _lib as L
from ppasr_b200.decoders.beam_search_decoder import BeamSearchDecoder
from ppasr_b200.decoders.ngram_lm import NGramLM, Scorer
from ppasr_b200 import weights as W
lib = L.load(); dev = torch.device('cuda:0'); torch.manual_seed(0)
V = 4233
vocab = W.make_vocab(V)

def ev_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

rng = np.random.RandomState(0)
sents = [[vocab[2 + int(z) % (V - 3)] for z in rng.zipf(1.3, size=rng.randint(5, 30))] for _ in range(20000)]
t0 = time.perf_counter(); lm = NGramLM.from_counts(sents, order=4); print("lm build s", time.perf_counter() - t0, len(lm.ngrams), flush=True)
for name, (B, T, beam, temp) in {"C5 b64 T62 beam20": (64, 62, 20, 3.0), "C3 b64 T748 beam10": (64, 748, 10, 3.0), "C5 peaked": (64, 62, 20, 8.0)}.items():
    probs = torch.softmax(torch.randn(B, T, V, device=dev) * temp, -1).contiguous()
    ws = torch.empty(lib.ppasr_b200_beam_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
    t_prune = ev_time(lambda: L.check(lib.ppasr_b200_op_ctc_prune(L.ptr(probs), B * T, V, 0.99, 40, L.ptr(ws), L.stream_ptr())))
    cnt = ws[:B * T * 4].view(torch.int32).float().mean().item()
    dec = BeamSearchDecoder(beam_size=beam, cutoff_prob=0.99, cutoff_top_n=40, vocab_list=vocab)
    st = dec._alloc_state(B, T)
    def adv(d, s):
        L.check(lib.ppasr_b200_beam_reset(L.ptr(s), B, T, beam, L.stream_ptr())); d._advance(s, probs, None, T)
    t_plain = ev_time(lambda: adv(dec, st))
    t0 = time.perf_counter(); decl = BeamSearchDecoder(alpha=2.2, beta=4.3, beam_size=beam, cutoff_prob=0.99, cutoff_top_n=40, vocab_list=vocab,
                                                       ext_scorer=Scorer(2.2, 4.3, None, vocab, lm=lm)); t_ctor = time.perf_counter() - t0
    stl = decl._alloc_state(B, T)
    t_lm = ev_time(lambda: adv(decl, stl), reps=3)
    t0 = time.perf_counter(); r = decl.decode_ids_batch(probs, nbest=1, approx=False); torch.cuda.synchronize(); t_wall = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); r = decl.decode_ids_batch(probs, nbest=1, approx=True); torch.cuda.synchronize(); t_wall_approx = (time.perf_counter() - t0) * 1e3
    print(name, f"mean candidates/frame {cnt:.1f}; prune {t_prune:.3f} ms; prune+prefix no LM {t_plain:.3f} ms; with LM {t_lm:.3f} ms; "
          f"decode_ids_batch wall (approx=False) {t_wall:.2f} ms, (approx=True) {t_wall_approx:.2f} ms; LM decoder ctor {t_ctor:.2f} s", flush=True)
