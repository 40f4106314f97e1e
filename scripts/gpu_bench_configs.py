"""Single-GPU measurements of the other BASELINE configs (parity-test cases, not the headline bench line):
C3 conformer non-streaming b64x30s beam 10; C4 shard squeezeformer b32x10s greedy; C5 shard efficient conformer b64x5s
beam 20 + 4-gram LM (offline forward); C1 deepspeech2 1x5s greedy. Writes gpurun_out/configs_r1.json."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from ppasr_b200.engine import ConformerEngine
from ppasr_b200.decoders.beam_search_decoder import BeamSearchDecoder
from ppasr_b200.decoders.ngram_lm import NGramLM, Scorer
from ppasr_b200 import weights as W

V = 4233
vocab = W.make_vocab(V)
out = {}

def timed(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

# ---- C3
cfg = W.ConformerConfig(vocab_size=V, streaming=False)
eng = ConformerEngine(cfg, W.init_conformer_weights(cfg))
fd = torch.from_numpy(W.synthetic_fbank(64, 2998)).cuda()
dec = BeamSearchDecoder(beam_size=10, cutoff_prob=0.99, cutoff_top_n=40, vocab_list=vocab)
def c3():
    eng.encode(fd); p = eng.ctc_probs(); return dec.decode_ids_batch(p, nbest=1)
ms_enc = timed(lambda: (eng.encode(fd), eng.ctc_probs()), reps=5)
ms_all = timed(c3, reps=3, warm=1)
out["C3 conformer non-streaming b64x30s beam10"] = {"encode+ctc_probs_ms": ms_enc, "with_beam_search_ms": ms_all,
                                                     "utt_per_s": 64 / ms_all * 1e3, "rtf": ms_all * 1e-3 / (64 * 30)}
print(out, flush=True)
eng.close(); del eng, fd
# ---- C4 shard
cfg = W.SqueezeformerConfig(vocab_size=V)
eng = ConformerEngine(cfg, W.init_squeezeformer_weights(cfg))
fd = torch.from_numpy(W.synthetic_fbank(32, 998)).cuda()
ms = timed(lambda: (eng.encode(fd), eng.ctc_greedy(to_host=False)), reps=20)
out["C4 shard squeezeformer streaming b32x10s greedy"] = {"ms_per_step": ms, "utt_per_s": 32 / ms * 1e3, "rtf": ms * 1e-3 / 320}
eng.close(); del eng, fd
# ---- C5 shard
cfg = W.EfficientConformerConfig(vocab_size=V)
eng = ConformerEngine(cfg, W.init_efficient_conformer_weights(cfg))
fd = torch.from_numpy(W.synthetic_fbank(64, 498)).cuda()
rng = np.random.RandomState(0)
sents = [[vocab[2 + int(z) % (V - 3)] for z in rng.zipf(1.3, size=rng.randint(5, 30))] for _ in range(20000)]
lm = NGramLM.from_counts(sents, order=4)
dec = BeamSearchDecoder(alpha=2.2, beta=4.3, beam_size=20, cutoff_prob=0.99, cutoff_top_n=40, vocab_list=vocab,
                        ext_scorer=Scorer(2.2, 4.3, None, vocab, lm=lm))
def c5():
    eng.encode(fd); p = eng.ctc_probs(); return dec.decode_batch_beam_search_offline(p)   # texts, like the reference's evaluate
ms_enc = timed(lambda: (eng.encode(fd), eng.ctc_probs()), reps=10)
ms_all = timed(c5, reps=3, warm=1)
out["C5 shard efficient_conformer b64x5s beam20 + 4-gram LM (offline forward)"] = {
    "lm_ngrams": len(lm.ngrams), "encode+ctc_probs_ms": ms_enc, "with_beam_search_ms": ms_all, "utt_per_s": 64 / ms_all * 1e3,
    "rtf": ms_all * 1e-3 / (64 * 5)}
eng.close(); del eng, fd
# ---- C1
cfg = W.DeepSpeech2Config(vocab_size=V)
eng = ConformerEngine(cfg, W.init_deepspeech2_weights(cfg))
fd = torch.from_numpy(W.synthetic_fbank(1, 498)).cuda()
ms = timed(lambda: (eng.encode(fd), eng.ctc_greedy(to_host=False)), reps=10)
out["C1 deepspeech2 non-streaming 1x5s greedy"] = {"ms_per_utt": ms, "rtf": ms * 1e-3 / 5}
eng.close(); del eng, fd
# ---- 8f: GPU fbank front end and ragged streaming sessions
from ppasr_b200.featurizer import GpuFbank
from ppasr_b200.infer_utils.inference_predictor import InferencePredictor
fb = GpuFbank()
audio = torch.randn(32, 160000, device="cuda") * 0.05
ms = timed(lambda: fb.featurize_batch(audio), reps=20)
out["8f fbank 32x10s waveforms -> [32,998,80] (dB norm + int16 + Kaldi fbank)"] = {
    "ms": ms, "audio_seconds_per_s": 320 / ms * 1e3, "bytes_in_out_GBps": (32 * 160000 * 4 * 2 + 32 * 998 * 80 * 4) / ms / 1e6}
cfg = W.ConformerConfig(vocab_size=V)
pred = InferencePredictor({"encoder_conf": cfg.to_dict(), "preprocess_conf": {"n_mels": 80}}, "conformer", streaming=True,
                          weights=W.init_conformer_weights(cfg))
S = 64
pred.engine.sessions_init(S)
chunk = torch.from_numpy(W.synthetic_fbank(S, 67)).cuda()
slots = list(range(S))
def sess_step():
    pred.engine.sessions_step(chunk, slots, -1); pred.engine.ctc_greedy(to_host=False)
for _ in range(10): sess_step()   # 10 chunks of history (160 cached keys)
ms = timed(sess_step, reps=20, warm=0)
out["8f ragged streaming sessions: 64 conformer streams, one 67-frame chunk each (0.64 s of audio per stream)"] = {
    "ms_per_step": ms, "real_time_streams_per_gpu": 64 * 0.64 / (ms * 1e-3)}
for k, v in out.items(): print(k, {a: (round(b, 5) if isinstance(b, float) else b) for a, b in v.items()}, flush=True)
json.dump(out, open('gpurun_out/configs_r1.json', 'w'), indent=1)
