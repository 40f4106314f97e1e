"""DeepSpeech2 engine vs the CPU oracle (logits): bi-LSTM / forward LSTM / GRU, ragged lengths, chunked streaming; C1 timing."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from ppasr_b200.engine import ConformerEngine, out_frames
from ppasr_b200.weights import DeepSpeech2Config, init_deepspeech2_weights, synthetic_fbank
from oracle.deepspeech2_oracle import DeepSpeech2Oracle, DeepSpeech2Conf

def run(nl, H, B, T, lens, streaming=False, gru=False, vocab=300):
    cfg = DeepSpeech2Config(num_rnn_layers=nl, rnn_size=H, vocab_size=vocab, streaming=streaming, use_gru=gru)
    w = init_deepspeech2_weights(cfg)
    feats = synthetic_fbank(B, T)
    for b in range(B): feats[b, lens[b]:] = 0
    eng = ConformerEngine(cfg, w)
    eng.encode(torch.from_numpy(feats).cuda(), lens)
    lg = eng.ctc_logits().float().cpu(); torch.cuda.synchronize()
    ref = DeepSpeech2Oracle(DeepSpeech2Conf(**cfg.to_dict()), w).get_encoder_out(torch.from_numpy(feats), torch.tensor(lens), return_logits=True)
    rel = ((lg - ref).abs().max() / ref.abs().max()).item()
    print(f"L={nl} H={H} B={B} T={T} streaming={streaming} gru={gru}: logits rel err {rel:.4g}", flush=True)
    eng.close()

run(1, 128, 2, 131, [131, 90])
run(2, 256, 3, 203, [203, 150, 99], streaming=True)
run(2, 256, 3, 203, [203, 150, 99], gru=True)
run(2, 128, 11, 131, [131] * 5 + [100] * 6)
run(5, 1024, 1, 498, [498], vocab=4233)
run(5, 1024, 4, 498, [498, 400, 300, 200], streaming=True, vocab=4233)
# chunked streaming == offline for the forward model (states carried on the device)
cfg = DeepSpeech2Config(num_rnn_layers=2, rnn_size=256, vocab_size=300, streaming=True)
w = init_deepspeech2_weights(cfg)
eng = ConformerEngine(cfg, w)
o = DeepSpeech2Oracle(DeepSpeech2Conf(**cfg.to_dict()), w)
x = synthetic_fbank(1, 67 + 64 * 3)
eng.stream_reset(1)
outs = []; h = c = None; refs = []
for s in range(0, x.shape[1] - 66, 64):
    ch = x[:, s:s + 67]
    eng.encode_chunk(ch)
    outs.append(eng.ctc_logits().float().cpu())
    r, _, h, c = o.get_encoder_out_chunk(torch.from_numpy(ch), torch.tensor([67]), h, c, return_logits=True)
    refs.append(r)
got, ref = torch.cat(outs, 1), torch.cat(refs, 1)
print("chunked forward LSTM vs oracle chunked:", ((got - ref).abs().max() / ref.abs().max()).item())
eng.close()
# C1: deepspeech2.yml non-streaming, 1 x 5 s
cfg = DeepSpeech2Config(vocab_size=4233)
eng = ConformerEngine(cfg, init_deepspeech2_weights(cfg))
fd = torch.from_numpy(synthetic_fbank(1, 498)).cuda()
for _ in range(3): eng.encode(fd); eng.ctc_greedy(to_host=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.encode(fd); eng.ctc_greedy(to_host=False)
e1.record(); torch.cuda.synchronize()
print(f"deepspeech2 C1 (1 x 5 s, bi-LSTM 5x1024): {e0.elapsed_time(e1)/10:.3f} ms/utt", flush=True)
eng.profile_enable(True)
for _ in range(3): eng.encode(fd); eng.ctc_greedy(to_host=False)
for k, v in sorted(eng.profile_read().items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:18s} {v[0]:4d} launches {v[1]/v[0]*1e3:8.1f} us")
print("DS2 DONE")
