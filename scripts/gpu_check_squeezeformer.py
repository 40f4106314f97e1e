"""Squeezeformer engine vs the CPU oracle (logits), small + full size, + timing at the C4 per-GPU shape."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from ppasr_b200.engine import ConformerEngine, out_frames
from ppasr_b200.weights import SqueezeformerConfig, init_squeezeformer_weights, synthetic_fbank
from oracle.squeezeformer_oracle import SqueezeformerOracle, SqueezeformerConf

def run(nb, B, T, lens, streaming=True, reduce_idx=None, recover_idx=None, vocab=300):
    cfg = SqueezeformerConfig(num_blocks=nb, vocab_size=vocab, streaming=streaming, reduce_idx=reduce_idx, recover_idx=recover_idx)
    w = init_squeezeformer_weights(cfg)
    feats = synthetic_fbank(B, T)
    for b in range(B): feats[b, lens[b]:] = 0
    eng = ConformerEngine(cfg, w)
    eng.encode(torch.from_numpy(feats).cuda(), lens)
    lg = eng.ctc_logits().float().cpu(); torch.cuda.synchronize()
    ref = SqueezeformerOracle(SqueezeformerConf(**cfg.to_dict()), w).get_encoder_out(torch.from_numpy(feats), torch.tensor(lens), return_logits=True)
    Tp = out_frames(T); vl = [min(Tp, (l + 3) // 4) for l in lens]
    rel = max((lg[b, :vl[b]] - ref[b, :vl[b]]).abs().max().item() for b in range(B)) / ref.abs().max().item()
    print(f"L={nb} B={B} T={T} streaming={streaming} reduce={reduce_idx}/{recover_idx}: logits rel err {rel:.4g}", flush=True)
    eng.close()

run(1, 2, 131, [131, 90])
run(2, 3, 203, [203, 150, 99], reduce_idx=None, recover_idx=None)
run(3, 3, 203, [203, 150, 99], reduce_idx=1, recover_idx=2)
run(3, 3, 207, [207, 150, 99], reduce_idx=1, recover_idx=2, streaming=False)
run(12, 4, 998, [998, 998, 900, 500], reduce_idx=5, recover_idx=11, vocab=4233)
run(12, 2, 523, [523, 300], reduce_idx=5, recover_idx=11, streaming=False)
cfg = SqueezeformerConfig(num_blocks=12, vocab_size=4233)
eng = ConformerEngine(cfg, init_squeezeformer_weights(cfg))
fd = torch.from_numpy(synthetic_fbank(32, 998)).cuda()
for _ in range(3): eng.encode(fd); eng.ctc_greedy(to_host=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): eng.encode(fd); eng.ctc_greedy(to_host=False)
e1.record(); torch.cuda.synchronize()
print(f"squeezeformer b32x10s: {e0.elapsed_time(e1)/20:.3f} ms/step", flush=True)
eng.profile_enable(True)
for _ in range(5): eng.encode(fd); eng.ctc_greedy(to_host=False)
prof = eng.profile_read()
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:18s} {v[0]:4d} launches {v[1]/v[0]*1e3:8.1f} us")
print("SQUEEZE DONE")
