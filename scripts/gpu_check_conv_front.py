"""Fused conv front end (conv_front.cu) vs the two-kernel path: logits agreement + timing."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from ppasr_b200.engine import ConformerEngine, out_frames
from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank
from oracle.conformer_oracle import ConformerOracle, ConformerConf

def logits(eng, feats, lens, fused):
    eng.set_option("fused_conv", fused)
    eng.encode(feats, lens)
    lg = eng.ctc_logits().float().cpu(); torch.cuda.synchronize()
    return lg

for nb, B, T, lens in [(1, 2, 131, [131, 90]), (2, 3, 523, [523, 333, 260]), (1, 5, 67, [67, 67, 50, 30, 67]), (1, 1, 998, [998])]:
    cfg = ConformerConfig(num_blocks=nb, vocab_size=4233)
    w = init_conformer_weights(cfg)
    feats = synthetic_fbank(B, T)
    for b in range(B): feats[b, lens[b]:] = 0
    eng = ConformerEngine(cfg, w)
    fd = torch.from_numpy(feats).cuda()
    a = logits(eng, fd, lens, 0); b_ = logits(eng, fd, lens, 1)
    ref = ConformerOracle(ConformerConf(**cfg.to_dict()), w).get_encoder_out(torch.from_numpy(feats), torch.tensor(lens), return_logits=True)
    Tp = out_frames(T); vl = [min(Tp, (l + 3) // 4) for l in lens]
    e01 = max((a[i, :vl[i]] - b_[i, :vl[i]]).abs().max().item() for i in range(B)) / ref.abs().max().item()
    e1r = max((b_[i, :vl[i]] - ref[i, :vl[i]]).abs().max().item() for i in range(B)) / ref.abs().max().item()
    e0r = max((a[i, :vl[i]] - ref[i, :vl[i]]).abs().max().item() for i in range(B)) / ref.abs().max().item()
    print(f"L={nb} B={B} T={T}: fused-vs-unfused {e01:.3g}  fused-vs-oracle {e1r:.3g}  unfused-vs-oracle {e0r:.3g}", flush=True)
    eng.close()

cfg = ConformerConfig(num_blocks=12, vocab_size=4233)
w = init_conformer_weights(cfg)
eng = ConformerEngine(cfg, w)
fd = torch.from_numpy(synthetic_fbank(32, 998)).cuda()
for fused in (0, 1, 0, 1):
    eng.set_option("fused_conv", fused)
    for _ in range(3): eng.encode(fd); eng.ctc_greedy(to_host=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): eng.encode(fd); eng.ctc_greedy(to_host=False)
    e1.record(); torch.cuda.synchronize()
    print(f"fused_conv={fused}: {e0.elapsed_time(e1)/20:.3f} ms/step", flush=True)
eng.profile_enable(True)
for _ in range(5): eng.encode(fd); eng.ctc_greedy(to_host=False)
prof = eng.profile_read()
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:18s} {v[0]:4d} launches {v[1]/v[0]*1e3:8.1f} us")
print("CONV FRONT DONE")
