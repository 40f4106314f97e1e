"""Efficient-Conformer engine vs the CPU oracle (logits): grouped attention, stride block, ragged lengths; timing."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from ppasr_b200.engine import ConformerEngine
from ppasr_b200.weights import EfficientConformerConfig, init_efficient_conformer_weights, synthetic_fbank
from oracle.efficient_conformer_oracle import EfficientConformerOracle, EfficientConformerConf

def run(nb, B, T, lens, streaming=True, vocab=300, **kw):
    cfg = EfficientConformerConfig(num_blocks=nb, vocab_size=vocab, streaming=streaming, **kw)
    w = init_efficient_conformer_weights(cfg)
    feats = synthetic_fbank(B, T)
    for b in range(B): feats[b, lens[b]:] = 0
    eng = ConformerEngine(cfg, w)
    eng.encode(torch.from_numpy(feats).cuda(), lens)
    lg = eng.ctc_logits().float().cpu(); torch.cuda.synchronize()
    ref, masks = None, None
    o = EfficientConformerOracle(EfficientConformerConf(**cfg.to_dict()), w)
    ref = o.get_encoder_out(torch.from_numpy(feats), torch.tensor(lens), return_logits=True)
    assert lg.shape == ref.shape, (lg.shape, ref.shape)
    _, m = o.encoder_forward(torch.from_numpy(feats), torch.tensor(lens))
    vl = m.squeeze(1).sum(1).tolist()
    rel = max((lg[b, :vl[b]] - ref[b, :vl[b]]).abs().max().item() for b in range(B)) / ref.abs().max().item()
    print(f"L={nb} B={B} T={T} streaming={streaming} {kw}: logits rel err {rel:.4g}", flush=True)
    eng.close()

run(1, 2, 131, [131, 90], group_layer_idx=(0,), stride_layer_idx=None)           # grouped attention only (T'=32: pad 1)
run(1, 2, 135, [135, 90], group_layer_idx=(), stride_layer_idx=0)                # stride block only (T'=33 odd)
run(2, 3, 203, [203, 150, 99], group_layer_idx=(0, 1), stride_layer_idx=1)
run(2, 3, 207, [207, 150, 5], group_layer_idx=(0, 1), stride_layer_idx=1, streaming=False)
run(3, 2, 1051, [1051, 700], group_layer_idx=(0, 1), stride_layer_idx=1)         # T'=262 -> 88 groups (2 key blocks)
run(12, 4, 498, [498, 498, 400, 250], vocab=4233)
run(12, 2, 998, [998, 600], vocab=4233, streaming=False)
cfg = EfficientConformerConfig(vocab_size=4233)
eng = ConformerEngine(cfg, init_efficient_conformer_weights(cfg))
fd = torch.from_numpy(synthetic_fbank(64, 498)).cuda()
for _ in range(3): eng.encode(fd); eng.ctc_greedy(to_host=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): eng.encode(fd); eng.ctc_greedy(to_host=False)
e1.record(); torch.cuda.synchronize()
print(f"efficient conformer b64x5s (C5 per-GPU shard, offline): {e0.elapsed_time(e1)/20:.3f} ms/step", flush=True)
eng.profile_enable(True)
for _ in range(5): eng.encode(fd); eng.ctc_greedy(to_host=False)
for k, v in sorted(eng.profile_read().items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:18s} {v[0]:4d} launches {v[1]/v[0]*1e3:8.1f} us")
print("EFFCONF DONE")
