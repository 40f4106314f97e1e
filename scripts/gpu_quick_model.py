"""Quick correctness probe: small + 12-layer model vs oracle, for both FFN implementations."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from ppasr_b200.engine import ConformerEngine, out_frames
from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank
from oracle.conformer_oracle import ConformerOracle, ConformerConf
def run(nb, B, T, lens, fused):
    cfg = ConformerConfig(num_blocks=nb, vocab_size=4233)
    w = init_conformer_weights(cfg)
    feats = synthetic_fbank(B, T)
    for b in range(B): feats[b, lens[b]:] = 0
    eng = ConformerEngine(cfg, w); eng.set_option("fused_ffn", fused); eng.set_option("fused_attn_out", fused)
    eng.encode(torch.from_numpy(feats).cuda(), lens)
    lg = eng.ctc_logits().cpu(); torch.cuda.synchronize()
    ref = ConformerOracle(ConformerConf(**cfg.to_dict()), w).get_encoder_out(torch.from_numpy(feats), torch.tensor(lens), return_logits=True)
    Tp = out_frames(T); vl = [min(Tp, (l + 3) // 4) for l in lens]
    rel = max((lg[b, :vl[b]] - ref[b, :vl[b]]).abs().max().item() for b in range(B)) / ref.abs().max().item()
    print(f"L={nb} B={B} T={T} fused={fused}: logits rel err {rel:.4g}", flush=True)
    eng.close()
for fused in (0, 1):
    run(1, 2, 131, [131, 90], fused)
    run(2, 3, 523, [523, 333, 260], fused)
run(12, 4, 998, [998, 998, 900, 500], 1)
print("QUICK DONE")
