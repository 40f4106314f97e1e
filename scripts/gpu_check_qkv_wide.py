import sys, torch
sys.path.insert(0,'.')
from ppasr_b200 import engine as E, weights as W
cfg = W.ConformerConfig(num_blocks=3, vocab_size=301)
w = W.init_conformer_weights(cfg)
feats = torch.from_numpy(W.synthetic_fbank(3, 523)).cuda(); lens=[523,3,260]
eng = E.ConformerEngine(cfg, w, device=0)
outs=[]
for v in (0,1):
    eng.set_option("qkv_wide", v); eng.encode(feats, lens); outs.append(eng.ctc_logits().float().cpu())
print("qkv wide vs narrow bit-identical:", torch.equal(outs[0], outs[1]), (outs[0]-outs[1]).abs().max().item())
