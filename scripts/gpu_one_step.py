"""One C2-size encode+greedy step of a 1-block model (for ncu captures of individual kernels)."""
import sys, torch
sys.path.insert(0, '.')
from ppasr_b200.engine import ConformerEngine
from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = ConformerConfig(num_blocks=nb, vocab_size=4233)
eng = ConformerEngine(cfg, init_conformer_weights(cfg))
fd = torch.from_numpy(synthetic_fbank(32, 998)).cuda()
for _ in range(2):
    eng.encode(fd); eng.ctc_greedy(to_host=False)
torch.cuda.synchronize()
print("ONE STEP DONE")
