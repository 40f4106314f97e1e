"""Which torch thread count is best for the CPU reference arm on this box?"""
import os, sys, time
import torch
sys.path.insert(0, '.')
from oracle.conformer_oracle import ConformerConf, ConformerOracle
from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "torch default threads", torch.get_num_threads())
try:
    print(open('/sys/fs/cgroup/cpu.max').read().strip())
except Exception as e:
    print("no cgroup cpu.max", e)
cfg = ConformerConfig(vocab_size=4233)
orc = ConformerOracle(ConformerConf(**cfg.to_dict()), init_conformer_weights(cfg))
x = torch.from_numpy(synthetic_fbank(4, 998)); lens = torch.tensor([998] * 4)
for nt in (torch.get_num_threads(), 8, 16, 32, 64, 128):
    torch.set_num_threads(nt)
    orc.get_encoder_out(x[:1], lens[:1])
    t = time.perf_counter(); orc.get_encoder_out(x, lens); dt = time.perf_counter() - t
    print(f"threads={nt}: {dt:.2f}s for 4 utt -> {4/dt:.2f} utt/s", flush=True)
