import sys, torch
sys.path.insert(0, '.')
from ppasr_b200.engine import ConformerEngine
from ppasr_b200.weights import DeepSpeech2Config, init_deepspeech2_weights, synthetic_fbank
cfg = DeepSpeech2Config(vocab_size=4233, num_rnn_layers=1)
eng = ConformerEngine(cfg, init_deepspeech2_weights(cfg))
fd = torch.from_numpy(synthetic_fbank(1, 498)).cuda()
for _ in range(2):
    eng.encode(fd); eng.ctc_greedy(to_host=False)
torch.cuda.synchronize()
print("ONE STEP DONE")
