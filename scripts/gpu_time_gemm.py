"""Stand-alone timing of the persistent GEMM (op_linear, bf16 store epilogue) at the small-K shapes of the encoder."""
import sys
import torch
sys.path.insert(0, '.')
from ppasr_b200 import _lib as L
lib = L.load()
dev = torch.device('cuda:0')
torch.manual_seed(0)
def go(M, N, K, bn, reps=200):
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    npad = (N + bn - 1) // bn * bn
    w = (torch.randn(npad, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(npad, device=dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    def run():
        L.check(lib.ppasr_b200_op_linear(L.ptr(a), K, L.ptr(w), npad, L.ptr(bias), L.ptr(out), N, M, N, K, 0, 0, 1.0, 0, None, 0, bn,
                                         L.stream_ptr()))
    for _ in range(5): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print("M %d N %d K %d bn %d: %.2f us  %.1f TFLOP/s" % (M, N, K, bn, us, 2.0 * M * N * K / us / 1e6))
go(7936, 768, 256, 128)
go(7936, 768, 256, 256)
go(7936, 256, 256, 128)
go(7936, 256, 256, 256)
go(7936, 2048, 256, 256)
go(7936, 768, 1024, 128)
go(7936, 4352, 256, 128)
