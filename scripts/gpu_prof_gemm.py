"""Launch a few representative GEMMs for ncu capture."""
import sys
import torch
sys.path.insert(0, '.')
from ppasr_b200 import _lib as L
lib = L.load()
dev = torch.device('cuda:0')
torch.manual_seed(0)
def go(M, N, K, epilogue, act, bn, alpha=1.0, residual=0):
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    npad = (N + bn - 1) // bn * bn
    w = (torch.randn(npad, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(npad, device=dev)
    if epilogue == 0: out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16); ldo = N
    elif epilogue == 1: out = torch.randn(M, N, device=dev); ldo = N
    elif epilogue == 2: out = torch.zeros(M, N // 2, device=dev, dtype=torch.bfloat16); ldo = N // 2
    else: ldo = (N + 3) // 4 * 4; out = torch.zeros(M, ldo, device=dev)
    for _ in range(3):
        L.check(lib.ppasr_b200_op_linear(L.ptr(a), K, L.ptr(w), npad, L.ptr(bias), L.ptr(out), ldo, M, N, K,
                                         epilogue, act, alpha, residual, None, 0, bn, L.stream_ptr()))
    torch.cuda.synchronize()
go(7936, 2048, 256, 0, 2, 256)
go(7936, 256, 2048, 1, 0, 128, 0.5, 1)
