"""GPU bring-up check for the tcgen05 GEMM: compares against torch.matmul (fp32 on bf16 inputs)."""
import sys, time
import torch
sys.path.insert(0, '.')
from ppasr_b200 import _lib as L

lib = L.load()
dev = torch.device('cuda:0')
torch.manual_seed(0)

def run_linear(M, N, K, epilogue, act=0, alpha=1.0, residual=0, block_n=128, lens=None, T=0, reps=0):
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    npad = (N + block_n - 1) // block_n * block_n
    w = torch.zeros(npad, K, device=dev, dtype=torch.bfloat16)
    w[:N] = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.zeros(npad, device=dev)
    bias[:N] = torch.randn(N, device=dev)
    ref = a.float() @ w[:N].float().t() + bias[:N]
    if epilogue == 0:
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        if act == 1: ref = torch.relu(ref)
        if act == 2: ref = ref * torch.sigmoid(ref)
        ldo = N
    elif epilogue == 1:
        out = torch.randn(M, N, device=dev)
        x0 = out.clone()
        ref = (x0 if residual else 0) + alpha * ref
        if lens is not None:
            rows = torch.arange(M, device=dev)
            b = rows // T; t = rows % T
            masked = t >= lens[b]
            ref = torch.where(masked[:, None], x0 if residual else torch.zeros_like(x0), ref)
        ldo = N
    elif epilogue == 2:
        out = torch.zeros(M, N // 2, device=dev, dtype=torch.bfloat16)
        ref = ref[:, 0::2] * torch.sigmoid(ref[:, 1::2])
        ldo = N // 2
    elif epilogue == 3:
        ldo = (N + 3) // 4 * 4
        out = torch.zeros(M, ldo, device=dev)
    st = L.stream_ptr()
    def call():
        L.check(lib.ppasr_b200_op_linear(L.ptr(a), K, L.ptr(w), npad, L.ptr(bias), L.ptr(out), ldo, M, N, K,
                                         epilogue, act, alpha, residual, L.ptr(lens), T, block_n, st))
    call()
    torch.cuda.synchronize()
    got = out.float()
    if epilogue == 3: got = got[:, :N]
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    ms = None
    if reps:
        if epilogue == 1: residual_backup = None
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        for _ in range(3): call()
        e0.record()
        for _ in range(reps): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    tf = (2.0 * M * N * K / (ms * 1e-3) / 1e12) if ms else None
    print(f"linear M={M} N={N} K={K} epi={epilogue} act={act} bn={block_n}: max_err={err:.4g} (scale {scale:.3g})"
          + (f"  {ms*1e3:.1f} us  {tf:.1f} TFLOP/s" if ms else ""), flush=True)
    return err / max(scale, 1e-6)

ok = True
cases = [
    (128, 128, 64, 0, 0, 128), (128, 256, 64, 0, 0, 256), (128, 128, 256, 0, 0, 128),
    (300, 256, 256, 0, 1, 128), (300, 512, 128, 0, 2, 256), (1000, 768, 256, 0, 0, 256),
    (7936, 2048, 256, 0, 2, 256), (7936, 2048, 256, 0, 2, 128),
]
for (M, N, K, epi, act, bn) in cases:
    r = run_linear(M, N, K, epi, act=act, block_n=bn, reps=20 if M > 5000 else 0)
    ok &= r < 2e-2
r = run_linear(7936, 256, 2048, 1, alpha=0.5, residual=1, block_n=128, reps=20); ok &= r < 1e-2
r = run_linear(7936, 256, 2048, 1, alpha=0.5, residual=1, block_n=256, reps=20); ok &= r < 1e-2
r = run_linear(7936, 256, 4864, 1, alpha=16.0, residual=0, block_n=128, reps=20); ok &= r < 1e-2
lens = torch.tensor([100, 248, 7, 200], device=dev, dtype=torch.int32)
r = run_linear(4 * 248, 256, 256, 1, alpha=1.0, residual=1, block_n=128, lens=lens, T=248); ok &= r < 1e-2
r = run_linear(7936, 512, 256, 2, block_n=128, reps=20); ok &= r < 2e-2
r = run_linear(7936, 512, 256, 2, block_n=256, reps=20); ok &= r < 2e-2
r = run_linear(7936, 4233, 256, 3, block_n=128, reps=10); ok &= r < 1e-2
r = run_linear(7936, 4233, 256, 3, block_n=256, reps=10); ok &= r < 1e-2
r = run_linear(16, 768, 256, 0, 0, block_n=256); ok &= r < 2e-2
r = run_linear(47872, 2048, 256, 0, 2, block_n=256, reps=10); ok &= r < 2e-2
r = run_linear(47872, 256, 2048, 1, alpha=0.5, residual=1, block_n=128, reps=10); ok &= r < 1e-2
print("GEMM_CHECK", "PASS" if ok else "FAIL")
