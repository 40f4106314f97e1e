"""ctc_prune_kernel alone on the C2 / C3 posteriors (peaked and flat): time and HBM fraction for the current
PPASR_B200_PRUNE_CTAS (resident CTAs per SM)."""
import json, os, sys, torch
sys.path.insert(0, '.')
from ppasr_b200 import _lib as L
lib = L.load(); dev = torch.device('cuda:0'); torch.manual_seed(0)
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for name, (B, T, V, temp) in {"C2": (32, 248, 4233, 8.0), "C3": (64, 748, 4233, 8.0), "C3 temp3": (64, 748, 4233, 3.0), "C3 flat": (64, 748, 4233, 1.0)}.items():
    probs = torch.softmax(torch.randn(B, T, V, device=dev) * temp, -1).contiguous()
    ws = torch.empty(lib.ppasr_b200_beam_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
    fn = lambda: L.check(lib.ppasr_b200_op_ctc_prune(L.ptr(probs), B * T, V, 0.99, 40, L.ptr(ws), L.stream_ptr()))
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    ms = tot / 10
    nbytes = probs.numel() * 4
    out[name] = {"ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / peak, 3)}
print("PRUNE_CTAS", os.environ.get("PPASR_B200_PRUNE_CTAS", "default(4)"), out, flush=True)
