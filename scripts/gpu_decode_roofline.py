"""HBM roofline of the stand-alone decode kernels on materialised posteriors (BASELINE configs[1] and [2] sizes)."""
import json, sys, torch
sys.path.insert(0, '.')
from ppasr_b200 import _lib as L
lib = L.load(); dev = torch.device('cuda:0'); torch.manual_seed(0)
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if __import__('os').path.exists('MEASURED_PEAKS.json') else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / reps
out = {}
for name, (B, T, V, beam, temp) in {"C2 [32,248,4233]": (32, 248, 4233, 10, 8.0), "C3 [64,748,4233]": (64, 748, 4233, 10, 8.0),
                                    "C3 flat posteriors": (64, 748, 4233, 10, 1.0)}.items():
    probs = torch.softmax(torch.randn(B, T, V, device=dev) * temp, -1).contiguous()
    nbytes = probs.numel() * 4
    ids = torch.zeros(B, T, dtype=torch.int32, device=dev); ol = torch.zeros(B, dtype=torch.int32, device=dev)
    sc = torch.zeros(B, device=dev); ti = torch.zeros(B * T, dtype=torch.int32, device=dev); tm = torch.zeros(B * T, device=dev)
    ms_g = timeit(lambda: L.check(lib.ppasr_b200_greedy_decode(L.ptr(probs), B, T, V, None, 0, L.ptr(ids), T, L.ptr(ol), L.ptr(sc), L.ptr(ti), L.ptr(tm), L.stream_ptr())))
    ws = torch.empty(lib.ppasr_b200_beam_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
    ms_p = timeit(lambda: L.check(lib.ppasr_b200_op_ctc_prune(L.ptr(probs), B * T, V, 0.99, 40, L.ptr(ws), L.stream_ptr())))
    st = torch.empty(lib.ppasr_b200_beam_state_bytes(B, T, beam), dtype=torch.uint8, device=dev)
    def beam_fn():
        L.check(lib.ppasr_b200_beam_reset(L.ptr(st), B, T, beam, L.stream_ptr()))
        L.check(lib.ppasr_b200_beam_advance(L.ptr(probs), B, T, V, None, beam, 0.99, 40, 0, L.ptr(st), T, L.ptr(ws), L.stream_ptr()))
    ms_b = timeit(beam_fn, reps=5)
    st300 = torch.empty(lib.ppasr_b200_beam_state_bytes(B, T, 300), dtype=torch.uint8, device=dev)
    def beam300_fn():
        L.check(lib.ppasr_b200_beam_reset(L.ptr(st300), B, T, 300, L.stream_ptr()))
        L.check(lib.ppasr_b200_beam_advance(L.ptr(probs), B, T, V, None, 300, 0.99, 40, 0, L.ptr(st300), T, L.ptr(ws), L.stream_ptr()))
    ms_b300 = timeit(beam300_fn, reps=2)
    del st300
    r = {"beam300_total_ms": ms_b300, "bytes": nbytes, "greedy_ms": ms_g, "greedy_gbs": nbytes / ms_g / 1e6, "greedy_frac": nbytes / ms_g / 1e6 / peak,
         "prune_ms": ms_p, "prune_gbs": nbytes / ms_p / 1e6, "prune_frac": nbytes / ms_p / 1e6 / peak,
         "beam10_total_ms": ms_b, "beam10_utt_per_s": B / ms_b * 1e3}
    out[name] = r
    print(name, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
json.dump({"hbm_peak_gbs": peak, "cases": out}, open('gpurun_out/decode_roofline_r2.json', 'w'), indent=1)
