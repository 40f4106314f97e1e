"""GPU bring-up check: every op vs a torch reference, then the whole encoder vs the CPU oracle."""
import sys, time, math
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from ppasr_b200 import _lib as L
from ppasr_b200.engine import ConformerEngine, out_frames
from ppasr_b200.weights import ConformerConfig, init_conformer_weights, synthetic_fbank
from oracle.conformer_oracle import ConformerOracle, ConformerConf
from oracle import decoders_oracle as DO

lib = L.load()
dev = torch.device('cuda:0')
torch.manual_seed(0)
ok = True
def report(name, got, ref, tol):
    global ok
    err = (got.float() - ref.float()).abs().max().item()
    sc = ref.float().abs().max().item()
    good = err <= tol * max(sc, 1e-6)
    ok &= good
    print(f"{name}: max_err={err:.4g} scale={sc:.4g} rel={err/max(sc,1e-9):.3g} {'OK' if good else 'FAIL'}", flush=True)

# ---------------- layernorm ----------------
M, D = 1000, 256
x = torch.randn(M, D, device=dev) * 3 + 0.5
g1, b1, g2, b2 = [torch.randn(D, device=dev) for _ in range(4)]
y = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
xs = x.clone()
L.check(lib.ppasr_b200_op_layernorm(L.ptr(xs), L.ptr(y), L.ptr(g1), L.ptr(b1), None, None, None, 0, M, D, 1e-5, L.stream_ptr()))
report("layernorm single", y, F.layer_norm(x, (D,), g1, b1, 1e-5), 1e-2)
xs = x.clone()
L.check(lib.ppasr_b200_op_layernorm(L.ptr(xs), L.ptr(y), L.ptr(g1), L.ptr(b1), L.ptr(g2), L.ptr(b2), None, 0, M, D, 1e-5, L.stream_ptr()))
r1 = F.layer_norm(x, (D,), g1, b1, 1e-5)
report("layernorm double x", xs, r1, 1e-5)
report("layernorm double y", y, F.layer_norm(r1, (D,), g2, b2, 1e-5), 1e-2)
lens = torch.tensor([100, 250, 3, 249], device=dev, dtype=torch.int32)
xs = x.clone()
L.check(lib.ppasr_b200_op_layernorm(L.ptr(xs), L.ptr(y), L.ptr(g1), L.ptr(b1), None, None, L.ptr(lens), 250, M, D, 1e-5, L.stream_ptr()))
ref = F.layer_norm(x, (D,), g1, b1, 1e-5).view(4, 250, D).clone()
for b in range(4): ref[b, lens[b]:] = 0
report("layernorm masked", y, ref.view(M, D), 1e-2)

# ---------------- dwconv + LN + swish ----------------
for (K, causal) in [(15, True), (15, False), (31, False), (7, True)]:
    B, T, C = 3, 77, 256
    g = torch.randn(B, T, C, device=dev).to(torch.bfloat16)
    w = torch.randn(C, K, device=dev) / K ** 0.5
    bias = torch.randn(C, device=dev) * 0.1
    gam, bet = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    pad = torch.randn(C, device=dev).to(torch.bfloat16).float()
    out = torch.zeros(B, T, C, device=dev, dtype=torch.bfloat16)
    lpad = K - 1 if causal else (K - 1) // 2
    L.check(lib.ppasr_b200_op_dwconv(L.ptr(g), L.ptr(w), L.ptr(bias), L.ptr(pad) if causal else None, L.ptr(gam), L.ptr(bet), 1,
                                     L.ptr(out), B, T, T, C, K, lpad, 1e-5, L.stream_ptr()))
    gi = g.float().transpose(1, 2)
    if causal:
        gi = torch.cat([pad.view(1, C, 1).expand(B, C, K - 1), gi], 2)
        cv = F.conv1d(gi, w.view(C, 1, K), bias, groups=C)
    else:
        cv = F.conv1d(gi, w.view(C, 1, K), bias, groups=C, padding=(K - 1) // 2)
    r = F.layer_norm(cv.transpose(1, 2), (C,), gam, bet, 1e-5)
    r = r * torch.sigmoid(r)
    report(f"dwconv K={K} causal={causal}", out, r, 1e-2)
# valid mode (chunk): Tin = lorder + Tout
B, Tout, C, K = 2, 16, 256, 15
g = torch.randn(B, Tout + K - 1, C, device=dev).to(torch.bfloat16)
out = torch.zeros(B, Tout, C, device=dev, dtype=torch.bfloat16)
L.check(lib.ppasr_b200_op_dwconv(L.ptr(g), L.ptr(w[:, :15].contiguous() if w.shape[1] >= 15 else w, ), L.ptr(bias), None, L.ptr(gam), L.ptr(bet), 1,
                                 L.ptr(out), B, Tout + K - 1, Tout, C, K, 0, 1e-5, L.stream_ptr())) if False else None

# ---------------- softmax ----------------
M, V = 500, 4233
ld = (V + 3) // 4 * 4
lg = torch.randn(M, ld, device=dev) * 4
pr = torch.zeros(M, V, device=dev)
L.check(lib.ppasr_b200_op_softmax(L.ptr(lg), ld, L.ptr(pr), M, V, L.stream_ptr()))
report("softmax", pr, torch.softmax(lg[:, :V], -1), 1e-5)

# ---------------- attention ----------------
def attn_ref(q, k, v, p, u, vb, klens):
    # q,k,v [B,H,T,64] fp32 ; p [H,T2,64]
    s = ((q + u[None, :, None, :]) @ k.transpose(-1, -2) + (q + vb[None, :, None, :]) @ p[None].transpose(-1, -2)) / 8.0
    T2 = k.shape[2]
    mask = torch.arange(T2, device=q.device)[None, :] >= klens[:, None]
    s = s.masked_fill(mask[:, None, None, :], float('-inf'))
    a = torch.softmax(s, -1).masked_fill(mask[:, None, None, :], 0.0)
    return a @ v
for (B, H, T1, T2) in [(2, 4, 248, 248), (3, 4, 100, 100), (1, 4, 16, 80), (2, 4, 300, 300)]:
    q = torch.randn(B, H, T1, 64, device=dev); k = torch.randn(B, H, T2, 64, device=dev); v = torch.randn(B, H, T2, 64, device=dev)
    pos_rows = 400; Lc = 2; l = 1; row0 = 7
    pos = (torch.randn(pos_rows, Lc * H * 64, device=dev)).to(torch.bfloat16)
    u = torch.randn(H, 64, device=dev) * 0.3; vb = torch.randn(H, 64, device=dev) * 0.3
    klens = torch.randint(max(1, T2 // 2), T2 + 1, (B,), device=dev, dtype=torch.int32); klens[0] = T2
    qb = q.to(torch.bfloat16)
    q2 = torch.cat([(q + u[None, :, None, :]), (q + vb[None, :, None, :])], -1).to(torch.bfloat16).contiguous()
    kb = k.to(torch.bfloat16).contiguous()
    T2p = (T2 + 63) // 64 * 64
    vt = torch.zeros(B, H, 64, T2p, device=dev, dtype=torch.bfloat16)
    vt[..., :T2] = v.to(torch.bfloat16).transpose(-1, -2)
    out = torch.zeros(B * T1, H * 64, device=dev, dtype=torch.bfloat16)
    L.check(lib.ppasr_b200_op_attention(L.ptr(q2), L.ptr(kb), L.ptr(vt), T2p, L.ptr(pos), pos_rows, Lc * H * 64, row0, l * H * 64,
                                        L.ptr(out), B, H, T1, T2, L.ptr(klens), L.stream_ptr()))
    torch.cuda.synchronize()
    p = pos[row0:row0 + T2, l * H * 64:(l + 1) * H * 64].float().view(T2, H, 64).transpose(0, 1)
    q2f = q2.float()
    s_ref_q_u = q2f[..., :64]; s_ref_q_v = q2f[..., 64:]
    s = (s_ref_q_u @ kb.float().transpose(-1, -2) + s_ref_q_v @ p[None].transpose(-1, -2)) / 8.0
    mask = torch.arange(T2, device=dev)[None, :] >= klens[:, None]
    s = s.masked_fill(mask[:, None, None, :], float('-inf'))
    a = torch.softmax(s, -1).masked_fill(mask[:, None, None, :], 0.0)
    r = (a @ v.to(torch.bfloat16).float()).transpose(1, 2).reshape(B * T1, H * 64)
    report(f"attention B={B} T1={T1} T2={T2}", out, r, 2e-2)

# ---------------- standalone greedy vs reference restatement ----------------
B, T, V = 4, 50, 4233
probs = torch.softmax(torch.randn(B, T, V, device=dev) * 3, -1)
probs[0, 5:9] = probs[0, 5:6]  # repeated frames
probs[1, :, 0] += 0.5          # many blanks
probs = probs.contiguous()
ids = torch.zeros(B, T, dtype=torch.int32, device=dev); ol = torch.zeros(B, dtype=torch.int32, device=dev)
sc = torch.zeros(B, device=dev); ti = torch.zeros(B * T, dtype=torch.int32, device=dev); tm = torch.zeros(B * T, device=dev)
L.check(lib.ppasr_b200_greedy_decode(L.ptr(probs), B, T, V, None, 0, L.ptr(ids), T, L.ptr(ol), L.ptr(sc), L.ptr(ti), L.ptr(tm), L.stream_ptr()))
torch.cuda.synchronize()
pn = probs.cpu().numpy()
vocab = [str(i) + ',' for i in range(V)]
good = True
for b in range(B):
    mi, coll, mp = DO.greedy_ids(pn[b])
    score_ref, _ = DO.greedy_decoder(pn[b], vocab)
    got = ids[b, :ol[b]].cpu().tolist()
    s_got = float(sc[b].item()) * 100.0
    if got != coll or s_got != score_ref:
        good = False
        print("greedy mismatch", b, got[:10], coll[:10], s_got, score_ref)
ok &= good
print("standalone greedy bit-exact:", good, flush=True)

# ---------------- whole encoder vs oracle ----------------
def run_model(num_blocks, B, T, lens, vocab=4233, streaming=True, norm="layer_norm", label=""):
    global ok
    cfg = ConformerConfig(num_blocks=num_blocks, vocab_size=vocab, streaming=streaming, cnn_module_norm=norm)
    w = init_conformer_weights(cfg)
    feats = synthetic_fbank(B, T)
    for b in range(B): feats[b, lens[b]:] = 0
    eng = ConformerEngine(cfg, w)
    t0 = time.time()
    eng.encode(torch.from_numpy(feats).to(dev), lens)
    logits = eng.ctc_logits()
    probs = eng.ctc_probs()
    ids, ol, sc, fi, fp = eng.ctc_greedy(to_host=True, with_frames=True)
    torch.cuda.synchronize()
    orc = ConformerOracle(ConformerConf(**cfg.to_dict()), w)
    t1 = time.time()
    ref_logits = orc.get_encoder_out(torch.from_numpy(feats), torch.tensor(lens), return_logits=True)
    ref_probs = torch.softmax(ref_logits, -1)
    t2 = time.time()
    Tp = out_frames(T)
    vl = [min(Tp, (l + 3) // 4) for l in lens]
    lg = logits.cpu()
    # compare only valid frames (pad frames are not meaningful), report both
    errs = []
    for b in range(B):
        e = (lg[b, :vl[b]] - ref_logits[b, :vl[b]]).abs().max().item()
        errs.append(e)
    scale = ref_logits.abs().max().item()
    rel = max(errs) / scale
    epad = (lg - ref_logits).abs().max().item() / scale
    fid_ref = ref_probs.argmax(-1)
    fid = torch.from_numpy(fi)
    agree = sum((fid[b, :vl[b]] == fid_ref[b, :vl[b]]).sum().item() for b in range(B)) / sum(vl)
    # margin-filtered agreement
    top2 = ref_logits.topk(2, -1).values
    margin = (top2[..., 0] - top2[..., 1])
    big = margin > 0.5
    m_agree = all(((fid[b, :vl[b]] == fid_ref[b, :vl[b]]) | ~big[b, :vl[b]]).all().item() for b in range(B))
    pe = (probs.cpu() - ref_probs).abs().max().item()
    good = rel < 1e-2 and m_agree
    ok &= good
    print(f"model{label} L={num_blocks} B={B} T={T} streaming={streaming} norm={norm}: logits rel_err(valid)={rel:.3g} (incl pad {epad:.3g}) "
          f"probs max_abs_err={pe:.3g} argmax agree={agree:.4f} margin>0.5 agree={m_agree} "
          f"gpu {t1-t0:.2f}s oracle {t2-t1:.2f}s {'OK' if good else 'FAIL'}", flush=True)
    # fused greedy vs greedy on our own probs
    pn = probs.cpu().numpy()
    g2 = True
    for b in range(B):
        mi, coll, mp = DO.greedy_ids(pn[b])
        if ids[b, :ol[b]].tolist() != coll:
            g2 = False
    print("  fused greedy ids == reference greedy on our probs:", g2, flush=True)
    ok &= g2
    eng.close()
    return eng

run_model(1, 2, 131, [131, 90], vocab=97, label="-tiny")
run_model(2, 3, 400, [400, 333, 250])
run_model(2, 2, 300, [300, 200], streaming=False, label="-nonstream")
run_model(2, 2, 300, [300, 200], streaming=False, norm="batch_norm", label="-bn")
run_model(12, 4, 998, [998, 998, 900, 500])

# ---------------- timing at the C2 shape ----------------
cfg = ConformerConfig()
w = init_conformer_weights(cfg)
eng = ConformerEngine(cfg, w)
feats = torch.from_numpy(synthetic_fbank(32, 998)).to(dev)
for _ in range(3):
    eng.encode(feats); eng.ctc_greedy(to_host=False)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.encode(feats); eng.ctc_greedy(to_host=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"C2 b32x10s encode+fused greedy: {ms:.3f} ms/step -> {32/ms*1e3:.0f} utt/s, RTF {ms*1e-3/320:.2e}, {741.9/ms:.0f} TFLOP/s", flush=True)
print("MODEL_CHECK", "PASS" if ok else "FAIL")
