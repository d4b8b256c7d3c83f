"""Runs each hand-written kernel a few times at the 4B workload's shapes (for ncu captures and
isolated CUDA-event timings with an L2 flush between launches).  Dev tool, not a bench."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unimp_b200 import ops

dev = "cuda"
bf = torch.bfloat16
torch.manual_seed(0)
B, T, Ti, n, H, dh, D, V = (int(os.environ.get(k, d)) for k, d in
                            (("B", 3), ("T", 256), ("TI", 2), ("N", 64), ("H", 8), ("DH", 64), ("D", 2560), ("V", 74053)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
reps = int(os.environ.get("REPS", 5))

def timeit(name, fn, alg_bytes=None, flops=None):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    us = sorted(ts)[len(ts) // 2]
    d = {"kernel": name, "us": round(us, 2)}
    if alg_bytes: d["GB/s"] = round(alg_bytes / us / 1e3, 1)
    if flops: d["TFLOP/s"] = round(flops / us / 1e6, 2)
    print(json.dumps(d))

# K1 masked cross attention
inner = H * dh
q = torch.randn(B, T, inner, device=dev, dtype=bf, requires_grad=True)
kv = torch.randn(B, Ti * n, 2 * inner, device=dev, dtype=bf, requires_grad=True)
loc = torch.zeros(B, T, dtype=torch.bool); 
for b in range(B):
    for j in range(Ti): loc[b, 1 + j * (T // Ti)] = True
tt = loc.cumsum(-1).to(torch.int32).to(dev)
go = torch.randn(B, T, inner, device=dev, dtype=bf)
xbytes = 2 * (2 * B * T * inner + 2 * B * Ti * n * inner) + 4 * B * T * H
xflops = 4.0 * H * dh * n * B * T
timeit("xattn_fwd", lambda: ops.masked_cross_attention(q, kv, tt, heads=H, n_latents=n, scale=0.125), xbytes, xflops)
o = ops.masked_cross_attention(q, kv, tt, heads=H, n_latents=n, scale=0.125)
timeit("xattn_bwd", lambda: torch.autograd.grad(o, (q, kv), go, retain_graph=True), 2.5 * xbytes, 2.5 * xflops)
# K3 ViT self attention (B*Ti images, 16 heads, 257 tokens)
N, L, Hv = B * Ti, 257, 16
qkv = torch.randn(N, L, 3 * Hv * dh, device=dev, dtype=bf)
vb = 2 * 4 * N * L * Hv * dh + 4 * N * L * Hv
vf = 4.0 * N * Hv * L * L * dh
timeit("vit_attn_fwd", lambda: ops.attention(qkv[..., :Hv * dh], qkv[..., Hv * dh:], heads=Hv, scale=0.125), vb, vf)
# K2 perceiver attention (64 x 320)
pq = torch.randn(N, 64, inner, device=dev, dtype=bf, requires_grad=True)
pkv = torch.randn(N, 320, 2 * inner, device=dev, dtype=bf, requires_grad=True)
pb_ = 2 * (2 * N * 64 * inner + 2 * N * 320 * inner) + 4 * N * 64 * H
pf = 4.0 * N * H * 64 * 320 * dh
timeit("perceiver_attn_fwd", lambda: ops.attention(pq, pkv, heads=H, scale=0.125), pb_, pf)
po = ops.attention(pq, pkv, heads=H, scale=0.125)
pg = torch.randn_like(po)
timeit("perceiver_attn_bwd", lambda: torch.autograd.grad(po, (pq, pkv), pg, retain_graph=True), 2.5 * pb_, 2.5 * pf)
# K5
rows = B * T
x = torch.randn(rows, D, device=dev, dtype=bf, requires_grad=True)
br = torch.randn(rows, D, device=dev, dtype=bf, requires_grad=True)
gate = torch.full((1,), 0.5, device=dev, dtype=bf, requires_grad=True)
gam = torch.ones(D, device=dev, dtype=bf, requires_grad=True)
bet = torch.zeros(D, device=dev, dtype=bf, requires_grad=True)
timeit("gate_residual_ln_fwd", lambda: ops.gate_residual_ln(br, x, gate, gam, bet), 4 * rows * D * 2)
xo, ln = ops.gate_residual_ln(br, x, gate, gam, bet)
g1, g2 = torch.randn_like(xo), torch.randn_like(ln)
timeit("gate_residual_ln_bwd", lambda: torch.autograd.grad((xo, ln), (br, x, gate, gam, bet), (g1, g2), retain_graph=True), 6 * rows * D * 2)
# the LM towers' variant: ungated residual, frozen LayerNorm
xo2, ln2 = ops.gate_residual_ln(br, x, None, gam.detach(), bet.detach())
timeit("residual_ln_bwd_frozen", lambda: torch.autograd.grad((xo2, ln2), (br, x), (g1, g2), retain_graph=True), 4 * rows * D * 2)
# exact GELU of the FeedForward blocks
h = torch.randn(rows, 4 * D, device=dev, dtype=bf, requires_grad=True)
timeit("gelu_fwd", lambda: ops.gelu(h), 2 * rows * 4 * D * 2)
hy = ops.gelu(h)
hg = torch.randn_like(hy)
timeit("gelu_bwd", lambda: torch.autograd.grad(hy, h, hg, retain_graph=True), 3 * rows * 4 * D * 2)
del h, hy, hg
# K6 (img_gen-like: many valid rows) and rec-like (few valid rows)
for nm, Tl, nvalid in (("focal_ce_rec", T, 4), ("focal_ce_imggen", 1024, 257)):
    z = torch.randn(B, Tl, V, device=dev, dtype=bf, requires_grad=True)
    y = torch.full((B, Tl), -100, device=dev, dtype=torch.int64)
    y[:, Tl - nvalid:] = torch.randint(0, V, (B, nvalid), device=dev)
    w = torch.ones(B, device=dev)
    timeit(nm + "_fwd", lambda: ops.focal_ce(z, y, w), B * nvalid * V * 2)
    l = ops.focal_ce(z, y, w)
    timeit(nm + "_bwd", lambda: torch.autograd.grad(l, z, retain_graph=True), B * nvalid * V * 2 + B * Tl * V * 2)
    del z

# kernel-level durations (CUDA-event timing above includes Python/autograd launch latency for
# the tiny kernels): torch profiler table of our kernels over one more round of calls
from torch.profiler import profile, ProfilerActivity
z = torch.randn(B, T, V, device=dev, dtype=bf, requires_grad=True)
y = torch.full((B, T), -100, device=dev, dtype=torch.int64); y[:, T - 4:] = 7
w = torch.ones(B, device=dev)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        flush.zero_()
        o = ops.masked_cross_attention(q, kv, tt, heads=H, n_latents=n, scale=0.125)
        torch.autograd.grad(o, (q, kv), go)
        ops.attention(qkv[..., :Hv * dh], qkv[..., Hv * dh:], heads=Hv, scale=0.125)
        po = ops.attention(pq, pkv, heads=H, scale=0.125)
        torch.autograd.grad(po, (pq, pkv), pg)
        xo, ln = ops.gate_residual_ln(br, x, gate, gam, bet)
        torch.autograd.grad((xo, ln), (br, x, gate, gam, bet), (g1, g2))
        l = ops.focal_ce(z, y, w)
        torch.autograd.grad(l, z)
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if "unimp" in e.key:
        print(json.dumps({"kernel": e.key[:90], "calls": e.count, "avg_us": round(e.device_time_total / e.count, 2)}))
