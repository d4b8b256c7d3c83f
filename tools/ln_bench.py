"""Cold-cache timings of the K5 (gate+residual+LN) and rotary / GELU kernels alone (dev tool).
Same method as unimp_b200/kbench.py."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unimp_b200 import ops
from unimp_b200.kbench import _time_graph, _k

dev, dt = "cuda", torch.bfloat16
# backward launches captured in a graph must not touch the legacy stream: run everything on a side stream
torch.cuda.set_stream(torch.cuda.Stream())
rows, D = int(os.environ.get("ROWS", 1536)), int(os.environ.get("D", 2560))
es = 2
res = {}
K = _k(4 * rows * D * es, cap=16)
xs = [torch.randn(rows, D, device=dev, dtype=dt, requires_grad=True) for _ in range(K)]
brs = [torch.randn(rows, D, device=dev, dtype=dt, requires_grad=True) for _ in range(K)]
gate = torch.full((1,), 0.5, device=dev, dtype=dt, requires_grad=True)
gam = torch.ones(D, device=dev, dtype=dt, requires_grad=True)
bet = torch.zeros(D, device=dev, dtype=dt, requires_grad=True)
g1s = [torch.randn(rows, D, device=dev, dtype=dt) for _ in range(K)]
res["ln_fwd_gated"] = _time_graph([lambda x=x, b=b: ops.gate_residual_ln(b, x, gate, gam, bet) for x, b in zip(xs, brs)])
outs = [ops.gate_residual_ln(b, x, gate, gam, bet) for x, b in zip(xs, brs)]
res["ln_bwd_gated_cols(+reduce)"] = _time_graph([lambda o=o, x=x, b=b, g=g: torch.autograd.grad(
    o, (b, x, gate, gam, bet), (g, g), retain_graph=True) for o, x, b, g in zip(outs, xs, brs, g1s)])
outs = [ops.gate_residual_ln(b, x, None, gam, bet) for x, b in zip(xs, brs)]
res["ln_bwd_ungated_cols(+reduce)"] = _time_graph([lambda o=o, x=x, b=b, g=g: torch.autograd.grad(
    o, (b, x, gam, bet), (g, g), retain_graph=True) for o, x, b, g in zip(outs, xs, brs, g1s)])
outs = [ops.layer_norm(x, gam, bet) for x in xs]
res["ln_only_bwd_cols(+reduce)"] = _time_graph([lambda o=o, x=x, g=g: torch.autograd.grad(
    o, (x, gam, bet), g, retain_graph=True) for o, x, g in zip(outs, xs, g1s)])
gf, bf = gam.detach(), bet.detach()
outs = [ops.gate_residual_ln(b, x, None, gf, bf) for x, b in zip(xs, brs)]
res["ln_bwd_ungated_frozen"] = _time_graph([lambda o=o, x=x, b=b, g=g: torch.autograd.grad(
    o, (b, x), (g, g), retain_graph=True) for o, x, b, g in zip(outs, xs, brs, g1s)])
del xs, brs, outs, g1s
# rotary at the 4B LM shape
B, T, H, dh = rows // 256 if rows % 256 == 0 else 1, 256 if rows % 256 == 0 else rows, 32, 80
qkvs = [torch.randn(B, T, 3 * H * dh, device=dev, dtype=dt, requires_grad=True) for _ in range(6)]
pos = torch.arange(T, device=dev, dtype=torch.float32)
inv = 1.0 / (10000 ** (torch.arange(0, dh, 2, device=dev, dtype=torch.float32) / dh))
fr = torch.outer(pos, inv)
emb = torch.cat([fr, fr], -1)[None]
cos, sin = emb.cos().to(dt), emb.sin().to(dt)
res["rotary_fwd"] = _time_graph([lambda q=q: ops.rotary_qkv(q, cos, sin, heads=H, head_dim=dh, rotary_dim=dh) for q in qkvs])
outs = [ops.rotary_qkv(q, cos, sin, heads=H, head_dim=dh, rotary_dim=dh) for q in qkvs]
gs = [tuple(torch.randn(B, H, T, dh, device=dev, dtype=dt) for _ in range(3)) for _ in qkvs]
res["rotary_bwd"] = _time_graph([lambda o=o, q=q, g=g: torch.autograd.grad(o, q, g, retain_graph=True)
                                 for o, q, g in zip(outs, qkvs, gs)])
print("LNBENCH " + json.dumps({k: round(v, 2) for k, v in res.items()}), flush=True)
