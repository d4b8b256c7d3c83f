"""One cold launch of each decode-step kernel at configs[3] shapes (dev tool: the target of an
`ncu --set full` capture; 6 projections, the LM attention step, the cross-attention step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unimp_b200 import ops

torch.cuda.set_device(0)
dev, bf = "cuda", torch.bfloat16
M, D, H, dh, Tmax, cur = 5, 2560, 32, 80, 768, 640
with torch.no_grad():
    for N, K in ((7680, 2560), (2560, 2560), (10240, 2560), (2560, 10240), (512, 2560), (2560, 512)):
        w = torch.randn(N, K, device=dev, dtype=bf)
        x = torch.randn(M, 1, K, device=dev, dtype=bf)
        b = torch.randn(N, device=dev, dtype=bf)
        torch.cuda.synchronize()
        ops.linear_rows(x, w, b, act_gelu=(N == 10240))
    kc, vc = torch.randn(M, H, Tmax, dh, device=dev, dtype=bf), torch.randn(M, H, Tmax, dh, device=dev, dtype=bf)
    qkv = torch.randn(M, 1, 3 * D, device=dev, dtype=bf)
    cos, sin = torch.randn(M, 1, dh, device=dev, dtype=bf), torch.randn(M, 1, dh, device=dev, dtype=bf)
    indir = torch.randint(0, M, (M, Tmax), device=dev, dtype=torch.int32)
    mask = torch.zeros(M, Tmax, device=dev, dtype=bf)
    mask[:, cur + 1:] = float("-inf")
    torch.cuda.synchronize()
    ops.lm_decode_attention(qkv, cos, sin, kc, vc, indir, mask, torch.tensor([cur], device=dev), heads=H, head_dim=dh,
                            rotary_dim=dh, scale=dh ** -0.5)
    q1 = torch.randn(M, 1, 512, device=dev, dtype=bf)
    kv = torch.randn(M, 5 * 64, 1024, device=dev, dtype=bf)
    torch.cuda.synchronize()
    ops.xattn_decode(q1, kv, torch.full((M,), 5, device=dev, dtype=torch.int32), heads=8, n_latents=64, scale=0.125)
    torch.cuda.synchronize()
    lp, idx = ops.beam_topk(torch.randn(M, 74053, device=dev), torch.zeros(1, M, device=dev), M, 10)
    torch.cuda.synchronize()
    # ragged shapes: N not a multiple of the row tile, K = 32, one beam
    ops.linear_rows(torch.randn(1, 1, 32, device=dev, dtype=bf), torch.randn(1005, 32, device=dev, dtype=bf))
    ops.linear_rows(torch.randn(3, 1, 10240, device=dev, dtype=bf), torch.randn(40, 10240, device=dev, dtype=bf))
    torch.cuda.synchronize()
print("ok")
