"""Per-shape timing of unimp_linear_small_m against cuBLAS (F.linear) on cold weights: NB distinct weight
buffers per shape (> L2 in total) launched round-robin from one CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from unimp_b200 import ops

torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
M = 5
shapes = [("qkv", 7680, 2560), ("dense", 2560, 2560), ("h_to_4h", 10240, 2560), ("4h_to_h", 2560, 10240),
          ("to_q", 512, 2560), ("to_out", 2560, 512)]
def timed(fn, n):
    for _ in range(2): fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * n)
with torch.no_grad():
    for name, N, K in shapes:
        nb = max(4, int(400e6 // (N * K * 2)) + 1)
        ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) for _ in range(nb)]
        b = torch.randn(N, device="cuda", dtype=torch.bfloat16)
        x = torch.randn(M, 1, K, device="cuda", dtype=torch.bfloat16)
        ours = timed(lambda: [ops.linear_rows(x, w, b) for w in ws], nb)
        ref = timed(lambda: [F.linear(x, w, b) for w in ws], nb) if not os.environ.get("NO_CUBLAS") else float("nan")
        mb = N * K * 2 / 1e6
        print(f"{name:8s} N={N:6d} K={K:6d} {mb:6.1f} MB  ours {ours:6.2f} us = {mb / ours * 1e3:6.0f} GB/s   cuBLAS {ref:6.2f} us = {mb / ref * 1e3:6.0f} GB/s")
