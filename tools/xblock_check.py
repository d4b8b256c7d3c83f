"""K1-fused diagnostics (dev tool, run under gpurun): per-output errors of unimp_xattn_block_fwd
(q, o, lse, y each checked on its own, so a wrong phase is named) and its isolated timing next to
the three-launch form (cuBLAS to_q + attention kernel + cuBLAS to_out) and to eager PyTorch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from unimp_b200 import _lib, ops
from unimp_b200._lib import BF16
from unimp_b200.kbench import _time_graph, _k, _eager_xattn

torch.cuda.set_stream(torch.cuda.Stream())
dev, bf = "cuda", torch.bfloat16
H, dh, n = 8, 64, 64


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def mk_tt(B, T, Ti):
    loc = torch.zeros(B, T, dtype=torch.bool)
    for b in range(B):
        for j in range(Ti):
            loc[b, 1 + j * (T // Ti)] = True
    return loc.cumsum(-1).to(torch.int32).to(dev)


def check(B, T, Ti, D):
    torch.manual_seed(B + T + D)
    x = torch.randn(B, T, D, device=dev, dtype=bf)
    wq = (torch.randn(H * dh, D, device=dev) * D ** -0.5).to(bf)
    wout = (torch.randn(D, H * dh, device=dev) * (H * dh) ** -0.5).to(bf)
    kv = torch.randn(B, Ti * n, 2 * H * dh, device=dev, dtype=bf)
    tt = mk_tt(B, T, Ti)
    inner = H * dh
    q = torch.full((B, T, inner), float("nan"), device=dev, dtype=bf)
    o = torch.full_like(q, float("nan"))
    lse = torch.full((B, H, T), float("nan"), device=dev)
    y = torch.full((B, T, D), float("nan"), device=dev, dtype=bf)
    k, v = kv[..., :inner], kv[..., inner:]
    rc = _lib.load().unimp_xattn_block_fwd(x.data_ptr(), wq.data_ptr(), ops._view3(k), ops._view3(v), tt.data_ptr(),
                                           wout.data_ptr(), q.data_ptr(), o.data_ptr(), lse.data_ptr(), y.data_ptr(),
                                           B, T, Ti, n, H, dh, D, dh ** -0.5, BF16,
                                           torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == 0, _lib.load().unimp_last_error_string()
    q_ref = torch.nn.functional.linear(x, wq)
    # phase 2 checked on the kernel's OWN q (isolates the attention core), phase 3 on its own o
    o_ref = ops.masked_cross_attention(q.clone(), kv, tt, heads=H, n_latents=n, scale=dh ** -0.5)
    y_ref = torch.nn.functional.linear(o.clone(), wout)
    y_e2e = torch.nn.functional.linear(
        ops.masked_cross_attention(q_ref, kv, tt, heads=H, n_latents=n, scale=dh ** -0.5), wout)
    print(f"XB check B={B} T={T} Ti={Ti} D={D}: q {rel(q, q_ref):.2e}  o|q {rel(o, o_ref):.2e}  "
          f"y|o {rel(y, y_ref):.2e}  y e2e {rel(y, y_e2e):.2e}  nan: q {int(q.isnan().sum())} o {int(o.isnan().sum())} "
          f"y {int(y.isnan().sum())} lse {int(lse.isnan().sum())}", flush=True)


def bench(B, T, Ti, D, tag):
    inner = H * dh
    fl = 2.0 * B * T * D * inner * 2 + 4.0 * H * dh * n * B * T
    byts = 2 * (2 * B * T * D + 2 * B * T * inner + 2 * B * Ti * n * inner + 2 * D * inner) + 4 * B * T * H
    K = _k(byts, cap=24)
    tt = mk_tt(B, T, Ti)
    xs = [torch.randn(B, T, D, device=dev, dtype=bf) for _ in range(K)]
    kvs = [torch.randn(B, Ti * n, 2 * inner, device=dev, dtype=bf) for _ in range(K)]
    wqs = [(torch.randn(inner, D, device=dev) * D ** -0.5).to(bf) for _ in range(K)]
    wos = [(torch.randn(D, inner, device=dev) * inner ** -0.5).to(bf) for _ in range(K)]
    with torch.no_grad():
        fused = _time_graph([lambda x=x, kv=kv, wq=wq, wo=wo: ops.xattn_block(x, wq, kv, tt, wo, heads=H, n_latents=n, scale=0.125)
                             for x, kv, wq, wo in zip(xs, kvs, wqs, wos)])
        three = _time_graph([lambda x=x, kv=kv, wq=wq, wo=wo: torch.nn.functional.linear(
            ops.masked_cross_attention(torch.nn.functional.linear(x, wq), kv, tt, heads=H, n_latents=n, scale=0.125), wo)
            for x, kv, wq, wo in zip(xs, kvs, wqs, wos)])
        eager = _time_graph([lambda x=x, kv=kv, wq=wq, wo=wo: torch.nn.functional.linear(
            _eager_xattn(torch.nn.functional.linear(x, wq), kv, tt, H, n, 0.125), wo)
            for x, kv, wq, wo in zip(xs, kvs, wqs, wos)])
    print(f"XB bench {tag} B={B} T={T} Ti={Ti} D={D}: fused {fused:.2f} us = {fl / fused / 1e6:.1f} TFLOP/s | "
          f"three launches {three:.2f} us | eager {eager:.2f} us | {fl / 1e9:.2f} GFLOP, {byts / 1e6:.1f} MB per launch",
          flush=True)


def timeline(B, T, Ti, D):
    """Per-CTA phase timestamps of one launch (unimp__xattn_block_debug test hook)."""
    import ctypes
    lib = _lib.load()
    inner = H * dh
    x = torch.randn(B, T, D, device=dev, dtype=bf)
    wq = (torch.randn(inner, D, device=dev) * D ** -0.5).to(bf)
    wo = (torch.randn(D, inner, device=dev) * inner ** -0.5).to(bf)
    kv = torch.randn(B, Ti * n, 2 * inner, device=dev, dtype=bf)
    tt = mk_tt(B, T, Ti)
    n_cta = 8 * ((T + 127) // 128) * B
    buf = torch.zeros(n_cta * 16 + 192, dtype=torch.int64, device=dev)
    f = lib.unimp__xattn_block_debug
    f.argtypes = [ctypes.c_void_p]
    f.restype = None
    with torch.no_grad():
        for _ in range(3):
            ops.xattn_block(x, wq, kv, tt, wo, heads=H, n_latents=n, scale=0.125)
        torch.cuda.synchronize()
        f(buf.data_ptr())
        ops.xattn_block(x, wq, kv, tt, wo, heads=H, n_latents=n, scale=0.125)
        torch.cuda.synchronize()
        f(None)
    t = buf[:n_cta * 16].view(n_cta, 16).cpu().double()
    t0 = t[:, 0].min()
    names = ["start", "csync1", "ph1 done", "attn done", "o stored", "csync2", "O tiles in", "to_out done",
             "y stored", "end"]
    print(f"XB timeline B={B} T={T} Ti={Ti} D={D} ({n_cta} CTAs); ns since the first CTA's start: median [min, max]")
    for i, nm in enumerate(names):
        col = t[:, i] - t0
        print(f"XB   {nm:12s} {col.median():9.0f} [{col.min():9.0f}, {col.max():9.0f}]")
    print("XB   per-CTA durations (median ns): csync1 %.0f | ph1 %.0f | attn %.0f | "
          "o store %.0f | csync2 %.0f | O exchange %.0f | to_out %.0f | y store %.0f" % (
              (t[:, 1] - t[:, 0]).median(), (t[:, 2] - t[:, 1]).median(),
              (t[:, 3] - t[:, 2]).median(), (t[:, 4] - t[:, 3]).median(),
              (t[:, 5] - t[:, 4]).median(), (t[:, 6] - t[:, 5]).median(), (t[:, 7] - t[:, 6]).median(),
              (t[:, 8] - t[:, 7]).median()), flush=True)


def core_timeline(B, T, Ti):
    """Per-CTA phase timestamps of one launch of the standalone core (unimp__xattn_fwd_debug hook)."""
    import ctypes
    lib = _lib.load()
    inner = H * dh
    q = torch.randn(B, T, inner, device=dev, dtype=bf)
    kv = torch.randn(B, Ti * n, 2 * inner, device=dev, dtype=bf)
    tt = mk_tt(B, T, Ti)
    n_cta = ((T + 127) // 128) * H * B
    buf = torch.zeros(n_cta * 8, dtype=torch.int64, device=dev)
    f = lib.unimp__xattn_fwd_debug
    f.argtypes = [ctypes.c_void_p]
    f.restype = None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        for _ in range(3):
            ops.masked_cross_attention(q, kv, tt, heads=H, n_latents=n, scale=0.125)
        flush.zero_()                      # cold L2, like the back-to-back cold sets of kbench
        torch.cuda.synchronize()
        f(buf.data_ptr())
        ops.masked_cross_attention(q, kv, tt, heads=H, n_latents=n, scale=0.125)
        torch.cuda.synchronize()
        f(None)
    t = buf.view(n_cta, 8).cpu().double()
    t0 = t[:, 0].min()
    names = ["start", "prologue", "S ready", "P written", "softmax", "O ready", "stored"]
    print(f"XF timeline B={B} T={T} Ti={Ti} ({n_cta} CTAs, cold L2); ns since the first CTA's start: median [min, max]")
    for i, nm in enumerate(names):
        col = t[:, i] - t0
        print(f"XF   {nm:10s} {col.median():8.0f} [{col.min():8.0f}, {col.max():8.0f}]")
    d = lambda i, j: (t[:, i] - t[:, j]).median()
    print("XF   per-CTA durations (median ns): prologue %.0f | loads+first S %.0f | softmax(1st) %.0f | rest %.0f | PV+wait %.0f | "
          "store %.0f | CTA total %.0f" % (d(1, 0), d(2, 1), d(3, 2), d(4, 3), d(5, 4), d(6, 5), d(6, 0)), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "check"):
        for shape in ((1, 128, 1, 128), (2, 32, 2, 128), (1, 128, 2, 2560), (3, 256, 2, 2560), (1, 513, 8, 512),
                      (6, 1024, 8, 2560)):
            check(*shape)
    if what in ("all", "core"):
        core_timeline(6, 256, 2)
        core_timeline(6, 1024, 8)
    if what in ("all", "timeline"):
        timeline(6, 256, 2, 2560)
        timeline(6, 1024, 8, 2560)
    if what in ("all", "bench"):
        bench(6, 256, 2, 2560, "C2")
        bench(6, 1024, 8, 2560, "C3")
