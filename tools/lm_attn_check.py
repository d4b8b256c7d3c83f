"""On-GPU check + timing of the K4 kernels (unimp_lm_attn_fwd/bwd) against fp64 dense / cuDNN SDPA.
usage: python tools/lm_attn_check.py [check] [bench]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unimp_b200 import ops  # noqa: E402

dev, bf, dh = "cuda", torch.bfloat16, 80


def views(pk, B, T, H):
    return tuple(pk.view(B, T, H, 3, dh)[:, :, :, i].transpose(1, 2) for i in range(3))


def ref(packed, H, km, scale):
    B, T, _ = packed.shape
    q, k, v = views(packed, B, T, H)
    sim = (q @ k.transpose(-1, -2)) * scale
    vis = torch.ones(T, T, dtype=torch.bool, device=packed.device).tril()[None, None]
    if km is not None:
        vis = vis & km.bool()[:, None, None, :]
    sim = sim.masked_fill(~vis, float("-inf"))
    none = ~vis.any(-1, keepdim=True)
    p = torch.softmax(sim.masked_fill(none, 0.0), -1).masked_fill(none, 0.0)
    return (p @ v).transpose(1, 2).reshape(B, T, H * dh)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def check(B, T, H, pad):
    torch.manual_seed(T)
    packed = torch.randn(B, T, H * 3 * dh, device=dev, dtype=bf)
    go = torch.randn(B, T, H * dh, device=dev, dtype=bf)
    km = None
    if pad:
        km = torch.ones(B, T, dtype=torch.int64, device=dev)
        for b in range(1, B):
            km[b, T - (b * 37) % (T // 2) - 1:] = 0
    bits = ops.key_bits(km) if pad else None
    pk = packed.clone().requires_grad_(True)
    o = ops.lm_attention(*views(pk, B, T, H), bits, scale=dh ** -0.5)
    o.backward(go)
    torch.cuda.synchronize()
    r = packed.double().requires_grad_(True)
    want = ref(r, H, km, dh ** -0.5)
    want.backward(go.double())
    g, gr = pk.grad.view(B, T, H, 3, dh), r.grad.view(B, T, H, 3, dh)
    print(f"LM check B={B} T={T} H={H} pad={pad}: o {rel(o, want):.2e}  dq {rel(g[..., 0, :], gr[..., 0, :]):.2e}  "
          f"dk {rel(g[..., 1, :], gr[..., 1, :]):.2e}  dv {rel(g[..., 2, :], gr[..., 2, :]):.2e}  "
          f"nan {int(torch.isnan(o).sum())}/{int(torch.isnan(pk.grad).sum())}", flush=True)


def bench(B, T, H, sets=8, iters=5):
    F = torch.nn.functional
    packs = [torch.randn(B, T, H * 3 * dh, device=dev, dtype=bf, requires_grad=True) for _ in range(sets)]
    gos = [torch.randn(B, T, H * dh, device=dev, dtype=bf) for _ in range(sets)]

    def run(fn):
        outs = [fn(p) for p in packs]                      # warm-up + graphs for backward timing
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        for _ in range(iters):
            outs = [fn(p) for p in packs]
        e[1].record()
        torch.cuda.synchronize()
        fwd = e[0].elapsed_time(e[1]) / (iters * sets) * 1e3
        e[2].record()
        for _ in range(iters):
            for o, p, g in zip(outs, packs, gos):
                torch.autograd.grad(o, p, g, retain_graph=True)
        e[3].record()
        torch.cuda.synchronize()
        return fwd, e[2].elapsed_time(e[3]) / (iters * sets) * 1e3

    ours = run(lambda p: ops.lm_attention(*views(p, B, T, H), None, scale=dh ** -0.5))
    sdpa = run(lambda p: F.scaled_dot_product_attention(*views(p, B, T, H), is_causal=True, scale=dh ** -0.5)
               .transpose(1, 2).reshape(B, T, H * dh))
    flops = 4 * B * H * T * T * dh / 2
    print(f"LM bench B={B} T={T} H={H}: ours fwd {ours[0]:.1f} us ({flops / ours[0] / 1e6:.0f} TFLOP/s) "
          f"fwd+bwd-launch {ours[1]:.1f} us | SDPA fwd {sdpa[0]:.1f} us bwd {sdpa[1]:.1f} us "
          f"(eager loops: launch overhead included on both sides)", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["check"]
    if "check" in what:
        for B, T, H, pad in [(1, 64, 1, False), (1, 128, 1, False), (2, 24, 4, False), (1, 200, 3, True),
                             (3, 256, 32, True), (2, 513, 2, True), (6, 1024, 32, True)]:
            check(B, T, H, pad)
    if "bench" in what:
        with torch.no_grad():
            pass
        bench(6, 256, 32)
        bench(6, 1024, 32, sets=4, iters=3)


def timeline(kind):
    """Per-CTA step stamps of one flash_fwd launch (unimp__flash_fwd_debug hook)."""
    import ctypes
    from unimp_b200 import _lib
    lib = _lib.load()
    f = lib.unimp__flash_fwd_debug
    f.argtypes = [ctypes.c_void_p]
    f.restype = None
    if kind == "lm":
        B, T, H = 6, 1024, 32
        pk = torch.randn(B, T, H * 3 * dh, device=dev, dtype=bf)
        run = lambda: ops.lm_attention(*views(pk, B, T, H), None, scale=dh ** -0.5)
        n_cta = (T // 128) * H * B
    else:
        Bt, L, Hh = 48, 257, 16
        q = torch.randn(Bt, L, Hh * 64, device=dev, dtype=bf)
        kv = torch.randn(Bt, L, 2 * Hh * 64, device=dev, dtype=bf)
        run = lambda: ops.attention(q, kv, heads=Hh, scale=0.125)
        n_cta = 3 * Hh * Bt
    buf = torch.zeros(n_cta * 64, dtype=torch.int64, device=dev)
    with torch.no_grad():
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        f(buf.data_ptr())
        run()
        torch.cuda.synchronize()
        f(None)
    t = buf.view(n_cta, 64).cpu().double()
    t0 = t[:, 0].min()
    span = (t[:, 3].max() - t0)
    occ = lib.unimp__flash_fwd_occupancy(80 if kind == "lm" else 64)
    print(f"FF occupancy query: {occ} CTAs/SM; measured concurrency "
          f"{float((t[:, 3] - t[:, 0]).sum() / span / 148):.2f} CTAs/SM (sum of CTA lifetimes / span / 148)")
    print(f"FF timeline {kind}: {n_cta} CTAs, first start to last end {span:.0f} ns; per-CTA start->ready "
          f"{(t[:, 1] - t[:, 0]).median():.0f}, ready->last PV {(t[:, 2] - t[:, 1]).median():.0f}, epilogue "
          f"{(t[:, 3] - t[:, 2]).median():.0f}")
    full = t[t[:, 8 + 4 * 6 + 5] > 0]      # CTAs with >= 5 steps
    print(f"FF   {len(full)} CTAs with >= 5 steps; medians over them, SM CYCLES since the CTA's 'ready' stamp:")
    for j in range(5):
        s = 8 + j * 6
        rel = lambda k: (full[:, s + k] - full[:, 4]).median()
        print(f"FF   step {j}: worker S ready {rel(2):7.0f} | exps done {rel(3):7.0f} | PV(j-1) seen "
              f"{rel(4) if j else float('nan'):7.0f} | arrived {rel(5):7.0f} || issuer got P {rel(0):7.0f} | "
              f"issuer step done {rel(1):7.0f}", flush=True)


def bwd_timeline(B, T, H):
    """Per-CTA pair stamps of one unimp_lm_attn_bwd launch (unimp__lm_bwd_debug hook)."""
    import ctypes
    from unimp_b200 import _lib
    lib = _lib.load()
    f = lib.unimp__lm_bwd_debug
    f.argtypes = [ctypes.c_void_p]
    f.restype = None
    pk = torch.randn(B, T, H * 3 * dh, device=dev, dtype=bf)
    qkv = tuple(x.requires_grad_() for x in views(pk, B, T, H))
    o = ops.lm_attention(*qkv, None, scale=dh ** -0.5)
    go = torch.randn_like(o)
    n_items = ((T + 63) // 64) * H * B
    n_cta = min(n_items, torch.cuda.get_device_properties(0).multi_processor_count)
    buf = torch.zeros(n_cta * 64, dtype=torch.int64, device=dev)
    for _ in range(3):
        torch.autograd.grad(o, qkv, go, retain_graph=True)
    torch.cuda.synchronize()
    f(buf.data_ptr())
    torch.autograd.grad(o, qkv, go, retain_graph=True)
    torch.cuda.synchronize()
    f(None)
    t = buf.view(n_cta, 64).cpu().double()
    span = t[:, 7].max() - t[:, 0].min()
    print(f"BW timeline B={B} T={T} H={H}: {n_items} items on {n_cta} persistent CTAs, span {span:.0f} ns "
          f"({span / (n_items / n_cta):.0f} ns per item per CTA); start->ready {(t[:, 1] - t[:, 0]).median():.0f} ns; "
          f"first item: ready -> end {(t[:, 3] - t[:, 1]).median():.0f} ns, last gradients -> end "
          f"{(t[:, 3] - t[:, 2]).median():.0f} ns ({(t[:, 6] - t[:, 5]).median():.0f} cycles)")
    sel = t
    print(f"BW   first item of each CTA (the heaviest key blocks); cycles since 'ready':")
    for i in range(6):
        s_ = 8 + i * 8
        if (sel[:, s_ + 5] > 0).sum() < len(sel) // 2:
            break
        rel = lambda k: (sel[:, s_ + k] - sel[:, 4]).median()
        print(f"BW     pair {i}: S/dP ready {rel(2):6.0f} | math done {rel(3):6.0f} | prev grads seen {rel(4):6.0f} | "
              f"arrived {rel(5):6.0f} | prev dQ staged {rel(6):6.0f} || MMA got P {rel(0):6.0f} | MMA issued {rel(1):6.0f}",
              flush=True)
    print(f"BW     flush start {(sel[:, 5] - sel[:, 4]).median():6.0f} | item end {(sel[:, 6] - sel[:, 4]).median():6.0f}")


if "small" in sys.argv[1:]:      # for compute-sanitizer runs: every kernel variant once, small shapes
    for B, T, H, pad in [(1, 200, 2, True), (2, 513, 1, True)]:
        check(B, T, H, pad)
    q = torch.randn(2, 257, 2 * 64, device=dev, dtype=bf)
    kv = torch.randn(2, 257, 2 * 2 * 64, device=dev, dtype=bf)
    o = ops.attention(q, kv, heads=2, scale=0.125)
    torch.cuda.synchronize()
    print("SMALL attention (dh 64) ok", float(o.float().abs().mean()))

if "bwd_timeline" in sys.argv[1:]:
    bwd_timeline(6, 256, 32)
    bwd_timeline(6, 1024, 32)

def timeline3(kind):
    """Per-tile step stamps of one flash_fwd3 launch (three query tiles per CTA)."""
    import ctypes
    from unimp_b200 import _lib
    lib = _lib.load()
    f = lib.unimp__flash_fwd_debug
    f.argtypes = [ctypes.c_void_p]
    f.restype = None
    if kind == "lm":
        B, T, H = 6, 1024, 32
        pk = torch.randn(B, T, H * 3 * dh, device=dev, dtype=bf)
        run = lambda: ops.lm_attention(*views(pk, B, T, H), None, scale=dh ** -0.5)
        n_items = ((T // 128 + 2) // 3) * H * B
    else:
        Bt, L, Hh = 48, 257, 16
        q = torch.randn(Bt, L, Hh * 64, device=dev, dtype=bf)
        kv = torch.randn(Bt, L, 2 * Hh * 64, device=dev, dtype=bf)
        run = lambda: ops.attention(q, kv, heads=Hh, scale=0.125)
        n_items = Hh * Bt
    n_cta = min(n_items, torch.cuda.get_device_properties(0).multi_processor_count)
    buf = torch.zeros(n_cta * 64, dtype=torch.int64, device=dev)
    with torch.no_grad():
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        f(buf.data_ptr())
        run()
        torch.cuda.synchronize()
        f(None)
    t = buf.view(n_cta, 64).cpu().double()
    span = t[:, 5].max() - t[:, 0].min()
    print(f"F3 timeline {kind}: {n_items} items on {n_cta} persistent CTAs, span {span:.0f} ns "
          f"({span / (n_items / n_cta):.0f} ns per item per CTA); first item: ready -> end "
          f"{(t[:, 3] - t[:, 1]).median():.0f} ns")
    full = t[(t[:, 8 + 2 * 16 + 3 * 4 + 2] > 0)] if kind == "lm" else t[(t[:, 8 + 1 * 16 + 3 * 4 + 2] > 0)]
    print(f"F3   {len(full)} CTAs whose FIRST item has >= 4 steps in every tile; medians, SM cycles since 'ready':")
    for j in range(4):
        for tl in range(3):
            s_ = 8 + tl * 16 + j * 4
            rel = lambda k: (full[:, s_ + k] - full[:, 4]).median()
            print(f"F3   step {j} tile {tl}: S ready {rel(0):7.0f} | math done {rel(1):7.0f} | arrived {rel(2):7.0f} | "
                  f"MMA lane served {rel(3):7.0f}", flush=True)


if "timeline3" in sys.argv[1:]:
    timeline3("lm")
    timeline3("vit")

if "timeline" in sys.argv[1:]:
    timeline("lm")
    timeline("vit")
