"""A/B harness for unimp_b200/csrc/wip/gate_ln_bwd_pipelined.cu (dev tool for the next round; needs a GPU).

Builds the WIP kernel into build/libunimp_wip.so (separate from the product library), runs it next to
the production K5 backward on the same inputs, compares d_x / d_branch / d_gamma / d_beta / d_gate, and
times both with the cold-cache graph method of unimp_b200/kbench.py.  Run under a short `timeout`:
a protocol bug in the mbarrier pipeline traps (bounded waits) rather than hangs, but be careful.

    python tools/wip_ln_pipelined_check.py            # rows=1536 D=2560 bf16
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unimp_b200 import ops  # noqa: E402
from unimp_b200.kbench import _k, _time_graph  # noqa: E402

src = os.path.join(ROOT, "unimp_b200", "csrc", "wip", "gate_ln_bwd_pipelined.cu")
so = os.path.join(ROOT, "build", "libunimp_wip.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--use_fast_math",
                       "-lineinfo", "-Xcompiler", "-fPIC", "-shared", "-o", so, src, "-lcudart"])
wip = C.CDLL(so)
f = wip.unimp__gate_residual_ln_bwd_pipelined_main
f.restype = C.c_int
f.argtypes = [C.c_void_p] * 10 + [C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]

torch.cuda.set_stream(torch.cuda.Stream())
rows, D = int(os.environ.get("ROWS", 1536)), int(os.environ.get("D", 2560))
dt, dev = torch.bfloat16, "cuda"
torch.manual_seed(0)


def ptr(t):
    return None if t is None else t.data_ptr()


def run_wip(g_xout, g_ln, branch, x_out, gate, gamma, mean, rstd, cols):
    d_x = torch.empty_like(x_out)
    d_branch = torch.empty_like(x_out) if (branch is not None and gate is not None) else None
    partial = torch.zeros(3 * 148 * (2 * D + 1), dtype=torch.float32, device=dev)
    G = f(ptr(g_xout), ptr(g_ln), ptr(branch), ptr(x_out), ptr(gate), ptr(gamma), ptr(mean), ptr(rstd), ptr(d_x),
          ptr(d_branch), int(cols), ptr(partial), rows, D, 1, torch.cuda.current_stream().cuda_stream)
    assert G > 0, f"wip launch failed: {G}"
    p = partial[: G * (2 * D + 1)].view(G, 2 * D + 1).sum(0)
    return d_x, d_branch, p[:D], p[D:2 * D], p[2 * D]


for gated in (True, False):
    x = torch.randn(rows, D, device=dev, dtype=dt, requires_grad=True)
    br = torch.randn(rows, D, device=dev, dtype=dt, requires_grad=True)
    gate = torch.full((1,), 0.5, device=dev, dtype=dt, requires_grad=True) if gated else None
    gam = (1 + 0.1 * torch.randn(D, device=dev)).to(dt).requires_grad_(True)
    bet = torch.zeros(D, device=dev, dtype=dt, requires_grad=True)
    g1, g2 = torch.randn(rows, D, device=dev, dtype=dt), torch.randn(rows, D, device=dev, dtype=dt)
    xo, ln = ops.gate_residual_ln(br, x, gate, gam, bet)
    want = torch.autograd.grad((xo, ln), (br, x, gam, bet) + ((gate,) if gated else ()), (g1, g2))
    # mean / rstd as the forward saved them
    mean = xo.float().mean(-1)
    rstd = (xo.float().var(-1, unbiased=False) + 1e-5).rsqrt()
    d_x, d_b, dg, db, dgate = run_wip(g1, g2, br.detach(), xo.detach(), gate.detach() if gated else None,
                                      gam.detach(), mean, rstd, True)
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))
    print(f"gated={gated}: d_x {rel(d_x, want[1]):.2e}  d_branch {rel(d_b if gated else d_x, want[0]):.2e}  "
          f"d_gamma {rel(dg, want[2]):.2e}  d_beta {rel(db, want[3]):.2e}" +
          (f"  d_gate {rel(dgate, want[4]):.2e}" if gated else ""))
    K = _k(6 * rows * D * 2, cap=8)
    sets = [tuple(torch.randn(rows, D, device=dev, dtype=dt) for _ in range(4)) for _ in range(K)]
    t_wip = _time_graph([lambda s=s: run_wip(s[0], s[1], s[2] if gated else None, s[3], gate.detach() if gated else None,
                                            gam.detach(), mean, rstd, True) for s in sets])
    print(f"   wip main pass (+ host-side partial sum): {t_wip:.2f} us per launch")
