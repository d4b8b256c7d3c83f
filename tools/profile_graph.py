"""Kernel-level durations inside the replayed CUDA graph of one 4B train step (dev tool)."""
import os, sys, copy, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from unimp_b200 import openflamingo_4b_config
from unimp_b200.config import WORKLOADS
from unimp_b200.factory import build_flamingo
from unimp_b200.synth import make_batch
from unimp_b200.train import FlatAdamW, get_grouped_params, GraphedTrainStep

torch.cuda.set_stream(torch.cuda.Stream())
cfg = openflamingo_4b_config()
wl = copy.copy(WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C2-rec"])
model = build_flamingo(cfg, dtype=torch.bfloat16, device="cuda", gate=0.5).train()
opt = FlatAdamW(get_grouped_params(model, 0.1), lr=2e-4)
mbs = [{k: v.cuda() for k, v in make_batch(cfg, wl, seed=i).items()} for i in range(2)]
g = GraphedTrainStep(model, cfg.tokens, opt, None, mbs, fuse_accum=os.environ.get('FUSE', '1') == '1')
for _ in range(3):
    g(mbs)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g(mbs)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type.name == "CUDA" or "cuda" in str(e.device_type).lower():
        agg[e.name][0] += 1
        agg[e.name][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"total kernel time {tot/1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t/1e3:8.3f} ms {n:5d} x {t/n:8.2f} us  {name[:110]}")
