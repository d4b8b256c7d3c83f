"""Which aten::copy_/contiguous calls in one eager fused step move big tensors? (dev tool)"""
import os, sys, copy, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from unimp_b200 import openflamingo_4b_config
from unimp_b200.config import WORKLOADS
from unimp_b200.factory import build_flamingo
from unimp_b200.synth import make_batch
from unimp_b200.train import FlatAdamW, get_grouped_params, train_step
cfg = openflamingo_4b_config(); wl = copy.copy(WORKLOADS["C2-rec"])
model = build_flamingo(cfg, dtype=torch.bfloat16, device="cuda", gate=0.5).train()
opt = FlatAdamW(get_grouped_params(model, 0.1), lr=2e-4)
mbs = [{k: v.cuda() for k, v in make_batch(cfg, wl, seed=i).items()} for i in range(2)]
for _ in range(2):
    train_step(model, None, cfg.tokens, opt, None, accum_steps=2, micro_batches=mbs, fuse_accum=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    train_step(model, None, cfg.tokens, opt, None, accum_steps=2, micro_batches=mbs, fuse_accum=True)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0, None])
for e in prof.events():
    if e.name in ("aten::copy_", "aten::contiguous", "aten::clone", "aten::fill_", "aten::zero_", "aten::add_", "aten::add"):
        key = (e.name, str(e.input_shapes)[:90])
        agg[key][0] += 1
        agg[key][1] += e.device_time_total
        if agg[key][2] is None and e.stack:
            agg[key][2] = [s for s in e.stack if "unimp_b200" in s or "transformers" in s][:3]
for (name, shp), (n, t, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{t/1e3:7.3f} ms {n:4d} x {name:18s} {shp}\n        {st}")
