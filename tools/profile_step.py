"""torch.profiler breakdown of one 4B train step (eager) (dev tool; numbers under a profiler are never bench values)."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from unimp_b200 import openflamingo_4b_config
from unimp_b200.config import WORKLOADS
from unimp_b200.factory import build_flamingo
from unimp_b200.synth import make_batch
from unimp_b200.train import FlatAdamW, get_grouped_params, train_step

cfg = openflamingo_4b_config()
wl = copy.copy(WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C2-rec"])
model = build_flamingo(cfg, dtype=torch.bfloat16, device="cuda", gate=0.5).train()
opt = FlatAdamW(get_grouped_params(model, 0.1), lr=2e-4)
mbs = [{k: v.cuda() for k, v in make_batch(cfg, wl, seed=i).items()} for i in range(2)]
for _ in range(3):
    train_step(model, None, cfg.tokens, opt, None, accum_steps=2, micro_batches=mbs)
torch.cuda.synchronize()
import time
t0 = time.time()
for _ in range(3):
    train_step(model, None, cfg.tokens, opt, None, accum_steps=2, micro_batches=mbs)
torch.cuda.synchronize()
print("wall ms/step", (time.time() - t0) / 3 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    train_step(model, None, cfg.tokens, opt, None, accum_steps=2, micro_batches=mbs)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
