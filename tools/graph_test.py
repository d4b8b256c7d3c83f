"""Does the whole train step capture into a CUDA graph? eager vs graphed timing + loss equality (dev tool)."""
import os, sys, copy, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unimp_b200 import openflamingo_4b_config, tiny_config
from unimp_b200.config import WORKLOADS
from unimp_b200.factory import build_flamingo
from unimp_b200.synth import make_batch
from unimp_b200.train import FlatAdamW, get_grouped_params, train_step, GraphedTrainStep

tiny = len(sys.argv) > 1 and sys.argv[1] == "tiny"
cfg = tiny_config() if tiny else openflamingo_4b_config()
wl = copy.copy(WORKLOADS["C1-tiny" if tiny else "C2-rec"])
def fresh():
    m = build_flamingo(cfg, dtype=torch.bfloat16, device="cuda", gate=0.5).train()
    return m, FlatAdamW(get_grouped_params(m, 0.1), lr=2e-4)
mbs = [{k: v.cuda() for k, v in make_batch(cfg, wl, seed=i).items()} for i in range(2)]
model, opt = fresh()
eager = []
for i in range(6):
    eager.append(float(train_step(model, None, cfg.tokens, opt, None, accum_steps=2, micro_batches=mbs)))
torch.cuda.synchronize(); t0 = time.time()
for i in range(5):
    train_step(model, None, cfg.tokens, opt, None, accum_steps=2, micro_batches=mbs)
torch.cuda.synchronize(); print("eager ms/step", (time.time() - t0) / 5 * 1e3)
del model, opt
model, opt = fresh()
g = GraphedTrainStep(model, cfg.tokens, opt, None, mbs, warmup_iters=3)   # consumes steps 1-3 eagerly + capture=4th
graphed = [float(g(mbs)) for _ in range(2)]
print("eager losses  ", eager)
print("graphed losses (steps 5,6)", graphed)
torch.cuda.synchronize(); t0 = time.time()
for i in range(10):
    g(mbs)
torch.cuda.synchronize(); print("graphed ms/step", (time.time() - t0) / 10 * 1e3)
