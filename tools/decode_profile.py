"""Kernel breakdown of ONE decode token (C4: 4B, batch 1, beams 5): torch.profiler around the graph
replays of GraphedDecoder.generate.  Usage: python tools/decode_profile.py [n_new_tokens]"""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unimp_b200.config import Workload, openflamingo_4b_config
from unimp_b200.decode import GraphedDecoder
from unimp_b200.factory import build_flamingo
from unimp_b200.synth import make_batch

n_new = int(sys.argv[1]) if len(sys.argv) > 1 else 34
cfg = openflamingo_4b_config()
torch.cuda.set_device(0)
torch.cuda.set_stream(torch.cuda.Stream())
model = build_flamingo(cfg, dtype=torch.bfloat16, device="cuda", gate=0.5).eval()
b = make_batch(cfg, Workload("C4-decode", B=1, Ti=5, T=512), seed=0)
L = int(b["attention_masks"][0].sum())
ids = b["input_ids"][:, :L].cuda()
vis = b["patch_images"].unsqueeze(2).cuda()
dec = GraphedDecoder(model)
kw = dict(num_beams=5, max_new_tokens=n_new, eos_token_id=-1, pad_token_id=cfg.tokens.pad, early_stopping=False)
dec.generate(vis, ids, torch.ones_like(ids), **kw)
print("timing", dec.last_timing)
prof = torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA])
state = {"n": 0}
orig = torch.cuda.CUDAGraph.replay
def replay(self):
    if state["n"] == 0:
        torch.cuda.synchronize()
        prof.start()
    state["n"] += 1
    return orig(self)
torch.cuda.CUDAGraph.replay = replay
dec.generate(vis, ids, torch.ones_like(ids), **kw)
torch.cuda.synchronize()
prof.stop()
torch.cuda.CUDAGraph.replay = orig
n = state["n"]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[e.name[:110]]
        a[0] += 1
        a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"replays {n}; kernels/token {sum(v[0] for v in agg.values()) / n:.0f}; summed kernel time/token {tot / n:.1f} us; "
      f"replay_ms/token {dec.last_timing['replay_ms'] / max(1, dec.last_timing['replays']):.3f}")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t / n:9.1f} us  {c / n:6.1f} x {t / c:7.2f} us  {name}")
