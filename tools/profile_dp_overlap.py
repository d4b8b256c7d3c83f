"""N>1: do the NCCL all-reduces inside the replayed graph overlap backward? (dev tool; torchrun)"""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.cuda.set_stream(torch.cuda.Stream())
from unimp_b200 import openflamingo_4b_config
from unimp_b200.config import WORKLOADS
from unimp_b200.factory import build_flamingo
from unimp_b200.synth import make_batch
from unimp_b200.train import FlatAdamW, BucketedAllReduce, ShardedDataParallel, get_grouped_params, GraphedTrainStep
cfg = openflamingo_4b_config(); wl = copy.copy(WORKLOADS["C2-rec"])
model = build_flamingo(cfg, dtype=torch.bfloat16, device="cuda", gate=0.5).train()
world = dist.get_world_size()
opt = FlatAdamW(get_grouped_params(model, 0.1), lr=2e-4, shard_world=world, allocate_states=False)
red = ShardedDataParallel(opt, deferred_gather_module=model.perceiver if os.environ.get("DEFER", "1") == "1" else None,
                          gather_start_module=model.vision_encoder if os.environ.get("UNIMP_AG_AT_VIT") else None)
mbs = [{k: v.cuda() for k, v in make_batch(cfg, wl, seed=rank * 10 + i).items()} for i in range(2)]
g = GraphedTrainStep(model, cfg.tokens, opt, red, mbs, fuse_accum=True)
for _ in range(3):
    g(mbs)
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g(mbs); torch.cuda.synchronize()
if rank == 0:
    ev = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if "cuda" in str(e.device_type).lower()]
    ev.sort()
    t0 = ev[0][0]; t1 = max(e[1] for e in ev)
    nccl = [(s, e) for s, e, n in ev if "nccl" in n.lower()]
    for s_, e_, n_ in ev:
        if "nccl" in n_.lower():
            ov = sum(max(0, min(e_, ce) - max(s_, cs)) for cs, ce, cn in ev if "nccl" not in cn.lower())
            print(f"   {(s_-t0)/1e3:7.2f} ms {(e_-s_)/1e3:6.2f} ms ov {ov/1e3:5.2f}  {n_[:60]}")
    print("first 40 events:")
    for s_, e_, n_ in ev[:40]:
        print(f"   @{(s_-t0)/1e3:7.3f} ms {(e_-s_)/1e3:6.3f} ms {n_[:70]}")
    firstp = [(s_ - t0) / 1e3 for s_, e_, n_ in ev if "gate_residual_ln" in n_][:1]
    print("first LN kernel at", firstp)
    print("pending after step:", red.pending, "armed", red.armed)
    comp = [(s, e) for s, e, n in ev if "nccl" not in n.lower()]
    print(f"buckets {len(red.buckets)} sizes MB {[round(b[0].numel()*2/2**20) for b in red.buckets]}")
    print(f"step span {(t1-t0)/1e3:.2f} ms; compute kernel time {sum(e-s for s,e in comp)/1e3:.2f} ms; nccl kernels {len(nccl)} time {sum(e-s for s,e in nccl)/1e3:.2f} ms")
    for s, e in nccl:
        ov = sum(max(0, min(e, ce) - max(s, cs)) for cs, ce in comp)
        print(f"  nccl start {(s-t0)/1e3:7.2f} ms dur {(e-s)/1e3:6.2f} ms overlapped with compute {ov/1e3:6.2f} ms")
    last_comp_before_opt = [n for s, e, n in ev if "adamw" in n]
    print("adamw start", [(s - t0) / 1e3 for s, e, n in ev if "adamw" in n])
sys.stdout.flush(); torch.cuda.synchronize(); dist.barrier(); os._exit(0)
