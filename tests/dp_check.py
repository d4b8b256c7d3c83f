"""N-GPU == 1-GPU gradient equality for the REAL model path (direct-grad accumulation + bucketed
NCCL all-reduce, eager and CUDA-graph).  Launched by tests/test_dp_gpu.py with torchrun on >= 2 GPUs:
every rank holds the same weights and a different micro-batch pair; after the reduce each rank's flat
gradient (x 1/world) must equal the average of the per-rank gradients computed WITHOUT any reducer."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from unimp_b200 import tiny_config  # noqa: E402
from unimp_b200.config import WORKLOADS  # noqa: E402
from unimp_b200.factory import build_flamingo  # noqa: E402
from unimp_b200.synth import make_batch  # noqa: E402
from unimp_b200.train import (BucketedAllReduce, FlatAdamW, GraphedTrainStep, ShardedDataParallel,  # noqa: E402
                                  get_grouped_params, train_step, unimp_loss)


def flat(opt):
    return torch.cat([g["flat_g"].float() for g in opt.groups])


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    torch.cuda.set_stream(torch.cuda.Stream())
    cfg = tiny_config()
    wl = WORKLOADS["C1-tiny"]
    dtype = torch.float32

    def batches(r):
        return [{k: v.cuda() for k, v in make_batch(cfg, wl, seed=100 * r + i).items()} for i in range(2)]

    def grads(model, opt, mbs, reducer):
        opt.zero_grad()
        for i, mb in enumerate(mbs):
            if reducer is not None:
                reducer.armed = i == len(mbs) - 1
            loss, _, _ = unimp_loss(model, mb, cfg.tokens)
            (loss / len(mbs)).backward()
        if reducer is not None:
            reducer.finish()
        opt._zero_unwritten()
        return flat(opt)

    # reference: every rank computes all ranks' local gradients without a reducer and averages
    model = build_flamingo(cfg, dtype=dtype, device="cuda", gate=0.5, seed=0)
    opt = FlatAdamW(get_grouped_params(model, 0.1), lr=0.0)
    want = sum(grads(model, opt, batches(r), None).clone() for r in range(world)) / world
    # data-parallel, eager, small buckets so several fire mid-backward
    model = build_flamingo(cfg, dtype=dtype, device="cuda", gate=0.5, seed=0)
    opt = FlatAdamW(get_grouped_params(model, 0.1), lr=0.0)
    red = BucketedAllReduce(opt, bucket_bytes=256 << 10)
    assert len(red.buckets) >= 4
    for rep in range(2):
        got = grads(model, opt, batches(rank), red) * red.grad_scale
        err = float((got - want).norm() / want.norm())
        assert err < 1e-5, f"eager rep {rep}: rel err {err}"
    # the same through the captured graph (lr = 0: weights stay put, gradients comparable)
    g = GraphedTrainStep(model, cfg.tokens, opt, red, batches(rank), warmup_iters=2)
    for rep in range(2):
        g(batches(rank))
        torch.cuda.synchronize()
        got = flat(opt) * red.grad_scale
        err = float((got - want).norm() / want.norm())
        assert err < 1e-5, f"graph rep {rep}: rel err {err}"
    # ---- sharded optimizer (reduce-scatter -> AdamW on 1/world -> all-gather) == all-reduce + full AdamW
    def run(mode, graph, deferred=False):
        m = build_flamingo(cfg, dtype=dtype, device="cuda", gate=0.5, seed=0)
        sharded = mode == "zero1"
        o = FlatAdamW(get_grouped_params(m, 0.1), lr=1e-3, shard_world=world if sharded else 1,
                      allocate_states=not sharded)
        if sharded:
            r = ShardedDataParallel(o, bucket_bytes=256 << 10,
                                    deferred_gather_module=m.perceiver if deferred else None,
                                    gather_start_module=m.vision_encoder if deferred else None)
        else:
            r = BucketedAllReduce(o, bucket_bytes=256 << 10)
        mbs = batches(rank)
        if graph:
            gs = GraphedTrainStep(m, cfg.tokens, o, r, mbs, warmup_iters=2, fuse_accum=True)  # consumes no step
            assert o.step_count == 0
            for _ in range(3):
                gs(mbs)                                                                         # steps 1-3
        else:
            for _ in range(3):
                train_step(m, None, cfg.tokens, o, r, accum_steps=2, micro_batches=mbs, fuse_accum=True)
        if sharded:
            r.sync_params()
        torch.cuda.synchronize()
        return torch.cat([g["flat_p"].float() for g in o.groups]), [n for g in o.groups for (n, _, _, _) in g["spans"]], o

    p_ar, _, o_ar = run("allreduce", False)
    for graph, deferred in ((False, False), (True, False), (False, True), (True, True)):
        p_z, names, o_z = run("zero1", graph, deferred)
        # layouts differ only by span padding: compare parameter by parameter
        for ga, gz in zip(o_ar.groups, o_z.groups):
            for (na, pa, _, _), (nz, pz, _, _) in zip(ga["spans"], gz["spans"]):
                assert na == nz
                err = float((pz.float() - pa.float()).norm() / pa.float().norm().clamp_min(1e-20))
                assert err < 2e-6, f"zero1 (graph={graph}, deferred={deferred}) vs allreduce: {na} rel err {err}"
        # every rank holds the same full parameters after the all-gather
        chk = p_z.clone()
        dist.all_reduce(chk, op=dist.ReduceOp.MAX)
        assert torch.equal(chk, p_z), "ranks diverged after the parameter all-gather"
    print(f"rank {rank}: DP gradient check OK (world {world})", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
