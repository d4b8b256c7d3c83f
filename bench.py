#!/usr/bin/env python
"""bench.py — train samples/s of the UniMP hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 8 --warmup 3            # ours, N=1
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # ours, N>1 (one rank per GPU)
    python bench.py --impl reference --steps 2 --warmup 1    # reference arithmetic on host CPU

One "step" = one optimizer step of the reference's canonical launch (README.md:56-57,
UniMP/unimp_task.sh:2-9): `accum` micro-batches of B samples per GPU, each = Flamingo forward +
focal loss + backward; then gradient all-reduce, clip 1.0, AdamW.  samples/s follows the
reference's definition `accum * B * world / step_time` (UniMP/mmrec.py:267-275).
Workload = BASELINE.json configs[1]: OpenFlamingo-4B-instruct, bf16, synthetic 2-image histories.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--mode", default="train", choices=["train", "decode"],
                    help="decode: configs[3] (C4) explanation generation, tokens/s (its own JSON line)")
    ap.add_argument("--soak-s", type=float, default=3.0,
                    help="untimed steps run for this many seconds before the timed region, so the "
                         "timed steps see sustained-power clocks even when K x step is < 1 s")
    ap.add_argument("--label-rows", default="auto",
                    help="head+loss fusion row capacity: auto (bound from the batches), off, or an int")
    ap.add_argument("--no-eager-baseline", action="store_true",
                    help="skip the eager-PyTorch-on-this-GPU baseline (oracle graph, bf16 autocast)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2-rec")
    ap.add_argument("--model", default="4b", choices=["4b", "tiny"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU micro-batch (default: workload's)")
    ap.add_argument("--accum", type=int, default=2, help="gradient accumulation (reference: 2)")
    ap.add_argument("--dp", default="zero1", choices=["zero1", "allreduce"],
                    help="N>1: sharded optimizer (reduce-scatter/all-gather) or plain all-reduce")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of a CUDA graph")
    ap.add_argument("--defer-optimizer", action="store_true",
                    help="N=1: run clip+AdamW of step k under the frozen ViT forward of step k+1 (measured: no "
                         "gain under the 1 kW power cap, DESIGN.md; off by default)")
    ap.add_argument("--no-fuse-accum", action="store_true",
                    help="run the accumulation window as sequential micro-batches (reference style)")
    ap.add_argument("--ncu-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed region (ncu --profile-from-start off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-profile", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=1)
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference / cpu_baseline arm: the oracle (restated reference arithmetic) on the host cores
# ---------------------------------------------------------------------------------------------

def build_oracle_model(cfg, device):
    """The oracle (plain-PyTorch restatement of the reference graph) with cheap random init, fp32,
    frozen like upstream, AdamW param groups by the reference's weight-decay rule."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import hf_configs

    from oracle.flamingo_oracle import Flamingo as OracleFlamingo
    from unimp_b200.train import apply_decay
    from transformers import CLIPVisionModel, GPTNeoXForCausalLM

    vc, lc = hf_configs(cfg)
    with torch.device("meta"):
        vis = CLIPVisionModel(vc)
        lm = GPTNeoXForCausalLM(lc)
        model = OracleFlamingo(vis, lm, cfg.tokens.endofchunk, cfg.tokens.media,
                               vis_dim=cfg.vis_width,
                               cross_attn_every_n_layers=cfg.cross_attn_every_n_layers)
    model.to_empty(device=device)
    g = torch.Generator(device=device).manual_seed(0)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() >= 2:
                p.uniform_(-0.02, 0.02, generator=g)
            elif "norm" in n.lower() and n.endswith("weight") or n.endswith(".0.weight"):
                p.fill_(1.0)
            else:
                p.zero_()
        for n, b in model.named_buffers():
            if "inv_freq" in n:
                dim = b.numel() * 2
                b.copy_(1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim)))
            elif "position_ids" in n:
                b.copy_(torch.arange(b.numel()).view_as(b))
        for blk in model.lang_encoder.gated_cross_attn_layers:
            if blk is not None:
                blk.attn_gate.fill_(0.5)
                blk.ff_gate.fill_(0.5)
    model.requires_grad_(False)
    model.perceiver.requires_grad_(True)
    model.lang_encoder.gated_cross_attn_layers.requires_grad_(True)
    model.lang_encoder.lm.get_input_embeddings().requires_grad_(True)
    wd, no_wd = [], []
    for n, p in model.named_parameters():
        if p.requires_grad:
            (wd if apply_decay(n) else no_wd).append(p)
    groups = [{"params": wd, "weight_decay": 0.1}, {"params": no_wd, "weight_decay": 0.0}]
    return model, groups


def cpu_oracle_samples_per_s(cfg, wl, *, batch, steps, warmup, accum=1):
    """Times the oracle train step (fwd + focal loss + bwd + clip + AdamW, fp32) on all host
    cores.  `batch` samples per step is the bounded sample of the workload."""
    import copy

    from oracle.loss_oracle import focal_loss, mask_labels
    from unimp_b200.synth import make_batch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t0 = time.time()
    model, groups = build_oracle_model(cfg, "cpu")
    opt = torch.optim.AdamW(groups, lr=2e-4)
    init_s = time.time() - t0
    wl2 = copy.copy(wl)
    wl2.B = batch
    tk = cfg.tokens
    times = []
    for it in range(warmup + steps):
        mbs = [make_batch(cfg, wl2, seed=1234 + it * accum + a) for a in range(accum)]
        t1 = time.time()
        opt.zero_grad(set_to_none=True)
        for b in mbs:
            labels = mask_labels(b["input_ids"], answer_token_id=tk.answer,
                                 endofchunk_token_id=tk.endofchunk, media_token_id=tk.media,
                                 pad_token_id=tk.pad)
            out = model(vision_x=b["patch_images"].unsqueeze(2), lang_x=b["input_ids"],
                        attention_mask=b["attention_masks"], labels=labels)
            loss = focal_loss(out["logits"], labels, b["weights"], gamma=wl.gamma)
            (loss / accum).backward()
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], 1.0)
        opt.step()
        dt = time.time() - t1
        if it >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    return {"value": accum * batch / per_step, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{steps} optimizer steps of {accum} x {batch} sample(s) of {wl.name} "
                      f"(T={wl.T}, Ti={wl.Ti}), fp32 eager PyTorch oracle, {warmup} warm-up; "
                      f"model init {init_s:.0f}s not timed",
            "ms_per_step": per_step * 1e3}


def gpu_eager_samples_per_s(cfg, wl, *, accum, steps=6, warmup=3):
    """The honest GPU baseline (BASELINE.md §2 column 2): the SAME oracle module
    graph the CPU arm runs, on this B200 in eager PyTorch, the way the reference trains it —
    fp32 weights, `torch.autocast(bf16)` around the model call (UniMP/mmrec.py:176), the loss lines
    :190-213 outside autocast, sequential micro-batches (:175), clip 1.0, fused torch AdamW.
    Labels are precomputed (the reference's per-element Python loop over a CUDA tensor, :146-156,
    would add ~B*T host syncs per micro-batch; leaving it out favours the baseline)."""
    from oracle.loss_oracle import focal_loss, mask_labels
    from unimp_b200.synth import make_batch

    torch.cuda.synchronize()
    model, groups = build_oracle_model(cfg, "cuda")
    opt = torch.optim.AdamW(groups, lr=2e-4, fused=True)
    params = [p for p in model.parameters() if p.requires_grad]
    tk = cfg.tokens
    data = []
    for i in range(accum * 2):
        b = make_batch(cfg, wl, seed=4321 + i)
        lab = mask_labels(b["input_ids"], answer_token_id=tk.answer, endofchunk_token_id=tk.endofchunk,
                          media_token_id=tk.media, pad_token_id=tk.pad)
        data.append(({k: v.cuda() for k, v in b.items()}, lab.cuda()))

    def step(i):
        opt.zero_grad(set_to_none=True)
        for a in range(accum):
            b, lab = data[(i * accum + a) % len(data)]
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = model(vision_x=b["patch_images"].unsqueeze(2), lang_x=b["input_ids"],
                            attention_mask=b["attention_masks"], labels=lab)
            loss = focal_loss(out["logits"], lab, b["weights"], gamma=wl.gamma)
            (loss / accum).backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        step(i)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    peak_gb = torch.cuda.max_memory_allocated() / 2 ** 30
    del model, opt, params, data
    torch.cuda.empty_cache()
    return {"value": accum * wl.B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
            "what": f"oracle module graph (restated open_flamingo v2.0.1 + UniMP loss), eager PyTorch on this "
                    f"GPU, fp32 weights + bf16 autocast, {accum} sequential micro-batches of {wl.B}, clip + "
                    f"fused torch AdamW; {steps} steps after {warmup} warm-up; labels precomputed",
            "peak_mem_gb": peak_gb}


def decode_bench(args, cfg):
    """configs[3] (C4): explanation generation as reference `UniMP/pipeline/eval/eval_exp.py:101-114`
    runs it — eval batch 1, num_beams 5, max_new_tokens 256 — on a T=512-token prompt with Ti=5
    images, cached vision latents + cached cross-attention K/V.  Prints its own JSON line."""
    import statistics as st

    from unimp_b200.config import Workload
    from unimp_b200.decode import GraphedDecoder
    from unimp_b200.factory import build_flamingo
    from unimp_b200.synth import make_batch

    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    wl = Workload("C4-decode", B=1, Ti=5, T=512)
    beams, new = 5, 256
    model = build_flamingo(cfg, dtype=torch.bfloat16, device="cuda", gate=0.5).eval()
    b = make_batch(cfg, wl, seed=0)
    L = int(b["attention_masks"][0].sum())
    ids = b["input_ids"][:, :L].cuda()
    vis = b["patch_images"].unsqueeze(2).cuda()
    dec = GraphedDecoder(model)
    kw = dict(num_beams=beams, max_new_tokens=new, eos_token_id=-1, pad_token_id=cfg.tokens.pad,
              early_stopping=False)   # eos -1: every run decodes exactly `new` tokens
    runs = []
    for i in range(args.warmup if args.warmup < 3 else 2):
        dec.generate(vis, ids, torch.ones_like(ids), **kw)
    reps = max(5, min(args.steps, 9))
    for i in range(reps):
        torch.cuda.synchronize()
        t0 = time.time()
        dec.generate(vis, ids, torch.ones_like(ids), **kw)
        torch.cuda.synchronize()
        t = dict(dec.last_timing)
        t["wall_ms"] = (time.time() - t0) * 1e3
        runs.append(t)
    med = {k: st.median(r[k] for r in runs) for k in runs[0]}
    per_tok = med["replay_ms"] / max(1, med["replays"])
    # the HF path (Flamingo.generate -> transformers beam search), same prompt, fewer tokens
    hf = None
    try:
        n_hf = 32
        hkw = dict(attention_mask=torch.ones_like(ids), num_beams=beams, max_new_tokens=n_hf,
                   min_new_tokens=n_hf, eos_token_id=cfg.tokens.endofchunk, pad_token_id=cfg.tokens.pad,
                   do_sample=False, early_stopping=False)
        model.generate(vision_x=vis, lang_x=ids, **hkw)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.time()
            model.generate(vision_x=vis, lang_x=ids, **hkw)
            torch.cuda.synchronize(); ts.append(time.time() - t0)
        hf = {"tokens_per_s_incl_prefill": n_hf / st.median(ts), "new_tokens": n_hf,
              "what": "Flamingo.generate -> transformers beam search (the reference's call chain) on our kernels"}
    except Exception as ex:  # noqa: BLE001
        hf = {"error": repr(ex)[:200]}
    # HBM floor of one token: every projection weight of the step once (GPT-NeoX layers, the x-attn
    # blocks' to_q / to_out / FF, the output head) + the K/V rows the beams attend (LM caches up to the
    # mean cursor, the last image's cached cross-attention K/V); peak = MEASURED_PEAKS.json copy bandwidth
    lm = model.lang_encoder
    w_bytes = 0
    for layer in lm._get_decoder_layers():
        blk, dl = layer.gated_cross_attn_layer, layer.decoder_layer
        ws = [dl.attention.query_key_value.weight, dl.attention.dense.weight, dl.mlp.dense_h_to_4h.weight,
              dl.mlp.dense_4h_to_h.weight]
        if blk is not None:
            ws += [blk.attn.to_q.weight, blk.attn.to_out.weight, blk.ff[1].weight, blk.ff[3].weight]
        w_bytes += sum(w.numel() * w.element_size() for w in ws)
    w_bytes += lm.embed_out.weight.numel() * lm.embed_out.weight.element_size()
    n_lm = len(list(lm._get_decoder_layers()))
    kv_bytes = int(2 * n_lm * beams * cfg.lm_hidden * (L + new / 2) * 2)
    _pk = load_peaks()
    peak_gbs, peak_src = _pk["hbm_gbs"], _pk["source"]
    floor_ms = (w_bytes + kv_bytes) / (peak_gbs * 1e9) * 1e3
    kernels = None
    if not args.no_kernel_profile:
        from unimp_b200 import kbench
        del dec
        torch.cuda.empty_cache()
        kernels = kbench.run(cfg, wl, _pk, only={"decode"}, eager=not args.no_eager_baseline)
    line = {"metric": "decode tokens/s (C4: explanation generation, beams 5)", "unit": "tokens/s",
            "value": 1e3 / per_tok, "higher_is_better": True, "n_gpus": 1, "dtype": "bf16",
            "data": "synthetic", "steps": reps, "warmup": 2,
            "config": {"workload": f"C4-decode: {cfg.name}, batch 1, beams {beams}, prompt {L} tokens "
                                   f"(T=512 padded), Ti=5 images, {new} new tokens, cached vision latents + "
                                   "cached cross-attention K/V, one CUDA-graph replay per token; beams re-ordered through an "
                                   "indirection table (unimp_lm_decode_attn), projections on unimp_linear_small_m, "
                                   "programmatic dependent launch between the step's kernels",
                       "value_is": "steady-state tokens/s of the graph replays (CUDA events around the replay "
                                   "loop, median of runs); prefill and graph capture reported separately"},
            "median_ms": med, "ms_per_token": per_tok,
            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak_gbs, "peak_source": peak_src,
                         "weight_bytes_per_token": w_bytes, "kv_bytes_per_token": kv_bytes,
                         "achieved": (w_bytes + kv_bytes) / (per_tok * 1e-3) / 1e9,
                         "frac": floor_ms / per_tok, "hbm_floor_ms_per_token": floor_ms,
                         "what": "whole decode step (one CUDA-graph replay: ~460 kernels) against streaming "
                                 "its weights and K/V rows once"},
            "end_to_end_tokens_per_s": {"incl_prefill_and_capture": new / (med["wall_ms"] * 1e-3),
                                        "excl_capture": new / ((med["prefill_ms"] + med["replay_ms"]) * 1e-3)},
            "runs": runs, "hf_generate_path": hf, "kernels": kernels}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# launch counting: how many of OUR kernels one step launches
# ---------------------------------------------------------------------------------------------

class LaunchCounter:
    """Counts the kernels launched through the C ABI during one eager step (the graph replays the
    same launches).  Kernels per C-ABI call are the launch lists of the .cu files."""
    # kernels per call: focal CE forward = row pass + fixed-order finish; LN backward = row pass
    # (+ the column/gate fold only when d_gate / d_gamma / d_beta are requested: args 10-12, g_ln = 1)
    PER_CALL = {"unimp_focal_ce_fwd": lambda a: 2, "unimp_focal_ce_rows_fwd": lambda a: 2,
                "unimp_lm_attn_bwd": lambda a: 2,     # delta = rowsum(dO o O), then the gradient kernel
                "unimp_gate_residual_ln_bwd": lambda a: 2 if (((a[11] or a[12]) and a[1]) or a[10]) else 1}

    def __init__(self):
        self.calls = {}

    def install(self):
        from unimp_b200 import _lib
        lib = _lib.load()
        self.lib, self.orig = lib, {}
        for name in _lib.SIGNATURES:
            if name in ("unimp_version", "unimp_last_error_string", "unimp_device_ok") or \
               name.endswith("_workspace") or name.endswith("_supported"):
                continue
            fn = getattr(lib, name)
            self.orig[name] = fn
            setattr(lib, name, self._wrap(name, fn))

    def uninstall(self):
        for n, f in self.orig.items():
            setattr(self.lib, n, f)

    def _wrap(self, name, fn):
        per = self.PER_CALL.get(name, lambda a: 1)

        def w(*a):
            self.calls[name] = self.calls.get(name, 0) + per(a)
            return fn(*a)
        return w

    def kernels(self):
        return dict(self.calls)


# ---------------------------------------------------------------------------------------------

def main():
    args = parse()
    from unimp_b200 import openflamingo_4b_config, tiny_config
    from unimp_b200.config import WORKLOADS
    import copy

    cfg = openflamingo_4b_config() if args.model == "4b" else tiny_config()
    wl = copy.copy(WORKLOADS[args.workload])
    if args.batch:
        wl.B = args.batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "train samples/s"
    config = {"workload": f"{wl.name}: {cfg.name}, per-GPU micro-batch {wl.B} x accum {args.accum}, "
                          f"Ti={wl.Ti} images/sample, T={wl.T} tokens, V={cfg.vocab}, gamma={wl.gamma}, "
                          f"fwd+focal loss+bwd+allreduce+clip+AdamW",
              "per_gpu_batch": wl.B, "accum": args.accum, "seq_len": wl.T, "images_per_sample": wl.Ti,
              "parallelism": f"dp{world}",
              "l2_policy": "per-step working set (weights 8 GB + activations) far exceeds the 126 MB L2"}

    if args.mode == "decode":
        if rank == 0:
            decode_bench(args, cfg)
        return

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_oracle_samples_per_s(cfg, wl, batch=args.cpu_batch, steps=max(1, args.steps),
                                     warmup=args.warmup, accum=1)
        # same workload / model / shapes as our arm; what ONE CPU STEP covers is a bounded sample of
        # it (the contract's "each step a bounded sample of that workload"): say so in `config`
        config = dict(config)
        config["workload"] = config["workload"].replace(
            f"per-GPU micro-batch {wl.B} x accum {args.accum}",
            f"[CPU arm: each step = {args.cpu_batch} sample x accum 1 of the named per-GPU micro-batch "
            f"{wl.B} x accum {args.accum}; samples/s normalises]")
        config["step_sample"] = {"per_step_batch": args.cpu_batch, "accum": 1,
                                 "of": {"per_gpu_batch": wl.B, "accum": args.accum}}
        config["parallelism"] = f"host CPU, {r['cores']} threads, fp32 eager PyTorch (oracle port)"
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": "samples/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import copy
    import torch.distributed as dist
    from unimp_b200 import _lib
    from unimp_b200.factory import build_flamingo
    from unimp_b200.synth import make_batch
    from unimp_b200.train import (BucketedAllReduce, FlatAdamW, GraphedTrainStep, ShardedDataParallel,
                                  get_grouped_params, train_step)

    assert torch.cuda.is_available(), "bench.py (ours) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # everything (model build, warm-up, capture, replay, timing events) runs on ONE non-default
    # stream: the legacy default stream cannot take part in a graph capture that contains NCCL
    # (high priority: at N=1 the deferred AdamW pass runs on a low-priority stream underneath it)
    torch.cuda.set_stream(torch.cuda.Stream(priority=-1))
    assert _lib.load().unimp_device_ok() == 1
    torch.backends.cuda.matmul.allow_tf32 = True
    peaks = load_peaks()

    model = build_flamingo(cfg, dtype=torch.bfloat16, device="cuda", seed=0, gate=0.5)
    model.train()
    sharded = world > 1 and args.dp == "zero1"
    opt = FlatAdamW(get_grouped_params(model, 0.1), lr=2e-4, shard_world=world if sharded else 1,
                    allocate_states=not sharded)
    reducer = None
    if world > 1:
        reducer = (ShardedDataParallel(opt, deferred_gather_module=model.perceiver,
                                       gather_start_module=model.vision_encoder) if sharded
                   else BucketedAllReduce(opt))
        config["parallelism"] = f"dp{world}, " + (
            "sharded optimizer: bucketed reduce-scatter overlapped with backward, AdamW on 1/N, "
            "parameter all-gather overlapped with the next step's ViT forward"
            if sharded else "bucketed all-reduce overlapped with backward")
    tk = cfg.tokens

    n_batches = args.accum * 4
    host = [make_batch(cfg, wl, seed=1234 + rank * 1000 + i) for i in range(n_batches)]
    for b in host:
        for k in b:
            b[k] = b[k].pin_memory()
    dev = [{k: v.cuda(non_blocking=True) for k, v in b.items()} for b in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values()) * args.accum

    # head + loss fusion: static capacity of the valid-row gather = a bound over the
    # batches this run feeds (a real loader bounds it by max answer tokens per sample x B)
    from unimp_b200 import ops as _ops
    label_rows = None
    if args.label_rows != "off":
        if args.label_rows == "auto":
            per_mb = []
            for b in dev:
                lab = _ops.mask_labels(b["input_ids"], answer_token_id=tk.answer,
                                       endofchunk_token_id=tk.endofchunk, media_token_id=tk.media,
                                       pad_token_id=tk.pad)
                per_mb.append(int((lab[:, 1:] != -100).sum()))
            per_fwd = max(per_mb) * (args.accum if (not args.no_fuse_accum and args.accum > 1) else 1)
            label_rows = max(64, (per_fwd + 63) // 64 * 64)
        else:
            label_rows = int(args.label_rows)
    config["head_loss_fusion"] = (
        f"valid-label rows gathered before embed_out, static capacity {label_rows} rows per forward "
        f"(of {wl.B * (args.accum if not args.no_fuse_accum else 1) * wl.T}); out[0] (HF mean CE, "
        "mmrec.py:182) is produced from the same rows" if label_rows else "off (dense (B,T,V) logits)")

    use_graph = not args.no_graph
    graphed = None
    if use_graph:
        # whole optimizer step captured once (DESIGN.md "launch overhead"); replays read the
        # step's inputs from static device buffers that are refilled before every replay
        try:
            graphed = GraphedTrainStep(model, tk, opt, reducer, dev[:args.accum], gamma=wl.gamma,
                                       fuse_accum=not args.no_fuse_accum, label_rows=label_rows,
                                       defer_optimizer=(world == 1 and args.defer_optimizer))
        except Exception as ex:  # noqa: BLE001  (never fail the bench on a capture problem: say so)
            if world > 1:
                raise
            graphed, use_graph = None, False
            config["graph_capture_error"] = repr(ex)[:300]
            torch.cuda.synchronize()
    config["launch"] = "cuda-graph replay of the whole step" if use_graph else "eager"
    if graphed is not None and graphed.deferred:
        config["optimizer_overlap"] = (
            "clip + AdamW of step k runs at the start of replay k+1 on a low-priority stream underneath the "
            "frozen ViT forward (the Perceiver waits for it): every replay still applies exactly one "
            "optimizer step; identical parameters after flush() (tests/test_model_gpu.py::"
            "test_deferred_optimizer_graph_equals_eager_steps)")
    config["accum_window"] = (
        f"{args.accum} micro-batches of {wl.B} run as ONE forward/backward over {args.accum * wl.B} samples "
        "(identical arithmetic: per-sample ops, per-micro-batch loss normalisation kept; "
        "tests/test_model_gpu.py::test_fused_accumulation_window_equals_sequential_micro_batches)"
        if (not args.no_fuse_accum and args.accum > 1) else "sequential micro-batches")

    fuse = not args.no_fuse_accum and args.accum > 1

    def run_step(mbs):
        if graphed is not None:
            return graphed(mbs)
        return train_step(model, None, tk, opt, reducer, gamma=wl.gamma, accum_steps=args.accum,
                          micro_batches=mbs, fuse_accum=fuse, label_rows=label_rows)

    def step_resident(i):
        # inputs already resident in HBM (the graphed path copies them device->device into its
        # static buffers: 7 MB, inside the timed region)
        return run_step([dev[(i * args.accum + a) % n_batches] for a in range(args.accum)])

    def step_e2e(i):
        # host (pinned) -> device copy of this step's inputs, then device -> host read of the loss
        if graphed is not None:
            loss = graphed([host[(i * args.accum + a) % n_batches] for a in range(args.accum)])
        else:
            mbs = [{k: v.cuda(non_blocking=True) for k, v in host[(i * args.accum + a) % n_batches].items()}
                   for a in range(args.accum)]
            loss = run_step(mbs)
        return loss.item()

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms)

    for i in range(args.warmup):
        step_resident(i)
    # soak: K x step can be < 1 s, too short for the clocks to settle under the 1 kW cap; run
    # untimed steps for --soak-s first so that the timed steps see the sustained state
    torch.cuda.synchronize()
    n_soak, t_soak = 0, time.time()
    while time.time() - t_soak < args.soak_s:
        for _ in range(4):
            step_resident(n_soak)
            n_soak += 1
        torch.cuda.synchronize()
    if world > 1:
        ns = torch.tensor([n_soak], device="cuda")
        dist.all_reduce(ns, op=dist.ReduceOp.MAX)
    config["soak"] = f"{n_soak} untimed steps ({args.soak_s:.1f} s) between warm-up and the timed region"
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if args.ncu_range:
        torch.cuda.profiler.start()
    ms = timed(step_resident, args.steps)
    if args.ncu_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if rank == 0 else None
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)

    samples_per_step = args.accum * wl.B * world
    value = samples_per_step * args.steps / (ms * 1e-3)
    e2e_value = samples_per_step * args.steps / (ms_e2e * 1e-3)

    kernels, roofline, roofline_dominant, roofline_core, launches, per_step = {}, None, None, None, None, None
    if graphed is not None:
        graphed.flush()
    if not args.no_kernel_profile:
        cnt = LaunchCounter()
        cnt.install()
        train_step(model, None, tk, opt, reducer, gamma=wl.gamma, accum_steps=args.accum,
                   micro_batches=[dev[a % n_batches] for a in range(args.accum)], fuse_accum=fuse,
                   label_rows=label_rows)
        torch.cuda.synchronize()
        cnt.uninstall()
        per_step = cnt.kernels()
        launches = int(sum(per_step.values()) * args.steps)
        if rank == 0:
            from unimp_b200 import kbench, ops
            lab = ops.mask_labels(dev[0]["input_ids"], answer_token_id=tk.answer,
                                  endofchunk_token_id=tk.endofchunk, media_token_id=tk.media,
                                  pad_token_id=tk.pad)
            n_valid = int((lab[:, 1:] != -100).sum())
            wl_k = copy.copy(wl)   # the shapes the timed region actually launches
            if fuse:
                wl_k.B, n_valid = wl.B * args.accum, n_valid * args.accum
            kernels = kbench.run(cfg, wl_k, peaks, n_valid_rows=n_valid)
            # DRAM traffic per launch of the same kernel at the same shape, from the committed
            # ncu --set full capture (profiles/r1_ncu_full_kernels.csv); null if the shape differs
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(tpath):
                tj = json.load(open(tpath))
                if tj.get("shape") == {"B": wl_k.B, "T": wl_k.T, "Ti": wl_k.Ti}:
                    traffic = tj["bytes_per_launch"].get("xattn_fwd")
            kx = kernels["xattn_fwd"]
            # the kernel the metric names; at this workload (2.4 MB, 100 MFLOP per launch) it is
            # latency-bound: AI = 42 FLOP/B is left of the ridge (~210), so the bound is HBM
            roofline_core = {"kernel": "xattn_fwd_tc_kernel (unimp_xattn_fwd: the bare masked media-located "
                                       "cross-attention core, tcgen05+TMEM+TMA; launched by the step only "
                                       "where the fused kernel does not apply)",
                        "bound": "hbm", "achieved": kx["GB/s"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": kx["frac_of_hbm_peak"], "traffic": traffic,
                        "alg_bytes_per_launch": kx["alg_bytes_per_launch"], "avg_us": kx["avg_us"],
                        "tflops": kx["TFLOP/s"], "frac_of_bf16_peak": kx["frac_of_bf16_peak"],
                        "peak_source": peaks["source"],
                        "how": "K cold input sets (> L2) launched back to back from one CUDA graph, "
                               "CUDA events around the replay, avg = elapsed / K"}

            # the dominant kernel of OUR share of the step by time: clip+AdamW over the 1.15 B trainable
            # parameters (HBM-bound, 28 B/param), timed on the real flat buffers (32 GB >> L2)
            if world == 1:
                n_par = sum(g["flat_p"].numel() for g in opt.groups)
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                s_.record()
                reps_ = 3
                for _ in range(reps_):
                    for g in opt.groups:
                        _ops.adamw_step_(g["master"], g["flat_p"], g["flat_g"], g["m"], g["v"],
                                         hyper=opt.hyper, beta1=opt.betas[0], beta2=opt.betas[1],
                                         eps=opt.eps, weight_decay=g["weight_decay"],
                                         gnorm_sq=opt.gnorm_sq, max_norm=opt.max_grad_norm)
                e_.record()
                torch.cuda.synchronize()
                us = s_.elapsed_time(e_) * 1e3 / reps_
                byts = n_par * (2 * 2 + 6 * 4)
                roofline_dominant = {
                    "kernel": "adamw_kernel (unimp_adamw_step: clip + AdamW + bf16 working copy, one pass)",
                    "bound": "hbm", "achieved": byts / us / 1e3, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": byts / us / 1e3 / peaks["hbm_gbs"], "traffic": None,
                    "alg_bytes_per_launch": byts, "avg_us": us, "launches_per_step": len(opt.groups),
                    "share_of_step": us / (ms / args.steps * 1e3), "peak_source": peaks["source"],
                    "how": f"{reps_} passes over the real flat buffers ({n_par} params, {byts / 1e9:.1f} GB "
                           "per pass >> L2), CUDA events on the launching stream"}

            roofline = roofline_core
            kf = kernels.get("xattn_block_fwd")
            if kf is not None and per_step.get("unimp_xattn_block_fwd"):
                # the variant of the metric's kernel whose bound is the tensor pipe: to_q GEMM + masked
                # attention + to_out GEMM in one cluster kernel (what the timed step launches 16x)
                roofline = {
                    "kernel": "xattn_block_fwd_kernel (unimp_xattn_block_fwd: to_q -> masked media-located "
                              "attention -> to_out, 8-CTA clusters, tcgen05+TMEM+TMA multicast)",
                    "bound": "tensor", "achieved": kf["TFLOP/s"], "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": kf["frac_of_bf16_peak"], "traffic": None,
                    "alg_flops_per_launch": kf["flops_per_launch"], "alg_bytes_per_launch": kf["alg_bytes_per_launch"],
                    "avg_us": kf["avg_us"], "launches_per_step": per_step.get("unimp_xattn_block_fwd"),
                    "three_launch_form_us": kernels.get("xattn_three_launches", {}).get("avg_us"),
                    "peak_source": peaks["source"],
                    "how": "K cold input sets (> L2) launched back to back from one CUDA graph, CUDA events "
                           "around the replay, avg = elapsed / K"}
                tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
                if os.path.exists(tpath):
                    tj = json.load(open(tpath))
                    if tj.get("shape") == {"B": wl_k.B, "T": wl_k.T, "Ti": wl_k.Ti}:
                        roofline["traffic"] = tj["bytes_per_launch"].get("xattn_block_fwd")

    def finish():
        # a process group cannot be torn down cleanly while a captured graph still references its
        # communicator: flush, meet at a barrier, and leave
        sys.stdout.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        finish()
        return

    gpu_eager = None
    if world == 1 and not args.no_eager_baseline:
        try:
            del graphed
            torch.cuda.empty_cache()
            gpu_eager = gpu_eager_samples_per_s(cfg, wl, accum=args.accum)
            gpu_eager["speedup_ours_vs_eager"] = value / gpu_eager["value"]
        except Exception as ex:  # noqa: BLE001
            gpu_eager = {"value": None, "unit": "samples/s", "what": f"failed: {ex!r}"[:300]}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu_baseline = cpu_oracle_samples_per_s(cfg, wl, batch=args.cpu_batch, steps=2, warmup=1)
            cpu_baseline = {k: cpu_baseline[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # noqa: BLE001
            cpu_baseline = {"value": None, "unit": "samples/s", "cores": os.cpu_count(),
                            "kind": "port", "sample": f"failed: {ex!r}"}

    line = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "gpu_launches_per_step": per_step, "clocks": clocks,
            "roofline": roofline, "roofline_core": roofline_core, "roofline_dominant": roofline_dominant,
            "cpu_baseline": cpu_baseline,
            "gpu_eager_baseline": gpu_eager, "kernels": kernels}
    print(json.dumps(line))
    finish()


if __name__ == "__main__":
    main()
