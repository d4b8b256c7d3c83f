"""Why does the graph decoder stop early on the 4B bf16 model? Tries mask fill values x SDPA backends (dev tool)."""
import os, sys, time, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.nn.attention import SDPBackend, sdpa_kernel
from unimp_b200 import openflamingo_4b_config
from unimp_b200.config import Workload
from unimp_b200.factory import build_flamingo
from unimp_b200.synth import make_batch
from unimp_b200.decode import GraphedDecoder

cfg = openflamingo_4b_config()
wl = Workload("C4-decode", B=1, Ti=5, T=512)
model = build_flamingo(cfg, dtype=torch.bfloat16, device="cuda", gate=0.5).eval()
b = make_batch(cfg, wl, seed=0)
L = int(b["attention_masks"][0].sum()) - 2
ids = b["input_ids"][:, :L].cuda(); vis = b["patch_images"].unsqueeze(2).cuda()
new = int(os.environ.get("NEW", 48))
for fill_name, fill in (("-inf", float("-inf")), ("finfo.min", torch.finfo(torch.bfloat16).min)):
    for bname, be in (("default", None), ("math", SDPBackend.MATH), ("efficient", SDPBackend.EFFICIENT_ATTENTION),
                      ("cudnn", SDPBackend.CUDNN_ATTENTION)):
        GraphedDecoder.MASK_FILL = fill
        dec = GraphedDecoder(model)
        ctx = sdpa_kernel([be]) if be is not None else contextlib.nullcontext()
        try:
            with ctx:
                for n in (new, 3 * new):
                    torch.cuda.synchronize(); t0 = time.time()
                    out = dec.generate(vis, ids, torch.ones_like(ids), num_beams=5, max_new_tokens=n, eos_token_id=-1,
                                       pad_token_id=cfg.tokens.pad, early_stopping=False)
                    torch.cuda.synchronize(); dt = time.time() - t0
                    print(f"fill={fill_name:9s} sdpa={bname:9s} new={n:4d}: {dt*1e3:8.1f} ms steps={dec.last_n_steps} "
                          f"out_len={out.shape[1]-L} finite={dec.last_logits_finite}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"fill={fill_name} sdpa={bname}: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
