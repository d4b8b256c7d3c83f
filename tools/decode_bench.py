"""Config 4 (explanation generation, reference eval_exp.py:101-114: beams 5) decode timing (dev tool).
Cached vision latents + cached x-attn K/V + single-token decode kernel vs recomputing to_kv(media) and
the full masked attention every step (what upstream does)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unimp_b200 import openflamingo_4b_config
from unimp_b200.config import Workload
from unimp_b200.factory import build_flamingo
from unimp_b200.synth import make_batch
from unimp_b200 import helpers

cfg = openflamingo_4b_config()
wl = Workload("C4-decode", B=1, Ti=5, T=512)
new = int(os.environ.get("NEW", 64)); beams = 5
model = build_flamingo(cfg, dtype=torch.bfloat16, device="cuda", gate=0.5).eval()
b = make_batch(cfg, wl, seed=0)
L = int(b["attention_masks"][0].sum()) - 2
ids = b["input_ids"][:, :L].cuda(); vis = b["patch_images"].unsqueeze(2).cuda()
kw = dict(attention_mask=torch.ones_like(ids), num_beams=beams, max_new_tokens=new, min_new_tokens=new,
          eos_token_id=cfg.tokens.endofchunk, pad_token_id=cfg.tokens.pad, do_sample=False, early_stopping=False)

def run():
    torch.cuda.synchronize(); t0 = time.time()
    out = model.generate(vision_x=vis, lang_x=ids, **kw)
    torch.cuda.synchronize()
    return time.time() - t0, out

run()
t_cached, out_a = run()
# upstream-style: no K/V cache, full masked attention for the single token
orig = helpers.MaskedCrossAttention.forward
def no_cache_forward(self, x, media, media_locations=None, use_cached_media=False, text_time=None, x_ln=None):
    self._kv_cache = None
    B, T, D = x.shape
    _, Ti, n = media.shape[:3]
    if text_time is None:
        from unimp_b200 import ops
        text_time = ops.text_time(media_locations.to(torch.int64), 1, use_cached=use_cached_media, T_out=T)
    from unimp_b200 import ops
    xl = ops.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
    q = torch.nn.functional.linear(xl, self.to_q.weight)
    kv = self.project_media(media)
    out = ops.masked_cross_attention(q, kv, text_time, heads=self.heads, n_latents=n, scale=self.scale)
    return torch.nn.functional.linear(out, self.to_out.weight)
helpers.MaskedCrossAttention.forward = no_cache_forward
run()
t_plain, out_b = run()
helpers.MaskedCrossAttention.forward = orig
# the CUDA-graph decoder (unimp_b200/decode.py): same decoding rules, one graph replay per token
from unimp_b200.decode import GraphedDecoder
dec = GraphedDecoder(model)
def run_graphed(n):
    torch.cuda.synchronize(); t0 = time.time()
    out = dec.generate(vis, ids, torch.ones_like(ids), num_beams=beams, max_new_tokens=n, eos_token_id=-1,
                       pad_token_id=cfg.tokens.pad, early_stopping=False)
    torch.cuda.synchronize()
    return time.time() - t0, out
try:
    run_graphed(new)
    t_graph, out_c = run_graphed(new)
    t_long, _ = run_graphed(3 * new)
    per_tok = (t_long - t_graph) / (2 * new)
    print(f"cuda-graph decoder         : {t_graph*1e3:8.1f} ms  ({new/t_graph:6.1f} tokens/s incl. prefill + capture; "
          f"steady state {per_tok*1e3:.2f} ms/token = {1/per_tok:.1f} tokens/s)"
          f"  same tokens as HF path: {out_c.shape == out_a.shape and torch.equal(out_c, out_a)}")
except Exception:
    import traceback; traceback.print_exc()
print(f"prompt {L} tokens, Ti={wl.Ti}, beams {beams}, {new} new tokens")
print(f"cached K/V + decode kernel : {t_cached*1e3:8.1f} ms  ({new/t_cached:6.1f} tokens/s)")
print(f"recompute to_kv every step : {t_plain*1e3:8.1f} ms  ({new/t_plain:6.1f} tokens/s)")
print("same tokens:", torch.equal(out_a, out_b))
