"""Isolated kernel timings (unimp_b200/kbench.py) without building the model (dev tool).
    python tools/kbench_cli.py --workload C3-multitask --accum 2 --only xattn vit"""
import argparse, copy, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from unimp_b200 import kbench, openflamingo_4b_config
from unimp_b200.config import WORKLOADS

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C2-rec")
ap.add_argument("--accum", type=int, default=2, help="window size: the fused window runs accum x B samples per launch")
ap.add_argument("--only", nargs="*", default=None)
ap.add_argument("--no-eager", action="store_true")
ap.add_argument("--tag", default="")
a = ap.parse_args()
torch.cuda.set_stream(torch.cuda.Stream())
cfg = openflamingo_4b_config()
wl = copy.copy(WORKLOADS[a.workload])
wl.B *= a.accum
res = kbench.run(cfg, wl, bench.load_peaks(), eager=not a.no_eager, only=a.only)
print(json.dumps({"tag": a.tag, "workload": a.workload, "B": wl.B, "T": wl.T, "Ti": wl.Ti, "kernels": res}))
for k, v in res.items():
    print(f"KB {a.tag} {a.workload} {k}: {v['avg_us']:.2f} us  {v['GB/s']:.0f} GB/s ({v['frac_of_hbm_peak']:.3f} of HBM)"
          + (f"  {v['TFLOP/s']:.1f} TFLOP/s" if 'TFLOP/s' in v else "")
          + (f"  x{v['speedup_vs_eager']:.2f} vs eager" if 'speedup_vs_eager' in v else ""), file=sys.stderr)
