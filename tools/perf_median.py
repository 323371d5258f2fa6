"""Quick device-resident timing of the median kernels (development aid, not the bench contract).

usage: python tools/perf_median.py [--T 180] [--V 64] [--N 230400] [--variant 0|1|2] [--iters 5] [--mixed]
"""
import argparse, os, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bgdebias_b200.ops  # noqa
from bgdebias_b200 import _cabi

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=180); ap.add_argument("--V", type=int, default=64)
ap.add_argument("--N", type=int, default=230400); ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--iters", type=int, default=5); ap.add_argument("--mixed", action="store_true")
ap.add_argument("--content", default="random", choices=["random", "static"],
                help="static: one background per video + camera-like noise of +-1 on 5 %% of the samples + a moving block")
a = ap.parse_args()
rng = np.random.default_rng(1)
Ts = rng.integers(120, 241, a.V) if a.mixed else np.full(a.V, a.T)
offs = torch.from_numpy(np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64))
rows = int(offs[-1])
fr = torch.randint(0, 256, (rows, a.N), dtype=torch.uint8, device="cuda")
if a.content == "static":
    for v in range(a.V):
        r0, r1 = int(offs[v]), int(offs[v + 1])
        bg = torch.randint(1, 255, (1, a.N), dtype=torch.int16, device="cuda")
        noise = torch.randint(-1, 2, (r1 - r0, a.N), dtype=torch.int16, device="cuda") * (torch.rand((r1 - r0, a.N), device="cuda") < 0.05)
        fr[r0:r1] = (bg + noise).to(torch.uint8)
        for t in range(r0, r1, 1):
            c0 = (977 * (t - r0)) % (a.N - 4800)
            fr[t, c0:c0 + 4800] = 255 - (t % 5)
_cabi.set_median_variant(a.variant)
out = torch.ops.bgdebias.temporal_median_varlen(fr, offs); torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = torch.ops.bgdebias.temporal_median_varlen(fr, offs); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = min(ts); by = (rows + a.V) * a.N
print(f"content={a.content} variant={a.variant} T={'mixed' if a.mixed else a.T} V={a.V} N={a.N} env R={os.environ.get('BGD_MEDIAN_TARGET_R')} thr={os.environ.get('BGD_MEDIAN_TARGET_THREADS')} ctas={os.environ.get('BGD_MEDIAN_CTAS_PER_SM')}: "
      f"{ms:.3f} ms  {by/ms/1e6:.1f} GB/s  {rows/ms*1e3/1e6:.2f} Mframes/s (median of iters {sorted(ts)[len(ts)//2]:.3f} ms)")
