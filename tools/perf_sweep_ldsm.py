"""Sweep of the transposing-load median kernel's planner knobs (development aid).
usage: python tools/perf_sweep_ldsm.py T [T ...]   with env COMBOS="strips:stages:blocks,..." (0 = planner default)"""
import os, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bgdebias_b200.ops  # noqa
from bgdebias_b200 import _cabi

def run(fr, offs, iters=int(os.environ.get('ITERS', 12))):
    torch.ops.bgdebias.temporal_median_varlen(fr, offs); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.ops.bgdebias.temporal_median_varlen(fr, offs); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return ts

N = int(os.environ.get("N", 230400)); V = int(os.environ.get("V", 64))
combos = [tuple(int(x) for x in c.split(":")) for c in os.environ.get("COMBOS", "1:0:0").split(",")]
_cabi.set_median_variant(int(os.environ.get("VARIANT", 4)))
rng = np.random.default_rng(1)
for case in sys.argv[1:]:
    Ts = rng.integers(120, 241, V) if case == "mixed" else np.full(V, int(case))
    offs = torch.from_numpy(np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64))
    rows = int(offs[-1])
    fr = torch.randint(0, 256, (rows, N), dtype=torch.uint8, device="cuda")
    by = (rows + V) * N
    res = {c: [] for c in combos}
    for rnd in range(int(os.environ.get('REPEAT', 1))):
        for c in (combos if rnd % 2 == 0 else combos[::-1]):
            strips, stages, blocks = c
            os.environ["BGD_LDSM_STRIPS"] = str(strips); os.environ["BGD_LDSM_STAGES"] = str(stages); os.environ["BGD_LDSM_BLOCKS"] = str(blocks)
            try:
                res[c] += run(fr, offs)
            except Exception as e:
                print(f"T={case} strips={strips} stages={stages} blocks={blocks}: ERROR {e}", flush=True)
    for (strips, stages, blocks), ts in res.items():
        if not ts: continue
        ts.sort(); ms, med = ts[0], ts[len(ts) // 2]
        print(f"T={case:>5s} strips={strips} stages={stages} blocks={blocks}: {ms:7.3f} ms {by/ms/1e6:7.1f} GB/s (median {by/med/1e6:7.1f}) {rows/ms/1e3:6.2f} Mframes/s", flush=True)
    del fr
