"""Randomised differential test of the AUTO median path against the C oracle: varlen batches of random
composition (T 1..560, ragged N, several value distributions).  usage: python tools/fuzz_median.py [seconds=60] [seed=0]"""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bgdebias_b200.ops  # noqa
from oracle import c_oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t_end = time.time() + budget
cases = videos = 0
while time.time() < t_end:
    V = int(rng.integers(1, 12))
    N = 16 * int(rng.integers(1, 400)) if rng.random() < 0.85 else int(rng.integers(1, 3000))
    hi = int(rng.choice([20, 70, 260, 560]))
    Ts = rng.integers(1, hi + 1, V)
    offs = np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64)
    kind = rng.integers(0, 5)
    rows = int(offs[-1])
    if kind == 0: fr = rng.integers(0, 256, (rows, N), dtype=np.uint8)
    elif kind == 1: fr = rng.integers(120, 124, (rows, N), dtype=np.uint8)
    elif kind == 2: fr = rng.choice(np.array([0, 255], np.uint8), (rows, N))
    elif kind == 3: fr = np.broadcast_to(rng.integers(0, 256, (1, N), dtype=np.uint8), (rows, N)).copy()
    else: fr = (rng.integers(0, 256, (1, N)) + rng.integers(-3, 4, (rows, N))).clip(0, 255).astype(np.uint8)
    out = torch.ops.bgdebias.temporal_median_varlen(torch.from_numpy(fr).cuda(), torch.from_numpy(offs)).cpu().numpy()
    for v in range(V):
        exp = c_oracle.temporal_median(fr[offs[v]:offs[v + 1]])
        if not np.array_equal(out[v], exp):
            bad = np.flatnonzero(out[v] != exp)
            print(f"MISMATCH V={V} N={N} Ts={Ts.tolist()} kind={kind} video={v} T={Ts[v]} first bad column {bad[0]} got {out[v][bad[0]]} exp {exp[bad[0]]} ({bad.size} bad)")
            sys.exit(1)
    cases += 1; videos += V
print(f"fuzz ok: {cases} calls, {videos} videos")
