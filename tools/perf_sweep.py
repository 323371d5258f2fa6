"""In-process sweep of the median planner knobs (development aid)."""
import os, sys, pathlib, itertools
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bgdebias_b200.ops  # noqa
from bgdebias_b200 import _cabi

def run(fr, offs, iters=4):
    torch.ops.bgdebias.temporal_median_varlen(fr, offs); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.ops.bgdebias.temporal_median_varlen(fr, offs); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

N = int(os.environ.get("N", 230400)); V = int(os.environ.get("V", 64))
cases = sys.argv[1:] or ["180", "181", "mixed"]
Rs = [int(x) for x in os.environ.get("RS", "6,8,10,12").split(",")]
THs = [int(x) for x in os.environ.get("THS", "192,256,384").split(",")]
Cs = [int(x) for x in os.environ.get("CS", "2,3").split(",")]
rng = np.random.default_rng(1)
for case in cases:
    Ts = rng.integers(120, 241, V) if case == "mixed" else np.full(V, int(case))
    offs = torch.from_numpy(np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64))
    rows = int(offs[-1])
    fr = torch.randint(0, 256, (rows, N), dtype=torch.uint8, device="cuda")
    by = (rows + V) * N
    for var in [int(v) for v in os.environ.get("VARIANTS", "3").split(",")]:
        if var == 2: continue
        _cabi.set_median_variant(var)
        for thr, tc2, tc4 in itertools.product([int(v) for v in os.environ.get("COLTHR", "128,256").split(",")],
                                               [int(v) for v in os.environ.get("TC2", "240").split(",")],
                                               [int(v) for v in os.environ.get("TC4", "96").split(",")]):
            os.environ["BGD_COL_THREADS_C2"] = str(thr); os.environ["BGD_COL_THREADS_C4"] = str(thr)
            os.environ["BGD_COL_T_C2"] = str(tc2); os.environ["BGD_COL_T_C4"] = str(tc4)
            try:
                ms = run(fr, offs)
                print(f"T={case:>5s} variant={var} col_threads={thr} t_c2={tc2} t_c4={tc4}: {ms:7.3f} ms {by/ms/1e6:7.1f} GB/s {rows/ms/1e3:6.2f} Mframes/s", flush=True)
            except Exception as e:
                print(f"T={case} variant={var} thr={thr}: ERROR {e}", flush=True)
    _cabi.set_median_variant(2)
    for R, TH, C in (itertools.product(Rs, THs, Cs) if "2" in os.environ.get("VARIANTS", "3").split(",") else []):
        os.environ["BGD_MEDIAN_TARGET_R"] = str(R); os.environ["BGD_MEDIAN_TARGET_THREADS"] = str(TH); os.environ["BGD_MEDIAN_CTAS_PER_SM"] = str(C)
        try:
            ms = run(fr, offs)
            print(f"T={case:>5s} R={R:2d} thr={TH:3d} ctas={C}: {ms:7.3f} ms {by/ms/1e6:7.1f} GB/s {rows/ms/1e3:6.2f} Mframes/s", flush=True)
        except Exception as e:
            print(f"T={case} R={R} thr={TH} ctas={C}: ERROR {e}", flush=True)
    del fr
