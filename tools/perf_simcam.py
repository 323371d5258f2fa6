"""Device-resident timing of the NaN-masked temporal median / mean (sim_cam extraction) next to numpy's
nanmedian / nanmean on the host (what the reference calls, cil_tools/extract_background.py:94-98).
usage: python tools/perf_simcam.py [V=256]"""
import sys, pathlib, time, json, warnings
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bgdebias_b200.ops  # noqa

V = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = 100 * 100 * 3
rng = np.random.default_rng(0)
Ts = rng.integers(50, 151, V)
offs = torch.from_numpy(np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64))
rows = int(offs[-1])
x = torch.rand((rows, N), device="cuda") * 255.0
x[torch.rand((rows, N), device="cuda") < 0.25] = 0.0
for avg, name in ((0, "median"), (1, "mean")):
    for _ in range(2): out = torch.ops.bgdebias.nan_temporal_reduce_varlen(x, offs, avg, True)
    torch.cuda.synchronize(); ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = torch.ops.bgdebias.nan_temporal_reduce_varlen(x, offs, avg, True); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]; by = rows * N * 4 + V * N
    # CPU arm: the reference's call on a bounded sample (8 folders), one thread
    k = min(8, V); t0 = time.perf_counter(); fr_cpu = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for v in range(k):
            f = x[int(offs[v]):int(offs[v + 1])].cpu().numpy().reshape(-1, 100, 100, 3).copy(); f[f == 0] = np.nan
            t1 = time.perf_counter()
            (np.nanmedian if avg == 0 else np.nanmean)(list(f), axis=0).astype(np.uint8)
            t0 += 0; fr_cpu += f.shape[0]; cpu_t = (cpu_t if v else 0.0) + (time.perf_counter() - t1)
    print(json.dumps({"op": "nan_" + name, "folders": V, "frames": rows, "ms": ms, "frames_per_s": rows / ms * 1e3,
                      "GB/s": by / ms / 1e6, "frac_of_6549.8": by / ms / 1e6 / 6549.8,
                      "cpu_frames_per_s_1thread": fr_cpu / cpu_t, "cpu_sample": f"{k} folders, {fr_cpu} frames"}), flush=True)
