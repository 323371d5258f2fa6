"""Latency of the per-video drop-in call (a list of decoded frames in pageable host memory -> background on the host),
what bg_extraction_tmf does after decoding, next to np.median on one core.  usage: python tools/perf_single_video.py [T=180]"""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
from bgdebias_b200 import extract_background as eb
T = int(sys.argv[1]) if len(sys.argv) > 1 else 180
rng = np.random.default_rng(0)
frames = [rng.integers(0, 256, (240, 320, 3), dtype=np.uint8) for _ in range(T)]
out = eb.temporal_median_frames(frames)           # warm-up: CUDA context, pinned slabs
ts = []
for _ in range(10):
    t0 = time.perf_counter(); out = eb.temporal_median_frames(frames); ts.append(time.perf_counter() - t0)
t0 = time.perf_counter(); ref = np.median(frames, axis=0).astype(np.uint8); t_cpu = time.perf_counter() - t0
print(f"T={T}: GPU path {min(ts) * 1e3:.2f} ms (median {sorted(ts)[5] * 1e3:.2f}) = {T / min(ts):.0f} frames/s; np.median {t_cpu * 1e3:.0f} ms = {T / t_cpu:.0f} frames/s; "
      f"speed-up {t_cpu / min(ts):.0f}x; identical {bool(np.array_equal(out, ref))}")
