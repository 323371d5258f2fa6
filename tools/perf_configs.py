"""Resident-shard timing of the other BASELINE.json extraction configs (parity-test cases, not the bench line):
  hmdb51 : configs[2]  6,766 videos / 8 GPUs, even T = 2 * U[30,80], 240x320x3
  sthv2  : configs[3]  ~220k clips / 8 GPUs, T = U[24,72], 240x427x3 (N = 307,440); timed on a 4,096-clip chunk
usage: python tools/perf_configs.py [hmdb51] [sthv2] [ucf101]"""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bgdebias_b200.ops  # noqa
from oracle import c_oracle

CFG = {
    "ucf101": dict(V=1665, N=240 * 320 * 3, T=lambda r, v: r.integers(120, 241, v), seed=1),
    "hmdb51": dict(V=846, N=240 * 320 * 3, T=lambda r, v: 2 * r.integers(30, 81, v), seed=2),
    "sthv2": dict(V=4096, N=240 * 427 * 3, T=lambda r, v: r.integers(24, 73, v), seed=3),
}
dev = torch.device("cuda:0")
for name in (sys.argv[1:] or ["hmdb51", "sthv2"]):
    c = CFG[name]; rng = np.random.default_rng(c["seed"])
    Ts = c["T"](rng, c["V"]); offs = torch.from_numpy(np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64))
    rows, N, V = int(offs[-1]), c["N"], c["V"]
    fr = torch.empty((rows, N), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(7); ch = max(1, (1 << 30) // N)
    for r0 in range(0, rows, ch):
        fr[r0:r0 + ch] = torch.randint(0, 256, (min(ch, rows - r0), N), dtype=torch.uint8, device=dev, generator=g)
    for _ in range(3): out = torch.ops.bgdebias.temporal_median_varlen(fr, offs)
    torch.cuda.synchronize(); ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = torch.ops.bgdebias.temporal_median_varlen(fr, offs); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]; by = (rows + V) * N
    v = int(np.argmin(Ts)); sl = slice(int(offs[v]), int(offs[v + 1]))
    ok = bool(np.array_equal(out[v].cpu().numpy(), c_oracle.temporal_median(fr[sl].cpu().numpy())))
    print(json.dumps({"config": name, "videos": V, "frames": rows, "N": N, "resident_gb": rows * N / 1e9, "ms": ms,
                      "frames_per_s": rows / ms * 1e3, "GB/s": by / ms / 1e6, "frac_of_6549.8": by / ms / 1e6 / 6549.8, "parity_spotcheck": ok}), flush=True)
    del fr, out
