"""Device-resident timing of the ActorCutMix blend (64 clips x 8 x 224 x 224 x 3, L2 flushed between iterations)
next to the reference's numpy loop on one host core.  usage: python tools/perf_cutmix.py"""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bgdebias_b200.ops  # noqa
B, T, H, W = 64, 8, 224, 224
n = B * T * H * W * 3
a, s = (torch.randint(0, 256, (B * T, H, W, 3), dtype=torch.uint8, device="cuda") for _ in range(2))
m = (torch.rand((B * T, H, W, 1), device="cuda") > 0.8).to(torch.uint8).expand(-1, -1, -1, 3).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): out, tot = torch.ops.bgdebias.actor_cut_mix(a, m, s)
ts = []
for _ in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out, tot = torch.ops.bgdebias.actor_cut_mix(a, m, s); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
an, mn, sn = (x[:T].cpu().numpy() for x in (a, m, s))
t0 = time.perf_counter()
for _ in range(5):
    for f in range(T): an[f] * mn[f] + sn[f] * (1 - mn[f])
    sum(mn[f][:, :, 0].sum() for f in range(T))
cpu = 5 / (time.perf_counter() - t0)
print(f"actor_cut_mix {B} clips: {ms * 1e3:.1f} us = {4 * n / ms / 1e6:.0f} GB/s ({4 * n / ms / 1e6 / 6549.8:.2f} of the measured HBM peak), "
      f"{B / ms * 1e3:.0f} clips/s; numpy loop on one core: {cpu:.0f} clips/s; mask sum ok: {int(tot) == int(m[..., 0].sum())}")
