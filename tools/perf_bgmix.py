"""Device-resident timing of the BG-mix blend (development aid): configs[4] shape, L2 flushed between iterations.
usage: [BGD_BGMIX_NO_PERSIST=1] python tools/perf_bgmix.py [B]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import bgdebias_b200.ops as ops

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Tm, Hm, Wm, P = 8, 224, 224, 1024
gm = torch.Generator(device=dev).manual_seed(4)
fg = torch.randint(0, 256, (B, Tm, Hm, Wm, 3), dtype=torch.uint8, device=dev, generator=gm)
pool = torch.rand((P, 3, 256, 341), device=dev, generator=gm) * 255.0
torch.manual_seed(0)
idx = torch.randint(0, P, (B,)).int().to(dev); top = torch.randint(0, 33, (B,)).int().to(dev)
left = torch.randint(0, 118, (B,)).int().to(dev); app = torch.ones(B, dtype=torch.uint8, device=dev)
mean, std = torch.tensor([123.675, 116.28, 103.53]), torch.tensor([58.395, 57.12, 57.375])
lut = ops.make_fg_lut(mean.tolist(), std.tolist(), dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
mix = lambda: torch.ops.bgdebias.bgmix_blend(fg, pool, idx, top, left, app, lut, mean, std, 0.5, "NTCHW")
for _ in range(3): mix()
ts = []
for _ in range(30):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); o = mix(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort(); by = B * (Tm * Hm * Wm * 3 * 5 + Hm * Wm * 3 * 4)
print(f"B={B}: min {ts[0]*1e3:.1f} us {by/ts[0]/1e6:.0f} GB/s | median {ts[len(ts)//2]*1e3:.1f} us {by/ts[len(ts)//2]/1e6:.0f} GB/s  checksum {float(o.double().sum()):.6f}")
