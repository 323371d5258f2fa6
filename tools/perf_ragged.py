"""Device-resident timing of the ragged-pool BG-mix blend (development aid): configs[4] shape, 1,024 uint8 backgrounds of
mixed widths, Resize(256) inside the launch; L2 flushed between iterations.  usage: python tools/perf_ragged.py [B]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bgdebias_b200.ops as ops
from bgdebias_b200.pool import RaggedPool

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Tm, Hm, Wm = 8, 224, 224
gm = torch.Generator(device=dev).manual_seed(4)
fg = torch.randint(0, 256, (B, Tm, Hm, Wm, 3), dtype=torch.uint8, device=dev, generator=gm)
g = torch.Generator().manual_seed(6)
sizes = [(240, 320), (240, 427), (240, 352), (240, 426)]
rp = RaggedPool(256, dev)
imgs = [torch.randint(0, 256, (3,) + sizes[i % 4], dtype=torch.uint8, generator=g) for i in range(64)]
for _ in range(16):
    rp.append(imgs)
rs = np.random.default_rng(9)
idx = rs.integers(0, len(rp), B); hw = [rp.hw(int(i)) for i in idx]
top = np.array([rs.integers(0, h - Hm + 1) for h, w in hw]); left = np.array([rs.integers(0, w - Wm + 1) for h, w in hw])
d = lambda a: torch.tensor(np.asarray(a), dtype=torch.int32, device=dev)
mean, std = torch.tensor([123.675, 116.28, 103.53]), torch.tensor([58.395, 57.12, 57.375])
lut = ops.make_fg_lut(mean.tolist(), std.tolist(), dev)
app = torch.ones(B, dtype=torch.uint8, device=dev)
args = (rp.data, rp.slots_tensor, rp.tables.tensor, d(idx), d(top), d(left), app, lut, mean, std, 0.5, "NTCHW")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
mix = lambda: torch.ops.bgdebias.bgmix_blend_ragged(fg, *args)
for _ in range(3): mix()
ts = []
for _ in range(30):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); o = mix(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort(); by = B * (Tm * Hm * Wm * 3 * 5 + 3 * 212 * 250)
print(f"ragged B={B}: min {ts[0]*1e3:.1f} us {by/ts[0]/1e6:.0f} GB/s | median {ts[len(ts)//2]*1e3:.1f} us {by/ts[len(ts)//2]/1e6:.0f} GB/s")
