"""Device-resident timing of the foreground tail (development aid): 64 clips x 8 frames of MultiScaleCrop-sized uint8 crops
-> 224x224, (a) fused resize+blend, (b) resize then blend (two launches), (c) the blend alone on pre-resized clips.
L2 flushed between iterations.  usage: python tools/perf_resize.py [B]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import bgdebias_b200.ops as ops

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Tm, Hm, Wm, P = 8, 224, 224, 1024
sizes = [(256, 256), (224, 256), (256, 224), (224, 224), (192, 224), (224, 192), (192, 192), (168, 192), (192, 168), (168, 168)]
g = torch.Generator().manual_seed(4)
clips = [torch.randint(0, 256, (Tm, *sizes[b % len(sizes)], 3), dtype=torch.uint8, generator=g) for b in range(B)]
buf, geom = ops.pack_clips(clips)
d_buf = buf.to(dev)
gm = torch.Generator(device=dev).manual_seed(4)
pool = torch.rand((P, 3, 256, 341), device=dev, generator=gm) * 255.0
torch.manual_seed(0)
idx = torch.randint(0, P, (B,)).int().to(dev); top = torch.randint(0, 33, (B,)).int().to(dev)
left = torch.randint(0, 118, (B,)).int().to(dev); app = torch.ones(B, dtype=torch.uint8, device=dev)
mean, std = torch.tensor([123.675, 116.28, 103.53]), torch.tensor([58.395, 57.12, 57.375])
lut = ops.make_fg_lut(mean.tolist(), std.tolist(), dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tail = (pool, idx, top, left, app, lut, mean, std, 0.5, "NTCHW")
fused = lambda: torch.ops.bgdebias.bgmix_resize_blend(d_buf, geom, Tm, Hm, Wm, *tail)
resize = lambda: torch.ops.bgdebias.resize_bilinear(d_buf, geom, Tm, Hm, Wm)
fg224 = resize()
blend = lambda: torch.ops.bgdebias.bgmix_blend(fg224, *tail)
two = lambda: torch.ops.bgdebias.bgmix_blend(resize(), *tail)
assert torch.equal(fused(), two())
src_bytes = sum(c.numel() for c in clips)
out_bytes = B * Tm * Hm * Wm * 3 * 4
bg_bytes = B * Hm * Wm * 3 * 4
for name, fn, by in (("fused resize+blend", fused, src_bytes + bg_bytes + out_bytes),
                     ("resize, then blend", two, src_bytes + 2 * B * Tm * Hm * Wm * 3 + bg_bytes + out_bytes),
                     ("resize only (u8)", resize, src_bytes + B * Tm * Hm * Wm * 3),
                     ("blend only", blend, B * Tm * Hm * Wm * 3 + bg_bytes + out_bytes)):
    for _ in range(3): fn()
    ts = []
    for _ in range(30):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); o = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"{name:20s} B={B}: min {ts[0]*1e3:7.1f} us {by/ts[0]/1e6:6.0f} GB/s | median {ts[len(ts)//2]*1e3:7.1f} us "
          f"{by/ts[len(ts)//2]/1e6:6.0f} GB/s  ({by/1e6:.1f} MB)", flush=True)
