"""Randomised differential test of the ragged-pool blend and the on-device Resize against torchvision itself:
random image sizes (up- and down-scaling, portrait, sizes Resize leaves alone), random crops and batches.
usage: python tools/fuzz_ragged.py [seconds=60] [seed=0]"""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from torchvision.transforms import Resize
import bgdebias_b200.ops as ops
from bgdebias_b200.pool import RaggedPool, resized_hw
from oracle import bgmix_oracle as bo

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
dev = torch.device("cuda")
lut = ops.make_fg_lut(bo.DEFAULT_MEAN, bo.DEFAULT_STD, dev)
mean, std = torch.tensor(bo.DEFAULT_MEAN), torch.tensor(bo.DEFAULT_STD)
t_end = time.time() + budget
cases = clips = images = 0
while time.time() < t_end:
    size = int(rng.choice([16, 24, 40, 64, 100]))
    H = int(rng.integers(4, size + 1)); W = int(rng.integers(1, size // 4 + 1)) * 4 if rng.random() < 0.8 else int(rng.integers(4, size + 1))
    n_img = int(rng.integers(1, 6))
    imgs = []
    for _ in range(n_img):
        k = rng.integers(0, 4)
        if k == 0: h, w = size, size + int(rng.integers(0, 40))                      # one axis kept
        elif k == 1: h, w = int(rng.integers(size, 3 * size)), int(rng.integers(size, 3 * size))     # down-scaling
        elif k == 2: h, w = int(rng.integers(max(4, size // 3), size + 1)), int(rng.integers(max(4, size // 3), 2 * size))   # up-scaling
        else: h, w = int(rng.integers(size, 6 * size)), int(rng.integers(size, 2 * size))            # many taps, portrait
        kind = rng.integers(0, 3)
        img = rng.integers(0, 256, (3, h, w), dtype=np.uint8) if kind == 0 else (
            np.broadcast_to(rng.integers(0, 256, (3, 1, 1), dtype=np.uint8), (3, h, w)).copy() if kind == 1 else
            rng.choice(np.array([0, 255], np.uint8), (3, h, w)))
        imgs.append(img)
    rp = RaggedPool(size, dev)
    rp.append(imgs)
    resized = [Resize(size)(torch.from_numpy(im).float()).numpy() for im in imgs]
    for i, r in enumerate(resized):
        got = rp.resized(i).cpu().numpy()
        if got.shape != r.shape or not np.array_equal(got.view(np.uint32), r.view(np.uint32)):
            print(f"RESIZE MISMATCH seed={seed} case={cases} image {imgs[i].shape} -> {r.shape}"); sys.exit(1)
    B, T = int(rng.integers(1, 7)), int(rng.integers(1, 5))
    fg = rng.integers(0, 256, (B, T, H, W, 3), dtype=np.uint8)
    ok_imgs = [i for i, r in enumerate(resized) if r.shape[1] >= H and r.shape[2] >= W]
    if not ok_imgs:
        continue
    idx = rng.choice(ok_imgs, B)
    top = np.array([rng.integers(0, resized[i].shape[1] - H + 1) for i in idx]); left = np.array([rng.integers(0, resized[i].shape[2] - W + 1) for i in idx])
    app = (rng.random(B) < 0.8).astype(np.uint8)
    alpha = float(rng.choice([0.5, 0.3, 0.7]))
    exp = np.stack([bo.mix_clip(fg[b], resized[idx[b]], int(top[b]), int(left[b]), (H, W), alpha, bool(app[b])) for b in range(B)])
    t32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.int32, device=dev)
    got = torch.ops.bgdebias.bgmix_blend_ragged(torch.from_numpy(fg).to(dev), rp.data, rp.slots_tensor, rp.tables.tensor, t32(idx), t32(top), t32(left),
                                                torch.from_numpy(app).to(dev), lut, mean, std, alpha, "NTCHW").cpu().numpy()
    if not np.array_equal(got.view(np.uint32), exp.view(np.uint32)):
        print(f"BLEND MISMATCH seed={seed} case={cases} H={H} W={W} size={size} imgs={[im.shape for im in imgs]} idx={idx.tolist()}"); sys.exit(1)
    cases += 1; clips += B; images += n_img
print(f"fuzz_ragged ok: {cases} batches, {clips} clips, {images} images bit-exact against torchvision Resize + the oracle blend (seed {seed})")
