"""Randomised differential test of the GPU Resize (bgdebias::resize_bilinear and the fused bgmix_resize_blend) against
cv2.resize(..., INTER_LINEAR): ragged batches of random source sizes, destination sizes, contents, strides and offsets.
usage: python tools/fuzz_resize.py [seconds=60] [seed=0]"""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import cv2, numpy as np, torch
import bgdebias_b200.ops as ops
from oracle import bgmix_oracle as bo

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda")
lut = ops.make_fg_lut(bo.DEFAULT_MEAN, bo.DEFAULT_STD, dev)
lut_h = lut.cpu().numpy()
t_end = time.time() + budget
cases = clips = 0
while time.time() < t_end:
    B, T = int(rng.integers(1, 7)), int(rng.integers(1, 5))
    hi = int(rng.choice([8, 40, 300, 700]))
    dh, dw = int(rng.integers(1, hi + 1)), int(rng.integers(1, hi + 1))
    srcs, geom, chunks, off = [], [], [], int(rng.integers(0, 7))
    chunks.append(np.zeros(off, np.uint8))
    for b in range(B):
        fh, fw = int(rng.integers(1, hi + 1)), int(rng.integers(1, hi + 1))         # decoded frame
        kind = rng.integers(0, 4)
        if kind == 0: fr = rng.integers(0, 256, (T, fh, fw, 3), dtype=np.uint8)
        elif kind == 1: fr = rng.choice(np.array([0, 255], np.uint8), (T, fh, fw, 3))
        elif kind == 2: fr = (np.arange(fh)[None, :, None, None] * 3 + np.arange(fw)[None, None, :, None] * 5 + rng.integers(0, 3, (T, fh, fw, 3))).astype(np.uint8)
        else: fr = np.full((T, fh, fw, 3), int(rng.integers(0, 256)), np.uint8)
        top, left = int(rng.integers(0, fh)), int(rng.integers(0, fw))               # crop inside it
        h, w = int(rng.integers(1, fh - top + 1)), int(rng.integers(1, fw - left + 1))
        geom.append([off + (top * fw + left) * 3, h, w, fw * 3, fh * fw * 3])
        srcs.append(fr[:, top:top + h, left:left + w])
        chunks.append(fr.reshape(-1)); off += fr.size
        pad = int(rng.integers(0, 5)); chunks.append(np.zeros(pad, np.uint8)); off += pad
    chunks.append(np.zeros(8 + (-off) % 4, np.uint8))
    buf = torch.from_numpy(np.concatenate(chunks)).to(dev)
    g = torch.tensor(geom, dtype=torch.int64)
    out = torch.ops.bgdebias.resize_bilinear(buf, g, T, dh, dw)
    z = torch.zeros(B, dtype=torch.int32, device=dev)
    fused = torch.ops.bgdebias.bgmix_resize_blend(buf, g, T, dh, dw, torch.zeros((1, 3, dh, dw), dtype=torch.uint8, device=dev), z, z, z,
                                                  torch.zeros(B, dtype=torch.uint8, device=dev), lut, torch.tensor(bo.DEFAULT_MEAN),
                                                  torch.tensor(bo.DEFAULT_STD), 0.5, "NTCHW").cpu().numpy()
    got = out.cpu().numpy()
    for b in range(B):
        for t in range(T):
            exp = cv2.resize(np.ascontiguousarray(srcs[b][t]), (dw, dh), interpolation=cv2.INTER_LINEAR).reshape(dh, dw, 3)
            if not np.array_equal(got[b, t], exp):
                print("MISMATCH resize", geom[b], (dh, dw), "max diff", int(np.abs(got[b, t].astype(int) - exp).max())); sys.exit(1)
            expf = lut_h[np.arange(3)[:, None, None], exp.transpose(2, 0, 1)]
            if not np.array_equal(fused[b, t].view(np.uint32), expf.view(np.uint32)):
                print("MISMATCH fused", geom[b], (dh, dw)); sys.exit(1)
        clips += 1
    cases += 1
print(f"fuzz_resize: {cases} batches, {clips} clips bit-exact against cv2.resize (seed {sys.argv[2] if len(sys.argv) > 2 else 0})")
