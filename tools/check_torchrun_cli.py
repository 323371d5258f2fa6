"""The sharded extraction + pool gather on real GPUs under torchrun (run on a multi-GPU box):

    torchrun --nproc-per-node 2 tools/check_torchrun_cli.py

Every rank calls the CLI entry point (`extract_background.main`, torchrun mode) on one shared folder of lossless videos with
`--gather_pool`; rank 0 then checks (a) every JPEG is byte-identical to cv2.imwrite of the oracle's median, (b) the gathered
pool on EVERY rank holds exactly torchvision's decode of every file of the directory, in sorted order, (c) a batch blended
from the attached pool equals the oracle.  Prints one JSON line."""
import json, os, pathlib, sys, tempfile
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import cv2, numpy as np, torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from bgdebias_b200 import comix_loader as cl, extract_background as eb
from oracle import aa_resize_oracle as ao, bgmix_oracle as bo, median_oracle as mo
from torchvision.io import ImageReadMode, read_image

holder = [tempfile.mkdtemp(prefix="bgd_torchrun_") if rank == 0 else None]
dist.broadcast_object_list(holder, src=0)
root = pathlib.Path(holder[0])
vdir, odir = root / "videos", root / "bg"
rng = np.random.default_rng(5)
videos = {}
sizes = [(48, 64), (48, 80), (56, 64)]
for i in range(11):                                      # 11 videos over `world` ranks: uneven shards, mixed frame sizes
    H, W = sizes[i % 3]
    T = int(rng.integers(3, 40))
    base = rng.integers(0, 256, (1, H, W, 3), dtype=np.uint8)
    videos[f"v{i:02d}"] = np.clip(base.astype(np.int16) + rng.integers(-25, 26, (T, H, W, 3)), 0, 255).astype(np.uint8)
if rank == 0:
    vdir.mkdir()
    for name, fr in videos.items():
        wr = cv2.VideoWriter(str(vdir / f"{name}.avi"), cv2.VideoWriter_fourcc(*"FFV1"), 25, (fr.shape[2], fr.shape[1]))
        for f in fr:
            wr.write(f)
        wr.release()
dist.barrier()
paths, pool = eb.main(["--video_dir", str(vdir), "--output_dir", str(odir), "--from_video", "--gather_pool", "--gather_pool_resize", "--size", "64"])
ok_jpeg = ok_pool = ok_mix = True
files = sorted(str(f) for f in odir.glob("*.jpg"))
ok_pool = paths == files and len(pool) == len(videos)
for f, s in zip(files, pool.slots):
    ref = read_image(f, mode=ImageReadMode.RGB).cuda()
    o = int(s["offset"])
    ok_pool = ok_pool and bool(torch.equal(pool.data[o:o + ref.numel()].view_as(ref), ref))
if rank == 0:
    for name, fr in videos.items():
        exp = root / "exp.jpg"
        cv2.imwrite(str(exp), mo.temporal_median_np(list(fr)))
        ok_jpeg = ok_jpeg and (odir / f"{name}.jpg").read_bytes() == exp.read_bytes()
# blend from the attached pool on every rank
T, crop, B = 2, (40, 40), 6
fg = np.random.default_rng(100 + rank).integers(0, 256, (B, T, 40, 40, 3), dtype=np.uint8)
names = sorted(videos)
infos = [dict(frame_dir=f"/x/{names[i % len(names)]}", total_frames=T, label=i, sample=i) for i in range(B)]
ds = cl.BackgroundMixDataset(infos, lambda info: dict(imgs=torch.from_numpy(fg[info["sample"]]), label=torch.tensor([0]), randAug=False),
                             bg_dir=str(odir), bg_resize=64, bg_crop_size=crop, with_randAug=True, device_mix=True,
                             device=f"cuda:{local}", bg_reader=lambda p: (_ for _ in ()).throw(RuntimeError("decoded again: " + p)))
ds.attach_pool(paths, pool)
torch.manual_seed(7 + rank)
samples = [ds.prepare_train_frames(i) for i in range(B)]
got = ds.gpu_collate(samples)["imgs"].cpu().numpy()
exp = []
for i, s in enumerate(samples):
    img = read_image(ds.bg_files[s["bg_idx"]], mode=ImageReadMode.RGB).numpy()
    exp.append(bo.mix_clip(fg[i], ao.aa_resize(img, 64), s["bg_top"], s["bg_left"], crop, 0.5, True))
ok_mix = bool(np.array_equal(got.view(np.uint32), np.stack(exp).view(np.uint32)))
flags = [None] * world
dist.all_gather_object(flags, (ok_pool, ok_mix))
if rank == 0:
    import shutil
    shutil.rmtree(root, ignore_errors=True)
    print(json.dumps({"world": world, "videos": len(videos), "jpegs_identical_to_reference_procedure": ok_jpeg,
                      "gathered_pool_equals_decoding_the_directory_on_every_rank": all(f[0] for f in flags),
                      "blend_from_attached_pool_bit_exact_on_every_rank": all(f[1] for f in flags)}))
dist.barrier()
dist.destroy_process_group()
