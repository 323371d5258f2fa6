"""GPU integration test of the extraction CLI drop-in (bgdebias_b200.extract_background.main): lossless
FFV1 videos in a folder -> one JPEG per video, byte-identical to cv2.imwrite of the oracle's median over
exactly the frames the reference's loop keeps (cil_tools/extract_background.py:51-60,73-74,108), with the
reference's skip-if-exists resume (:119-126).  Single shard in-process and two spawned shards."""
import pathlib
import sys

import cv2
import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import median_oracle as mo                     # noqa: E402

pytestmark = pytest.mark.gpu
H, W = 48, 64


def _write(path, frames):
    wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"FFV1"), 25, (W, H))
    assert wr.isOpened()
    for f in frames:
        wr.write(f)
    wr.release()


def _expected_jpeg(frames, interval, max_frames, tmp):
    kept = mo.select_frame_indices(len(frames), interval, max_frames)
    med = mo.temporal_median_np([frames[i] for i in kept])
    p = tmp / "exp.jpg"
    cv2.imwrite(str(p), med)
    return p.read_bytes()


@pytest.mark.parametrize("workers", [1, 2])
def test_cli_folder_of_videos(tmp_path, workers):
    from bgdebias_b200 import extract_background as eb
    rng = np.random.default_rng(5)
    vdir, odir = tmp_path / "videos", tmp_path / "bg"
    vdir.mkdir()
    videos = {}
    for name, T in [("a", 9), ("b", 30), ("c", 1), ("d", 44), ("e", 17)]:
        base = rng.integers(0, 256, (1, H, W, 3), dtype=np.uint8)
        fr = np.clip(base.astype(np.int16) + rng.integers(-40, 41, (T, H, W, 3)), 0, 255).astype(np.uint8)
        _write(vdir / f"{name}.avi", fr)
        videos[name] = fr
    argv = ["--video_dir", str(vdir), "--output_dir", str(odir), "--from_video", "--num_workers", str(workers),
            "--interval", "2", "--max_frames", "10"]
    eb.main(argv)
    for name, fr in videos.items():
        got = (odir / f"{name}.jpg").read_bytes()
        assert got == _expected_jpeg(list(fr), 2, 10, tmp_path), name
    # the index persisted next to the JPEGs: sorted names + sizes; the pool can be rebuilt from it alone
    import json
    assert sorted(p.name for p in odir.iterdir()) == sorted(f"{n}.jpg" for n in videos)      # nothing but backgrounds in bg_dir
    index = json.loads(eb.index_path(odir).read_text())
    assert [e["name"] for e in index] == sorted(videos) and all((e["height"], e["width"]) == (H, W) for e in index)
    from bgdebias_b200.pool import BackgroundPool
    pool = BackgroundPool.from_index(odir, bg_resize=None)
    assert len(pool) == len(videos) and pool.hw == (H, W)
    assert [pathlib.Path(n).stem for n in pool.names] == sorted(videos)
    # resume: existing outputs are skipped, not rewritten
    marker = odir / "a.jpg"
    marker.write_bytes(b"kept")
    eb.main(argv)
    assert marker.read_bytes() == b"kept"


def test_per_video_functions_keep_the_reference_signatures(tmp_path):
    """extract_background.bg_extraction_tmf(data_path, dest, from_video, interval, max_frames, avg_method)
    (reference :42-43,108) and comix_loader.bg_extraction_tmf(data_path, dest) (reference
    comix_loader.py:148-164): same return value (the uint8 median frame) and the JPEG on disk."""
    from bgdebias_b200 import comix_loader as cl, extract_background as eb
    rng = np.random.default_rng(9)
    fr = rng.integers(0, 256, (23, H, W, 3), dtype=np.uint8)
    _write(tmp_path / "v.avi", fr)
    out = eb.bg_extraction_tmf(tmp_path / "v.avi", tmp_path / "v.jpg", True, 3, 5, 0)
    kept = mo.select_frame_indices(23, 3, 5)
    exp = mo.temporal_median_np([fr[i] for i in kept])
    assert out.dtype == np.uint8 and np.array_equal(out, exp)
    assert np.array_equal(cv2.imread(str(tmp_path / "v.jpg")), cv2.imdecode(cv2.imencode(".jpg", exp)[1], cv2.IMREAD_COLOR))

    # raw-frame folder: every image of the folder, decoded by cv2.imread like the reference (:157-160)
    folder = tmp_path / "frames"
    folder.mkdir()
    for i, f in enumerate(fr[:12]):
        cv2.imwrite(str(folder / f"img_{i + 1:05d}.jpg"), f)
    decoded = [cv2.imread(str(p)) for p in folder.glob("*")]
    out2 = cl.bg_extraction_tmf(folder, tmp_path / "f.jpg")
    assert np.array_equal(out2, mo.temporal_median_np(decoded))
    assert (tmp_path / "f.jpg").exists()
    with pytest.raises(NotImplementedError):
        cl.bg_extraction_tmf(folder, tmp_path / "g.jpg", from_video=True)


def test_many_threads_small_slabs_mixed_sizes(tmp_path):
    """Decoder threads reserve rows and pack concurrently while slabs rotate (1 MB slabs hold a few videos), frame
    sizes alternate (a slab holds one size) and one video is larger than a slab."""
    from bgdebias_b200 import extract_background as eb
    rng = np.random.default_rng(11)
    vdir, odir = tmp_path / "videos", tmp_path / "bg"
    vdir.mkdir()
    videos = {}
    for k in range(40):
        h, w = (H, W) if k % 3 else (32, 48)
        T = 150 if k == 7 else int(rng.integers(1, 40))
        fr = rng.integers(0, 256, (T, h, w, 3), dtype=np.uint8)
        wr = cv2.VideoWriter(str(vdir / f"v{k:02d}.avi"), cv2.VideoWriter_fourcc(*"FFV1"), 25, (w, h))
        for f in fr:
            wr.write(f)
        wr.release()
        videos[f"v{k:02d}"] = fr
    eb.main(["--video_dir", str(vdir), "--output_dir", str(odir), "--from_video", "--num_workers", "1",
             "--decode_threads", "8", "--slab_mb", "1"])
    for name, fr in videos.items():
        assert (odir / f"{name}.jpg").read_bytes() == _expected_jpeg(list(fr), 1, 500, tmp_path), name


def test_sharded_extraction_then_gathered_pool_feeds_the_mix(tmp_path, monkeypatch):
    """The product flow of the path's one collective, as `torchrun -m bgdebias_b200.extract_background --gather_pool`
    runs it (here: world size 1 over NCCL in-process): extraction writes the JPEGs (byte-identical to the reference's),
    the ranks all-gather the JPEG-round-tripped pixels into a resident pool, and a BackgroundMixDataset fed with that
    pool blends exactly like one that decodes the directory itself (comix_loader.py:84-103,126-131)."""
    import socket
    import torch
    import torch.distributed as dist
    from bgdebias_b200 import comix_loader as cl, extract_background as eb
    from oracle import bgmix_oracle as bo
    rng = np.random.default_rng(11)
    vdir, odir = tmp_path / "videos", tmp_path / "bg"
    vdir.mkdir()
    names = ["p", "q", "r", "s"]
    for name, T in zip(names, [5, 12, 8, 3]):
        base = rng.integers(0, 256, (1, H, W, 3), dtype=np.uint8)
        _write(vdir / f"{name}.avi", np.clip(base.astype(np.int16) + rng.integers(-30, 31, (T, H, W, 3)), 0, 255).astype(np.uint8))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    for k, v in dict(RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port)).items():
        monkeypatch.setenv(k, v)
    assert not dist.is_initialized()
    paths, ragged = eb.main(["--video_dir", str(vdir), "--output_dir", str(odir), "--from_video", "--gather_pool",
                             "--gather_pool_resize", "--size", "56"])
    assert not dist.is_initialized()
    assert [pathlib.Path(p).name for p in paths] == [f"{n}.jpg" for n in names] and len(ragged) == 4

    T, crop, B = 2, (40, 40), 6
    fg = rng.integers(0, 256, (B, T, 40, 40, 3), dtype=np.uint8)
    infos = [dict(frame_dir=f"/x/{names[i % 4]}", total_frames=T, label=i, sample=i) for i in range(B)]
    pipeline = lambda info: dict(imgs=torch.from_numpy(fg[info["sample"]]), label=torch.tensor([0]), randAug=False)
    outs = []
    for attach in (True, False):
        calls = []
        from torchvision.io import ImageReadMode, read_image
        ds = cl.BackgroundMixDataset(infos, pipeline, bg_dir=str(odir), bg_resize=56, bg_crop_size=crop, with_randAug=True,
                                     device_mix=True, bg_reader=lambda p: (calls.append(p), read_image(p, mode=ImageReadMode.RGB))[1])
        if attach:
            ds.attach_pool(paths, ragged)
        torch.manual_seed(3)
        samples = [ds.prepare_train_frames(i) for i in range(B)]
        outs.append(ds.gpu_collate(samples)["imgs"].cpu().numpy())
        assert (len(calls) == 0) == attach                      # the gathered pool is used as is: nothing is decoded again
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    # and against the oracle on the decoded JPEGs
    from oracle import aa_resize_oracle as ao
    torch.manual_seed(3)
    exp = []
    for i in range(B):
        b, t, l = bo.draw_bg_params(len(ds.bg_files), *ao.resized_hw(H, W, 56), crop)
        img = read_image(str(odir / f"{pathlib.Path(ds.bg_files[b]).stem}.jpg"), mode=ImageReadMode.RGB).numpy()
        exp.append(bo.mix_clip(fg[i], ao.aa_resize(img, 56), t, l, crop, 0.5, True))
    assert np.array_equal(outs[0].view(np.uint32), np.stack(exp).view(np.uint32))
