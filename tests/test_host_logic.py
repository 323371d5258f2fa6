"""CPU tests of the host-side mirror: shard split, frame selection, RNG draw order, the LUT, pool
geometry, and the multi-rank pool all-gather (gloo, world_size 2)."""
import os
import pathlib
import random

import numpy as np
import pytest
import torch

from oracle import bgmix_oracle as bo, median_oracle as mo


def test_contiguous_splits_match_reference():
    from bgdebias_b200 import shard
    for n, w in [(10, 4), (3, 4), (0, 2), (13320, 8), (7, 7), (8, 3)]:
        ours = shard.contiguous_splits(list(range(n)), w)
        ref = mo.reference_contiguous_splits(n, w)
        assert [list(r) for r in ref] == ours
        assert sum(len(s) for s in ours) == n
    assert shard.rank_slice(list(range(10)), 3, 4) == [9]
    assert shard.device_for(5, 4) == 1
    with pytest.raises(RuntimeError):
        shard.device_for(0, 0)


def _write_ffv1(path, frames):
    import cv2
    T, H, W, _ = frames.shape
    wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"FFV1"), 25, (W, H))
    for f in frames:
        wr.write(f)
    wr.release()


@pytest.mark.parametrize("n,interval,max_frames", [(20, 1, 4), (40, 2, 5), (50, 3, 500), (9, 1, 500), (9, 4, 0)])
def test_read_frames_selection_matches_oracle(tmp_path, n, interval, max_frames):
    """read_frames keeps exactly the frames the reference's loop keeps (extract_background.py:52-60)."""
    from bgdebias_b200 import extract_background as eb
    fr = np.random.default_rng(n).integers(0, 256, (n, 16, 24, 3), dtype=np.uint8)
    vid = tmp_path / "v.avi"
    _write_ffv1(vid, fr)
    got = eb.read_frames(vid, True, interval, max_frames)
    idx = mo.select_frame_indices(n, interval, max_frames)
    assert len(got) == len(idx)
    for g, i in zip(got, idx):
        np.testing.assert_array_equal(g, fr[i])


@pytest.mark.parametrize("n,interval,max_frames", [(20, 1, 4), (40, 2, 5), (50, 3, 500), (9, 1, 500), (9, 4, 0)])
def test_read_frames_packed_equals_read_frames(tmp_path, n, interval, max_frames):
    """The batched path's decoder (frames straight into a reusable per-thread scratch) keeps the same frames,
    also when the scratch is reused for a video of another size and by several threads at once."""
    from concurrent.futures import ThreadPoolExecutor
    from bgdebias_b200 import extract_background as eb
    rng = np.random.default_rng(n)
    vids = []
    for k, (h, w) in enumerate([(16, 24), (16, 24), (12, 8), (16, 24)]):
        fr = rng.integers(0, 256, (n + k, h, w, 3), dtype=np.uint8)
        _write_ffv1(tmp_path / f"v{k}.avi", fr)
        vids.append((tmp_path / f"v{k}.avi", fr))

    def check(item):
        path, fr = item
        got = eb.read_frames_packed(path, True, interval, max_frames).copy()
        idx = mo.select_frame_indices(len(fr), interval, max_frames)
        assert got.shape == (len(idx),) + fr.shape[1:]
        np.testing.assert_array_equal(got, fr[idx])
        return True

    assert all(check(v) for v in vids)                         # one thread, scratch reused across sizes
    with ThreadPoolExecutor(3) as ex:
        assert all(ex.map(check, vids * 3))


def test_read_frames_image_folder(tmp_path):
    import cv2
    from bgdebias_b200 import extract_background as eb
    fr = np.random.default_rng(1).integers(0, 256, (7, 8, 8, 3), dtype=np.uint8)
    for t, f in enumerate(fr):
        cv2.imwrite(str(tmp_path / f"img_{t + 1:05}.png"), f)
    got = eb.read_frames(tmp_path, False, 2, 500)
    assert len(got) == 4
    np.testing.assert_array_equal(got[1], fr[2])


def test_cli_flags_are_the_reference_flags():
    from bgdebias_b200 import extract_background as eb
    a = eb.parse_args(["--video_dir", "v", "--output_dir", "o"])
    assert (a.glob_pattern, a.num_workers, a.from_video, a.image_suffix, a.interval, a.max_frames, a.size, a.method,
            a.avg_method) == ('*', 4, False, '.jpg', 1, 500, 256, 'tmf', 'median')


def test_fg_lut_bitwise_equals_oracle():
    from bgdebias_b200 import ops
    for mean, std in [(bo.DEFAULT_MEAN, bo.DEFAULT_STD), ((10.5, 200.25, 0.0), (1.0, 33.3, 255.0))]:
        np.testing.assert_array_equal(ops.make_fg_lut(mean, std).numpy().view(np.uint32), bo.fg_lut(mean, std).view(np.uint32))


def test_pool_geometry_and_draw_order():
    from bgdebias_b200 import comix_loader as cl, pool
    for h, w, s in [(240, 320, 256), (240, 427, 256), (320, 240, 256), (256, 256, 256), (90, 120, 36)]:
        assert pool.resized_hw(h, w, s) == bo.resized_hw(h, w, s)
        assert tuple(pool.resize_like_reference(torch.zeros(3, h, w), s).shape[1:]) == bo.resized_hw(h, w, s)
    torch.manual_seed(5)
    a = (int(torch.randint(9, (1,)).item()),) + cl.draw_crop(256, 341, (224, 224))
    torch.manual_seed(5)
    assert a == bo.draw_bg_params(9, 256, 341, (224, 224))
    torch.manual_seed(5)
    st = torch.get_rng_state()
    assert cl.draw_crop(224, 224, (224, 224)) == (0, 0)          # nothing drawn when sizes match
    assert torch.equal(st, torch.get_rng_state())
    with pytest.raises(ValueError):
        cl.draw_crop(100, 300, (224, 224))


def test_dataset_pool_modes_without_gpu(tmp_path):
    """Pool assembly (comix_loader.py:84-103) needs no device."""
    from bgdebias_b200 import comix_loader as cl
    bg = tmp_path / "bg"
    bg.mkdir()
    for n in ("a", "b", "zz"):
        (bg / f"{n}.jpg").write_bytes(b"x")
    infos = [dict(frame_dir=f"/d/{n}", total_frames=3, label=0) for n in ("a", "b", "c")]
    ident = lambda r: r
    ds = cl.BackgroundMixDataset(infos, ident, bg_dir=str(bg), extract_bg_if_not_found=False)
    assert [pathlib.Path(p).name for p in ds.bg_files] == ["a.jpg", "b.jpg"]
    ds2 = cl.BackgroundMixDataset(infos, ident, bg_dir=str(bg), map_bg_to_video=False, merge_bg_files=False)
    assert sorted(pathlib.Path(p).name for p in ds2.bg_files) == ["a.jpg", "b.jpg", "zz.jpg"]
    ds3 = cl.BackgroundMixDataset(infos, ident, bg_dir=str(bg), back_ground_from_bg_dir=False)
    assert ds3.bg_files == []
    assert ds.merge_bg_files is True and ds.test_mode is False and len(ds.video_infos) == 3
    # gate: with_randAug and randAug=True -> untouched, bg_idx -1, no GPU needed
    ds4 = cl.BackgroundMixDataset(infos, lambda r: dict(r, imgs=torch.zeros(2, 3, 4, 4), randAug=True),
                                  bg_dir=str(bg), extract_bg_if_not_found=False, with_randAug=True)
    r = ds4.prepare_train_frames(0)
    assert r["bg_idx"] == -1 and torch.equal(r["imgs"], torch.zeros(2, 3, 4, 4))
    random.seed(0)
    ds5 = cl.BackgroundMixDataset(infos, lambda r: dict(r, imgs=torch.zeros(2, 3, 4, 4)), bg_dir=str(bg),
                                  extract_bg_if_not_found=False, prob=0.0)
    assert ds5.prepare_train_frames(1)["bg_idx"] == -1


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bgdebias_b200.pool import BackgroundPool
    n = 3 if rank == 0 else 1                                   # ragged shards
    names = [f"r{rank}_v{i}" for i in range(n)]
    bgs = torch.full((n, 4, 5, 3), rank + 1, dtype=torch.uint8) + torch.arange(n, dtype=torch.uint8).view(n, 1, 1, 1)
    all_names, all_bgs = BackgroundPool.all_gather(names, bgs)
    index = BackgroundPool.all_gather_index(names)
    assert [n for n, _, _ in index] == all_names and [(r, i) for _, r, i in index] == [(0, 0), (0, 1), (0, 2), (1, 0)]
    q.put((rank, all_names, all_bgs.numpy()))
    dist.destroy_process_group()


def test_pool_all_gather_two_ranks():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_names = ["r0_v0", "r0_v1", "r0_v2", "r1_v0"]
    for rank, names, bgs in res:
        assert names == exp_names
        assert bgs.shape == (4, 4, 5, 3)
        assert [int(bgs[i, 0, 0, 0]) for i in range(4)] == [1, 2, 3, 2]


def _gather_dir_worker(rank, world, port, bg_dir, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bgdebias_b200.pool import BackgroundPool, gather_extracted_backgrounds
    paths, pool = gather_extracted_backgrounds(bg_dir, ".jpg", 40, "cpu")
    imgs = []
    for s in pool.slots:
        o, h, w = int(s["offset"]), int(s["h"]), int(s["w"])
        imgs.append(pool.data[o:o + 3 * h * w].view(3, h, w).numpy().copy())
    # shards in an order the extraction's split never produces (first shard shorter): the compacting branch
    n = 1 if rank == 0 else 3
    names, bgs = BackgroundPool.all_gather([f"r{rank}_{i}" for i in range(n)], torch.full((n, 2, 2), 10 * rank, dtype=torch.uint8) + torch.arange(n, dtype=torch.uint8).view(n, 1, 1))
    q.put((rank, paths, imgs, [(int(s["Hb"]), int(s["Wb"])) for s in pool.slots], names, bgs.numpy()))
    dist.destroy_process_group()


def test_gather_of_extracted_backgrounds_equals_decoding_the_directory(tmp_path):
    """The path's one collective as the product runs it after a sharded extraction (world size 2, gloo, CPU): every rank
    decodes the JPEGs of its own slice and the all-gathered pool holds, in sorted-name order, exactly the pixels
    torchvision.io.read_image returns for every file of the directory -- what BackgroundPool.from_files decodes and what
    the reference's _get_bg_image reads (comix_loader.py:130)."""
    import multiprocessing as mp
    import cv2
    from torchvision.io import ImageReadMode, read_image
    rng = np.random.default_rng(3)
    sizes = [(36, 48), (36, 64), (40, 40), (36, 48), (50, 30)]
    for i, (h, w) in enumerate(sizes):
        cv2.imwrite(str(tmp_path / f"v{i:02d}.jpg"), rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gather_dir_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    files = sorted(str(f) for f in tmp_path.glob("*.jpg"))
    from bgdebias_b200.pool import resized_hw
    for rank, paths, imgs, hws, names, bgs in res:
        assert paths == files
        for f, im, hw in zip(files, imgs, hws):
            ref = read_image(f, mode=ImageReadMode.RGB).numpy()
            assert np.array_equal(im, ref) and hw == resized_hw(ref.shape[1], ref.shape[2], 40)
        assert names == ["r0_0", "r1_0", "r1_1", "r1_2"] and [int(b[0, 0]) for b in bgs] == [0, 10, 11, 12]


def _gather_empty_worker(rank, world, port, bg_dir, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bgdebias_b200.pool import gather_extracted_backgrounds
    paths, pool = gather_extracted_backgrounds(bg_dir, ".jpg", None, "cpu")
    q.put((rank, paths, len(pool), [int(s["offset"]) for s in pool.slots], int(pool.slots["xtab"][0]) if len(pool) else None))
    dist.destroy_process_group()


def test_gather_with_an_empty_shard(tmp_path):
    """More ranks than files (the ceil split leaves later ranks empty, extract_background.py:128-133): the empty rank sends
    nothing and still ends up with the whole pool; bg_resize=None means no resize tables at all."""
    import multiprocessing as mp
    import cv2
    cv2.imwrite(str(tmp_path / "only.jpg"), np.full((20, 24, 3), 77, np.uint8))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gather_empty_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, paths, n, offs, xtab in res:
        assert paths == [str(tmp_path / "only.jpg")] and n == 1 and offs == [0] and xtab == -1


def test_background_store_slots_without_gpu():
    """BackgroundStore on the CPU device: names are decoded once, pools are slot tables over shared pixels."""
    import numpy as np
    import torch
    from bgdebias_b200.pool import BackgroundStore
    rng = np.random.default_rng(0)
    imgs = {f"n{i}": torch.from_numpy(rng.integers(0, 256, (3, 8, 10), dtype=np.uint8)) for i in range(6)}
    seen = []
    store = BackgroundStore(bg_resize=None, device="cpu")
    store.ensure(["n0", "n1", "n1"], lambda n: (seen.append(n), imgs[n])[1])
    assert seen == ["n0", "n1"] and len(store) == 2
    a = store.view(["n1", "n0", "n1"])
    assert len(a) == 3 and a.slots.tolist() == [1, 0, 1] and a.hw == (8, 10)
    store.ensure(["n1", "n4", "n2"], lambda n: (seen.append(n), imgs[n])[1])
    assert seen == ["n0", "n1", "n4", "n2"] and store.decoded == 4
    b = store.view(["n2", "n4", "n0"])
    assert b.slots.tolist() == [3, 2, 0]
    assert torch.equal(b.tensor[b.rows(torch.tensor([1]))[0]], imgs["n4"].float())
    assert torch.equal(a.tensor[a.slots[0].item()], imgs["n1"].float())      # growth kept the old pixels
    import pytest
    with pytest.raises(KeyError):
        store.view(["n5"])
    # a second image size is accepted (the reference crops every background at its own size, comix_loader.py:139-141):
    # the dense cache is gone, sizes are per image
    store.ensure(["other"], lambda n: torch.zeros(3, 9, 9, dtype=torch.uint8))
    c = store.view(["other", "n0"])
    assert c.tensor is None and c.hw_of(0) == (9, 9) and c.hw_of(1) == (8, 10) and len(store) == 5
    with pytest.raises(ValueError):
        c.hw


def _fake_person_detector():
    """Detector plug-in for the type B/C shim: 'person' iff the image's top-left pixel is white."""
    return lambda img: [0, 17] if img[0, 0, 0] > 200 else [17]


def test_type_b_and_c_shim_selection_rule(tmp_path, monkeypatch):
    """type_b_and_c_bg.py:42-54: copy the backgrounds in which class 0 is not predicted; out_dir must be new."""
    import cv2
    from bgdebias_b200 import type_b_and_c_bg as tb
    src = tmp_path / "bg"; src.mkdir()
    for i in range(5):
        img = np.full((8, 8, 3), 255 if i % 2 else 10, np.uint8)
        cv2.imwrite(str(src / f"v{i}.png"), img)
    monkeypatch.chdir(tmp_path)
    out = tmp_path / "type_c"
    rec = tb.main(["-i", str(src), "-o", str(out), "--glob_pattern", "*.png", "--detector", "test_host_logic:_fake_person_detector"])
    assert sorted(p.name for p in out.iterdir()) == ["v0.png", "v2.png", "v4.png"]
    assert [r["copied"] for r in rec] == [True, False, True, False, True] and (tmp_path / "detection.json").exists()
    with pytest.raises(FileExistsError):
        tb.main(["-i", str(src), "-o", str(out), "--detector", "test_host_logic:_fake_person_detector"])
    with pytest.raises(RuntimeError, match="detector"):
        tb.main(["-i", str(src), "-o", str(tmp_path / "x")])           # detectron2 is absent here: loud failure
