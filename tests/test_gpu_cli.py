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
    # resume: existing outputs are skipped, not rewritten
    marker = odir / "a.jpg"
    marker.write_bytes(b"kept")
    eb.main(argv)
    assert marker.read_bytes() == b"kept"
