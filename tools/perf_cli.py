"""End-to-end timing of the extraction CLI on a folder of synthetic lossless videos (decode included), next to the
reference's per-video procedure (decode + np.median, one process per core, extract_background.py:42-75,154-162)
restated with the oracle.  usage: [FOURCC=HFYU|mp4v|XVID] [CONTENT=noise|smooth] [WORKERS=gpus] python tools/perf_cli.py [n_videos=48] [frames=150]
FOURCC=HFYU CONTENT=noise (default) is the worst case for the decoder (incompressible frames, lossless codec);
FOURCC=mp4v CONTENT=smooth is what UCF101 / HMDB51 look like (MPEG-4 part 2 in .avi, natural-image statistics)."""
import json, os, pathlib, sys, tempfile, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import cv2, numpy as np

H, W = 240, 320
n_videos = int(sys.argv[1]) if len(sys.argv) > 1 else 48
T = int(sys.argv[2]) if len(sys.argv) > 2 else 150


def _cpu_one(path):
    from oracle import median_oracle as mo
    cap = cv2.VideoCapture(path); frames = []
    while cap.isOpened() and len(frames) <= 500:
        ok, f = cap.read()
        if not ok: break
        frames.append(f)
    cap.release()
    out = mo.temporal_median_np(frames)
    cv2.imwrite(path + ".cpu.jpg", out)
    return len(frames)


def main():
    from concurrent.futures import ProcessPoolExecutor
    from bgdebias_b200 import extract_background as eb
    rng = np.random.default_rng(0)
    cores = len(os.sched_getaffinity(0))
    with tempfile.TemporaryDirectory() as tmp:
        vdir, odir = pathlib.Path(tmp) / "videos", pathlib.Path(tmp) / "bg"
        vdir.mkdir()
        base = rng.integers(0, 256, (4, H, W, 3), dtype=np.uint8)
        if os.environ.get("CONTENT", "noise") == "smooth":
            base = np.stack([cv2.GaussianBlur(b, (31, 31), 0) for b in base])
        t0 = time.perf_counter()
        for v in range(n_videos):
            wr = cv2.VideoWriter(str(vdir / f"v{v:04d}.avi"), cv2.VideoWriter_fourcc(*os.environ.get("FOURCC", "HFYU")), 25, (W, H))
            for t in range(T):
                f = base[(v + t) % 4].copy(); f[(7 * t) % (H - 40):(7 * t) % (H - 40) + 40, (11 * t) % (W - 40):(11 * t) % (W - 40) + 40] = 255 - (t % 5)
                wr.write(f)
            wr.release()
        print(f"wrote {n_videos} x {T} frames in {time.perf_counter() - t0:.1f} s", file=sys.stderr)
        workers = int(os.environ.get("WORKERS", "1"))               # GPU shards (one process per GPU)
        argv = ["--video_dir", str(vdir), "--output_dir", str(odir), "--from_video", "--num_workers", str(workers),
                "--decode_threads", str(max(1, cores // workers))]
        eb.main(argv)                                           # warm-up run (CUDA context, pinned slabs) ...
        for f in odir.glob("*.jpg"): f.unlink()
        t0 = time.perf_counter(); eb.main(argv); t_gpu = time.perf_counter() - t0
        paths = [str(p) for p in sorted(vdir.glob("*.avi"))]
        with ProcessPoolExecutor(cores) as ex:
            list(ex.map(_cpu_one, paths[:cores]))               # warm-up
            t0 = time.perf_counter(); n = sum(ex.map(_cpu_one, paths)); t_cpu = time.perf_counter() - t0
        same = all((odir / (pathlib.Path(p).stem + ".jpg")).read_bytes() == pathlib.Path(p + ".cpu.jpg").read_bytes() for p in paths)
    print(json.dumps({"videos": n_videos, "frames": n, "host_cores": cores, "gpu_shards": workers,
                      "ours_cli_s": round(t_gpu, 3), "ours_frames_per_s": round(n / t_gpu, 1),
                      "reference_procedure_s": round(t_cpu, 3), "reference_frames_per_s": round(n / t_cpu, 1),
                      "speedup": round(t_cpu / t_gpu, 2), "identical_jpegs": same,
                      "codec": os.environ.get("FOURCC", "HFYU"), "content": os.environ.get("CONTENT", "noise"),
                      "note": "decode (OpenCV/FFmpeg) is on the host in both arms; ours decodes with a thread pool into a pinned slab and reduces many videos per launch"}))


if __name__ == "__main__":
    main()
