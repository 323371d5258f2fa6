#!/usr/bin/env python
"""bench.py -- background-extraction frames/s (and BG-mix clips/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): the UCF101-shaped set -- 13,320 videos, T ~ U[120,240] frames
of 240x320x3 uint8 -- sharded over 8 GPUs: 1,665 videos (~69 GB) per GPU, one process per GPU, no
collective on the data path (weak scaling: every rank always owns one 1,665-video shard).  A "step"
is one pass of the temporal-median extraction over the rank's resident shard.

One JSON line on stdout (rank 0):
  value      frames/s over all ranks, inputs resident in HBM (CUDA events, max over ranks)
  e2e        frames/s through the C-ABI host call (pinned host frames -> H2D -> kernel -> D2H)
  roofline   algorithmic bytes of a step / event time of a step vs the measured HBM peak
  cpu_baseline  the reference's np.median path on this box's cores, bounded sample
  bgmix      BASELINE.json configs[4]: fused blend, 64 x 8 x 224 x 224 clips per step
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# The bench line is configs[1] (ucf101).  The other extraction configs of BASELINE.json are parity-test cases; they
# can be timed through the same code with --workload (the JSON line then names that workload).
WORKLOADS = {
    "ucf101": dict(H=240, W=320, t_lo=120, t_hi=241, t_mul=1, videos=13320 // 8, t_text="U[120,240]",
                   label="configs[1]: UCF101-shaped set, one 1/8 shard per GPU"),
    "hmdb51": dict(H=240, W=320, t_lo=30, t_hi=81, t_mul=2, videos=846, t_text="2*U[30,80] (even)",
                   label="configs[2]: HMDB51-shaped set (even T), one 1/8 shard per GPU"),
    "sthv2": dict(H=240, W=427, t_lo=24, t_hi=73, t_mul=1, videos=4096, t_text="U[24,72]",
                  label="configs[3]: Sth-Sth-v2-shaped clips, a resident 4,096-clip chunk of the 27,606-clip 1/8 shard"),
}
# Legs reported INSIDE the default line under "workloads" (bounded resident chunks, same code path, same run):
# the other extraction configs of BASELINE.json and the structured-content variant of SURVEY.md section 8d.
EXTRA_LEGS = {
    "hmdb51": dict(workload="hmdb51", videos=846, content="random"),
    "sthv2": dict(workload="sthv2", videos=4096, content="random"),
    "ucf101_structured": dict(workload="ucf101", videos=416, content="structured"),
}


def _select_workload(name: str) -> None:
    global WL, WL_NAME, H, W, N_BYTES
    WL_NAME, WL = name, WORKLOADS[name]
    H, W = WL["H"], WL["W"]
    N_BYTES = H * W * 3
    os.environ["BGD_BENCH_WORKLOAD"] = name          # spawned CPU-arm workers re-import this module


def draw_T(rng, n=None):
    return WL["t_mul"] * rng.integers(WL["t_lo"], WL["t_hi"], n)


_select_workload(os.environ.get("BGD_BENCH_WORKLOAD", "ucf101"))


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner
# to stdout under NCCL_DEBUG=VERSION), so keep a private handle on the real stdout for the JSON line and
# point file descriptor 1 at stderr for everything else.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own computation, np.median(frames, axis=0).astype(uint8)
# (cil_tools/extract_background.py:73), one video per call, fanned out over processes like
# extract_background.py:154-162.  Lives in oracle/ (test infrastructure); only timed here.
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, n_videos = args
    import numpy as np
    from oracle import _ref_import, median_oracle as mo
    rng = np.random.default_rng(seed)
    vids = []
    for _ in range(n_videos):
        T = int(draw_T(rng))
        vids.append([rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(T)])   # the `frames` list
    # where the reference tree exists (the build container) the reference's own bg_extraction_tmf runs, file I/O stubbed;
    # on the GPU box it is absent and the oracle's restatement of the same line runs instead
    median = _ref_import.reference_median_of_frames() if _ref_import.available() else mo.temporal_median_np
    t0 = time.perf_counter()
    for frames in vids:
        median(frames)
    return time.perf_counter() - t0, sum(len(v) for v in vids)


def cpu_kind() -> str:
    from oracle import _ref_import
    return "reference" if _ref_import.available() else "port"


def cpu_reference_fps(n_procs: int, videos_per_proc: int, seed: int = 0):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(n_procs) as pool:
        res = pool.map(_cpu_worker, [(seed * 1000 + i, videos_per_proc) for i in range(n_procs)])
    t = max(r[0] for r in res)
    frames = sum(r[1] for r in res)
    return frames / t, frames, t


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:           # pragma: no cover
            self.nv = None
            log("NVML unavailable:", e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.thread is not None:
            self.thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}



def traffic_ratio(name):
    """DRAM bytes per algorithmic byte of the kernel, from the committed `ncu --set full` capture (profiles/r2_traffic*.json)."""
    try:
        return float(json.loads((ROOT / "profiles" / name).read_text())["dram_bytes_per_algorithmic_byte"])
    except Exception:
        return None


TRAFFIC_SOURCE = "profiles ncu capture: DRAM bytes per algorithmic byte of one `ncu --set full` launch of this kernel x this step's algorithmic bytes"


def fill_frames(frames, Ts, offs, Hh, Ww, content, seed, torch):
    """Synthetic frames on the device.  "random": uniform bytes (the selection's worst case, every column has T distinct
    candidates).  "structured": what the temporal median is for -- one static background per video,
    +-1 sensor noise on 5 % of the samples and a bright 40x40 block that moves a few pixels per frame."""
    dev = frames.device
    n_bytes = frames.shape[1]
    g = torch.Generator(device=dev).manual_seed(seed)
    rows = frames.shape[0]
    if content == "random":
        chunk = max(1, (1 << 30) // n_bytes)
        for r0 in range(0, rows, chunk):
            r1 = min(rows, r0 + chunk)
            frames[r0:r1] = torch.randint(0, 256, (r1 - r0, n_bytes), dtype=torch.uint8, device=dev, generator=g)
        return
    for v, T in enumerate(Ts):
        r0 = int(offs[v])
        T = int(T)
        base = torch.randint(40, 200, (1, n_bytes), dtype=torch.int16, device=dev, generator=g)
        noise = (torch.rand((T, n_bytes), device=dev, generator=g) < 0.05).to(torch.int16) * \
            (torch.randint(0, 2, (T, n_bytes), dtype=torch.int16, device=dev, generator=g) * 2 - 1)
        fr = (base + noise).clamp_(0, 255).to(torch.uint8).view(T, Hh, Ww, 3)
        t = torch.arange(T, device=dev)
        x0 = (t * 3) % (Ww - 40)
        y0 = (t * 2) % (Hh - 40)
        yy = (y0[:, None, None] + torch.arange(40, device=dev)[None, :, None]).expand(T, 40, 40)
        xx = (x0[:, None, None] + torch.arange(40, device=dev)[None, None, :]).expand(T, 40, 40)
        tt = t[:, None, None].expand(T, 40, 40)
        fr[tt, yy, xx] = (255 - (t % 7)).to(torch.uint8)[:, None, None, None].expand(T, 40, 40, 3)
        frames[r0:r0 + T] = fr.view(T, n_bytes)


def spot_check(frames, offs, Ts, out, v_chk, torch):
    """One video of the timed path against the definition in plain torch: (s[(T-1)//2] + s[T//2]) >> 1 over the sorted column."""
    sl = slice(int(offs[v_chk]), int(offs[v_chk + 1]))
    srt = torch.sort(frames[sl].to(torch.int16), dim=0).values
    Tc = srt.shape[0]
    return bool(torch.equal(out[v_chk], ((srt[(Tc - 1) // 2] + srt[Tc // 2]) >> 1).to(torch.uint8)))


def extraction_leg(leg_name, spec, rank, world, dev, steps, barrier):
    """One bounded, resident chunk of another extraction workload through the same op: frames/s, roofline, spot check."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from bgdebias_b200 import _cabi
    wl = WORKLOADS[spec["workload"]]
    Hh, Ww = wl["H"], wl["W"]
    n_bytes = Hh * Ww * 3
    V = spec["videos"]
    rng = np.random.default_rng(50 + rank)
    Ts = wl["t_mul"] * rng.integers(wl["t_lo"], wl["t_hi"], V)
    free, _ = torch.cuda.mem_get_info(dev)
    while int(Ts.sum()) * n_bytes + (4 << 30) > free and V > 16:
        V //= 2
        Ts = Ts[:V]
    offs = torch.from_numpy(np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64))
    rows = int(offs[-1])
    frames = torch.empty((rows, n_bytes), dtype=torch.uint8, device=dev)
    fill_frames(frames, Ts, offs, Hh, Ww, spec["content"], 200 + rank, torch)
    torch.cuda.synchronize()
    out = None
    for _ in range(3):
        out = torch.ops.bgdebias.temporal_median_varlen(frames, offs)
    barrier()
    l0 = _cabi.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = torch.ops.bgdebias.temporal_median_varlen(frames, offs)
    e1.record()
    barrier()
    launches = _cabi.kernel_launch_count() - l0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(rows)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(t.item()) / steps
    ok = spot_check(frames, offs, Ts, out, int(np.argmin(Ts)), torch) and spot_check(frames, offs, Ts, out, int(np.argmax(Ts)), torch)
    step_bytes = (rows + V) * n_bytes
    peak, _ = measured_peak_gbs()
    achieved = step_bytes / (ms * 1e-3) / 1e9
    rec = {"metric": "bg_extraction_frames_per_sec", "value": float(tot.item()) / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
           "steps": steps, "gpu_launches_per_step": launches // steps,
           "config": {"workload": wl["label"] + f"; resident chunk of {V} videos per GPU, content: {spec['content']}", "videos_per_gpu": V,
                      "frames_per_gpu": rows, "frame_shape": [Hh, Ww, 3], "T": wl["t_text"], "resident_gb": rows * n_bytes / 1e9},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "frac_of_nominal_8TBs": achieved / 8000.0, "algorithmic_bytes_per_step": step_bytes,
                        "traffic": (lambda r: r * step_bytes if r else None)(traffic_ratio(
                            {"hmdb51": "r2_traffic_T180.json", "sthv2": "r2_traffic_sthv2_T48.json"}.get(leg_name, "r2_traffic.json"))),
                        "traffic_source": TRAFFIC_SOURCE},
           "parity_spotcheck": ok}
    if leg_name == "sthv2" and world > 1:
        # Sth-Sth-v2's pool (67.9 GB of uint8 backgrounds) is not replicated: the ranks all-gather the INDEX
        # (name, owner rank, slot on owner) and the pixels stay where they were extracted
        from bgdebias_b200.pool import BackgroundPool
        names = [f"r{rank}_clip{i:06d}" for i in range(V)]
        BackgroundPool.all_gather_index(names)
        barrier()
        t0 = time.perf_counter()
        index = BackgroundPool.all_gather_index(names)
        dt = time.perf_counter() - t0
        rec["index_all_gather"] = {"ms": dt * 1e3, "entries": len(index), "parity": index[rank * V][0] == names[0] and len(index) == world * V}
    del frames, out
    torch.cuda.empty_cache()
    return rec



def time_back_to_back(calls, flush, reps, torch):
    """Median ms per call of `calls` (each on its own input / output buffers, together far larger than L2) enqueued back to
    back inside one CUDA-event bracket, L2 flushed before every bracket: the steady-state cost of a batch in a training loop.
    A bracket around a single ~80 us launch also holds ~8 us of launch latency with the GPU idle."""
    for c in calls:
        c()
    times = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for c in calls:
            c()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b) / len(calls))
    return sorted(times)[len(times) // 2]


def ragged_mix_legs(rank, world, dev):
    """The blend reading a RAGGED uint8 pool with Resize(256) inside the launch (csrc/raggedmix.cu): (a) 1,024 backgrounds
    of the widths HMDB51 / Sth-Sth-v2 mix, (b) a pool of Sth-Sth-v2's SIZE -- 220,847 backgrounds of 240x427 = 67.9 GB of
    uint8 resident on the GPU (the resized fp32 form would be 308 GB)."""
    import ctypes
    import numpy as np
    import torch
    import bgdebias_b200.ops as ops
    from bgdebias_b200 import _cabi
    from bgdebias_b200.pool import RaggedPool
    L = _cabi.lib()
    B, Tm, Hm, Wm = 64, 8, 224, 224
    gm = torch.Generator(device=dev).manual_seed(5)
    fg = torch.randint(0, 256, (B, Tm, Hm, Wm, 3), dtype=torch.uint8, device=dev, generator=gm)
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    lut = ops.make_fg_lut(mean, std, dev)
    c_mean, c_std = _cabi.f32x3(mean), _cabi.f32x3(std)
    o = torch.empty((B, Tm, 3, Hm, Wm), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    app = torch.ones(B, dtype=torch.uint8, device=dev)
    cur = torch.cuda.current_stream(dev).cuda_stream
    peak, _ = measured_peak_gbs()

    def leg(rp, label, check):
        rs = np.random.default_rng(9)
        idx = rs.integers(0, len(rp), B)
        hw = [rp.hw(int(i)) for i in idx]
        top = np.array([rs.integers(0, h - Hm + 1) for h, w in hw]); left = np.array([rs.integers(0, w - Wm + 1) for h, w in hw])
        d = lambda a: torch.tensor(np.asarray(a), dtype=torch.int32, device=dev)
        d_idx, d_top, d_left = d(idx), d(top), d(left)
        slots, tables = rp.slots_tensor, rp.tables.tensor
        mix = lambda: _cabi.check(L.bgd_bgmix_blend_ragged_f32(fg.data_ptr(), B, Tm, Hm, Wm, rp.data.data_ptr(), slots.data_ptr(), len(rp),
                                                                tables.data_ptr(), d_idx.data_ptr(), d_top.data_ptr(), d_left.data_ptr(),
                                                                app.data_ptr(), lut.data_ptr(), c_mean, c_std, 0.5, 0, o.data_ptr(), cur))
        for _ in range(3):
            mix()
        fgs = [fg] + [torch.randint(0, 256, fg.shape, dtype=torch.uint8, device=dev, generator=gm) for _ in range(3)]
        outs4 = [o] + [torch.empty_like(o) for _ in range(3)]
        mk = lambda f_, o_: (lambda: _cabi.check(L.bgd_bgmix_blend_ragged_f32(f_.data_ptr(), B, Tm, Hm, Wm, rp.data.data_ptr(), slots.data_ptr(), len(rp),
                                                                               tables.data_ptr(), d_idx.data_ptr(), d_top.data_ptr(), d_left.data_ptr(),
                                                                               app.data_ptr(), lut.data_ptr(), c_mean, c_std, 0.5, 0, o_.data_ptr(), cur)))
        ms = time_back_to_back([mk(f_, o_) for f_, o_ in zip(fgs, outs4)], flush, 7, torch)
        del fgs, outs4
        # algorithmic bytes: fg uint8 in, fp32 out, and the source window of each crop (3 planes of ~(224 h/Hb + 2) x (224 w/Wb + 2) bytes)
        bg_bytes = 0
        for i, (Hb, Wb) in zip(idx, hw):
            s_ = rp.slots[int(i)]
            bg_bytes += 3 * int(np.ceil(Hm * int(s_["h"]) / Hb + 2)) * int(np.ceil(Wm * int(s_["w"]) / Wb + 2))
        by = B * Tm * Hm * Wm * 3 * 5 + bg_bytes
        ok = None
        if check:        # against the dense path on the same images: Resize on the device (bgd_aa_resize_u8_f32), then the fp32-pool blend
            dense = torch.stack([rp.resized(int(i)) for i in idx]) if len({h for h in hw}) == 1 else None
            if dense is not None:
                ref = torch.ops.bgdebias.bgmix_blend(fg, dense, torch.arange(B, dtype=torch.int32, device=dev), d_top, d_left, app, lut,
                                                     torch.tensor(mean), torch.tensor(std), 0.5, "NTCHW")
                ok = bool(torch.equal(ref, o))
        return {"metric": "bgmix_clips_per_sec", "value": world * B / (ms * 1e-3), "unit": "clips/s", "ms_per_step": ms,
                "config": {"workload": label, "pool_images": len(rp), "pool_resident_gb": rp.used / 1e9,
                           "l2": "256 MB flush write between iterations", "note": "output stays on the device (it is the training tensor)"},
                "roofline": {"bound": "hbm", "achieved": by / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": by / (ms * 1e-3) / 1e9 / peak,
                             "traffic": (lambda r: r * by if r else None)(traffic_ratio("r2_traffic_bgmix_ragged.json")),
                             "traffic_source": TRAFFIC_SOURCE, "algorithmic_bytes_per_step": by,
                             "note": "instruction-bound: the separable Resize of the background window runs inside the launch"},
                "parity_spotcheck": ok}

    out = {}
    g = torch.Generator().manual_seed(6)
    sizes = [(240, 320), (240, 427), (240, 352), (240, 426)]
    rp = RaggedPool(256, dev)
    imgs = [torch.randint(0, 256, (3,) + sizes[i % 4], dtype=torch.uint8, generator=g) for i in range(64)]
    for _ in range(16):
        rp.append(imgs)
    out["mixed_widths"] = leg(rp, "configs[4] from a ragged uint8 pool of 1,024 backgrounds of 240x{320,427,352,426}, Resize(256) inside the launch", False)
    del rp
    n, h, w = 220_847, 240, 427
    need = n * 3 * h * w
    free, _ = torch.cuda.mem_get_info(dev)
    if need + (6 << 30) > free:
        n = int((free - (6 << 30)) // (3 * h * w))
        need = n * 3 * h * w
    data = torch.empty(need, dtype=torch.uint8, device=dev)
    gd = torch.Generator(device=dev).manual_seed(7)
    for o0 in range(0, need, 1 << 30):
        o1 = min(need, o0 + (1 << 30))
        data[o0:o1] = torch.randint(0, 256, (o1 - o0,), dtype=torch.uint8, device=dev, generator=gd)
    rp = RaggedPool.from_uniform_buffer(data, n, h, w, 256)
    out["sthv2_sized_pool"] = leg(rp, f"configs[4] from a Sth-Sth-v2-sized uint8 pool: {n} backgrounds of 240x427 resident on the GPU "
                                      "(configs[3]'s pool after the all-gather), Resize(256) inside the launch", True)
    del rp, data, flush
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    procs = max(1, min(cores, 64))
    vpp = 1
    for _ in range(args.warmup):
        cpu_reference_fps(procs, vpp, seed=99)
    t_all, f_all = 0.0, 0
    for s in range(args.steps):
        fps, frames, t = cpu_reference_fps(procs, vpp, seed=s)
        t_all += t
        f_all += frames
    value = f_all / t_all
    sample = f"{procs} videos per step (one per process), T~{WL['t_text']}, {H}x{W}x3, np.median only (no decode)"
    emit({
        "impl": "reference", "metric": "bg_extraction_frames_per_sec", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WL["label"] + f" (T~{WL['t_text']}, {H}x{W}x3 uint8); bounded sample per step",
                   "videos_per_step": procs},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": procs, "kind": cpu_kind(), "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import bgdebias_b200.ops as ops
    from bgdebias_b200 import _cabi
    import ctypes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bgdebias_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.lib()

    # ---- the rank's shard, resident in HBM --------------------------------------------------
    V = int(os.environ.get("BGD_BENCH_VIDEOS", WL["videos"]))
    rng = np.random.default_rng(1 + rank)
    Ts = draw_T(rng, V)
    free, _total = torch.cuda.mem_get_info(dev)
    while int(Ts.sum()) * N_BYTES + (4 << 30) > free and V > 16:          # another tenant on the GPU: shrink, say so
        V //= 2
        Ts = Ts[:V]
    offs = torch.from_numpy(np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64))
    rows = int(offs[-1])
    frames = torch.empty((rows, N_BYTES), dtype=torch.uint8, device=dev)
    fill_frames(frames, Ts, offs, H, W, args.content, 100 + rank, torch)
    torch.cuda.synchronize()
    step_bytes = (rows + V) * N_BYTES                      # every frame byte read once + one frame written per video

    def step():
        return torch.ops.bgdebias.temporal_median_varlen(frames, offs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    launches0 = _cabi.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        nvml_index = torch.cuda._get_nvml_device_index(local_rank)
    except Exception:
        nvml_index = local_rank
    with ClockSampler(nvml_index) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            out = step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = _cabi.kernel_launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(rows)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max, frames_total = float(t.item()), float(tot.item())
    value = frames_total * args.steps / (ms_max * 1e-3)
    ms_per_step = ms_max / args.steps

    # spot-check the timed path (outside the timed region) against the definition evaluated with plain torch:
    # (s[(T-1)//2] + s[T//2]) >> 1 over the sorted column -- what np.median(...).astype(uint8) gives
    ok = spot_check(frames, offs, Ts, out, int(np.argmin(Ts)), torch) and spot_check(frames, offs, Ts, out, int(np.argmax(Ts)), torch)

    # ---- end to end through the C ABI with host buffers ----------------------------------------
    Ve = int(os.environ.get("BGD_BENCH_E2E_VIDEOS", 48))
    Te = Ts[:Ve]
    offs_e = np.concatenate([[0], np.cumsum(Te)]).astype(np.int64)
    rows_e = int(offs_e[-1])
    h_frames = torch.empty((rows_e, N_BYTES), dtype=torch.uint8, pin_memory=True)
    h_frames.copy_(frames[:rows_e])
    h_out = torch.empty((Ve, N_BYTES), dtype=torch.uint8, pin_memory=True)
    L = _cabi.lib()
    optr = offs_e.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))

    def e2e_step():
        _cabi.check(L.bgd_temporal_median_varlen_u8_host(h_frames.data_ptr(), optr, Ve, N_BYTES, h_out.data_ptr(),
                                                         local_rank))
    for _ in range(2):
        e2e_step()
    barrier()
    k_e2e = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        e2e_step()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * rows_e * k_e2e / float(te.item())
    e2e_ok = bool(torch.equal(h_out.to(dev), out[:Ve]))
    # the ceiling of that path on this box: the same pinned bytes through plain cudaMemcpyAsync (torch's copy_), all ranks
    # at once -- what the PCIe links and the host's memory system give N concurrent H2D streams, no kernel, no D2H
    scratch = torch.empty((rows_e, N_BYTES), dtype=torch.uint8, device=dev)
    scratch.copy_(h_frames, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        scratch.copy_(h_frames, non_blocking=True)
    torch.cuda.synchronize()
    tc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    h2d_ceiling_gbs = world * rows_e * N_BYTES * k_e2e / float(tc.item()) / 1e9
    del scratch

    # ---- the path's one collective: all-gather of the background index + pixels for the mix pool
    #      (pool.BackgroundPool.all_gather).  Outside `value`; reported beside it.
    gather = None
    if world > 1:
        try:
            from bgdebias_b200.pool import BackgroundPool
            names = [f"r{rank}_v{i:05d}" for i in range(V)]
            all_names, all_bgs = BackgroundPool.all_gather(names, out)       # warm-up at full size (NCCL channels, buffer registration)
            del all_bgs
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_host = time.perf_counter()
            all_names, all_bgs = BackgroundPool.all_gather(names, out, events=(g0, g1))
            barrier()
            t_host = time.perf_counter() - t_host
            tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            same = bool(torch.equal(all_bgs[rank * V:(rank + 1) * V], out)) and len(all_names) == world * V
            gather = {"ms": float(tg.item()), "backgrounds": int(all_bgs.shape[0]), "bytes_per_rank_out": int(all_bgs.numel()),
                      "GB/s_per_rank": all_bgs.numel() / (float(tg.item()) * 1e-3) / 1e9, "backend": "nccl", "parity": same,
                      "in_place": True, "ms_with_name_exchange_host_clock": t_host * 1e3}
            del all_bgs
            # the product flow at small scale: every rank writes 16 of its backgrounds as JPEGs (what the CLI does),
            # the ranks gather the decoded files (pool.gather_extracted_backgrounds) and rank 0 compares the gathered
            # pool with its own decode of the whole directory (what BackgroundPool.from_files reads)
            import cv2, tempfile, shutil
            from torchvision.io import ImageReadMode, read_image
            from bgdebias_b200.pool import gather_extracted_backgrounds
            holder = [tempfile.mkdtemp(prefix="bgd_bench_pool_") if rank == 0 else None]
            dist.broadcast_object_list(holder, src=0)
            small = out[:16].cpu().numpy().reshape(16, H, W, 3)
            for i in range(16):
                cv2.imwrite(os.path.join(holder[0], f"r{rank}_v{i:03d}.jpg"), small[i])
            barrier()
            paths, rp = gather_extracted_backgrounds(holder[0], ".jpg", 256, dev)
            flow_ok = len(paths) == 16 * world
            if rank == 0:
                for k in range(0, len(paths), max(1, len(paths) // 8)):
                    s_ = rp.slots[k]
                    ref = read_image(paths[k], mode=ImageReadMode.RGB).to(dev)
                    o_ = int(s_["offset"])
                    flow_ok = flow_ok and bool(torch.equal(rp.data[o_:o_ + ref.numel()].view_as(ref), ref))
            barrier()
            if rank == 0:
                shutil.rmtree(holder[0], ignore_errors=True)
            gather["product_flow"] = {"files": len(paths), "equals_decoding_the_directory": flow_ok}
        except Exception as e:
            gather = {"error": repr(e)}

    # ---- BG-mix (configs[4]): 64 clips x 8 x 224 x 224 per step --------------------------------
    bgmix = None
    try:
        B, Tm, Hm, Wm, P = 64, 8, 224, 224, 1024
        gm = torch.Generator(device=dev).manual_seed(4)
        fg = torch.randint(0, 256, (B, Tm, Hm, Wm, 3), dtype=torch.uint8, device=dev, generator=gm)
        pool = torch.rand((P, 3, 256, 341), device=dev, generator=gm) * 255.0
        torch.manual_seed(0)
        idx = torch.randint(0, P, (B,)).int().to(dev)
        top = torch.randint(0, 33, (B,)).int().to(dev)
        left = torch.randint(0, 118, (B,)).int().to(dev)
        app = torch.ones(B, dtype=torch.uint8, device=dev)
        mean, std = torch.tensor([123.675, 116.28, 103.53]), torch.tensor([58.395, 57.12, 57.375])
        lut = ops.make_fg_lut(mean.tolist(), std.tolist(), dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > L2 (126 MB)
        # the timed call goes straight to the C ABI on torch's current stream (what the custom op does after its checks):
        # the bracket then holds the launch, not ~10 us of Python dispatch with the GPU idle
        o = torch.empty((B, Tm, 3, Hm, Wm), dtype=torch.float32, device=dev)
        c_mean, c_std = _cabi.f32x3(mean.tolist()), _cabi.f32x3(std.tolist())
        cur = torch.cuda.current_stream(dev).cuda_stream
        mix = lambda: _cabi.check(L.bgd_bgmix_blend_f32(fg.data_ptr(), B, Tm, Hm, Wm, pool.data_ptr(), P, 256, 341, idx.data_ptr(),
                                                        top.data_ptr(), left.data_ptr(), app.data_ptr(), lut.data_ptr(), c_mean, c_std,
                                                        0.5, 0, o.data_ptr(), cur))
        for _ in range(3):
            mix()
        assert torch.equal(o, torch.ops.bgdebias.bgmix_blend(fg, pool, idx, top, left, app, lut, mean, std, 0.5, "NTCHW"))
        times = []
        for _ in range(max(5, args.steps)):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); mix(); b.record(); torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        mix_ms_single = sorted(times)[len(times) // 2]
        # steady state: 4 batches (own clips, own output tensors: 1.5 GB, twelve times the L2) back to back per bracket
        fgs = [fg] + [torch.randint(0, 256, fg.shape, dtype=torch.uint8, device=dev, generator=gm) for _ in range(3)]
        outs4 = [o] + [torch.empty_like(o) for _ in range(3)]
        mk = lambda f_, o_: (lambda: _cabi.check(L.bgd_bgmix_blend_f32(f_.data_ptr(), B, Tm, Hm, Wm, pool.data_ptr(), P, 256, 341, idx.data_ptr(),
                                                                        top.data_ptr(), left.data_ptr(), app.data_ptr(), lut.data_ptr(), c_mean, c_std,
                                                                        0.5, 0, o_.data_ptr(), cur)))
        mix_ms = time_back_to_back([mk(f_, o_) for f_, o_ in zip(fgs, outs4)], flush, max(5, args.steps), torch)
        del fgs, outs4
        mix_bytes = B * (Tm * Hm * Wm * 3 * 5 + Hm * Wm * 3 * 4)
        # end to end: pinned uint8 clips + draws in host memory -> training tensor on device + checksum back
        h_fg = torch.empty(fg.shape, dtype=torch.uint8, pin_memory=True); h_fg.copy_(fg)
        hi, ht, hl, ha = (x.cpu().numpy().copy() for x in (idx, top, left, app))
        d_out = torch.empty((B, Tm, 3, Hm, Wm), dtype=torch.float32, device=dev)
        chk = ctypes.c_double()
        call = lambda: _cabi.check(L.bgd_bgmix_blend_f32_host(
            h_fg.data_ptr(), B, Tm, Hm, Wm, pool.data_ptr(), P, 256, 341, hi.ctypes.data, ht.ctypes.data, hl.ctypes.data,
            ha.ctypes.data, lut.data_ptr(), _cabi.f32x3(mean.tolist()), _cabi.f32x3(std.tolist()), 0.5, 0,
            d_out.data_ptr(), ctypes.byref(chk), local_rank))
        call(); call()
        t0 = time.perf_counter()
        for _ in range(5):
            call()
        mix_e2e = 5 * B / (time.perf_counter() - t0)
        # CPU arm of the mix (rank 0, N=1): what a DataLoader worker of the reference does per mixed clip
        # (comix_loader.py:138-145 + :72-75): Resize(256) of a 240x320 background, RandomCrop(224), Normalize,
        # blend with the already normalised fp32 clip -- the oracle's restatement, one thread, bounded sample
        mix_cpu = None
        if rank == 0 and world == 1:
            try:
                from oracle import bgmix_oracle as bo
                torch.set_num_threads(1)
                rs = np.random.default_rng(3)
                n_cpu = 48
                fgn = [torch.from_numpy(bo.fg_normalize(rs.integers(0, 256, (Tm, Hm, Wm, 3), dtype=np.uint8), bo.fg_lut())) for _ in range(4)]
                bgs = [torch.from_numpy(rs.integers(0, 256, (3, 240, 320), dtype=np.uint8)) for _ in range(4)]
                bo.mix_clip_like_reference(fgn[0], bgs[0])                     # imports + first-call set-up
                tc = time.perf_counter()
                for i in range(n_cpu):
                    bo.mix_clip_like_reference(fgn[i % 4], bgs[i % 4])
                tc = time.perf_counter() - tc
                mix_cpu = {"value": n_cpu / tc, "unit": "clips/s", "cores": 1, "kind": "port",
                           "sample": f"{n_cpu} clips of [8,3,224,224] fp32 + 240x320 background, Resize(256)+RandomCrop+Normalize+blend per clip, one thread (one DataLoader worker; the shipped UCF101 config runs 4 per GPU)"}
            except Exception as e:
                mix_cpu = {"value": None, "unit": "clips/s", "cores": 0, "kind": "port", "sample": "failed: " + repr(e)}
        peak, _ = measured_peak_gbs()
        mix_traffic = None
        try:
            mix_traffic = json.loads((ROOT / "profiles" / "r2_traffic_bgmix.json").read_text())["dram_bytes_per_algorithmic_byte"] * mix_bytes
        except Exception:
            pass
        bgmix = {"metric": "bgmix_clips_per_sec", "value": world * B / (mix_ms * 1e-3), "unit": "clips/s",
                 "ms_per_step": mix_ms, "ms_single_launch_bracket": mix_ms_single,
                 "timing": "4 batches with their own clips and output tensors (1.5 GB) back to back per CUDA-event bracket, L2 flushed before each bracket, median; ms_single_launch_bracket = the same kernel bracketed alone (holds the launch latency)", "config": {"workload": "configs[4]: fg u8 [64,8,224,224,3], fp32 pool 1024x3x256x341, alpha 0.5, all samples mixed, out fp32 [64,8,3,224,224]", "l2": "256 MB flush write between iterations",
                            "e2e_note": "e2e ships pinned uint8 clips + draws to the device and leaves the fp32 training tensor ON the device, where the model consumes it; only a checksum (8 bytes) comes back"},
                 "roofline": {"bound": "hbm", "achieved": mix_bytes / (mix_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": mix_bytes / (mix_ms * 1e-3) / 1e9 / peak, "traffic": mix_traffic},
                 "e2e": {"value": world * mix_e2e, "unit": "clips/s", "h2d_bytes_per_step": int(h_fg.numel() + B * 13),
                         "d2h_bytes_per_step": 8},
                 "cpu_baseline": mix_cpu,
                 "parity_spotcheck": bool(torch.equal(d_out, o))}
        # the same step when the DataLoader ships MultiScaleCrop's uint8 crops and the pipeline's last
        # Resize((224, 224), keep_ratio=False) (config :136) runs inside the blend launch (cv2 INTER_LINEAR, bit-exact)
        try:
            sizes = [(256, 256), (224, 256), (256, 224), (224, 224), (192, 224), (224, 192), (192, 192), (168, 192), (192, 168), (168, 168)]
            gc = torch.Generator().manual_seed(4)
            crops = [torch.randint(0, 256, (Tm, *sizes[i % len(sizes)], 3), dtype=torch.uint8, generator=gc) for i in range(B)]
            buf, geom = ops.pack_clips(crops)
            d_buf = buf.to(dev)
            o2 = torch.empty((B, Tm, 3, Hm, Wm), dtype=torch.float32, device=dev)
            gptr = ctypes.cast(geom.data_ptr(), ctypes.POINTER(ctypes.c_int64))
            tail = lambda: _cabi.check(L.bgd_bgmix_resize_blend_f32(
                d_buf.data_ptr(), d_buf.numel(), gptr, B, Tm, Hm, Wm, pool.data_ptr(), 0, P, 256, 341, idx.data_ptr(), top.data_ptr(),
                left.data_ptr(), app.data_ptr(), lut.data_ptr(), c_mean, c_std, 0.5, 0, o2.data_ptr(), cur))
            for _ in range(3):
                tail()
            bufs = [d_buf] + [d_buf.clone() for _ in range(3)]
            outs4 = [o2] + [torch.empty_like(o2) for _ in range(3)]
            mkt = lambda s_, o_: (lambda: _cabi.check(L.bgd_bgmix_resize_blend_f32(
                s_.data_ptr(), s_.numel(), gptr, B, Tm, Hm, Wm, pool.data_ptr(), 0, P, 256, 341, idx.data_ptr(), top.data_ptr(),
                left.data_ptr(), app.data_ptr(), lut.data_ptr(), c_mean, c_std, 0.5, 0, o_.data_ptr(), cur)))
            tail_ms = time_back_to_back([mkt(s_, o_) for s_, o_ in zip(bufs, outs4)], flush, max(5, args.steps), torch)
            del bufs, outs4
            tail_bytes = sum(c.numel() for c in crops) + B * Hm * Wm * 3 * 4 * (1 + Tm)
            two = torch.ops.bgdebias.bgmix_blend(torch.ops.bgdebias.resize_bilinear(d_buf, geom, Tm, Hm, Wm), pool, idx, top, left,
                                                 app, lut, mean, std, 0.5, "NTCHW")
            # end to end: pinned packed crops + draws in host memory -> training tensor on device + checksum back
            h_buf = torch.empty(buf.shape, dtype=torch.uint8, pin_memory=True); h_buf.copy_(buf)
            call2 = lambda: _cabi.check(L.bgd_bgmix_resize_blend_f32_host(
                h_buf.data_ptr(), h_buf.numel(), gptr, B, Tm, Hm, Wm, pool.data_ptr(), P, 256, 341, hi.ctypes.data, ht.ctypes.data,
                hl.ctypes.data, ha.ctypes.data, lut.data_ptr(), c_mean, c_std, 0.5, 0, d_out.data_ptr(), ctypes.byref(chk), local_rank))
            call2(); call2()
            t0 = time.perf_counter()
            for _ in range(5):
                call2()
            tail_e2e = 5 * B / (time.perf_counter() - t0)
            tail_e2e_ok = bool(torch.equal(d_out, o2))
            tail_cpu = None
            if rank == 0 and world == 1:
                import cv2
                cv2.setNumThreads(1)
                frames = [c[t].numpy() for c in crops[:10] for t in range(Tm)]
                cv2.resize(frames[0], (Wm, Hm), interpolation=cv2.INTER_LINEAR)
                tc = time.perf_counter()
                for _ in range(4):
                    for fr in frames:
                        cv2.resize(fr, (Wm, Hm), interpolation=cv2.INTER_LINEAR)
                tc = time.perf_counter() - tc
                tail_cpu = {"value": 4 * len(frames) / Tm / tc, "unit": "clips/s", "cores": 1, "kind": "reference",
                            "sample": f"{4 * len(frames)} cv2.resize(frame, (224, 224), INTER_LINEAR) calls (the Resize step alone, 8 per clip), one thread"}
            tail_traffic = None
            try:
                tail_traffic = json.loads((ROOT / "profiles" / "r2_traffic_resize_blend.json").read_text())["dram_bytes_per_algorithmic_byte"] * tail_bytes
            except Exception:
                pass
            bgmix["with_resize"] = {
                "metric": "bgmix_clips_per_sec", "value": world * B / (tail_ms * 1e-3), "unit": "clips/s", "ms_per_step": tail_ms,
                "config": {"workload": "configs[4] with the foreground as MultiScaleCrop-sized uint8 crops (168..256 px, packed), "
                                       "Resize((224,224)) + Normalize + FormatShape + blend in one launch"},
                "roofline": {"bound": "hbm", "achieved": tail_bytes / (tail_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": tail_bytes / (tail_ms * 1e-3) / 1e9 / peak, "traffic": tail_traffic,
                             "note": "integer-ALU bound (issue slots ~70 % busy, profiles/r1_ncu_resize_blend.txt), not HBM bound"},
                "e2e": {"value": world * tail_e2e, "unit": "clips/s", "h2d_bytes_per_step": int(h_buf.numel() + B * 13),
                        "d2h_bytes_per_step": 8},
                "cpu_baseline": tail_cpu, "parity_spotcheck": bool(torch.equal(o2, two)) and tail_e2e_ok}
            del d_buf, o2, two
        except Exception as e:
            bgmix["with_resize"] = {"error": repr(e)}
        del fg, pool, flush, d_out, h_fg
    except Exception as e:           # the headline metric stands even if the secondary bench cannot run
        bgmix = {"error": repr(e)}

    # ---- the other extraction configs + structured content, as bounded resident chunks in the same run --------
    legs = {}
    if not args.no_legs and WL_NAME == "ucf101":
        del frames, out, h_frames, h_out
        torch.cuda.empty_cache()
        for leg_name, spec in EXTRA_LEGS.items():
            try:
                legs[leg_name] = extraction_leg(leg_name, spec, rank, world, dev, max(3, min(args.steps, 8)), barrier)
            except Exception as e:
                legs[leg_name] = {"error": repr(e)}
                torch.cuda.empty_cache()

    # ---- BG-mix from ragged uint8 pools: mixed image sizes (HMDB51 / Sth-Sth-v2 widths) and a Sth-Sth-v2-SIZED pool ----
    if isinstance(bgmix, dict) and "error" not in bgmix and not args.no_legs:
        try:
            bgmix["ragged"] = ragged_mix_legs(rank, world, dev)
        except Exception as e:
            bgmix["ragged"] = {"error": repr(e)}
            torch.cuda.empty_cache()

    if rank == 0:
        cb = None
        try:
            if world > 1:
                raise RuntimeError("reported at N=1 only")
            procs = max(1, min(host_cores(), 32))
            fps, nfr, tt = cpu_reference_fps(procs, 2, seed=7)
            cb = {"value": fps, "unit": "frames/s", "cores": procs, "kind": cpu_kind(),
                  "sample": f"{2 * procs} videos of the same workload ({nfr} frames, {tt:.1f} s), np.median(frames,0).astype(uint8) per video "
                            "(extract_background.py:73; the reference's own function where its tree is present, else the oracle's restatement), one process per core"}
        except Exception as e:
            cb = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port",
                  "sample": ("not run: " if world > 1 else "failed: ") + str(e)}
        peak, peak_src = measured_peak_gbs()
        achieved = step_bytes / (ms_per_step * 1e-3) / 1e9          # this rank's step bytes / max step time
        traffic = None
        tf = ROOT / "profiles" / "r2_traffic.json"
        if tf.exists():
            try:
                ratio = json.loads(tf.read_text())["dram_bytes_per_algorithmic_byte"]
                traffic = ratio * step_bytes
            except Exception:
                traffic = None
        emit({
            "metric": "bg_extraction_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WL["label"], "videos_per_gpu": V,
                       "frames_per_gpu": rows, "frame_shape": [H, W, 3], "T": WL["t_text"], "resident_gb": rows * N_BYTES / 1e9,
                       "l2": "resident inputs (tens of GB per GPU) exceed L2; no flush needed", "kernel": "median_ldsm for T<=512, median_colplane above (AUTO)", "content": args.content,
                       "parallelism": f"shard x{world}, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": rows_e * N_BYTES,
                    "d2h_bytes_per_step": Ve * N_BYTES, "videos_per_step": Ve, "parity": e2e_ok,
                    "h2d_GBs": e2e_value * N_BYTES / 1e9,
                    "h2d_ceiling_GBs": h2d_ceiling_gbs, "frac_of_h2d_ceiling": e2e_value * N_BYTES / 1e9 / h2d_ceiling_gbs,
                    "ceiling": "the same pinned bytes by plain cudaMemcpyAsync on all ranks at once (no kernel, no D2H), measured in this run"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "profiles ncu capture (dram bytes per algorithmic byte of one --set full capture) x this step's algorithmic bytes",
                         "peak_source": peak_src, "algorithmic_bytes_per_step": step_bytes,
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "cpu_baseline": cb,
            "clocks": clk.summary(),
            "parity_spotcheck": ok,
            "bgmix": bgmix,
            "pool_all_gather": gather,
            "workloads": legs,
        })
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=WL_NAME, choices=sorted(WORKLOADS))
    ap.add_argument("--content", default="random", choices=["random", "structured"], help="synthetic frame content of the main leg")
    ap.add_argument("--no-legs", dest="no_legs", action="store_true", help="skip the extra workload legs (hmdb51, sthv2, structured)")
    args = ap.parse_args()
    _select_workload(args.workload)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
