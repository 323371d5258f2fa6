"""Concurrent pinned host-to-device copy ceiling of the box (development aid for the e2e numbers).

    python tools/h2d_ceiling.py                     one GPU
    torchrun --nproc-per-node 8 tools/h2d_ceiling.py     8 ranks copying at once

Per rank: one pinned host buffer (default flags, then write-combined: cudaHostAllocWriteCombined -- the host only writes
the staging slabs, the device only reads them), copied to the device with plain cudaMemcpyAsync, one call per copy, in
chunks of 256 MB (the staging slab size of the host pipeline).  Prints per-rank and aggregate GB/s."""
import ctypes, os, sys, time
import torch
import torch.distributed as dist

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rt = ctypes.CDLL("libcudart.so.12")
GB = int(os.environ.get("GB", 2))
nbytes = GB << 30
chunk = 256 << 20
dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(flags, label):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags)) == 0
    ctypes.memset(p, 1, nbytes)
    stream = torch.cuda.current_stream().cuda_stream
    def once():
        for o in range(0, nbytes, chunk):
            assert rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr() + o), ctypes.c_void_p(p.value + o), ctypes.c_size_t(min(chunk, nbytes - o)),
                                      ctypes.c_int(1), ctypes.c_void_p(stream)) == 0
    once(); barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mine = 5 * nbytes / dt / 1e9
    agg = world * 5 * nbytes / float(t.item()) / 1e9
    allr = [None] * world
    if world > 1:
        dist.all_gather_object(allr, round(mine, 1))
    else:
        allr = [round(mine, 1)]
    if rank == 0:
        print(f"{label:28s} ranks={world}  aggregate {agg:7.1f} GB/s  per rank {allr}", flush=True)
    barrier()
    rt.cudaFreeHost(p)


try:
    cores = len(os.sched_getaffinity(0))
except Exception:
    cores = os.cpu_count()
if rank == 0:
    numa = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")] if os.path.isdir("/sys/devices/system/node") else []
    print(f"host cores {cores}, NUMA nodes {len(numa)}, buffer {GB} GB per rank, chunks of 256 MB", flush=True)
run(0, "pinned (default)")
run(4, "pinned write-combined")
run(0, "pinned (default) again")
if world > 1:
    dist.destroy_process_group()
