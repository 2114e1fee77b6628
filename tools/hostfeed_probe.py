"""Host-feed probe for the N-GPU end-to-end curve: every rank copies frame-sized planes between ITS page-locked host memory and
ITS GPU, first alone (ranks take turns), then all ranks at once.  The ratio tells whether the host side (memory channels, NUMA
placement, PCIe root complexes shared between GPUs) or the GPU side limits e2e frames/s at N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/hostfeed_probe.py [--bind]

--bind: CPU affinity + preferred memory node of the rank's GPU before the page-locked planes are allocated (what bench.py does).
Prints one JSON line on rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
numa = None
if "--bind" in sys.argv:
    import bench
    numa = bench.bind_to_gpu_numa_node(local)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
OUT, IN = 3840 * 2160 * 3 // 2, 1920 * 1080 * 3 // 2                     # bytes of a 4K / 1080p yuv420p frame
NB = 8
h_out = [torch.empty(OUT, dtype=torch.uint8).pin_memory() for _ in range(NB)]
h_in = [torch.empty(IN, dtype=torch.uint8).pin_memory() for _ in range(NB)]
d_out = torch.empty(OUT, dtype=torch.uint8, device=dev)
d_in = torch.empty(IN, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def frames(n):
    """n frames worth of traffic: 1080p frame in, 4K frame out, both directions at once (like the engine's pipeline)"""
    for i in range(n):
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in[i % NB], non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out[i % NB].copy_(d_out, non_blocking=True)
    s_in.synchronize(); s_out.synchronize()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(n=200):
    frames(20)
    t0 = time.perf_counter()
    frames(n)
    return n / (time.perf_counter() - t0)


solo = 0.0
for r in range(world):                                                     # one rank at a time
    barrier()
    if r == rank:
        solo = timed()
barrier()
together = timed()                                                         # all ranks at once
barrier()
vals = torch.tensor([solo, together], dtype=torch.float64, device=dev)
allv = [torch.zeros_like(vals) for _ in range(world)]
if world > 1:
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    solo_l = [float(v[0]) for v in allv]
    tog_l = [float(v[1]) for v in allv]
    gb = (IN + OUT) / 1e9
    print(json.dumps({"ranks": world, "bind": "--bind" in sys.argv, "numa_rank0": numa,
                      "frames_per_s_alone": [round(x, 1) for x in solo_l], "frames_per_s_all_at_once": [round(x, 1) for x in tog_l],
                      "GBps_alone_per_rank": [round(x * gb, 1) for x in solo_l], "GBps_all_at_once_total": round(sum(tog_l) * gb, 1),
                      "feed_limit_frames_per_s_total": round(sum(tog_l), 1),
                      "cpus_allowed": len(os.sched_getaffinity(0)),
                      "numa_nodes": sorted(n for n in os.listdir("/sys/devices/system/node") if n.startswith("node")) if os.path.isdir("/sys/devices/system/node") else None}))
if world > 1:
    dist.destroy_process_group()
