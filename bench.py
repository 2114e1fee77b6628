#!/usr/bin/env python
"""bench.py -- frames/s of the RAISR luma+chroma hot path on B200 (BASELINE.json configs[1]):
1080p -> 4K yuv420p, filters_2x/filters_lowres, passes=1, bits=8.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one batch of FRAMES_PER_STEP frames through the whole per-frame path (luma pass kernel + two chroma
resizes).  `value` = frames/s with the planes already resident in HBM (device entry point of the C ABI, CUDA events
on the launching stream); `e2e` = frames/s through the blocking host-pointer call RNLHandler_Process binds
(pinned host planes, H2D and D2H inside the timed region).  Multi-GPU = frame-parallel shards (each rank its own
frames, no data-path collective): weak scaling.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import importlib.util
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import raisr_testlib as T  # noqa: E402  (synthetic frames + the reference/oracle bindings for the CPU legs)

IN_W, IN_H, OUT_W, OUT_H = 1920, 1080, 3840, 2160
FOLDER = "filters_2x/filters_lowres"
FRAMES_PER_STEP = 16
NBUF = 12                      # distinct in/out frame sets rotated through: 12 x 15.55 MB = 187 MB > 126 MB L2
BYTES_Y = IN_W * IN_H + OUT_W * OUT_H                       # algorithmic bytes of the luma kernel per launch
BYTES_FRAME = BYTES_Y + 2 * (IN_W // 2 * IN_H // 2 + OUT_W // 2 * OUT_H // 2)
WORKLOAD = "1080p->4K yuv420p, filters_2x/filters_lowres, passes=1, bits=8"


def load_binding():
    spec = importlib.util.spec_from_file_location("raisr_binding", os.path.join(T.PKG_DIR, "binding.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 7]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[0]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][1])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        seen = set()
        for r in rows:
            for n, v in zip(names, r[4:8]):
                if v.strip().lower().startswith("active"):
                    seen.add(n)
        out["reasons"] = sorted(seen)
        out["samples"] = len(rows)
        return out


def reference_arm(args, rank):
    """The reference's own CPU implementation of the path (oracle/_ref = untouched sources + IPP stand-in),
    all host threads, bounded sample per step."""
    if rank != 0:
        return
    L = T.handler_lib(T.ref_lib_path())
    threads = os.cpu_count() or 1
    frames = 2                                                 # frames per step: bounded sample of the workload
    folder = T.filter_folder(FOLDER)
    y = T.synth_frame(IN_W, IN_H, 8, 1234)
    u, v = T.synth_chroma(IN_W // 2, IN_H // 2, 8, 1), T.synth_chroma(IN_W // 2, IN_H // 2, 8, 2)
    oy = np.zeros((OUT_H, OUT_W), np.uint8)
    ou = np.zeros((OUT_H // 2, OUT_W // 2), np.uint8)
    ov = np.zeros_like(ou)
    vs = [T.vdt(a) for a in (y, u, v, oy, ou, ov)]
    refs = [ctypes.byref(x) for x in vs]
    sys.stdout.flush()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)                                        # the library prints a banner on stdout
    try:
        assert L.RNLHandler_Init(folder.encode(), 2.0, 8, T.VideoRange, threads, T.AVX512, 1, 1) == 0
        assert L.RNLHandler_SetRes(*refs) == 0
        for _ in range(args.warmup):
            for _ in range(frames):
                L.RNLHandler_Process(*refs, T.CountOfBitsChanged)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for _ in range(frames):
                L.RNLHandler_Process(*refs, T.CountOfBitsChanged)
        dt = time.perf_counter() - t0
        L.RNLHandler_Deinit()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
    fps = args.steps * frames / dt
    line = {
        "impl": "reference", "metric": "frames/sec 1080p->4K 2x RAISR (yuv420p frame: Y pass + chroma resize)",
        "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": frames},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "reference",
                         "sample": "%d frames per step of the same 1080p->4K workload; untouched reference sources, "
                                   "AVX512 fp32 path, threadcount=%d, IPP replaced by oracle/ipp_standin" % (frames, threads)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def cpu_baseline_leg():
    """oracle/_ref timed on this host's cores: bounded sample (about 10-30 s of CPU work)."""
    if not T.have_ref():
        # fall back to the C restatement (single thread)
        folder = T.filter_folder(FOLDER)
        m = T.OracleModel(folder, 8)
        y = T.synth_frame(IN_W // 2, IN_H // 2, 8, 1234)
        t0 = time.perf_counter()
        T.oracle_process_y(y, IN_W, IN_H, m)
        dt = time.perf_counter() - t0
        return {"value": 0.25 / dt, "unit": "frames/s", "cores": 1, "kind": "port",
                "sample": "one 540p->1080p luma frame (1/4 of the workload's pixels) through oracle/raisr_oracle.c, scaled by 1/4"}
    L = T.handler_lib(T.ref_lib_path())
    threads = os.cpu_count() or 1
    folder = T.filter_folder(FOLDER)
    y = T.synth_frame(IN_W, IN_H, 8, 1234)
    u, v = T.synth_chroma(IN_W // 2, IN_H // 2, 8, 1), T.synth_chroma(IN_W // 2, IN_H // 2, 8, 2)
    oy = np.zeros((OUT_H, OUT_W), np.uint8)
    ou = np.zeros((OUT_H // 2, OUT_W // 2), np.uint8)
    ov = np.zeros_like(ou)
    refs = [ctypes.byref(x) for x in [T.vdt(a) for a in (y, u, v, oy, ou, ov)]]
    sys.stdout.flush()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        assert L.RNLHandler_Init(folder.encode(), 2.0, 8, T.VideoRange, threads, T.AVX512, 1, 1) == 0
        assert L.RNLHandler_SetRes(*refs) == 0
        L.RNLHandler_Process(*refs, T.CountOfBitsChanged)
        n, t0 = 0, time.perf_counter()
        while n < 12 or (time.perf_counter() - t0 < 2.0 and n < 64):
            L.RNLHandler_Process(*refs, T.CountOfBitsChanged)
            n += 1
        dt = time.perf_counter() - t0
        L.RNLHandler_Deinit()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
    return {"value": n / dt, "unit": "frames/s", "cores": threads, "kind": "reference",
            "sample": "%d frames of the same 1080p->4K yuv420p workload after 1 warm-up; untouched reference sources "
                      "(AVX512 fp32 path, threadcount=%d), IPP replaced by oracle/ipp_standin" % (n, threads)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the RAISR engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B = load_binding()
    folder = T.filter_folder(FOLDER)
    eng = B.Engine(folder, 2.0, 8, T.VideoRange, 1, 1, device=local, numerics=B.NUMERICS_AUTO)
    eng.set_res(IN_W, IN_H, OUT_W, OUT_H, IN_W // 2, IN_H // 2, OUT_W // 2, OUT_H // 2)

    # ---- synthetic frames: NBUF distinct sets, pinned on the host and resident on the device -----------------
    dev = torch.device("cuda", local)
    h_in, d_in, h_out, d_out = [], [], [], []
    for i in range(NBUF):
        seed = 1234 + 97 * rank + i
        planes = [T.synth_frame(IN_W, IN_H, 8, seed), T.synth_chroma(IN_W // 2, IN_H // 2, 8, seed + 1),
                  T.synth_chroma(IN_W // 2, IN_H // 2, 8, seed + 2)]
        hp = [torch.from_numpy(p).pin_memory() for p in planes]
        h_in.append(hp)
        d_in.append([p.to(dev) for p in hp])
        ho = [torch.empty((OUT_H, OUT_W), dtype=torch.uint8).pin_memory(),
              torch.empty((OUT_H // 2, OUT_W // 2), dtype=torch.uint8).pin_memory(),
              torch.empty((OUT_H // 2, OUT_W // 2), dtype=torch.uint8).pin_memory()]
        h_out.append(ho)
        d_out.append([torch.empty_like(o, device=dev) for o in ho])
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)

    def dev_frame(i):
        a, o = d_in[i % NBUF], d_out[i % NBUF]
        rc = eng.process_device(a[0].data_ptr(), a[0].stride(0), o[0].data_ptr(), o[0].stride(0),
                                a[1].data_ptr(), a[1].stride(0), a[2].data_ptr(), a[2].stride(0),
                                o[1].data_ptr(), o[1].stride(0), o[2].data_ptr(), o[2].stride(0), 2, sptr)
        assert rc == 0

    def luma_only(i):
        a, o = d_in[i % NBUF], d_out[i % NBUF]
        rc = eng.process_device_rows(a[0].data_ptr(), a[0].stride(0), o[0].data_ptr(), o[0].stride(0), 0, OUT_H, 2, sptr)
        assert rc == 0

    def host_frame(i):
        a, o = h_in[i % NBUF], h_out[i % NBUF]
        rc = eng.L.raisr_cuda_process_host(eng.h, a[0].data_ptr(), a[0].stride(0), a[1].data_ptr(), a[1].stride(0),
                                           a[2].data_ptr(), a[2].stride(0), o[0].data_ptr(), o[0].stride(0),
                                           o[1].data_ptr(), o[1].stride(0), o[2].data_ptr(), o[2].stride(0), 2)
        assert rc == 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value") --------------------------------------------------------------
    n = 0
    for _ in range(args.warmup):
        for _ in range(FRAMES_PER_STEP):
            dev_frame(n); n += 1
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        for _ in range(FRAMES_PER_STEP):
            dev_frame(n); n += 1
    e1.record(stream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count() - launches0
    if world > 1:
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())

    # ---- dominant kernel alone (roofline): luma pass launches, CUDA events on the launching stream -----------
    for _ in range(FRAMES_PER_STEP):
        luma_only(n); n += 1
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nk = args.steps * FRAMES_PER_STEP
    k0.record(stream)
    for _ in range(nk):
        luma_only(n); n += 1
    k1.record(stream)
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / nk
    clocks = sampler.stop() if sampler else None

    # ---- end to end through the host-pointer C ABI ("e2e") ---------------------------------------------------
    for _ in range(max(1, args.warmup)):
        for _ in range(FRAMES_PER_STEP):
            host_frame(n); n += 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(FRAMES_PER_STEP):
            host_frame(n); n += 1
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)

    # spot check of the last host frame against nothing but itself being written (non-zero, in range)
    last = h_out[(n - 1) % NBUF][0]
    assert int(last.max()) > 0

    frames_total = world * args.steps * FRAMES_PER_STEP
    value = frames_total / (dev_ms * 1e-3)
    e2e = frames_total / e2e_s
    if rank == 0:
        peak, how = measured_peaks()
        achieved = BYTES_Y / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": "frames/sec 1080p->4K 2x RAISR (yuv420p frame: Y pass + chroma resize)",
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": FRAMES_PER_STEP, "sharding": "frame-parallel, no collective",
                       "l2": "rotating %d distinct frame sets (%.0f MB > 126 MB L2)" % (NBUF, NBUF * BYTES_FRAME / 1e6),
                       "numerics": "x86-exact (bit-identical to the compiled reference)" if eng.numerics() == 1 else "ieee"},
            "e2e": {"value": e2e, "unit": "frames/s",
                    "h2d_bytes_per_step": world * FRAMES_PER_STEP * (IN_W * IN_H + 2 * (IN_W // 2) * (IN_H // 2)),
                    "d2h_bytes_per_step": world * FRAMES_PER_STEP * (OUT_W * OUT_H + 2 * (OUT_W // 2) * (OUT_H // 2))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "raisr_pass_kernel<uint8_t>" if os.environ.get("RAISR_CUDA_KERNEL") == "tile" else "raisr_pass_pipe_kernel<uint8_t,4,1>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": how,
                         "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": BYTES_Y,
                         "note": "compute-bound stencil: fp32 issue + shared-memory gather, see DESIGN.md"},
            "cpu_baseline": cpu_baseline_leg() if world == 1 else None,      # timed at N=1 only
        }
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            try:
                tj = json.load(open(traffic_file))
                line["roofline"]["traffic"] = tj.get("dram_bytes_per_launch")
                # the resources that actually bound the kernel: per-launch counts from the committed ncu
                # capture, rates from this run's kernel time and the SM clock sampled during the timed region
                mhz = (clocks or {}).get("sm_mhz") or 1965.0
                sms = torch.cuda.get_device_properties(local).multi_processor_count
                wf, wi = tj.get("smem_wavefronts_per_launch"), tj.get("warp_instructions_per_launch")
                if wf and wi:
                    line["roofline"]["onchip"] = {
                        "shared_memory_pipe": {"wavefronts_per_launch": wf, "peak_per_s": sms * mhz * 1e6,
                                               "frac": wf / (kern_ms * 1e-3) / (sms * mhz * 1e6)},
                        "issue_slots": {"warp_instructions_per_launch": wi, "peak_per_s": 4 * sms * mhz * 1e6,
                                        "frac": wi / (kern_ms * 1e-3) / (4 * sms * mhz * 1e6)},
                        "source": tj.get("source")}
            except Exception:
                pass
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
