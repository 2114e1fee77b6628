"""Kernel-level timing of the BASELINE.json configurations on the GPU box (luma passes, device-resident planes, CUDA events).
usage: python tools/kbench_configs.py"""
import os, sys, importlib.util
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import raisr_testlib as T
spec = importlib.util.spec_from_file_location("b", os.path.join(T.PKG_DIR, "binding.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
CONFIGS = [
    ("configs[0] 540p->1080p lowres p1 8b", "filters_2x/filters_lowres", 2.0, 8, 1, 1, 960, 540),
    ("configs[1] 1080p->4K lowres p1 8b", "filters_2x/filters_lowres", 2.0, 8, 1, 1, 1920, 1080),
    ("configs[2] 1080p->4K highres p2 8b", "filters_2x/filters_highres", 2.0, 8, 2, 1, 1920, 1080),
    ("configs[3] 4K->8K denoise p2 mode2 10b", "filters_2x/filters_denoise", 2.0, 10, 2, 2, 3840, 2160),
    ("configs[4] 720p->1080p 1.5x highres p1 8b (the folder has no pass-2 tables)", "filters_1.5x/filters_highres", 1.5, 8, 1, 1, 1280, 720),
    ("configs[4] geometry, 1.5x denoise p2 mode2 8b", "filters_1.5x/filters_denoise", 1.5, 8, 2, 2, 1280, 720),
]
NB = 6
for name, folder, ratio, bits, passes, mode, w, h in CONFIGS:
    oW, oH = int(w * ratio), int(h * ratio)
    tdt = torch.uint8 if bits == 8 else torch.int16
    ys = [torch.from_numpy(T.synth_frame(w, h, bits, 1234 + i).view(np.int16) if bits != 8 else T.synth_frame(w, h, bits, 1234 + i)).cuda() for i in range(NB)]
    outs = [torch.empty((oH, oW), dtype=tdt, device="cuda") for _ in range(NB)]
    eng = B.Engine(T.filter_folder(folder), ratio, bits, 1, passes, mode, device=0, numerics=int(os.environ.get("RAISR_KB_NUMERICS", B.NUMERICS_AUTO)))
    eng.set_res(w, h, oW, oH)
    bps = 1 if bits == 8 else 2
    def run(i):
        rc = eng.process_device_rows(ys[i % NB].data_ptr(), ys[i % NB].stride(0) * bps, outs[i % NB].data_ptr(), outs[i % NB].stride(0) * bps, 0, oH, 2, None)
        assert rc == 0
    for i in range(NB):
        run(i)
    torch.cuda.synchronize()
    reps = 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for i in range(reps):
            run(i)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    by = (w * h + oW * oH) * bps
    print("%-75s %8.4f ms  %8.1f frames/s  %6.1f GB/s algorithmic" % (name, best, 1e3 / best, by / best / 1e6))
    eng.close()
