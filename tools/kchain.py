"""Times the two-pass configurations (device-resident luma passes) with a given library build; prints ms and an output digest.
usage: python tools/kchain.py <lib.so>   (RAISR_CUDA_CHAIN=0 for one launch per pass)"""
import os, sys, hashlib, importlib.util
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import raisr_testlib as T
spec = importlib.util.spec_from_file_location("b", os.path.join(T.PKG_DIR, "binding.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
if len(sys.argv) > 1:
    B.LIB_PATH = os.path.abspath(sys.argv[1])
CONFIGS = [
    ("configs[2] 1080p->4K highres p2m1 8b", "filters_2x/filters_highres", 2.0, 8, 2, 1, 1920, 1080),
    ("1080p->4K denoise p2m2 8b", "filters_2x/filters_denoise", 2.0, 8, 2, 2, 1920, 1080),
    ("configs[3] 4K->8K denoise p2m2 10b", "filters_2x/filters_denoise", 2.0, 10, 2, 2, 3840, 2160),
]
NB = 6
fd = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1)
for name, folder, ratio, bits, passes, mode, w, h in CONFIGS:
    oW, oH = int(w * ratio), int(h * ratio)
    tdt = torch.uint8 if bits == 8 else torch.int16
    fr = [T.synth_frame(w, h, bits, 1234 + i) for i in range(2)]
    ys = [torch.from_numpy(fr[i % 2].view(np.int16) if bits != 8 else fr[i % 2]).cuda() for i in range(NB)]
    outs = [torch.empty((oH, oW), dtype=tdt, device="cuda") for _ in range(NB)]
    os.dup2(fd, 1)
    eng = B.Engine(T.filter_folder(folder), ratio, bits, 1, passes, mode, device=0, numerics=int(os.environ.get("RAISR_KB_NUMERICS", B.NUMERICS_AUTO)))
    os.dup2(saved, 1)
    eng.set_res(w, h, oW, oH)
    bps = 1 if bits == 8 else 2
    def run(i):
        assert eng.process_device_rows(ys[i % NB].data_ptr(), ys[i % NB].stride(0) * bps, outs[i % NB].data_ptr(), outs[i % NB].stride(0) * bps, 0, oH, 2, None) == 0
    for i in range(NB): run(i)
    torch.cuda.synchronize()
    dig = hashlib.sha1(outs[0].cpu().numpy().tobytes() + outs[1].cpu().numpy().tobytes()).hexdigest()[:12]
    n0 = eng.launch_count()
    reps = 60
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record()
        for i in range(reps): run(i)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    print("%-40s %8.4f ms  launches/frame %d  sha1 %s" % (name, min(ts), (eng.launch_count() - n0) // (3 * reps), dig), flush=True)
    eng.close()
