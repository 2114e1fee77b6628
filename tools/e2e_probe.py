"""End-to-end probe on the GPU box: frames/s of the blocking host-pointer call (pinned planes) under engine switches.
usage: python tools/e2e_probe.py"""
import os, sys, time, importlib.util
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import raisr_testlib as T
spec = importlib.util.spec_from_file_location("b", os.path.join(T.PKG_DIR, "binding.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
W, H = 1920, 1080
NB = 12
hin = [[torch.from_numpy(T.synth_frame(W, H, 8, 1234 + i)).pin_memory(), torch.from_numpy(T.synth_chroma(W // 2, H // 2, 8, i + 1)).pin_memory(),
        torch.from_numpy(T.synth_chroma(W // 2, H // 2, 8, i + 2)).pin_memory()] for i in range(NB)]
hout = [[torch.empty((2 * H, 2 * W), dtype=torch.uint8).pin_memory(), torch.empty((H, W), dtype=torch.uint8).pin_memory(),
         torch.empty((H, W), dtype=torch.uint8).pin_memory()] for _ in range(NB)]
ref = None
for name, env in (("default", {}), ("split_h2d=0", {"RAISR_CUDA_SPLIT_H2D": "0"}), ("tail_in_place=0", {"RAISR_CUDA_TAIL_IN_PLACE": "0"}),
                  ("luma only default", {"_LUMA": "1"}), ("luma only split=0", {"_LUMA": "1", "RAISR_CUDA_SPLIT_H2D": "0"})):
    for k in ("RAISR_CUDA_TAIL_IN_PLACE", "RAISR_CUDA_SPLIT_H2D"):
        os.environ.pop(k, None)
    for k, v in env.items():
        if not k.startswith("_"):
            os.environ[k] = v
    eng = B.Engine(T.filter_folder("filters_2x/filters_lowres"), 2.0, 8, 1, 1, 1, device=0, numerics=B.NUMERICS_AUTO)
    eng.set_res(W, H, 2 * W, 2 * H, W // 2, H // 2, W, H)
    luma_only = "_LUMA" in env
    def frame(i):
        a, o = hin[i % NB], hout[i % NB]
        z = 0
        rc = eng.L.raisr_cuda_process_host(eng.h, a[0].data_ptr(), a[0].stride(0), z if luma_only else a[1].data_ptr(), a[1].stride(0),
                                           z if luma_only else a[2].data_ptr(), a[2].stride(0), o[0].data_ptr(), o[0].stride(0),
                                           z if luma_only else o[1].data_ptr(), o[1].stride(0), z if luma_only else o[2].data_ptr(), o[2].stride(0), 2)
        assert rc == 0
    for i in range(24):
        frame(i)
    if ref is None:
        ref = hout[0][0].clone()
    else:
        assert torch.equal(ref, hout[0][0]), name
    if not luma_only:
        if "refc" not in globals():
            refc = [hout[0][1].clone(), hout[0][2].clone()]
        else:
            assert torch.equal(refc[0], hout[0][1]) and torch.equal(refc[1], hout[0][2]), name + " chroma"
    n = 300
    t0 = time.perf_counter()
    for i in range(n):
        frame(i)
    dt = time.perf_counter() - t0
    print("%-20s %8.1f frames/s  %.3f ms/frame" % (name, n / dt, 1e3 * dt / n))
    eng.close()
