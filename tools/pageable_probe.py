"""Pageable-plane probe on the GPU box: frames/s of the blocking host-pointer call with malloc'ed planes (what FFmpeg's software
frames are), checked against the page-locked result.  The copy switches are per process (environment):
usage: for t in 2 4 8; do for nt in 0 1; do RAISR_CUDA_COPY_THREADS=$t RAISR_CUDA_NT_COPY=$nt python tools/pageable_probe.py; done; done"""
import os, sys, time, importlib.util
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import raisr_testlib as T
spec = importlib.util.spec_from_file_location("b", os.path.join(T.PKG_DIR, "binding.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
W, H = 1920, 1080
NB = 12
src = [[T.synth_frame(W, H, 8, 1234 + i), T.synth_chroma(W // 2, H // 2, 8, i + 1), T.synth_chroma(W // 2, H // 2, 8, i + 2)] for i in range(NB)]
pin_in = [[torch.from_numpy(a).pin_memory() for a in f] for f in src]
pin_out = [[torch.empty((2 * H, 2 * W), dtype=torch.uint8).pin_memory(), torch.empty((H, W), dtype=torch.uint8).pin_memory(),
            torch.empty((H, W), dtype=torch.uint8).pin_memory()] for _ in range(NB)]
pg_in = [[np.array(a, copy=True) for a in f] for f in src]
pg_out = [[np.zeros((2 * H, 2 * W), np.uint8), np.zeros((H, W), np.uint8), np.zeros((H, W), np.uint8)] for _ in range(NB)]
ref = None
res = []
for pinned in (True, False):
    eng = B.Engine(T.filter_folder("filters_2x/filters_lowres"), 2.0, 8, 1, 1, 1, device=0, numerics=B.NUMERICS_AUTO)
    eng.set_res(W, H, 2 * W, 2 * H, W // 2, H // 2, W, H)
    def ptr(x): return x.data_ptr() if pinned else x.ctypes.data
    def step(x): return x.stride(0) if pinned else x.strides[0]
    ins, outs = (pin_in, pin_out) if pinned else (pg_in, pg_out)
    def frame(i):
        a, o = ins[i % NB], outs[i % NB]
        rc = eng.L.raisr_cuda_process_host(eng.h, ptr(a[0]), step(a[0]), ptr(a[1]), step(a[1]), ptr(a[2]), step(a[2]),
                                           ptr(o[0]), step(o[0]), ptr(o[1]), step(o[1]), ptr(o[2]), step(o[2]), 2)
        assert rc == 0
    for i in range(24):
        frame(i)
    if pinned:
        ref = [[t.numpy().copy() for t in f] for f in pin_out]
    else:
        for f in range(NB):
            for k in range(3):
                assert np.array_equal(ref[f][k], pg_out[f][k]), (f, k)
    n = 300
    t0 = time.perf_counter()
    for i in range(n):
        frame(i)
    dt = time.perf_counter() - t0
    res.append("%s %7.1f frames/s (%.3f ms)" % ("page-locked" if pinned else "pageable", n / dt, 1e3 * dt / n))
    eng.close()
print("threads=%s nt=%s : %s" % (os.environ.get("RAISR_CUDA_COPY_THREADS", "default"), os.environ.get("RAISR_CUDA_NT_COPY", "default"), " | ".join(res)))
