"""GPU tests of the host-side frame pipeline around the pass kernel (csrc/raisr_engine.cu: raisr_cuda_process_host /
raisr_cuda_process_device): whatever the memory kind of the caller's planes and whichever copy strategy is selected, a frame
comes out bit for bit the same -- pageable vs page-locked planes (split H2D, band-signalled D2H, in-place tail rows), the
measurement switches, the device entry point with chroma planes, the phase-sequential kernel.  The reference's contract is
simply "Process fills the six planes" (Raisr.cpp:1294-1390); these tests pin that the overlap machinery never changes a byte."""
import importlib.util
import os

import numpy as np
import pytest

import raisr_testlib as T

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("raisr_binding", os.path.join(T.PKG_DIR, "binding.py"))
B = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(B)

CONFIGS = [
    # folder, ratio, bits, passes, mode, (w, h)
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1, (1920, 1080)),      # BASELINE configs[1]
    ("filters_2x/filters_highres", 2.0, 8, 2, 1, (640, 362)),       # two passes, chroma rides with pass 1
    ("filters_2x/filters_denoise", 2.0, 10, 2, 2, (480, 270)),      # 16-bit samples, pass 1 at input resolution
    ("filters_1.5x/filters_highres", 1.5, 8, 1, 1, (640, 360)),     # generic-ratio chroma path inside the kernel
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1, (150, 66)),         # fewer tiles than SMs: chroma slices after the tile loop
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1, (333, 190)),        # odd sizes: no vector stores, every tile row written in place when pinned, chroma ratio != 2
    ("filters_2x/filters_highres", 2.0, 8, 2, 1, (1000, 600)),      # several rounds of tiles: some rows band-copied, the last ones in place
]


def planes(w, h, bits, seed):
    dt = np.uint8 if bits == 8 else np.uint16
    y = T.synth_frame(w, h, bits, seed)
    u = T.synth_chroma(w // 2, h // 2, bits, seed + 1).astype(dt)
    v = T.synth_chroma(w // 2, h // 2, bits, seed + 2).astype(dt)
    return y, u, v


def run_host(cfg, src, pinned, env=None, monkeypatch=None):
    import torch
    folder, ratio, bits, passes, mode, (w, h) = cfg
    oW, oH = int(w * ratio), int(h * ratio)
    if env:
        for k, v in env.items():
            monkeypatch.setenv(k, v)
    eng = B.Engine(T.filter_folder(folder), ratio, bits, T.VideoRange, passes, mode, numerics=B.NUMERICS_AUTO)
    eng.set_res(w, h, oW, oH, w // 2, h // 2, oW // 2, oH // 2)       # (odd sizes: chroma ratio is not exactly the luma ratio)
    dt = src[0].dtype
    outs = [np.zeros((oH, oW), dt), np.zeros((oH // 2, oW // 2), dt), np.zeros((oH // 2, oW // 2), dt)]
    ins = list(src)
    keep = []
    if pinned:
        tdt = torch.uint8 if dt == np.uint8 else torch.int16
        def pin(a):
            t = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).clone().pin_memory()
            keep.append(t)
            return t.numpy().view(a.dtype)
        ins = [pin(a) for a in ins]
        outs = [pin(a) for a in outs]
        assert tdt is not None
    for _ in range(2):                                               # twice: running counters / sequence numbers carry over
        for o in outs:
            o[...] = 0
        assert eng.process_host(ins[0], outs[0], ins[1], ins[2], outs[1], outs[2]) == 0
    res = [o.copy() for o in outs]
    if env:
        for k in env:
            monkeypatch.delenv(k)
    eng.close()
    return res


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "%s-%dx%d-p%d" % (c[0].split("/")[-1], c[5][0], c[5][1], c[3]))
def test_pinned_and_pageable_planes_give_the_same_frame(cfg, monkeypatch):
    src = planes(cfg[5][0], cfg[5][1], cfg[2], seed=77)
    base = run_host(cfg, src, pinned=False)
    pin = run_host(cfg, src, pinned=True)
    for a, b, n in zip(base, pin, "YUV"):
        assert np.array_equal(a, b), "%s plane differs on %d samples" % (n, (a != b).sum())
    # chroma is the plain cheap upscale (Raisr.cpp:1373-1388)
    oH, oW = base[1].shape
    assert np.array_equal(base[1], T.oracle_resize(src[1], oW, oH)) and np.array_equal(base[2], T.oracle_resize(src[2], oW, oH))


@pytest.mark.parametrize("env", [{"RAISR_CUDA_SPLIT_H2D": "0"}, {"RAISR_CUDA_ZERO_COPY": "0"}, {"RAISR_CUDA_ZERO_COPY": "2"},
                                 {"RAISR_CUDA_ZERO_COPY": "7"}, {"RAISR_CUDA_NO_BAND_PIPELINE": "1"}, {"RAISR_CUDA_KERNEL": "tile"},
                                 {"RAISR_CUDA_NO_MEMOPS": "1"}],
                         ids=lambda e: ",".join("%s=%s" % kv for kv in e.items()))
def test_copy_strategies_are_invisible(env, monkeypatch):
    cfg = CONFIGS[1]
    src = planes(cfg[5][0], cfg[5][1], cfg[2], seed=5)
    base = run_host(cfg, src, pinned=True)
    alt = run_host(cfg, src, pinned=True, env=env, monkeypatch=monkeypatch)
    for a, b, n in zip(base, alt, "YUV"):
        assert np.array_equal(a, b), "%s plane differs with %r" % (n, env)


def test_device_entry_with_chroma_equals_host_entry():
    import torch
    cfg = CONFIGS[0]
    folder, ratio, bits, passes, mode, (w, h) = cfg
    src = planes(w, h, bits, seed=9)
    host = run_host(cfg, src, pinned=False)
    eng = B.Engine(T.filter_folder(folder), ratio, bits, T.VideoRange, passes, mode, numerics=B.NUMERICS_AUTO)
    eng.set_res(w, h, 2 * w, 2 * h, w // 2, h // 2, w, h)
    d_in = [torch.from_numpy(a).cuda() for a in src]
    d_out = [torch.zeros((2 * h, 2 * w), dtype=torch.uint8, device="cuda"), torch.zeros((h, w), dtype=torch.uint8, device="cuda"),
             torch.zeros((h, w), dtype=torch.uint8, device="cuda")]
    n0 = eng.launch_count()
    rc = eng.process_device(d_in[0].data_ptr(), d_in[0].stride(0), d_out[0].data_ptr(), d_out[0].stride(0),
                            d_in[1].data_ptr(), d_in[1].stride(0), d_in[2].data_ptr(), d_in[2].stride(0),
                            d_out[1].data_ptr(), d_out[1].stride(0), d_out[2].data_ptr(), d_out[2].stride(0), 2, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert eng.launch_count() - n0 == 1, "a yuv420p frame is one launch of the pipelined kernel"
    for a, b, n in zip(host, d_out, "YUV"):
        assert np.array_equal(a, b.cpu().numpy()), n
    eng.close()
