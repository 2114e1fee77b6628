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


@pytest.mark.parametrize("env", [{"RAISR_CUDA_SPLIT_H2D": "0"}, {"RAISR_CUDA_TAIL_IN_PLACE": "0"}, {"RAISR_CUDA_NO_BAND_PIPELINE": "1"},
                                 {"RAISR_CUDA_KERNEL": "tile"}, {"RAISR_CUDA_NO_MEMOPS": "1"}],
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


# ---- hazards of the host engine (round-1 review): counters that start at zero, state that is per engine, races of the split H2D ----
def _pin(a, keep):
    import torch
    t = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).clone().pin_memory()
    keep.append(t)
    return t.numpy().view(a.dtype)


def test_many_create_destroy_cycles_with_the_band_pipeline():
    """200 engines in one process on page-locked planes: every engine's band / chroma / input counters must start from zero
    (recycled device allocations are not zeroed by cudaMalloc), or a band is copied out before it is computed."""
    cfg = CONFIGS[6]
    folder, ratio, bits, passes, mode, (w, h) = cfg
    src = planes(w, h, bits, seed=21)
    want = run_host(cfg, src, pinned=False)
    keep = []
    ins = [_pin(a, keep) for a in src]
    oW, oH = int(w * ratio), int(h * ratio)
    outs = [_pin(np.zeros((oH, oW), np.uint8), keep), _pin(np.zeros((oH // 2, oW // 2), np.uint8), keep),
            _pin(np.zeros((oH // 2, oW // 2), np.uint8), keep)]
    for cycle in range(200):
        eng = B.Engine(T.filter_folder(folder), ratio, bits, T.VideoRange, passes, mode, numerics=B.NUMERICS_AUTO)
        eng.set_res(w, h, oW, oH, w // 2, h // 2, oW // 2, oH // 2)
        for o in outs:
            o[...] = 0
        assert eng.process_host(ins[0], outs[0], ins[1], ins[2], outs[1], outs[2]) == 0
        eng.close()
        for a, b, n in zip(want, outs, "YUV"):
            assert np.array_equal(a, b), "cycle %d: %s plane differs on %d samples" % (cycle, n, (a != b).sum())


def test_two_live_engines_with_different_bit_depths_interleaved():
    """An 8-bit and a 10-bit engine alive on one device, frames interleaved: the Gaussian weights (bit-depth dependent
    normalisation, Raisr_globals.h:204-206) are per launch, not process-global device state."""
    f8, f10 = T.filter_folder("filters_2x/filters_lowres"), T.filter_folder("filters_2x/filters_highres")
    w, h = 322, 182
    i8, i10 = T.synth_frame(w, h, 8, seed=31), T.synth_frame(w, h, 10, seed=32, kind="edges")
    ref8 = T.oracle_process_y(i8, 2 * w, 2 * h, T.OracleModel(f8, 8))
    ref10 = T.oracle_process_y(i10, 2 * w, 2 * h, T.OracleModel(f10, 10))
    e8 = B.Engine(f8, 2.0, 8, T.VideoRange, 1, 1, numerics=B.NUMERICS_IEEE)
    e10 = B.Engine(f10, 2.0, 10, T.VideoRange, 1, 1, numerics=B.NUMERICS_IEEE)     # created second: would overwrite shared weights
    e8.set_res(w, h, 2 * w, 2 * h)
    e10.set_res(w, h, 2 * w, 2 * h)
    for _ in range(3):
        o8, o10 = np.zeros((2 * h, 2 * w), np.uint8), np.zeros((2 * h, 2 * w), np.uint16)
        assert e8.process_host(i8, o8) == 0
        assert e10.process_host(i10, o10) == 0
        assert np.array_equal(o8, ref8), "8-bit engine: %d px differ" % (o8 != ref8).sum()
        assert np.array_equal(o10, ref10), "10-bit engine: %d px differ" % (o10 != ref10).sum()
    e8.close()
    e10.close()


@pytest.mark.parametrize("size", [(96, 256), (128, 300), (640, 512), (1000, 600)], ids=lambda s: "%dx%d" % s)
def test_split_h2d_on_small_tall_frames_with_pinned_planes(size):
    """in_h >= 256 with fewer tiles than SMs: every tile is a CTA's first, so the filter warps read input rows behind the split
    at kernel start -- they must be ordered behind the watermark the H2D stream writes (page-locked planes make the copy truly
    async).  1000x600: several rounds of tiles, the chain warps wait once per tile row until the watermark has passed the plane."""
    w, h = size
    f = T.filter_folder("filters_2x/filters_lowres")
    img = T.synth_frame(w, h, 8, seed=w + h, kind="noise")
    ref = T.oracle_process_y(img, 2 * w, 2 * h, T.OracleModel(f, 8))
    keep = []
    pin_in, pin_out = _pin(img, keep), _pin(np.zeros((2 * h, 2 * w), np.uint8), keep)
    eng = B.Engine(f, 2.0, 8, T.VideoRange, 1, 1, numerics=B.NUMERICS_IEEE)
    eng.set_res(w, h, 2 * w, 2 * h)
    for k in range(20):
        pin_in[...] = 0 if k % 2 else img                     # alternate with a blank frame: stale rows would show
        pin_out[...] = 0
        assert eng.process_host(pin_in, pin_out) == 0
        if k % 2 == 0:
            assert np.array_equal(pin_out, ref), "frame %d: %d px differ" % (k, (pin_out != ref).sum())
    eng.close()


def test_setres_rejects_cb_geometry_that_differs_from_cr():
    import ctypes as C
    L = T.handler_lib(T.product_lib_path())
    f = T.filter_folder("filters_2x/filters_lowres")
    y, u, v = planes(64, 48, 8, seed=3)
    oy, ou, ov = np.zeros((96, 128), np.uint8), np.zeros((48, 64), np.uint8), np.zeros((48, 64), np.uint8)
    assert L.RNLHandler_Init(f.encode(), 2.0, 8, T.VideoRange, 1, T.AVX512, 1, 1) == 0
    try:
        good = [T.vdt(a) for a in (y, u, v, oy, ou, ov)]
        assert L.RNLHandler_SetRes(*[C.byref(x) for x in good]) == 0
        bad = [T.vdt(a) for a in (y, u, v[:, :16], oy, ou, ov)]
        assert L.RNLHandler_SetRes(*[C.byref(x) for x in bad]) == T.RNLErrorBadParameter
        bad = [T.vdt(a) for a in (y, u, v, oy, ou, ov[:24])]
        assert L.RNLHandler_SetRes(*[C.byref(x) for x in bad]) == T.RNLErrorBadParameter
    finally:
        L.RNLHandler_Deinit()


@pytest.mark.parametrize("env", [{}, {"RAISR_CUDA_STAGE_PAGEABLE": "0"}, {"RAISR_CUDA_COPY_THREADS": "0"}, {"RAISR_CUDA_COPY_THREADS": "1"},
                                 {"RAISR_CUDA_NO_MEMOPS": "1"}, {"RAISR_CUDA_NO_BAND_PIPELINE": "1"}, {"RAISR_CUDA_NT_COPY": "0"}],
                         ids=lambda e: ",".join("%s=%s" % kv for kv in e.items()) or "default")
def test_pageable_planes_with_padded_steps_through_the_staging_pipeline(env, monkeypatch):
    """Pageable caller planes (what av_frame_get_buffer hands a software filter) travel through the engine's page-locked staging
    planes, moved by the copy threads band by band.  Whatever the thread count or fallback, the frame is the page-locked result,
    and the bytes between width and step stay untouched."""
    cfg = CONFIGS[6]
    folder, ratio, bits, passes, mode, (w, h) = cfg
    src = planes(w, h, bits, seed=41)
    want = run_host(cfg, src, pinned=True)
    oW, oH = int(w * ratio), int(h * ratio)
    pad = 48
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    eng = B.Engine(T.filter_folder(folder), ratio, bits, T.VideoRange, passes, mode, numerics=B.NUMERICS_AUTO)
    eng.set_res(w, h, oW, oH, w // 2, h // 2, oW // 2, oH // 2)
    inb = [np.full((a.shape[0], a.shape[1] + pad), 0x5A, a.dtype) for a in src]
    for b, a in zip(inb, src):
        b[:, :a.shape[1]] = a
    ins = [b[:, :a.shape[1]] for b, a in zip(inb, src)]
    outb = [np.full((oH, oW + pad), 0xA5, np.uint8), np.full((oH // 2, oW // 2 + pad), 0xA5, np.uint8), np.full((oH // 2, oW // 2 + pad), 0xA5, np.uint8)]
    outs = [outb[0][:, :oW], outb[1][:, :oW // 2], outb[2][:, :oW // 2]]
    for _ in range(3):
        for o in outs:
            o[...] = 0
        assert eng.process_host(ins[0], outs[0], ins[1], ins[2], outs[1], outs[2]) == 0
        for a, b, n in zip(want, outs, "YUV"):
            assert np.array_equal(a, b), "%s plane differs on %d samples with %r" % (n, (a != b).sum(), env)
    for b, o in zip(outb, outs):
        assert (b[:, o.shape[1]:] == 0xA5).all(), "bytes beyond the row width were written"
    eng.close()


def test_lost_copy_flag_times_out_and_the_engine_recovers(monkeypatch, capfd):
    """The kernel waits in-kernel for the flag behind the late input copy.  If that flag never comes (a failed copy; a profiler that
    serialises the GPU so that the copy cannot run under the kernel) the wait gives up after ~2 s, the frame is re-run in plain
    stream order and the engine stays there: the caller gets the right frame and a warning, never a hung device."""
    import time
    cfg = CONFIGS[0]
    folder, ratio, bits, passes, mode, (w, h) = cfg
    src = planes(w, h, bits, seed=88)
    want = run_host(cfg, src, pinned=True)
    monkeypatch.setenv("RAISR_CUDA_TEST_DROP_IN_FLAG", "1")
    eng = B.Engine(T.filter_folder(folder), ratio, bits, T.VideoRange, passes, mode, numerics=B.NUMERICS_AUTO)
    eng.set_res(w, h, 2 * w, 2 * h, w // 2, h // 2, w, h)
    keep = []
    ins = [_pin(a, keep) for a in src]
    outs = [_pin(np.zeros((2 * h, 2 * w), np.uint8), keep), _pin(np.zeros((h, w), np.uint8), keep), _pin(np.zeros((h, w), np.uint8), keep)]
    t0 = time.perf_counter()
    assert eng.process_host(ins[0], outs[0], ins[1], ins[2], outs[1], outs[2]) == 0
    first = time.perf_counter() - t0
    for a, b, n in zip(want, outs, "YUV"):
        assert np.array_equal(a, b), "%s plane differs after the recovery" % n
    assert 1.5 < first < 10.0, "the first frame should have run into the ~2 s bound (took %.2f s)" % first
    t0 = time.perf_counter()
    for _ in range(5):
        assert eng.process_host(ins[0], outs[0], ins[1], ins[2], outs[1], outs[2]) == 0
    assert (time.perf_counter() - t0) / 5 < 0.05
    for a, b, n in zip(want, outs, "YUV"):
        assert np.array_equal(a, b)
    eng.close()
    assert "plain stream order" in capfd.readouterr().out
