"""GPU edge cases: tiny and ragged frames, 16-bit depth, other chroma samplings, odd 1.5x sizes, the remaining
BASELINE configurations at full size (properties + oracle slabs)."""
import importlib.util
import os
import shutil

import numpy as np
import pytest

import raisr_testlib as T

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("raisr_binding", os.path.join(T.PKG_DIR, "binding.py"))
B = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(B)

X86 = T.have_avx512()
NUM = B.NUMERICS_X86 if X86 else B.NUMERICS_IEEE
SM = 1 if X86 else 0


def engine_run(folder, img, ratio, bits, passes=1, mode=1, rng=T.VideoRange, out_hw=None):
    h, w = img.shape
    oW, oH = out_hw if out_hw else (int(w * ratio), int(h * ratio))
    eng = B.Engine(folder, ratio, bits, rng, passes, mode, numerics=NUM)
    eng.set_res(w, h, oW, oH)
    out = np.zeros((oH, oW), img.dtype)
    assert eng.process_host(img, out) == 0
    eng.close()
    return out


def oracle_run(folder, img, ratio, bits, passes=1, mode=1, rng=T.VideoRange, out_hw=None):
    h, w = img.shape
    oW, oH = out_hw if out_hw else (int(w * ratio), int(h * ratio))
    m1 = T.OracleModel(folder, bits, False, rng, SM)
    m2 = T.OracleModel(folder, bits, True, rng, SM) if passes == 2 else None
    return T.oracle_process_y(img, oW, oH, m1, m2, passes, mode)


@pytest.mark.parametrize("size", [(4, 4), (10, 10), (13, 7), (14, 20), (15, 15), (31, 17), (57, 9), (101, 57), (117, 35), (119, 64)])
def test_tiny_and_ragged_2x(size):
    """Frames narrower than one 16-block hash nothing (c_end == 6); tiles with 1..4 valid columns; odd widths (no vector stores)."""
    w, h = size
    f = T.filter_folder("filters_2x/filters_lowres")
    img = T.synth_frame(w, h, 8, seed=w * 131 + h, kind="noise")
    assert np.array_equal(engine_run(f, img, 2.0, 8), oracle_run(f, img, 2.0, 8))


@pytest.mark.parametrize("size", [(22, 14), (100, 66), (202, 118)])
def test_ratio_15_sizes(size):
    w, h = size
    f = T.filter_folder("filters_1.5x/filters_highres")
    img = T.synth_frame(w, h, 8, seed=w + h, kind="mix")
    assert np.array_equal(engine_run(f, img, 1.5, 8), oracle_run(f, img, 1.5, 8))


@pytest.mark.skipif(not (T.have_ref() and X86), reason="needs the compiled reference on an AVX-512 host")
def test_odd_15x_geometry_vs_live_reference():
    """101x67 -> 151x100: the reference's resize spec maps (int)(outH/ratio) = 66 source rows, not 67 (Raisr.cpp:1801-1803)."""
    img = T.synth_frame(101, 67, 8, seed=9, kind="mix")
    ref_y, _ = T.run_ref_subprocess("filters_1.5x/filters_highres", img, 1.5, 8)
    out = engine_run(T.filter_folder("filters_1.5x/filters_highres"), img, 1.5, 8)
    assert out.shape == ref_y.shape and np.array_equal(out, ref_y), "Y differs on %d px" % (out != ref_y).sum()


def test_16bit_depth(tmp_path):
    """No 16-bit model is shipped; a folder with the 10-bit tables renamed exercises the 16-bit code path
    (NF_16 Gaussian, 0..65535 range, u16 planes, Raisr.cpp:1462-1468)."""
    src = T.filter_folder("filters_2x/filters_highres")
    dst = tmp_path / "model16"
    os.makedirs(dst)
    shutil.copy(os.path.join(src, "config"), dst / "config")
    for stem in ("filterbin_2_", "Qfactor_strbin_2_", "Qfactor_cohbin_2_"):
        shutil.copy(os.path.join(src, stem + "10"), dst / (stem + "16"))
    rs = np.random.RandomState(5)
    base = T.synth_frame(120, 80, 10, seed=3).astype(np.uint32) * 64
    img = np.clip(base + rs.randint(0, 64, size=base.shape), 0, 65535).astype(np.uint16)
    assert np.array_equal(engine_run(str(dst), img, 2.0, 16), oracle_run(str(dst), img, 2.0, 16))


@pytest.mark.parametrize("shift", [(1, 1), (1, 0), (0, 0)])     # 4:2:0, 4:2:2, 4:4:4 (vf_raisr.c:158-162,191-204)
def test_chroma_samplings_through_handler(shift):
    f = T.filter_folder("filters_2x/filters_lowres")
    w, h = 128, 72
    cw, ch = w >> shift[0], h >> shift[1]
    y = T.synth_frame(w, h, 8, seed=1)
    u, v = T.synth_chroma(cw, ch, 8, 2), T.synth_chroma(cw, ch, 8, 3)
    L = T.handler_lib(T.product_lib_path())
    oy, ou, ov = T.run_handler(L, f, y, inU=u, inV=v, chroma_shift=shift)
    assert ou.shape == (2 * ch, 2 * cw)
    assert np.array_equal(ou, T.oracle_resize(u, 2 * cw, 2 * ch)) and np.array_equal(ov, T.oracle_resize(v, 2 * cw, 2 * ch))


def slab_check(folder, img, out, ratio, bits, passes, mode, lr0, lr_rows, margin_lr):
    """Runs the oracle on input rows [lr0, lr0+lr_rows) and compares the output rows that cannot see the slab edges."""
    ref = oracle_run(folder, img[lr0:lr0 + lr_rows], ratio, bits, passes, mode)
    a = int(margin_lr * ratio)
    o0 = int(lr0 * ratio)
    got = out[o0 + a:o0 + ref.shape[0] - a]
    assert np.array_equal(got, ref[a:-a]), "slab differs on %d px" % (got != ref[a:-a]).sum()


def test_config3_1080p_to_4k_highres_two_pass():
    """BASELINE configs[2] at full size: invariants + a slab against the oracle (two passes: 14 output rows of reach per pass)."""
    f = T.filter_folder("filters_2x/filters_highres")
    img = T.synth_frame(1920, 1080, 8, seed=21)
    out = engine_run(f, img, 2.0, 8, passes=2, mode=1)
    assert out[1:-1, 1:-1].min() >= 16 and out[1:-1, 1:-1].max() <= 235
    slab_check(f, img, out, 2.0, 8, 2, 1, lr0=500, lr_rows=96, margin_lr=16)


def test_config4_4k_to_8k_denoise_10bit_two_pass_mode2():
    """BASELINE configs[3] at full size (single GPU): 3840x2160 -> 7680x4320, 10-bit, passes=2 mode=2."""
    f = T.filter_folder("filters_2x/filters_denoise")
    img = T.synth_frame(3840, 2160, 10, seed=22)
    out = engine_run(f, img, 2.0, 10, passes=2, mode=2)
    assert out.shape == (4320, 7680)
    assert out[1:-1, 1:-1].min() >= 64 and out[1:-1, 1:-1].max() <= 940
    slab_check(f, img, out, 2.0, 10, 2, 2, lr0=1000, lr_rows=64, margin_lr=20)


def test_config5_720p_to_1080p_15x():
    """BASELINE configs[4] geometry with a folder that ships the needed tables (filters_1.5x/filters_denoise, 2 passes)."""
    f = T.filter_folder("filters_1.5x/filters_denoise")
    img = T.synth_frame(1280, 720, 8, seed=23)
    out = engine_run(f, img, 1.5, 8, passes=2, mode=2)
    assert out.shape == (1080, 1920)
    slab_check(f, img, out, 1.5, 8, 2, 2, lr0=300, lr_rows=96, margin_lr=24)
    # the folder BASELINE names for this config has no second-pass tables: same failure as the reference
    with pytest.raises(RuntimeError):
        B.Engine(T.filter_folder("filters_1.5x/filters_highres"), 1.5, 8, T.VideoRange, 2, 1)


@pytest.mark.parametrize("folder,ratio,bits,passes,mode", [
    ("filters_2x/filters_highres", 2.0, 8, 2, 1),
    ("filters_2x/filters_denoise", 2.0, 10, 2, 2),
    ("filters_1.5x/filters_denoise", 1.5, 8, 2, 2),
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1),
])
def test_row_bands_reproduce_full_frame(folder, ratio, bits, passes, mode):
    """Row-band launches (the multi-GPU shard unit) == one full-frame launch, bit for bit, including two-pass
    configurations where pass 1 is recomputed on the rows pass 2 can reach."""
    import torch
    f = T.filter_folder(folder)
    w, h = 400, 260
    img = T.synth_frame(w, h, bits, seed=77, kind="mix")
    oW, oH = int(w * ratio), int(h * ratio)
    full = engine_run(f, img, ratio, bits, passes, mode)
    eng = B.Engine(f, ratio, bits, T.VideoRange, passes, mode, numerics=NUM)
    eng.set_res(w, h, oW, oH)
    tdt = torch.uint8 if bits == 8 else torch.int16
    d_in = torch.from_numpy(img.view(np.int16) if bits != 8 else img).cuda()
    d_out = torch.zeros((oH, oW), dtype=tdt, device="cuda")
    for r0, r1 in [(0, 100), (100, 102), (102, 258), (258, oH)]:
        assert eng.process_device_rows(d_in.data_ptr(), d_in.stride(0) * d_in.element_size(), d_out.data_ptr(),
                                       d_out.stride(0) * d_out.element_size(), r0, r1) == 0
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(img.dtype)
    eng.close()
    assert np.array_equal(got, full), "bands differ from the full frame on %d px" % (got != full).sum()


@pytest.mark.parametrize("folder,ratio,bits,passes,mode,size,shift", [
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1, (640, 360), 0),            # NV12
    ("filters_2x/filters_denoise", 2.0, 10, 2, 2, (640, 360), 6),          # P010: value in the high 10 bits, two passes chained
    ("filters_1.5x/filters_highres", 1.5, 8, 1, 1, (426, 240), 0),         # NV12 at 1.5x: generic chroma path
    ("filters_2x/filters_highres", 2.0, 10, 2, 1, (322, 182), 6),          # P010, odd chroma sizes: no vector stores
], ids=["nv12", "p010-p2m2", "nv12-1.5x", "p010-odd"])
def test_semiplanar_device_frames_equal_planar_frames(folder, ratio, bits, passes, mode, size, shift):
    """NV12 / P010 device frames (what NVDEC and NVENC use; two of the three formats of vf_raisr_opencl.c:166-169) through
    raisr_cuda_process_device_semiplanar == the planar result, sample for sample: interleaved chroma in, interleaved chroma out,
    and for P010 every sample read as word >> 6 and written as value << 6."""
    import torch
    w, h = size
    oW, oH = int(w * ratio), int(h * ratio)
    cw, ch, ocw, och = (w + 1) // 2, (h + 1) // 2, (oW + 1) // 2, (oH + 1) // 2
    dt = np.uint8 if bits == 8 else np.uint16
    tdt = torch.uint8 if bits == 8 else torch.int16
    y = T.synth_frame(w, h, bits, seed=321)
    u, v = T.synth_chroma(cw, ch, bits, 5).astype(dt), T.synth_chroma(cw, ch, bits, 6).astype(dt)
    f = T.filter_folder(folder)
    # planar reference result through the host entry
    eng = B.Engine(f, ratio, bits, T.VideoRange, passes, mode, numerics=NUM)
    eng.set_res(w, h, oW, oH, cw, ch, ocw, och)
    py, pu, pv = np.zeros((oH, oW), dt), np.zeros((och, ocw), dt), np.zeros((och, ocw), dt)
    assert eng.process_host(y, py, u, v, pu, pv) == 0
    # semi-planar device frames
    uv = np.empty((ch, 2 * cw), dt)
    uv[:, 0::2], uv[:, 1::2] = u, v
    as_t = lambda a: torch.from_numpy((a.astype(np.uint32) << shift).astype(dt).view(np.int16) if bits != 8 else a).cuda()
    d_y, d_uv = as_t(y), as_t(uv)
    o_y = torch.zeros((oH, oW), dtype=tdt, device="cuda")
    o_uv = torch.zeros((och, 2 * ocw), dtype=tdt, device="cuda")
    bps = 1 if bits == 8 else 2
    for _ in range(2):
        assert eng.process_device_semiplanar(d_y.data_ptr(), d_y.stride(0) * bps, d_uv.data_ptr(), d_uv.stride(0) * bps,
                                             o_y.data_ptr(), o_y.stride(0) * bps, o_uv.data_ptr(), o_uv.stride(0) * bps, shift) == 0
    torch.cuda.synchronize()
    gy, guv = o_y.cpu().numpy().view(dt), o_uv.cpu().numpy().view(dt)
    eng.close()
    if shift:
        assert (gy & ((1 << shift) - 1)).max() == 0 and (guv & ((1 << shift) - 1)).max() == 0, "low bits must stay clear"
    assert np.array_equal(gy >> shift, py), "Y differs on %d px" % ((gy >> shift) != py).sum()
    assert np.array_equal(guv[:, 0::2] >> shift, pu) and np.array_equal(guv[:, 1::2] >> shift, pv), "chroma differs"


@pytest.mark.parametrize("bits", [8, 10])
def test_tma_staged_input_equals_scalar_staging(bits):
    """Stage A of exact-2x passes fetches the low-res window of interior tiles by a tensor-map TMA load (16-byte aligned planes); frame
    border tiles, unaligned planes and RAISR_CUDA_TMA=0 use clamped scalar loads.  All three must give the same frame -- and it must be
    the oracle's.  (A device plane whose base is off by 2 bytes / whose pitch is not a multiple of 16 exercises the fallback.)"""
    import torch
    w, h = 640, 360                                      # 12 x 16 tiles of which the inner ones take the TMA path
    f = T.filter_folder("filters_2x/filters_lowres")
    img = T.synth_frame(w, h, bits, seed=77, kind="mix")
    tdt = torch.uint8 if bits == 8 else torch.int16
    bps = 1 if bits == 8 else 2
    src = torch.from_numpy(img if bits == 8 else img.view(np.int16)).cuda()
    outs = {}
    for name, env, shift in (("tma", None, 0), ("scalar", "0", 0), ("unaligned", None, 2 // bps)):
        if env is None:
            os.environ.pop("RAISR_CUDA_TMA", None)
        else:
            os.environ["RAISR_CUDA_TMA"] = env
        try:
            eng = B.Engine(f, 2.0, bits, T.VideoRange, 1, 1, numerics=NUM)
        finally:
            os.environ.pop("RAISR_CUDA_TMA", None)
        eng.set_res(w, h, 2 * w, 2 * h)
        pitch_el = w + 32 + shift                            # elements per row of the padded device plane (a multiple of 16 bytes iff shift == 0)
        buf = torch.zeros((h + 1) * pitch_el + 64, dtype=tdt, device="cuda")
        plane = buf[shift:shift + h * pitch_el].view(h, pitch_el)[:, :w]
        plane.copy_(src)
        out = torch.empty((2 * h, 2 * w), dtype=tdt, device="cuda")
        assert eng.process_device_rows(plane.data_ptr(), pitch_el * bps, out.data_ptr(), out.stride(0) * bps, 0, 2 * h, 2, None) == 0
        torch.cuda.synchronize()
        outs[name] = out.cpu().numpy()
        eng.close()
    ref = oracle_run(f, img, 2.0, bits)
    for name, o in outs.items():
        got = o if bits == 8 else o.view(np.uint16)
        assert np.array_equal(got, ref), name
