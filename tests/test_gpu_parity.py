"""GPU parity tests: the CUDA engine, called through the C ABI, against the oracle on the same seeded inputs.
Bit-exact is the bar (integer output planes and bucket indices)."""
import importlib.util
import os

import numpy as np
import pytest

import raisr_testlib as T

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("raisr_binding", os.path.join(T.PKG_DIR, "binding.py"))
B = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(B)


def run_engine(folder, img, ratio=2.0, bits=8, passes=1, mode=1, rng=T.VideoRange, numerics=B.NUMERICS_IEEE,
               want_hash=True, out_pad=0):
    H, W = img.shape
    oW, oH = int(W * ratio), int(H * ratio)
    eng = B.Engine(folder, ratio, bits, rng, passes, mode, numerics=numerics, keep_hash=want_hash)
    eng.set_res(W, H, oW, oH)
    outb = np.zeros((oH, oW + out_pad), img.dtype)
    out = outb[:, :oW]
    rc = eng.process_host(img, out)
    assert rc == 0
    hashes = []
    if want_hash:
        for i in range(passes):
            lr = passes == 2 and mode == 2 and i == 0
            hashes.append(eng.read_hash(i, W if lr else oW, H if lr else oH))
    n = eng.launch_count()
    eng.close()
    assert n >= 1            # (two-pass configurations are ONE chained launch by default)
    return out.copy(), hashes


def oracle_models(folder, bits, passes, rng=T.VideoRange, sqrt_mode=0):
    m1 = T.OracleModel(folder, bits, False, rng, sqrt_mode)
    m2 = T.OracleModel(folder, bits, True, rng, sqrt_mode) if passes == 2 else None
    return m1, m2


CASES = [
    # folder, ratio, bits, passes, mode, (w, h), kind
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1, (480, 270), "mix"),
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1, (250, 131), "noise"),      # ragged: width not a multiple of 8/16
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1, (64, 40), "edges"),
    ("filters_2x/filters_highres", 2.0, 8, 2, 1, (320, 180), "mix"),
    ("filters_2x/filters_denoise", 2.0, 8, 2, 2, (320, 180), "mix"),
    ("filters_2x/filters_denoise", 2.0, 10, 2, 2, (320, 180), "mix"),
    ("filters_2x/filters_highres", 2.0, 10, 1, 1, (322, 182), "edges"),
    ("filters_1.5x/filters_highres", 1.5, 8, 1, 1, (320, 180), "mix"),
    ("filters_1.5x/filters_denoise", 1.5, 8, 2, 2, (426, 240), "noise"),
]


@pytest.mark.parametrize("folder,ratio,bits,passes,mode,size,kind", CASES)
def test_bit_exact_vs_oracle(folder, ratio, bits, passes, mode, size, kind):
    w, h = size
    f = T.filter_folder(folder)
    img = T.synth_frame(w, h, bits, seed=1234 + w, kind=kind)
    out, hashes = run_engine(f, img, ratio, bits, passes, mode)
    m1, m2 = oracle_models(f, bits, passes)
    ref, h1, h2 = T.oracle_process_y(img, int(w * ratio), int(h * ratio), m1, m2, passes, mode, want_hash=True)
    assert np.array_equal(hashes[0], h1), "pass-1 buckets differ: %d" % (hashes[0] != h1).sum()
    if passes == 2:
        assert np.array_equal(hashes[1], h2), "pass-2 buckets differ: %d" % (hashes[1] != h2).sum()
    assert np.array_equal(out, ref), "Y differs on %d px, max %d" % (
        (out != ref).sum(), np.abs(out.astype(int) - ref.astype(int)).max())


def test_full_range_and_padded_step():
    f = T.filter_folder("filters_2x/filters_lowres")
    img = T.synth_frame(200, 120, 8, seed=5, kind="noise")
    out, _ = run_engine(f, img, rng=T.FullRange, want_hash=False, out_pad=64)
    m1, _ = oracle_models(f, 8, 1, rng=T.FullRange)
    ref = T.oracle_process_y(img, 400, 240, m1)
    assert np.array_equal(out, ref)


def test_handler_api_yuv420():
    """RNLHandler_* call sequence of vf_raisr.c with chroma planes; chroma = exact-rational bilinear."""
    f = T.filter_folder("filters_2x/filters_lowres")
    w, h = 256, 144
    img = T.synth_frame(w, h, 8, seed=77)
    u, v = T.synth_chroma(w // 2, h // 2, 8, 1), T.synth_chroma(w // 2, h // 2, 8, 2)
    os.environ["RAISR_CUDA_NUMERICS"] = "0"
    L = T.handler_lib(T.product_lib_path())
    oy, ou, ov = T.run_handler(L, f, img, inU=u, inV=v, frames=2)
    m1, _ = oracle_models(f, 8, 1)
    assert np.array_equal(oy, T.oracle_process_y(img, 2 * w, 2 * h, m1))
    assert np.array_equal(ou, T.oracle_resize(u, w, h))
    assert np.array_equal(ov, T.oracle_resize(v, w, h))


def test_constant_frame_invariants():
    """constant input v -> unfiltered border = clamp(v), 1-px frame = v."""
    f = T.filter_folder("filters_2x/filters_lowres")
    img = np.full((60, 80), 7, np.uint8)        # below video-range minimum 16
    out, _ = run_engine(f, img, want_hash=False)
    assert (out[0, :] == 7).all() and (out[-1, :] == 7).all() and (out[:, 0] == 7).all() and (out[:, -1] == 7).all()
    assert (out[1:6, 1:-1] == 16).all()


# ---- x86-exact numerics: the CUDA engine against the COMPILED REFERENCE ---------------------------------------
@pytest.mark.parametrize("name", T.golden_names())
def test_x86_mode_bit_identical_to_reference_golden(name):
    """RAISR_NUMERICS_X86 vs tests/golden (outputs of the untouched reference sources built with their own flags):
    bit-identical buckets, Y and chroma."""
    g = T.load_golden(name)
    h, w = g["in_y"].shape
    oW, oH = int(w * g["ratio"]), int(h * g["ratio"])
    eng = B.Engine(T.filter_folder(g["folder"]), g["ratio"], g["bits"], g["rng"], g["passes"], g["mode"],
                   numerics=B.NUMERICS_X86, keep_hash=True)
    eng.set_res(w, h, oW, oH, g["in_u"].shape[1], g["in_u"].shape[0], g["out_u"].shape[1], g["out_u"].shape[0])
    oy, ou, ov = np.zeros_like(g["out_y"]), np.zeros_like(g["out_u"]), np.zeros_like(g["out_v"])
    assert eng.process_host(g["in_y"], oy, g["in_u"], g["in_v"], ou, ov, blending=g["blending"]) == 0
    for i in range(g["passes"]):
        hv = eng.read_hash(i, g["hash"][i].shape[1], g["hash"][i].shape[0])
        assert np.array_equal(hv, g["hash"][i]), "pass %d buckets: %d differ" % (i, (hv != g["hash"][i]).sum())
    eng.close()
    assert np.array_equal(oy, g["out_y"]), "Y differs on %d px" % (oy != g["out_y"]).sum()
    assert np.array_equal(ou, g["out_u"]) and np.array_equal(ov, g["out_v"])


@pytest.mark.skipif(not T.have_avx512(), reason="ORACLE_SQRT_X86 executes vrcp14ps/vrsqrt14ps on the host")
@pytest.mark.parametrize("folder,ratio,bits,passes,mode,size,kind", [
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1, (960, 540), "mix"),          # BASELINE configs[0]: 540p -> 1080p
    ("filters_2x/filters_highres", 2.0, 8, 2, 1, (480, 270), "noise"),
    ("filters_2x/filters_denoise", 2.0, 10, 2, 2, (480, 270), "edges"),
    ("filters_1.5x/filters_highres", 1.5, 8, 1, 1, (640, 360), "mix"),
])
def test_x86_mode_vs_oracle_x86(folder, ratio, bits, passes, mode, size, kind):
    """Larger frames: CUDA x86 mode vs the oracle's as-compiled hash (itself pinned to the reference binary)."""
    w, h = size
    f = T.filter_folder(folder)
    img = T.synth_frame(w, h, bits, seed=99 + w, kind=kind)
    out, hashes = run_engine(f, img, ratio, bits, passes, mode, numerics=B.NUMERICS_X86)
    m1, m2 = oracle_models(f, bits, passes, sqrt_mode=1)
    ref, h1, h2 = T.oracle_process_y(img, int(w * ratio), int(h * ratio), m1, m2, passes, mode, want_hash=True)
    assert np.array_equal(hashes[0], h1), "pass-1 buckets differ: %d" % (hashes[0] != h1).sum()
    if passes == 2:
        assert np.array_equal(hashes[1], h2), "pass-2 buckets differ: %d" % (hashes[1] != h2).sum()
    assert np.array_equal(out, ref)


@pytest.mark.skipif(not (T.have_ref() and T.have_avx512()), reason="oracle/_ref not usable on this host")
def test_x86_mode_vs_live_reference_1080p():
    """configs[0] against the reference library run live on this host's CPU (AVX512 path, 1 thread)."""
    img = T.synth_frame(960, 540, 8, seed=2024)
    ref_y, _ = T.run_ref_subprocess("filters_2x/filters_lowres", img)
    out, _ = run_engine(T.filter_folder("filters_2x/filters_lowres"), img, numerics=B.NUMERICS_X86, want_hash=False)
    assert np.array_equal(out, ref_y), "Y differs on %d px" % (out != ref_y).sum()


def test_full_size_properties_4k():
    """BASELINE configs[1] size (1080p -> 4K), size-independent properties instead of a full oracle run:
    (1) row-band launches reproduce the full-frame launch bit for bit (band independence);
    (2) device entry point == host entry point; (3) 1-px frame equals the cheap upscale; (4) range clamp holds inside."""
    import torch
    f = T.filter_folder("filters_2x/filters_lowres")
    img = T.synth_frame(1920, 1080, 8, seed=7)
    eng = B.Engine(f, 2.0, 8, T.VideoRange, 1, 1, numerics=B.NUMERICS_AUTO)
    eng.set_res(1920, 1080, 3840, 2160)
    full = np.zeros((2160, 3840), np.uint8)
    assert eng.process_host(img, full) == 0
    d_in = torch.from_numpy(img).cuda()
    d_out = torch.zeros((2160, 3840), dtype=torch.uint8, device="cuda")
    bands = [(0, 540), (540, 1082), (1082, 2160)]
    for r0, r1 in bands:
        assert eng.process_device_rows(d_in.data_ptr(), d_in.stride(0), d_out.data_ptr(), d_out.stride(0), r0, r1) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), full)
    up = T.oracle_resize(img, 3840, 2160)
    assert np.array_equal(full[0], up[0]) and np.array_equal(full[-1], up[-1])
    assert np.array_equal(full[:, 0], up[:, 0]) and np.array_equal(full[:, -1], up[:, -1])
    inner = full[1:-1, 1:-1]
    assert inner.min() >= 16 and inner.max() <= 235
    # a 256-row slab checked against the oracle (rows far from the slab do not influence it)
    m = T.OracleModel(f, 8, sqrt_mode=1 if T.have_avx512() else 0)
    if T.have_avx512():
        slab_in = img[400:400 + 160]                      # LR rows 400..559 -> HR rows 800..1119
        ref = T.oracle_process_y(slab_in, 3840, 320, m)
        assert np.array_equal(full[800 + 16:1120 - 16], ref[16:-16])
    eng.close()


def test_phase_sequential_kernel_variant_is_bit_identical(monkeypatch):
    """RAISR_CUDA_KERNEL=tile (phase-sequential schedule) must give exactly the default (pipelined, warp-specialised) kernel's
    buckets and pixels."""
    f = T.filter_folder("filters_2x/filters_highres")
    img = T.synth_frame(500, 300, 8, seed=31, kind="mix")
    base, hb = run_engine(f, img, 2.0, 8, 2, 1, numerics=B.NUMERICS_AUTO)
    monkeypatch.setenv("RAISR_CUDA_KERNEL", "tile")
    tile, ht = run_engine(f, img, 2.0, 8, 2, 1, numerics=B.NUMERICS_AUTO)
    assert np.array_equal(base, tile) and all(np.array_equal(a, b) for a, b in zip(hb, ht))


@pytest.mark.parametrize("numerics", [B.NUMERICS_IEEE, B.NUMERICS_X86])
@pytest.mark.parametrize("folder,ratio,bits,passes,mode,size", [
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1, (250, 140)),       # has pixels the reference leaves unwritten: defined as the upscale
    ("filters_2x/filters_highres", 2.0, 10, 2, 1, (200, 120)),
    ("filters_1.5x/filters_denoise", 1.5, 8, 2, 2, (240, 136)),
])
def test_randomness_blending_vs_oracle(folder, ratio, bits, passes, mode, size, numerics):
    """blending = 1 (Randomness, Raisr.cpp:1203-1242) against the oracle, both numerics."""
    if numerics == B.NUMERICS_X86 and not T.have_avx512():
        pytest.skip("ORACLE_SQRT_X86 needs AVX-512 on the host")
    w, h = size
    f = T.filter_folder(folder)
    img = T.synth_frame(w, h, bits, seed=808 + w, kind="mix")
    oW, oH = int(w * ratio), int(h * ratio)
    eng = B.Engine(f, ratio, bits, T.VideoRange, passes, mode, numerics=numerics)
    eng.set_res(w, h, oW, oH)
    out = np.zeros((oH, oW), img.dtype)
    assert eng.process_host(img, out, blending=T.Randomness) == 0
    eng.close()
    sm = 1 if numerics == B.NUMERICS_X86 else 0
    m1 = T.OracleModel(f, bits, False, T.VideoRange, sm, T.Randomness)
    m2 = T.OracleModel(f, bits, True, T.VideoRange, sm, T.Randomness) if passes == 2 else None
    ref = T.oracle_process_y(img, oW, oH, m1, m2, passes, mode)
    assert np.array_equal(out, ref), "Y differs on %d px" % (out != ref).sum()


# ---- two-pass configurations in ONE persistent launch (chained passes) ------------------------------------------------------
@pytest.mark.parametrize("folder,ratio,bits,passes,mode,size", [
    ("filters_2x/filters_highres", 2.0, 8, 2, 1, (1000, 600)),          # mode 1: exact-2x upscale in pass 1 -> plain pass 2
    ("filters_2x/filters_denoise", 2.0, 10, 2, 2, (960, 540)),          # mode 2: pass 1 at input resolution -> upscaling pass 2
    ("filters_1.5x/filters_denoise", 1.5, 8, 2, 2, (640, 360)),         # one pixel type, axis-map upscale in pass 2
    ("filters_2x/filters_highres", 2.0, 8, 2, 1, (150, 66)),            # fewer tiles than SMs: CTAs without a pass-1 tile
])
def test_chained_passes_equal_one_launch_per_pass(folder, ratio, bits, passes, mode, size, monkeypatch):
    """Both passes in one cooperative launch (pass-2 tiles wait for the pass-1 tile rows they read; the default for mode 2 and 1.5x,
    RAISR_CUDA_CHAIN=1 for every pair).  Must be bit-identical to RAISR_CUDA_CHAIN=0 (a kernel boundary between the passes) --
    buckets of both passes and Y -- and use a single launch."""
    w, h = size
    f = T.filter_folder(folder)
    img = T.synth_frame(w, h, bits, seed=909 + w, kind="mix")
    oW, oH = int(w * ratio), int(h * ratio)

    def run():
        eng = B.Engine(f, ratio, bits, T.VideoRange, passes, mode, numerics=B.NUMERICS_AUTO, keep_hash=True)
        eng.set_res(w, h, oW, oH)
        out = np.zeros((oH, oW), img.dtype)
        n0 = eng.launch_count()
        for _ in range(3):                                           # the tile-row counters are re-armed per frame
            out[...] = 0
            assert eng.process_host(img, out) == 0
        n = (eng.launch_count() - n0) // 3
        hs = [eng.read_hash(i, w if (mode == 2 and i == 0) else oW, h if (mode == 2 and i == 0) else oH) for i in range(2)]
        eng.close()
        return out, hs, n

    monkeypatch.setenv("RAISR_CUDA_CHAIN", "1")                      # (the default chains only where it was measured faster)
    chained, hc, nc = run()
    monkeypatch.setenv("RAISR_CUDA_CHAIN", "0")
    split, hs, ns = run()
    assert ns == 2 and nc == 1, "launches per frame: chained %d, split %d" % (nc, ns)
    assert all(np.array_equal(a, b) for a, b in zip(hc, hs)), "buckets differ"
    assert np.array_equal(chained, split), "Y differs on %d px" % (chained != split).sum()
