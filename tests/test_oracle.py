"""CPU tests (no GPU): the oracle (oracle/raisr_oracle.c) against the golden vectors produced by the COMPILED REFERENCE
(tests/golden/*.npz, tools/make_golden.py) and, when oracle/_ref is present, against the reference run live."""
import numpy as np
import pytest

import raisr_testlib as T

needs_filters = pytest.mark.skipif(not T.have_filters(), reason="trained filter folders not staged (run make -C oracle ref)")
needs_avx512 = pytest.mark.skipif(not T.have_avx512(), reason="ORACLE_SQRT_X86 executes vrcp14ps/vrsqrt14ps on the host")


def run_oracle(g, sqrt_mode):
    f = T.filter_folder(g["folder"])
    m1 = T.OracleModel(f, g["bits"], False, g["rng"], sqrt_mode, g["blending"])
    m2 = T.OracleModel(f, g["bits"], True, g["rng"], sqrt_mode, g["blending"]) if g["passes"] == 2 else None
    h, w = g["in_y"].shape
    return T.oracle_process_y(g["in_y"], int(w * g["ratio"]), int(h * g["ratio"]), m1, m2, g["passes"], g["mode"], want_hash=True)


@needs_filters
@needs_avx512
@pytest.mark.parametrize("name", T.golden_names())
def test_oracle_x86_mode_is_bit_identical_to_reference(name):
    """ORACLE_SQRT_X86 (the hash as compiled) reproduces the reference binary exactly: buckets and Y."""
    g = T.load_golden(name)
    out, h1, h2 = run_oracle(g, 1)
    assert np.array_equal(h1, g["hash"][0]), "pass-1 buckets: %d differ" % (h1 != g["hash"][0]).sum()
    if g["passes"] == 2:
        assert np.array_equal(h2, g["hash"][1]), "pass-2 buckets: %d differ" % (h2 != g["hash"][1]).sum()
    assert np.array_equal(out, g["out_y"])


@needs_filters
@pytest.mark.parametrize("name", T.golden_names())
def test_oracle_ieee_mode_close_to_reference(name):
    """ORACLE_SQRT_IEEE (source semantics with exact sqrt/div) differs from the binary only where the x86
    approximations flip a bucket: a small fraction of pixels (0.1-1.3 %)."""
    g = T.load_golden(name)
    out, h1, _ = run_oracle(g, 0)
    hashed = g["hash"][0] >= 0
    assert np.array_equal(h1 >= 0, hashed), "hashed set differs"
    frac_bucket = (h1 != g["hash"][0])[hashed].mean() if hashed.any() else 0.0
    frac_y = (out != g["out_y"]).mean()
    assert frac_bucket < 0.02 and frac_y < 0.03, (frac_bucket, frac_y)


@pytest.mark.parametrize("name", T.golden_names())
def test_oracle_resize_matches_reference_chroma(name):
    """Chroma planes are the plain cheap upscale (Raisr.cpp:1373-1388): oracle_resize == ipp stand-in inside the reference."""
    g = T.load_golden(name)
    for i, o in ((g["in_u"], g["out_u"]), (g["in_v"], g["out_v"])):
        assert np.array_equal(T.oracle_resize(i, o.shape[1], o.shape[0]), o)


def test_resize_2x_closed_form():
    """2x: weights {1/4,3/4}^2 -> (9a+3b+3c+d+8)>>4 with replicate border."""
    rs = np.random.RandomState(3)
    a = rs.randint(0, 1024, size=(9, 13)).astype(np.uint16)
    up = T.oracle_resize(a, 26, 18).astype(np.int64)
    p = np.pad(a.astype(np.int64), 1, mode="edge")
    for Y in range(18):
        for X in range(26):
            j, i = Y >> 1, X >> 1
            ya, yb, wa, wb = (j, j + 1, 3, 1) if Y & 1 else (j - 1, j, 1, 3)
            xa, xb, va, vb = (i, i + 1, 3, 1) if X & 1 else (i - 1, i, 1, 3)
            s = wa * (va * p[ya + 1, xa + 1] + vb * p[ya + 1, xb + 1]) + wb * (va * p[yb + 1, xa + 1] + vb * p[yb + 1, xb + 1])
            assert up[Y, X] == (s + 8) >> 4


def test_exported_standin_resize_equals_the_oracle_resize():
    """oracle/_build/libipp_standin.so (the stand-in's resize on its own, timed by bench.py's cpu_baseline context) is the same
    arithmetic as the oracle's resize: 2x and 1.5x, odd sizes."""
    import ctypes as C, os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_build", "libipp_standin.so")
    if not os.path.exists(path):
        T.build_oracle()
    S = C.CDLL(path)
    rs = np.random.RandomState(5)
    for (w, h, ow, oh) in ((64, 40, 128, 80), (37, 21, 74, 42), (64, 40, 96, 60), (50, 34, 75, 51)):
        a = np.ascontiguousarray(rs.randint(0, 256, size=(h, w)).astype(np.uint8))
        up = np.zeros((oh, ow), np.uint8)
        assert S.standin_resize_8u(C.c_void_p(a.ctypes.data), w, h, w, C.c_void_p(up.ctypes.data), ow, oh, ow) == 0
        assert np.array_equal(up, T.oracle_resize(a, ow, oh)), (w, h, ow, oh)


def test_hashed_column_range():
    """Column loop of processSegment (Raisr.cpp:1065-1066,1246-1250): c_end = 6 + 8*floor((W-12)/8) for W >= 28."""
    import ctypes as C
    L = T.oracle_lib()
    for W in range(12, 400):
        ce, ts = C.c_int(), C.c_int()
        L.oracle_hashed_cols(W, C.byref(ce), C.byref(ts))
        if W >= 28:
            assert ce.value == 6 + 8 * ((W - 12) // 8), W
            assert ts.value <= ce.value and (ce.value - ts.value) in (8, 16), (W, ts.value, ce.value)
        else:
            assert ce.value == 6        # narrower than one 16-block: nothing is hashed


def test_gaussian_table_symmetry():
    import ctypes as C
    L = T.oracle_lib()
    for bits in (8, 10, 16):
        w = np.zeros(121, np.float32)
        L.oracle_gaussian_weights(bits, w.ctypes.data_as(C.c_void_p))
        w = w.reshape(11, 11)
        assert np.array_equal(w, w.T) and np.array_equal(w, w[::-1]) and np.array_equal(w, w[:, ::-1])
        M = {8: 255.0, 10: 1023.0, 16: 65535.0}[bits]
        assert abs(w[5, 5] * (M * M * 4) - 0.0402265) < 1e-8


@needs_filters
def test_constant_frame_invariants():
    """constant v -> 1-px frame = v (unclamped), unfiltered border = clamp(v)."""
    m = T.OracleModel(T.filter_folder("filters_2x/filters_lowres"), 8)
    out = T.oracle_process_y(np.full((40, 60), 7, np.uint8), 120, 80, m)
    assert (out[0] == 7).all() and (out[-1] == 7).all() and (out[:, 0] == 7).all() and (out[:, -1] == 7).all()
    assert (out[1:6, 1:-1] == 16).all() and (out[1:-1, 1:6] == 16).all()


@needs_filters
@needs_avx512
@pytest.mark.skipif(not T.have_ref(dbg=True), reason="oracle/_ref not built")
@pytest.mark.parametrize("folder,ratio,bits,passes,mode,size", [
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1, (202, 118)),
    ("filters_2x/filters_highres", 2.0, 8, 2, 1, (176, 100)),
    ("filters_1.5x/filters_denoise", 1.5, 8, 2, 2, (200, 112)),
])
def test_oracle_vs_live_reference(folder, ratio, bits, passes, mode, size):
    """Fresh seeds through the compiled reference (separate process) and the oracle: bit-identical buckets and Y."""
    w, h = size
    img = T.synth_frame(w, h, bits, seed=4242 + w, kind="mix")
    ref_y, ref_h = T.run_ref_subprocess(folder, img, ratio, bits, passes=passes, mode=mode, want_hash=True)
    f = T.filter_folder(folder)
    m1 = T.OracleModel(f, bits, False, T.VideoRange, 1)
    m2 = T.OracleModel(f, bits, True, T.VideoRange, 1) if passes == 2 else None
    out, h1, h2 = T.oracle_process_y(img, int(w * ratio), int(h * ratio), m1, m2, passes, mode, want_hash=True)
    assert np.array_equal(h1, ref_h[0])
    if passes == 2:
        assert np.array_equal(h2, ref_h[1])
    assert np.array_equal(out, ref_y)


@pytest.mark.skipif(not (T.have_ref() and T.have_avx512()), reason="oracle/_ref not usable on this host")
@pytest.mark.parametrize("folder,passes,mode,size", [
    ("filters_1.5x/filters_highres", 1, 1, (211, 135)),        # 135 * 1.5 = 202.5 -> 202 output rows: the resize reads 134 source rows
    ("filters_1.5x/filters_denoise", 2, 2, (211, 135)),
    ("filters_1.5x/filters_denoise", 2, 1, (98, 51)),          # (an output width = 1 mod 8 would hit the reference's blend-loop overshoot)
])
def test_truncated_output_height_reads_fewer_source_rows_like_the_reference(folder, passes, mode, size):
    """Raisr.cpp:1801-1803: the luma resize spec is {inW, (int)(outH / ratio)} -> {outW, outH}.  With an odd input height at 1.5x the
    output height is truncated and the last input row is never read -- pinned here against the compiled reference."""
    w, h = size
    img = T.synth_frame(w, h, 8, seed=7000 + w, kind="mix")
    ref_y, _ = T.run_ref_subprocess(folder, img, ratio=1.5, passes=passes, mode=mode)
    oW, oH = int(w * 1.5), int(h * 1.5)
    assert ref_y.shape == (oH, oW)
    f = T.filter_folder(folder)
    m1 = T.OracleModel(f, 8, False, T.VideoRange, 1)
    m2 = T.OracleModel(f, 8, True, T.VideoRange, 1) if passes == 2 else None
    got = T.oracle_process_y(img, oW, oH, m1, m2, passes, mode, ratio=1.5)
    assert np.array_equal(got, ref_y), "%d px differ" % (got != ref_y).sum()


def test_standin_vs_real_ipp(capsys):
    """The cheap-upscale stage is DEFINED by oracle/ipp_standin (IPP is closed source and absent here: parity for this one stage
    is unpinned).  When tests/golden/ipp_*.npz exist (tools/ipp_pin/ run on a machine with oneAPI IPP) this
    reports how many upscaled samples of the stand-in differ from real IPP, and by how much; it asserts only |d| <= 1 LSB."""
    names = T.ipp_golden_names()
    if not names:
        pytest.skip("no real-IPP vectors committed (tools/ipp_pin/make_ipp_golden.py needs oneAPI IPP): the stage stays unpinned")
    for name in names:
        g = T.load_golden(name)
        z = np.load(os.path.join(T.ROOT, "tests", "golden", "ipp_" + name + ".npz"))
        oh, ow = g["out_y"].shape
        src_h = min(g["in_y"].shape[0], int(oh / g["ratio"]))
        mine = T.oracle_resize(g["in_y"][:src_h], ow, oh)
        d = np.abs(mine.astype(np.int64) - z["up_y"].astype(np.int64))
        with capsys.disabled():
            print("\\n[ipp pin] %s: stand-in differs from IPP on %.3f %% of upscaled luma samples, max %d LSB" % (name, 100.0 * (d != 0).mean(), d.max()))
        assert d.max() <= 1
