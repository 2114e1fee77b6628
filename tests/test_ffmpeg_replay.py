"""The FFmpeg-side contract without FFmpeg: tests/harness/vf_raisr_replay.c is a plain-C program that drives libraisr.so exactly
like ffmpeg/vf_raisr.c:98-337 does (Init, AV_CEIL_RSHIFT plane geometry, evenoutput, SetRes on the first frame, Process per frame
on freshly allocated PAGEABLE planes with 64-byte-aligned, padded linesize, Deinit).

CPU part: the harness compiles as C99 against include/raisr/*.h and links against libraisr.so (the headers are C-clean, the five
plugin symbols resolve).  GPU part: its output equals the oracle for every pixel format family of vf_raisr.c:158-162."""
import os
import subprocess

import numpy as np
import pytest

import raisr_testlib as T

HARNESS_SRC = os.path.join(T.ROOT, "tests", "harness", "vf_raisr_replay.c")


def build_harness(tmp):
    exe = os.path.join(str(tmp), "vf_raisr_replay")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O2", "-I" + os.path.join(T.ROOT, "include"), HARNESS_SRC, "-o", exe,
                           "-L" + T.PKG_DIR, "-lraisr", "-Wl,-rpath," + T.PKG_DIR])
    return exe


def test_harness_compiles_as_c_and_links_the_plugin_symbols(tmp_path):
    exe = build_harness(tmp_path)
    out = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True, check=True).stdout
    for sym in ("RNLHandler_Init", "RNLHandler_SetRes", "RNLHandler_Process", "RNLHandler_Deinit"):
        assert sym in out, sym
    assert subprocess.run([exe], capture_output=True).returncode == 2          # usage: no compute without arguments


CASES = [
    # folder, ratio, bits, passes, mode, blending, pixfmt, evenoutput, (w, h)
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1, 2, 420, 0, (322, 182)),
    ("filters_2x/filters_lowres", 2.0, 8, 1, 1, 2, 422, 0, (161, 91)),        # odd sizes: AV_CEIL_RSHIFT chroma planes
    ("filters_2x/filters_highres", 2.0, 10, 2, 1, 2, 444, 0, (200, 120)),
    ("filters_1.5x/filters_denoise", 1.5, 8, 2, 2, 1, 420, 1, (211, 135)),    # 1.5x of odd sizes: evenoutput trims 316x202 -> 316x202 / 316.5
    ("filters_2x/filters_denoise", 2.0, 10, 2, 2, 2, 420, 0, (256, 144)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("folder,ratio,bits,passes,mode,blending,pixfmt,even,size", CASES,
                         ids=lambda v: str(v).replace("filters_", "").replace("/", "-") if isinstance(v, str) else None)
def test_replay_of_vf_raisr_matches_the_oracle(folder, ratio, bits, passes, mode, blending, pixfmt, even, size, tmp_path):
    exe = build_harness(tmp_path)
    w, h = size
    dt = np.uint8 if bits == 8 else np.dtype("<u2")
    hs, vs = (0 if pixfmt == 444 else 1), (1 if pixfmt == 420 else 0)
    cw, ch = -((-w) >> hs), -((-h) >> vs)
    oW, oH = int(w * ratio), int(h * ratio)
    if even:
        oW, oH = oW - oW % 2, oH - oH % 2
    ocw, och = -((-oW) >> hs), -((-oH) >> vs)
    frames = 3
    ins = []
    with open(tmp_path / "in.yuv", "wb") as f:
        for n in range(frames):
            y = T.synth_frame(w, h, bits, seed=600 + n, kind=("mix", "noise", "edges")[n % 3])
            u, v = T.synth_chroma(cw, ch, bits, 10 + n), T.synth_chroma(cw, ch, bits, 20 + n)
            ins.append((y, u, v))
            for p in (y, u, v):
                f.write(np.ascontiguousarray(p, dtype=dt).tobytes())
    env = dict(os.environ, RAISR_CUDA_NUMERICS="0")
    r = subprocess.run([exe, T.filter_folder(folder), str(w), str(h), str(bits), str(ratio), str(passes), str(mode), str(blending), str(pixfmt),
                        str(even), str(frames), str(tmp_path / "in.yuv"), str(tmp_path / "out.yuv")], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    raw = np.fromfile(tmp_path / "out.yuv", dtype=dt)
    per = oW * oH + 2 * ocw * och
    assert raw.size == frames * per
    m1 = T.OracleModel(T.filter_folder(folder), bits, False, T.VideoRange, 0, blending)
    m2 = T.OracleModel(T.filter_folder(folder), bits, True, T.VideoRange, 0, blending) if passes == 2 else None
    for n, (y, u, v) in enumerate(ins):
        fr = raw[n * per:(n + 1) * per]
        oy = fr[:oW * oH].reshape(oH, oW)
        ou = fr[oW * oH:oW * oH + ocw * och].reshape(och, ocw)
        ov = fr[oW * oH + ocw * och:].reshape(och, ocw)
        ref = T.oracle_process_y(y, oW, oH, m1, m2, passes, mode)
        assert np.array_equal(oy, ref), "frame %d: Y differs on %d px" % (n, (oy != ref).sum())
        assert np.array_equal(ou, T.oracle_resize(u, ocw, och)) and np.array_equal(ov, T.oracle_resize(v, ocw, och)), "frame %d chroma" % n


def test_cuda_hwframe_filter_source_type_checks_against_ffmpeg_stubs():
    """ffmpeg/vf_raisr_cuda.c cannot be built here (no libav headers).  tests/harness/ffstub restates the FFmpeg declarations it
    uses; `gcc -fsyntax-only` against those and include/raisr_cuda.h catches typos, wrong member names and wrong C-ABI calls."""
    src = os.path.join(T.ROOT, "ffmpeg", "vf_raisr_cuda.c")
    r = subprocess.run(["gcc", "-std=gnu11", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter",
                        "-Wno-missing-field-initializers", "-I" + os.path.join(T.ROOT, "tests", "harness", "ffstub"),
                        "-I" + os.path.join(T.ROOT, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
