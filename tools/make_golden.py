#!/usr/bin/env python
"""Generates tests/golden/*.npz: outputs of the COMPILED REFERENCE (oracle/_ref/libraisr_ref_dbg.so = untouched sources
+ IPP stand-in, AVX512 fp32 path, threadcount=1) on small seeded frames, with the bucket planes captured by the debug
hook.  These pin the oracle (tests/test_oracle.py) and the CUDA engine's x86-exact mode (tests/test_gpu_parity.py)
without needing /root/reference or an AVX-512 host at test time.

    python tools/make_golden.py        (needs oracle/_ref built here: make -C oracle ref)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import raisr_testlib as T  # noqa: E402

CASES = [
    # name, folder, ratio, bits, passes, mode, range, (w,h), kind, seed
    ("lowres_2x_8b_p1", "filters_2x/filters_lowres", 2.0, 8, 1, 1, T.VideoRange, (160, 90), "mix", 11),
    ("lowres_2x_8b_p1_ragged_full", "filters_2x/filters_lowres", 2.0, 8, 1, 1, T.FullRange, (125, 67), "noise", 12),
    ("highres_2x_8b_p2m1", "filters_2x/filters_highres", 2.0, 8, 2, 1, T.VideoRange, (160, 90), "mix", 13),
    ("denoise_2x_8b_p2m2", "filters_2x/filters_denoise", 2.0, 8, 2, 2, T.VideoRange, (160, 90), "edges", 14),
    ("denoise_2x_10b_p2m2", "filters_2x/filters_denoise", 2.0, 10, 2, 2, T.VideoRange, (160, 90), "mix", 15),
    ("highres_2x_10b_p1", "filters_2x/filters_highres", 2.0, 10, 1, 1, T.FullRange, (130, 74), "mix", 16),
    ("highres_15x_8b_p1", "filters_1.5x/filters_highres", 1.5, 8, 1, 1, T.VideoRange, (160, 90), "mix", 17),
    ("denoise_15x_8b_p2m2", "filters_1.5x/filters_denoise", 1.5, 8, 2, 2, T.VideoRange, (142, 80), "noise", 18),
    ("lowres_2x_8b_flat", "filters_2x/filters_lowres", 2.0, 8, 1, 1, T.VideoRange, (96, 64), "flat", 19),
    # blending = 1 (Randomness); width chosen so that the pixels the reference leaves unwritten (row H-7, columns
    # [c_end, W-6)) do not exist: c_end == W-6
    ("lowres_2x_8b_p1_randomness", "filters_2x/filters_lowres", 2.0, 8, 1, 1, T.VideoRange, (154, 88), "mix", 20, T.Randomness),
    ("highres_15x_8b_p1_randomness", "filters_1.5x/filters_highres", 1.5, 8, 1, 1, T.FullRange, (168, 90), "noise", 21, T.Randomness),
]


def main():
    # The reference keeps its configuration in process globals that RNLInit does not reset (gPasses, gTwoPassMode,
    # gUsePixelType ... Raisr_globals.h:140-203), so every case runs in a fresh process.
    if len(sys.argv) == 1:
        import subprocess
        for c in CASES:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), c[0]], stdout=subprocess.DEVNULL)
            z = np.load(os.path.join(ROOT, "tests", "golden", c[0] + ".npz"))
            print(c[0], z["out_y"].shape, "hashed px", int((z["hash0"] >= 0).sum()))
        return
    L = T.handler_lib(T.ref_lib_path(dbg=True))
    hp = (C.c_void_p * 2).in_dll(L, "g_raisr_dbg_hash")
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for case in [c for c in CASES if c[0] == sys.argv[1]]:
        name, folder, ratio, bits, passes, mode, rng, (w, h), kind, seed = case[:10]
        blending = case[10] if len(case) > 10 else T.CountOfBitsChanged
        img = T.synth_frame(w, h, bits, seed, kind)
        u, v = T.synth_chroma(w // 2, h // 2, bits, seed + 100), T.synth_chroma(w // 2, h // 2, bits, seed + 200)
        oW, oH = int(w * ratio), int(h * ratio)
        planes = []
        for i in range(passes):
            lr = passes == 2 and mode == 2 and i == 0
            planes.append(np.full((h, w) if lr else (oH, oW), -1, np.int32))
            hp[i] = planes[-1].ctypes.data
        for i in range(passes, 2):
            hp[i] = None
        oy, ou, ov = T.run_handler(L, T.filter_folder(folder), img, ratio, bits, rng, 1, T.AVX512, passes, mode, blending=blending, inU=u, inV=v)
        hp[0] = hp[1] = None
        d = {"in_y": img, "in_u": u, "in_v": v, "out_y": oy, "out_u": ou, "out_v": ov,
             "meta": np.array([ratio, bits, passes, mode, rng, seed, blending], np.float64), "folder": np.array(folder), "kind": np.array(kind)}
        for i, pl in enumerate(planes):
            d["hash%d" % i] = pl.astype(np.int16)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)


if __name__ == "__main__":
    main()
