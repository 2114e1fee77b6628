"""Full-frame parity at BASELINE.json's sizes: the CUDA engine through the plugin symbols (RNLHandler_*) against the
UNTOUCHED reference sources (oracle/_ref, AVX-512 fp32 path) run live on this host with all its threads -- every pixel of the
3840x2160 frames of configs[1] and configs[2] and of the 7680x4320 frame of configs[3], luma and chroma, bit for bit.  (The
reference takes ~25 ms per 4K frame on 16 threads, so whole frames are affordable; slabs are not needed.)"""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

import raisr_testlib as T

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (T.have_ref() and T.have_avx512()), reason="oracle/_ref not usable on this host")]

_spec = importlib.util.spec_from_file_location("raisr_binding", os.path.join(T.PKG_DIR, "binding.py"))
B = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(B)

FULL = [
    # BASELINE index, folder, bits, passes, mode, (w, h)
    ("configs[1]", "filters_2x/filters_lowres", 8, 1, 1, (1920, 1080)),
    ("configs[2]", "filters_2x/filters_highres", 8, 2, 1, (1920, 1080)),
    ("configs[3]", "filters_2x/filters_denoise", 10, 2, 2, (3840, 2160)),
]


def reference_frame(folder, y, u, v, bits, passes, mode):
    """yuv420p frame through the compiled reference in a fresh process (threadcount = all host threads)."""
    import pickle
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        np.savez(os.path.join(td, "in.npz"), y=y, u=u, v=v)
        code = (
            "import sys, os, numpy as np\n"
            "sys.path.insert(0, %r)\n"
            "import raisr_testlib as T\n"
            "z = np.load(%r)\n"
            "L = T.handler_lib(T.ref_lib_path())\n"
            "out = T.run_handler(L, T.filter_folder(%r), z['y'], 2.0, %d, T.VideoRange, os.cpu_count() or 1, T.AVX512, %d, %d, inU=z['u'], inV=z['v'])\n"
            "np.savez(%r, y=out[0], u=out[1], v=out[2])\n"
        ) % (os.path.dirname(os.path.abspath(__file__)), os.path.join(td, "in.npz"), folder, bits, passes, mode, os.path.join(td, "out.npz"))
        subprocess.check_call([sys.executable, "-c", code], stdout=subprocess.DEVNULL)
        z = np.load(os.path.join(td, "out.npz"))
        return z["y"], z["u"], z["v"]


@pytest.mark.parametrize("name,folder,bits,passes,mode,size", FULL, ids=[f[0] for f in FULL])
def test_whole_frame_bit_identical_to_the_live_reference(name, folder, bits, passes, mode, size, monkeypatch):
    w, h = size
    y = T.synth_frame(w, h, bits, seed=4000 + w + passes)
    u, v = T.synth_chroma(w // 2, h // 2, bits, 11), T.synth_chroma(w // 2, h // 2, bits, 12)
    ry, ru, rv = reference_frame(folder, y, u, v, bits, passes, mode)
    assert ry.shape == (2 * h, 2 * w)
    monkeypatch.setenv("RAISR_CUDA_NUMERICS", "1")                  # x86-exact: the compiled reference is the target
    L = T.handler_lib(T.product_lib_path())
    oy, ou, ov = T.run_handler(L, T.filter_folder(folder), y, 2.0, bits, T.VideoRange, 1, T.AVX512, passes, mode, inU=u, inV=v, frames=2)
    assert np.array_equal(oy, ry), "%s: Y differs on %d of %d px (max %d)" % (
        name, (oy != ry).sum(), ry.size, np.abs(oy.astype(np.int64) - ry.astype(np.int64)).max())
    assert np.array_equal(ou, ru) and np.array_equal(ov, rv), "%s: chroma differs" % name


def test_whole_4k_frame_buckets_identical_to_the_live_reference():
    """configs[1]: every bucket index of the 3840x2160 frame (the reference's debug build exports its hash plane)."""
    if not T.have_ref(dbg=True):
        pytest.skip("debug build of the reference not present")
    y = T.synth_frame(1920, 1080, 8, seed=4100)
    ry, hashes = T.run_ref_subprocess("filters_2x/filters_lowres", y, threads=1, want_hash=True)
    eng = B.Engine(T.filter_folder("filters_2x/filters_lowres"), 2.0, 8, T.VideoRange, 1, 1, numerics=B.NUMERICS_X86, keep_hash=True)
    eng.set_res(1920, 1080, 3840, 2160)
    out = np.zeros((2160, 3840), np.uint8)
    assert eng.process_host(y, out) == 0
    hv = eng.read_hash(0, 3840, 2160)
    eng.close()
    assert np.array_equal(hv, hashes[0]), "buckets differ on %d px" % (hv != hashes[0]).sum()
    assert np.array_equal(out, ry)
