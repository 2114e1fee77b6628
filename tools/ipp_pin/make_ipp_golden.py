"""Runs ipp_dump (real Intel IPP, see ipp_dump.c) on the luma and chroma inputs of every golden case and stores IPP's upscaled
planes as tests/golden/ipp_<case>.npz.  tests/test_oracle.py::test_standin_vs_real_ipp then REPORTS (does not assert) where the
stand-in differs.  usage: python tools/ipp_pin/make_ipp_golden.py ./ipp_dump"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import raisr_testlib as T

exe = os.path.abspath(sys.argv[1])
for name in T.golden_names():
    if name.startswith("ipp_"):
        continue
    g = T.load_golden(name)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for key, shape_of in (("in_y", "out_y"), ("in_u", "out_u"), ("in_v", "out_v")):
            a = np.ascontiguousarray(g[key])
            oh, ow = g[shape_of].shape
            src_h = min(a.shape[0], int(oh / g["ratio"])) if key == "in_y" else a.shape[0]      # Raisr.cpp:1801-1803 / :1821
            a[:src_h].tofile(os.path.join(td, "in.raw"))
            subprocess.check_call([exe, "8" if g["bits"] == 8 else "16", str(a.shape[1]), str(src_h), str(ow), str(oh),
                                   os.path.join(td, "in.raw"), os.path.join(td, "out.raw")])
            out["up_" + key[3:]] = np.fromfile(os.path.join(td, "out.raw"), dtype=a.dtype).reshape(oh, ow)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ipp_" + name + ".npz"), **out)
    print("wrote ipp_%s.npz" % name)
