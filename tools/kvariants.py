"""Times the 1080p->4K luma pass of every library in video-super-resolution-library_b200/variants/ (plus the regular build) on the GPU box and
checks that all of them produce the same bytes.   usage: python tools/kvariants.py [reps]"""
import glob, hashlib, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(HERE, "..", "video-super-resolution-library_b200")
ONE = r'''
import os, sys, hashlib, importlib.util
sys.path.insert(0, os.path.join(%r, "..", "tests"))
import numpy as np, torch
import raisr_testlib as T
spec = importlib.util.spec_from_file_location("b", os.path.join(T.PKG_DIR, "binding.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
B.LIB_PATH = os.path.abspath(sys.argv[1]); reps = int(sys.argv[2])
NB = 12
ys = [torch.from_numpy(T.synth_frame(1920, 1080, 8, 1234 + i)).cuda() for i in range(NB)]
outs = [torch.empty((2160, 3840), dtype=torch.uint8, device="cuda") for _ in range(NB)]
eng = B.Engine(T.filter_folder("filters_2x/filters_lowres"), 2.0, 8, 1, 1, 1, device=0, numerics=1)
eng.set_res(1920, 1080, 3840, 2160)
def run(i):
    eng.process_device_rows(ys[i %% NB].data_ptr(), ys[i %% NB].stride(0), outs[i %% NB].data_ptr(), outs[i %% NB].stride(0), 0, 2160, 2, None)
for i in range(NB): run(i)
torch.cuda.synchronize()
dig = hashlib.sha1(b"".join(o.cpu().numpy().tobytes() for o in outs[:3])).hexdigest()[:12]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    e0.record()
    for i in range(reps): run(i)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / reps)
ts.sort()
print("%%.4f ms (median of 5, best %%.4f)  sha1 %%s" %% (ts[2], ts[0], dig))
eng.close()
''' % HERE
reps = sys.argv[1] if len(sys.argv) > 1 else "200"
libs = [os.path.join(PKG, "libraisr.so")] + sorted(glob.glob(os.path.join(PKG, "variants", "libraisr_*.so")))
for rnd in range(2):                       # two rounds: shows run-to-run noise
    for lib in libs:
        out = subprocess.run([sys.executable, "-c", ONE, lib, reps], capture_output=True, text=True)
        print("%-28s %s" % (os.path.basename(lib), out.stdout.strip().splitlines()[-1] if out.returncode == 0 else "FAILED " + out.stderr[-300:]), flush=True)
