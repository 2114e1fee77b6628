"""Small two-pass yuv420p frame through the host entry point, for compute-sanitizer runs on the GPU box:
    RAISR_CUDA_NO_MEMOPS=1 compute-sanitizer --tool memcheck python tools/sanitize_probe.py
The sanitizer serialises GPU work, so the in-kernel flag waits of the overlapped copy pipeline must be off (set below)."""
import os
os.environ.setdefault("RAISR_CUDA_NO_MEMOPS", "1")
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, importlib.util
import raisr_testlib as T
spec = importlib.util.spec_from_file_location("b", os.path.join(T.PKG_DIR, "binding.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (300, 200)
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
y = T.synth_frame(w, h, 8, 3); u = T.synth_chroma(w//2, h//2, 8, 4); v = T.synth_chroma(w//2, h//2, 8, 5)
eng = B.Engine(T.filter_folder("filters_2x/filters_highres"), 2.0, 8, 1, passes, 1, device=0, numerics=int(sys.argv[4]) if len(sys.argv) > 4 else B.NUMERICS_AUTO)
eng.set_res(w, h, 2*w, 2*h, w//2, h//2, w, h)
oy = np.zeros((2*h, 2*w), np.uint8); ou = np.zeros((h, w), np.uint8); ov = np.zeros((h, w), np.uint8)
for _ in range(1 if len(sys.argv) > 2 else 2):
    assert eng.process_host(y, oy, u, v, ou, ov) == 0
print("sum", int(oy.sum()), int(ou.sum()))
eng.close()
