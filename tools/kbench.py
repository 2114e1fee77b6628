"""Quick kernel timing on the GPU box: 1080p->4K luma pass, pipe kernel vs tile kernel (bit-identity check + CUDA-event timing).
usage: python tools/kbench.py [reps]"""
import os, sys, importlib.util
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import raisr_testlib as T
spec = importlib.util.spec_from_file_location("b", os.path.join(T.PKG_DIR, "binding.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
if os.environ.get("KBENCH_LIB"):
    B.LIB_PATH = os.path.abspath(os.environ["KBENCH_LIB"])        # experiment builds
NB = 12
ys = [torch.from_numpy(T.synth_frame(1920, 1080, 8, 1234 + i)).cuda() for i in range(NB)]
outs = [torch.empty((2160, 3840), dtype=torch.uint8, device="cuda") for _ in range(NB)]
res = {}
for kern in (("pipe",) if os.environ.get("KBENCH_LIB") else ("tile", "pipe")):
    os.environ["RAISR_CUDA_KERNEL"] = kern
    eng = B.Engine(T.filter_folder("filters_2x/filters_lowres"), 2.0, 8, 1, 1, 1, device=0, numerics=int(os.environ.get("NUM", "1")))
    eng.set_res(1920, 1080, 3840, 2160)
    def run(i):
        eng.process_device_rows(ys[i % NB].data_ptr(), ys[i % NB].stride(0), outs[i % NB].data_ptr(), outs[i % NB].stride(0), 0, 2160, 2, None)
    for i in range(NB):
        run(i)
    torch.cuda.synchronize()
    res[kern] = [o.clone() for o in outs[:3]]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for i in range(reps):
            run(i)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    print("%s kernel: %.4f ms per 4K frame" % (kern, best))
    eng.close()
same = all(torch.equal(a, b) for a, b in zip(res.get("tile", res["pipe"]), res["pipe"]))
print("pipe == tile:", same)
sys.exit(0 if same else 1)
