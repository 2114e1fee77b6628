import sys, os, ctypes, importlib.util
sys.path.insert(0, 'tests')
import numpy as np, torch
import raisr_testlib as T
spec = importlib.util.spec_from_file_location("b", os.path.join(T.PKG_DIR, "binding.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
eng = B.Engine(T.filter_folder("filters_2x/filters_lowres"), 2.0, 8, 1, 1, 1, device=0, numerics=int(os.environ.get("NUM","0")))
eng.set_res(1920,1080,3840,2160)
y = torch.from_numpy(T.synth_frame(1920,1080,8,1234)).cuda()
o = torch.empty((2160,3840),dtype=torch.uint8,device='cuda')
for i in range(4):
    eng.process_device_rows(y.data_ptr(), y.stride(0), o.data_ptr(), o.stride(0), 0, 2160, 2, None)
torch.cuda.synchronize()
print("ok")
