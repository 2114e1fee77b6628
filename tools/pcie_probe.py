import torch, time
d = torch.empty(2160*3840, dtype=torch.uint8, device='cuda'); h = torch.empty(2160*3840, dtype=torch.uint8).pin_memory()
d2 = torch.empty(1080*1920, dtype=torch.uint8, device='cuda'); h2 = torch.empty(1080*1920, dtype=torch.uint8).pin_memory()
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
us = t(lambda: h.copy_(d, non_blocking=True)); print("D2H 8.3MB: %.1f us  %.1f GB/s" % (us, 8.2944e6/us/1e3))
us = t(lambda: d.copy_(h, non_blocking=True)); print("H2D 8.3MB: %.1f us  %.1f GB/s" % (us, 8.2944e6/us/1e3))
us = t(lambda: h2.copy_(d2, non_blocking=True)); print("D2H 2MB: %.1f us  %.1f GB/s" % (us, 2.0736e6/us/1e3))
us = t(lambda: d2.copy_(h2, non_blocking=True)); print("H2D 2MB: %.1f us  %.1f GB/s" % (us, 2.0736e6/us/1e3))
s2 = torch.cuda.Stream()
def both():
    h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
us = t(both); print("D2H 8.3MB + H2D 2MB concurrently: %.1f us" % us)
