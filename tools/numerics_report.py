"""The opt-in numerics against the exact fp32 path: bucket agreement and the |dY| histogram per configuration.
  --mode 3 (default)  fp16 filter stage (RAISR_NUMERICS_FP16_FILTER); when the host CPU has AVX512-FP16 also the same comparison for
                      the reference's own asm=avx512fp16 path (compiled reference, oracle/_ref)
  --mode 4            separable fast hash (RAISR_NUMERICS_FAST_HASH)
Run on the GPU box:  python tools/numerics_report.py [--mode 3|4] [--full]  (--full adds the 4K->8K frame)."""
import importlib.util
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import raisr_testlib as T

spec = importlib.util.spec_from_file_location("b", os.path.join(T.PKG_DIR, "binding.py"))
B = importlib.util.module_from_spec(spec)
spec.loader.exec_module(B)

CASES = [
    ("1080p->4K lowres p1 8b", "filters_2x/filters_lowres", 2.0, 8, 1, 1, 1920, 1080),
    ("1080p->4K highres p2m1 8b", "filters_2x/filters_highres", 2.0, 8, 2, 1, 1920, 1080),
    ("1080p->4K denoise p2m2 10b", "filters_2x/filters_denoise", 2.0, 10, 2, 2, 1920, 1080),
    ("720p->1080p 1.5x denoise p2m2 8b", "filters_1.5x/filters_denoise", 1.5, 8, 2, 2, 1280, 720),
]
MODE = int(sys.argv[sys.argv.index("--mode") + 1]) if "--mode" in sys.argv else 3
if "--full" in sys.argv:
    CASES.append(("4K->8K denoise p2m2 10b (configs[3])", "filters_2x/filters_denoise", 2.0, 10, 2, 2, 3840, 2160))


def run(folder, img, ratio, bits, passes, mode, numerics):
    h, w = img.shape
    oW, oH = int(w * ratio), int(h * ratio)
    eng = B.Engine(T.filter_folder(folder), ratio, bits, T.VideoRange, passes, mode, numerics=numerics, keep_hash=True)
    eng.set_res(w, h, oW, oH)
    out = np.zeros((oH, oW), img.dtype)
    assert eng.process_host(img, out) == 0
    hs = []
    for i in range(passes):
        lr = passes == 2 and mode == 2 and i == 0
        hs.append(eng.read_hash(i, w if lr else oW, h if lr else oH))
    eng.close()
    return out, hs


def hist(a, b):
    d = np.abs(a.astype(np.int64) - b.astype(np.int64))
    return {"differ_pct": 100.0 * float((d != 0).mean()), "max": int(d.max()), "mean": float(d.mean()),
            "hist_0_1_2_3_4plus_pct": [100.0 * float((d == k).mean()) for k in range(4)] + [100.0 * float((d >= 4).mean())]}


res = {}
for name, folder, ratio, bits, passes, mode, w, h in CASES:
    img = T.synth_frame(w, h, bits, seed=5150)
    y32, h32 = run(folder, img, ratio, bits, passes, mode, B.NUMERICS_AUTO)
    y16, h16 = run(folder, img, ratio, bits, passes, mode, MODE)
    # pass-1 buckets see the same input in both modes; pass-2 buckets see the (slightly different) pass-1 output
    agree = [100.0 * float((a == b).mean()) for a, b in zip(h32, h16)]
    r = {"bucket_agreement_pct_per_pass": agree, "dY_vs_exact_fp32": hist(y16, y32)}
    if MODE == 3 and T.have_ref() and "avx512_fp16" in open("/proc/cpuinfo").read():
        ry32, _ = T.run_ref_subprocess(folder, img, ratio, bits, threads=os.cpu_count(), asm=T.AVX512, passes=passes, mode=mode)
        ry16, _ = T.run_ref_subprocess(folder, img, ratio, bits, threads=os.cpu_count(), asm=T.AVX512_FP16, passes=passes, mode=mode)
        r["reference_avx512fp16_vs_its_fp32"] = hist(ry16, ry32)
        r["dY_fp16_filter_vs_reference_avx512fp16"] = hist(y16, ry16)
        r["fp32_path_equals_reference_fp32"] = bool(np.array_equal(y32, ry32))
    res[name] = r
    print(name, json.dumps(r))
json.dump(res, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "numerics_report_mode%d.json" % MODE), "w"), indent=1)
