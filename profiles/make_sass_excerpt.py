"""Writes profiles/r2_sass_hot_loops.txt: excerpts of the SASS of the shipped hot instantiation (cuobjdump of the built object).
usage: python profiles/make_sass_excerpt.py"""
import collections, os, re, subprocess
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OBJ = os.path.join(ROOT, "video-super-resolution-library_b200", "build", "raisr_pipe_u8x.o")
FUN = "_ZN5raisr23raisr_frame_pipe_kernelIhLi4ELi1ELin1ELi4EEEvNS_10PassParamsES1_"
raw = subprocess.run(["cuobjdump", "-sass", "-fun", FUN, OBJ], capture_output=True, text=True).stdout
L = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", l).rstrip() for l in raw.splitlines() if re.match(r"^\s+/\*[0-9a-f]{4,5}\*/", l)]
ops = collections.Counter()
for l in L:
    m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m:
        ops[m.group(1).split(".")[0]] += 1


def find(pat, start=0):
    for i in range(start, len(L)):
        if re.search(pat, L[i]):
            return i
    return -1


out = ["SASS of the shipped hot instantiation raisr_frame_pipe_kernel<uint8_t, 4, 1, -1, 4> (1080p->4K, 8-bit, exact 2x; one pass, exact numerics)",
       "cuobjdump -sass -fun %s video-super-resolution-library_b200/build/raisr_pipe_u8x.o" % FUN,
       "(nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo); regenerate with profiles/make_sass_excerpt.py",
       "%d instructions.  Static opcode census (whole kernel):" % len(L),
       "  " + ", ".join("%s:%d" % kv for kv in ops.most_common(26)),
       "  sm_100-specific: FFMA2 %d + FMUL2 %d (packed fp32 pairs: two IEEE operations per issue slot), UBLKCP %d (cp.async.bulk = TMA bulk copies of the filter slices), UTMALDG %d (tensor-map TMA load of the input window),"
       % (ops["FFMA2"], ops["FMUL2"], ops["UBLKCP"], ops["UTMALDG"]),
       "  SYNCS %d (mbarrier), USETMAXREG %d (setmaxnreg), BAR %d (named barriers, immediate id and thread count), uniform datapath R2UR/LDCU/U* %d"
       % (ops["SYNCS"], ops["USETMAXREG"], ops["BAR"], sum(v for k, v in ops.items() if k[0] == "U" or k in ("R2UR", "LDCU"))), ""]
out.append("---- role split: the three warp roles re-balance the register file (setmaxnreg 56 / 56 / 88) --------------------------------")
out += [L[k] for k, l in enumerate(L) if "USETMAXREG" in l] + [""]
i = find(r"FMUL2")
out.append("---- stage B (chain warps): structure-tensor column chains, one (row, column) position per thread ----------------------------")
out.append("     per patch row: 2 FADD (gradients), 3 x (2 FMUL2 + 3 FFMA2) for the weight-column pairs; the Gaussian weights are uniform operands (URx)")
out += L[i - 6:i + 60] + ["     ...", ""]
j = find(r"BAR\.ARV", i)
out.append("---- bucket warps -> filter warps: bucket tile complete = bar.arrive on a named barrier (the filter side waits in bar.sync) ------")
out += L[j - 2:j + 2] + [""]
t = find(r"UTMALDG")
if t >= 0:
    out.append("---- stage A: low-res window of an interior tile by ONE tensor-map TMA load (cp.async.bulk.tensor.2d -> UTMALDG.2D) ---------------")
    out += L[t - 16:t + 8] + [""]
k = find(r"UBLKCP")
out.append("---- stage D: filter slice of one pixel type by TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx), 4 pieces, then the wait --")
out += L[k - 10:k + 34]
w = find(r"SYNCS\.PHASECHK", k)
out += ["     ..."] + L[w - 1:w + 2] + [""]
d = find(r"LDS\.128", w)
out.append("---- stage D, sliding-window item: a pixel group (8 lanes) walks down a column of same-type pixels, 8 steps per item -----------------")
out.append("     per step: 4 x LDS.128 (the pixel's coefficient units, chain pair (q - 3N) & 7 of its row), the 1-2 tap pairs that enter the")
out.append("     strip window (LDS via LEA.HI / LOP3 of the packed per-lane offsets), 2 x 8 FMUL2/FFMA2 (the chain in both alignments, SEL picks")
out.append("     the lane's own), level 1 of the lane tree (FSEL / SHFL.BFLY / FADD); every 4 steps the folded levels 2-4 with rotated shuffle")
out.append("     sources (SHFL.IDX, sources from c_slide_tbl), the strict range test and the STS of the 4 results")
out += L[d - 40:d + 300] + ["     ..."]
open(os.path.join(ROOT, "profiles", "r2_sass_hot_loops.txt"), "w").write("\n".join(out) + "\n")
print("wrote profiles/r2_sass_hot_loops.txt (%d lines)" % len(out))
