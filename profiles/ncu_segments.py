"""Summarise an ncu report of raisr_pass_kernel: per-barrier-delimited segment instruction counts and key metrics.
usage: python profiles/ncu_segments.py gpurun_out/prof.ncu-rep"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U, V = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "launch__grid_size"]
for h, u, v in zip(H, U, V):
    if h in want:
        print("%-75s %-8s %s" % (h, u, v))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
H = rows[1]
ia, ie, isamp, iw = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples"), H.index("L1 Wavefronts Shared")
seg, segs, tot = 0, collections.OrderedDict(), 0
for r in rows[2:]:
    if len(r) <= ie:
        continue
    src = r[ia]; n = int(r[ie] or 0); s = int(r[isamp] or 0); w = int(r[iw] or 0)
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    d = segs.setdefault(seg, {"n": 0, "samp": 0, "wf": 0, "ops": collections.Counter()})
    d["n"] += n; d["samp"] += s; d["wf"] += w; d["ops"][op.split(".")[0]] += n
    tot += n
    if "BAR.SYNC" in src or ("SYNCS" in src and "TRYWAIT" in src):
        seg += 1
print("total warp-instructions", tot)
for k, d in segs.items():
    if d["n"] == 0:
        continue
    top = ", ".join("%s:%.1f" % (o, c / 1e6) for o, c in d["ops"].most_common(9))
    print("seg%2d n=%7.1fM (%4.1f%%) samples=%5d smem_wavefronts=%6.1fM | %s" % (k, d["n"] / 1e6, 100 * d["n"] / tot, d["samp"], d["wf"] / 1e6, top))
