"""Per-segment stall-reason breakdown of an ncu report (segments delimited by BAR.SYNC / SYNCS.TRYWAIT in SASS order).
usage: python profiles/ncu_stalls.py report.ncu-rep"""
import collections, csv, io, subprocess, sys
sass = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
H = rows[1]
ia, ie, isamp, iw = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples"), H.index("L1 Wavefronts Shared")
stalls = [(i, h) for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
seg = 0
segs = collections.OrderedDict()
for r in rows[2:]:
    if len(r) <= ie:
        continue
    d = segs.setdefault(seg, {"n": 0, "samp": 0, "wf": 0, "st": collections.Counter(), "first": r[H.index("Address")]})
    d["n"] += int(r[ie] or 0); d["samp"] += int(r[isamp] or 0); d["wf"] += int(r[iw] or 0)
    for i, h in stalls:
        d["st"][h[6:]] += int(r[i] or 0)
    if "BAR.SYNC" in r[ia] or ("SYNCS" in r[ia] and "TRYWAIT" in r[ia]):
        seg += 1
tot = sum(d["samp"] for d in segs.values())
for k, d in segs.items():
    if d["samp"] < 50:
        continue
    top = ", ".join("%s:%d" % (o, c) for o, c in d["st"].most_common(7))
    print("seg%2d n=%7.1fM samples=%6d (%4.1f%%) wf=%6.1fM | %s" % (k, d["n"] / 1e6, d["samp"], 100.0 * d["samp"] / tot, d["wf"] / 1e6, top))
