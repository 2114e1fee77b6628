"""usage: python profiles/summarize_launches.py gpurun_out/launches.csv  -> per-kernel count / average / share"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value"); ui = H.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg[r[ki][:90]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print("command: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-92s n=%4d avg=%10.1f %s share=%.3f" % (k, len(v), sum(v) / len(v), rows[hdr + 1][ui], sum(v) / tot))
