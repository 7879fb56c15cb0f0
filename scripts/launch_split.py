"""Per-kernel split of an ncu launch list (--metrics gpu__time_duration.sum --csv): python scripts/launch_split.py file.csv [kernel ...]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
per = collections.defaultdict(list)
for x in csv.DictReader(lines):
    v = float(x["Metric Value"].replace(",", ""))
    v = v / 1e3 if x["Metric Unit"] == "ns" else (v * 1e3 if x["Metric Unit"] == "ms" else v)
    per[re.sub(r"\(.*", "", x["Kernel Name"])].append(v)
tot = sum(sum(v) for v in per.values())
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:28s} n={len(v):4d} total {sum(v):9.1f} us  max {max(v):8.1f}  share {100 * sum(v) / tot:5.1f}%")
print(f"total {tot:.1f} us")
for k in sys.argv[2:]:
    print(k, [round(x) for x in per[k]])
