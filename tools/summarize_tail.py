"""Summarise the LAST n launches of an ncu launch list (one warmed clip)."""
import collections, csv, sys
path, n = sys.argv[1], int(sys.argv[2])
rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))[-n:]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    k = r["Kernel Name"].split("(")[0][-48:]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("launches", len(rows), "total ms %.2f" % (tot / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%-50s n=%5d %9.1f us %5.1f%% avg %.1f" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
