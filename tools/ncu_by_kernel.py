"""ncu --csv log of a whole run (one row per launch and metric, or wide) -> per-kernel summary: launches, total / median
duration, DRAM bytes, achieved DRAM GB/s vs the measured peak, tensor-pipe activity.  usage: ncu_by_kernel.py in.csv out_prefix"""
import collections, csv, json, os, re, statistics, sys
src, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(src, errors="replace") if not l.startswith("==")]
rows = list(csv.reader(lines))
hdr = rows[0]
peak = 6544.3
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
per = collections.OrderedDict()
if "Metric Name" in hdr:      # long format: ID, ..., Kernel Name, ..., Metric Name, Metric Unit, Metric Value
    iid, ik, im, iu, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        d = launches.setdefault(r[iid], {"Kernel Name": r[ik]})
        try:
            val = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        u = r[iu]
        if m := re.match(r"(k|K|M|G)?byte", u):
            val *= {"k": 1e3, "K": 1e3, "M": 1e6, "G": 1e9, None: 1.0}[m.group(1)]
        if u in ("us", "usecond"): val *= 1e3
        if u in ("ms", "msecond"): val *= 1e6
        if u in ("s", "second"): val *= 1e9
        d[r[im]] = val
    launch_list = list(launches.values())
else:
    units = rows[1]
    launch_list = []
    for r in rows[2:]:
        d = {"Kernel Name": r[hdr.index("Kernel Name")]}
        for i, k in enumerate(hdr):
            try:
                d[k] = float(r[i].replace(",", ""))
            except ValueError:
                pass
        launch_list.append(d)
def short(n):
    n = re.sub(r"keep::\(anonymous namespace\)::", "", n)
    return re.sub(r"\(.*$", "", n)
for d in launch_list:
    per.setdefault(short(d["Kernel Name"]), []).append(d)
summary = []
for name, ls in per.items():
    t = [l.get("gpu__time_duration.sum", 0.0) for l in ls]            # ns
    rd = sum(l.get("dram__bytes_read.sum", 0.0) for l in ls)
    wr = sum(l.get("dram__bytes_write.sum", 0.0) for l in ls)
    tot = sum(t)
    big = max(ls, key=lambda l: l.get("gpu__time_duration.sum", 0.0))
    summary.append({
        "kernel": name, "launches": len(ls), "total_us": tot / 1e3, "median_us": statistics.median(t) / 1e3, "max_us": max(t) / 1e3,
        "dram_read_mb": rd / 1e6, "dram_write_mb": wr / 1e6, "dram_gbs": (rd + wr) / max(tot, 1.0), "dram_frac_of_peak": (rd + wr) / max(tot, 1.0) / peak,
        "longest_launch": {k: big.get(k) for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                                    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                                                    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                                                    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
                                                    "launch__block_size", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")},
    })
summary.sort(key=lambda s: -s["total_us"])
total = sum(s["total_us"] for s in summary)
json.dump({"source": "ncu --metrics ... --clock-control none, every launch of: python tools/run_clip.py --frames 2 --clips 2 --mode tc3 "
                     "(per-launch times are cold-cache and serialised: shares, not absolutes)", "hbm_peak_gbs": peak,
           "total_us": total, "kernels": summary}, open(out + ".json", "w"), indent=1)
with open(out + ".md", "w") as f:
    f.write("| kernel | launches | total µs | share | median µs | DRAM GB/s (frac of %.0f) | tensor pipe %% (longest launch) | regs | dyn smem |\n|---|---|---|---|---|---|---|---|---|\n" % peak)
    for s in summary:
        b = s["longest_launch"]
        f.write("| `%s` | %d | %.0f | %.1f %% | %.1f | %.0f (%.2f) | %s | %s | %s |\n" % (
            s["kernel"], s["launches"], s["total_us"], 100 * s["total_us"] / max(total, 1e-9), s["median_us"], s["dram_gbs"], s["dram_frac_of_peak"],
            "%.1f" % b["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"] if b.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") is not None else "-",
            "%d" % b["launch__registers_per_thread"] if b.get("launch__registers_per_thread") is not None else "-",
            "%d" % b["launch__shared_mem_per_block_dynamic"] if b.get("launch__shared_mem_per_block_dynamic") is not None else "-"))
print("kernels:", len(summary), "launches:", len(launch_list), "total us: %.0f" % total)
