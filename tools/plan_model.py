"""CPU-only: a floor model of one call from the engine's host-side plan (keep_plan_dump; no GPU needed).

For every conv / linear / GEMM / attention / norm op of the plan: algorithmic FLOPs and bytes, the time floor of the op
= max(FLOPs / tensor ceiling of the engine mode, bytes / HBM) and a fixed per-launch slot (graph replay + PDL: ~2 us,
profiles/r1_experiments.md).  Prints where a clip's time must go under the CURRENT decomposition into kernels -- the part
no kernel tuning can remove -- for one clip and for lockstep groups.
usage: python tools/plan_model.py [--frames 20] [--clips 1] [--config KEEP|Asian]"""
import argparse
import collections
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import keep_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=20)
ap.add_argument("--clips", type=int, default=1)
ap.add_argument("--config", default="KEEP")
ap.add_argument("--slot-us", type=float, default=2.0)
a = ap.parse_args()

pk = {"tflops": 1383.2, "hbm": 6544.3}
mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(mp):
    d = json.load(open(mp))
    pk["tflops"] = float(d.get("bf16_tflops_sustained", d.get("tflops_sustained", pk["tflops"])))
    pk["hbm"] = float(d.get("hbm_gbs", d.get("hbm_copy_gbs", pk["hbm"])))
ceil3 = pk["tflops"] / 3.0          # split-precision mode: three MMAs per MAC

lib = keep_b200.keep_net.load_library()
kn = keep_b200.keep_net
cfg = kn.KEEP_ASIAN_CFG if a.config == "Asian" else {}
net = keep_b200.KeepNetB200(flags=kn.TC3_FLAGS, **(dict(cfg, batch_clips=a.clips) if a.clips > 1 else cfg))
net.load_state_dict(keep_b200.synth.make_state_dict(seed=0, config=a.config), strict=True)
h = net._make_engine(flags=256 | net._flags)
path = os.path.join(tempfile.gettempdir(), "plan_model_%d.txt" % os.getpid())
assert lib.keep_plan_dump(h, a.clips, a.frames, path.encode()) == 0, lib.keep_last_error()
lines = open(path).read().splitlines()
os.unlink(path)

rows = []   # (class, flops, bytes, launches)
for l in lines:
    kind = l.split()[0]
    f = dict(kv.split("=") for kv in l.split()[1:])
    if kind == "conv":
        n, hh, w, c0, c1, co, k, st, up = (int(f[x]) for x in ("n", "h", "w", "c0", "c1", "cout", "k", "stride", "up"))
        M = n * (hh * up // st) * (w * up // st)
        fl = 2.0 * M * co * k * k * (c0 + c1)
        by = 4.0 * (n * hh * w * (c0 + c1) + M * co * (2 if f["res"] == "1" else 1)) + 4.0 * co * k * k * (c0 + c1)
        sk = int(f["splitk"])
        cls = "conv full-grid (>=148 tiles)" if M // 128 >= 148 else ("conv small, split-K" if sk > 1 else "conv small")
        rows.append((cls, fl, by + (8.0 * sk * M * co if sk > 1 else 0.0), 2 if sk > 1 else 1))
    elif kind == "gemm":
        nb, M, K, N = (int(f[x]) for x in ("nb", "M", "K", "N"))
        rows.append(("attention GEMM (tcgen05)", 2.0 * nb * M * K * N, 4.0 * nb * (M * K + 1.5 * N * K + M * N), 2))
    elif kind == "attention":
        nb, Lq, Lk, hd, dh = (int(f[x]) for x in ("nb", "Lq", "Lk", "heads", "dh"))
        rows.append(("attention small (CUDA cores)", 4.0 * nb * hd * Lq * Lk * dh, 4.0 * nb * hd * (2 * Lq * dh + 2 * Lk * dh + 3 * Lq * Lk), 3))
    elif kind == "groupnorm":
        n, hw, c = int(f["n"]), int(f["hw"]), int(f["c"])
        rows.append(("GroupNorm statistics", 0.0, 4.0 * n * hw * c, 1 if hw * c <= (1 << 20) else 2))
    elif kind == "layernorm":
        rows.append(("LayerNorm", 0.0, 8.0 * int(f["rows"]) * int(f["c"]), 1))

agg = collections.OrderedDict()
for cls, fl, by, nl in rows:
    t = max(fl / (ceil3 * 1e12), by / (pk["hbm"] * 1e9)) * 1e3        # ms
    e = agg.setdefault(cls, [0, 0, 0.0, 0.0, 0.0])
    e[0] += 1; e[1] += nl; e[2] += fl / 1e9; e[3] += by / 1e9; e[4] += t
frames = a.frames * a.clips
print("%s config, %d clip(s) x %d frames; ceilings: %.0f TFLOP/s (tensor / 3 passes), %.0f GB/s HBM, %.1f us per launch slot"
      % (a.config, a.clips, a.frames, ceil3, pk["hbm"], a.slot_us))
print("%-32s %6s %8s %10s %8s %10s %10s" % ("class", "ops", "launches", "GFLOP", "GB", "floor ms", "slots ms"))
tot = [0, 0, 0.0, 0.0, 0.0]
for cls, e in agg.items():
    print("%-32s %6d %8d %10.1f %8.2f %10.2f %10.2f" % (cls, e[0], e[1], e[2], e[3], e[4], e[1] * a.slot_us * 1e-3))
    for i in range(5):
        tot[i] += e[i]
slots = tot[1] * a.slot_us * 1e-3
print("%-32s %6d %8d %10.1f %8.2f %10.2f %10.2f" % ("total", tot[0], tot[1], tot[2], tot[3], tot[4], slots))
print("roofline floor  %.1f ms per call -> %.0f frames/s;  launch-slot floor %.1f ms -> %.0f frames/s;  both serialised %.1f ms -> %.0f frames/s"
      % (tot[4], frames / tot[4] * 1e3, slots, frames / slots * 1e3, tot[4] + slots, frames / (tot[4] + slots) * 1e3))
