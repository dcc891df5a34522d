"""Kernel-start timeline of one graph-replayed clip: every kernel stamps %globaltimer when its grid may start; the host-side
launch log of an eager run of the same clip supplies the names.  Start-to-start deltas = the wall-clock slot each kernel
occupies on the stream INCLUDING launch gaps -- what ncu's per-kernel durations cannot show.  Single stream only: run with
KEEP_DEBUG_SKIP_FLOW=1 (zero flows, no GMFlow) or KEEP_NO_SIDE=1 (GMFlow inline on the main stream).  Usage: python tools/timeline.py --frames 4 --out gpurun_out/tl"""
import argparse, collections, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, keep_b200

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=4); ap.add_argument("--mode", default="tc3")
ap.add_argument("--out", default="gpurun_out/timeline")
ap.add_argument("--clips", type=int, default=1, help="clips per call; with --batch-clips N they walk the recurrence in lockstep groups of N")
ap.add_argument("--batch-clips", type=int, default=1)
a = ap.parse_args()
assert os.environ.get("KEEP_DEBUG_SKIP_FLOW") or os.environ.get("KEEP_NO_SIDE"), "set KEEP_DEBUG_SKIP_FLOW=1 or KEEP_NO_SIDE=1 (the timeline needs a single stream)"
kn = keep_b200.keep_net
lib = kn.load_library()
flags = {"fp32": 0, "tc": kn.FLAG_TCGEN05, "tc3": kn.TC3_FLAGS}[a.mode]
sd = keep_b200.synth.make_state_dict(0)
x = torch.cat([keep_b200.synth.make_clip(a.frames, seed=1234 + 100 * i) for i in range(a.clips)], 0).cuda()
bk = dict(batch_clips=min(a.clips, a.batch_clips)) if a.batch_clips > 1 else {}

# 1. names: eager engine, launch log on for the third clip (weights packed, arenas warm)
eager = keep_b200.KeepNetB200(flags=flags, **bk); eager.load_state_dict(sd); eager.eval().to("cuda")
eager(x, need_upscale=False); eager(x, need_upscale=False); torch.cuda.synchronize()
lib.keepop_launch_log(1)
eager(x, need_upscale=False); torch.cuda.synchronize()
n_names = lib.keepop_launch_log_dump((a.out + "_names.txt").encode())
lib.keepop_launch_log(0)
names = [l.rstrip("\n").split("\t") for l in open(a.out + "_names.txt")]
del eager

# 2. stamps: graph engine, replay once with the stamp buffer set
net = keep_b200.KeepNetB200(flags=flags | kn.FLAG_CUDA_GRAPH, **bk); net.load_state_dict(sd); net.eval().to("cuda")
for _ in range(4): net(x, need_upscale=False)
torch.cuda.synchronize()
buf = torch.zeros(1 + 65536, dtype=torch.int64, device="cuda")
lib.keepop_kernel_stamps.argtypes = [ctypes.c_void_p]
lib.keepop_kernel_stamps(ctypes.c_void_p(buf.data_ptr()))
net(x, need_upscale=False); torch.cuda.synchronize()
lib.keepop_kernel_stamps(None)
b = buf.cpu()
cnt = int(b[0]); st = b[1:1 + cnt].tolist()
print("launch log: %d kernels; stamps: %d" % (n_names, cnt))
st.sort()   # arrival order == start order on a single stream (atomic slot order can differ by a few ns)
n = min(cnt, len(names))
if cnt != len(names): print("WARNING: counts differ; aligning the first %d" % n)
dt = [(st[i + 1] - st[i]) / 1e3 for i in range(n - 1)] + [0.0]
with open(a.out + "_slots.csv", "w") as f:
    f.write("idx,kernel,grid,block,start_us,slot_us\n")
    for i in range(n): f.write("%d,%s,%s,%s,%.3f,%.3f\n" % (i, names[i][0].replace(",", ";"), names[i][1].replace(",", "x"), names[i][2], (st[i] - st[0]) / 1e3, dt[i]))
total = (st[n - 1] - st[0]) / 1e3
print("call: %.1f us from first to last kernel start (T=%d, %d clip(s), lockstep groups of %d)" % (total, a.frames, a.clips, max(1, a.batch_clips)))
def short(nm):
    nm = nm.split("(")[0]
    return nm[nm.rfind("::") + 2:] if "::" in nm else nm
# last frame = after the second-to-last nhwc_to_nchw (a lockstep group writes one per clip back to back: step over the group)
ends = [i for i in range(n) if "nhwc_to_nchw" in names[i][0]]
g = max(1, min(a.clips, a.batch_clips))
lo = ends[-1 - g] + 1 if len(ends) >= 1 + g else 0
hi = ends[-1] + 1 if ends else n
for title, (p, q) in (("whole clip", (0, n)), ("last frame", (lo, hi))):
    agg = collections.defaultdict(lambda: [0, 0.0])
    for i in range(p, q):
        k = short(names[i][0]); agg[k][0] += 1; agg[k][1] += dt[i]
    tot = sum(v[1] for v in agg.values())
    print("\n== %s: %d kernels, %.1f us" % (title, q - p, tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        print("%-44s n=%5d %10.1f us  avg %7.2f  %5.1f%%" % (k[:44], v[0], v[1], v[1] / v[0], 100 * v[1] / max(tot, 1e-9)))
# conv_tc slots in the last frame by grid size bucket
print("\n== last frame, conv_tc_kernel slots by grid")
agg = collections.defaultdict(lambda: [0, 0.0])
for i in range(lo, hi):
    if "conv_tc_kernel" in names[i][0]:
        g = int(names[i][1].split(",")[0]); k = "grid<=32" if g <= 32 else ("grid<=96" if g <= 96 else ("grid<148" if g < 148 else "grid=148"))
        agg[k][0] += 1; agg[k][1] += dt[i]
for k, v in sorted(agg.items()): print("%-10s n=%4d %9.1f us avg %6.2f" % (k, v[0], v[1], v[1] / v[0]))
