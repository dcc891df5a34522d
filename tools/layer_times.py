"""Per-layer conv/GEMM timing (CUDA events around every launch) grouped by shape."""
import argparse, collections, csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, keep_b200
ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=3); ap.add_argument("--mode", default="tc"); ap.add_argument("--out", default="gpurun_out/layers.csv")
a = ap.parse_args()
kn = keep_b200.keep_net
flags = {"fp32": 0, "tc": kn.FLAG_TCGEN05, "tc3": kn.TC3_FLAGS}[a.mode]
net = keep_b200.KeepNetB200(flags=flags); net.load_state_dict(keep_b200.synth.make_state_dict(0)); net.eval().to("cuda")
x = keep_b200.synth.make_clip(a.frames, seed=1234).cuda()
net(x, need_upscale=False); net(x, need_upscale=False); torch.cuda.synchronize()
net.profile(True); net(x, need_upscale=False); net.profile_dump(a.out); net.profile(False)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in csv.DictReader(open(a.out)):
    k = (r["tag"], r["M"], r["K"], r["N"], r["kh_stride"], r["splitk"], r["bn"])
    agg[k][0] += 1; agg[k][1] += float(r["ms"]); agg[k][2] += float(r["gflop"])
tot = sum(v[1] for v in agg.values())
print("mode %s T=%d total conv/gemm ms %.2f" % (a.mode, a.frames, tot))
print("tag M K N kh*10+stride splitk bn | n ms_total us_each TFLOP/s share")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(" ".join(k), "|", v[0], "%.3f" % v[1], "%.1f" % (1e3 * v[1] / v[0]), "%.1f" % (v[2] / max(v[1], 1e-9)), "%.1f%%" % (100 * v[1] / tot))
