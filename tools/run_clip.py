"""Run N clips of T frames through keep_net (for ncu captures: short, no timing)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import keep_b200

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--clips", type=int, default=1)
ap.add_argument("--mode", default="fp32")
ap.add_argument("--graph", action="store_true")
a = ap.parse_args()
kn = keep_b200.keep_net
flags = {"fp32": 0, "tc": kn.FLAG_TCGEN05, "tc3": kn.TC3_FLAGS}[a.mode]
if a.graph:
    flags |= kn.FLAG_CUDA_GRAPH
net = keep_b200.KeepNetB200(flags=flags)
net.load_state_dict(keep_b200.synth.make_state_dict(0))
net.eval().to("cuda")
x = keep_b200.synth.make_clip(a.frames, seed=1234).cuda()
for _ in range(a.clips):
    out = net(x, need_upscale=False)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()), net.launch_count())
