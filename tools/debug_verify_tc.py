"""Debug: one eager forward with KEEP_DEBUG_VERIFY_TC=1 -- every tcgen05 conv / linear layer is re-run on the exact-fp32
CUDA-core kernel and mismatching layer configurations are printed (stderr).  usage: python tools/debug_verify_tc.py [KEEP|Asian] [T]"""
import os
import sys

os.environ["KEEP_DEBUG_VERIFY_TC"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import keep_b200  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "KEEP"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kn = keep_b200.keep_net
sd = keep_b200.synth.make_state_dict(seed=0, config=config)
net = keep_b200.KeepNetB200(flags=kn.TC3_FLAGS, **(kn.KEEP_ASIAN_CFG if config == "Asian" else kn.KEEP_GENERAL_CFG))
net.load_state_dict(sd, strict=True)
net.eval().to("cuda")
x = keep_b200.synth.make_clip(T, seed=1234, coherent=True).cuda()
out = net(x, need_upscale=False)
torch.cuda.synchronize()
print("config", config, "T", T, "nonfinite outputs per frame:", [int((~torch.isfinite(out[0, i])).sum()) for i in range(T)],
      "absmax", float(out[torch.isfinite(out)].abs().max()))
