"""Per-role timeline (clock64) of CTA 0 of the tcgen05 conv kernel for one layer shape."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, keep_b200
from test_gpu_ops import run_conv
lib = keep_b200.keep_net.load_library()
n, cin, h, w, cout = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (1, 64, 512, 512, 64))]
mode = int(sys.argv[6]) if len(sys.argv) > 6 else 1
ksz = int(sys.argv[7]) if len(sys.argv) > 7 else 3          # 3: 3x3 conv, 1: 1x1 / linear
act = sys.argv[8] if len(sys.argv) > 8 else "swish"
pads = (1, 1, 1, 1) if ksz == 3 else (0, 0, 0, 0)
x = torch.randn((n, cin, h, w), device="cuda"); wt = torch.randn((cout, cin, ksz, ksz), device="cuda") / math.sqrt(cin * ksz * ksz); b = torch.randn(cout, device="cuda")
pre = (torch.ones((n, cin), device="cuda"), torch.zeros((n, cin), device="cuda"))
run_conv(lib, x, wt, b, 1, pads, 1, pre if act != "none" else None, act, "none", None, use_tc=mode)
gn = os.environ.get("TC_TRACE_GN") == "1"     # trace the variant whose epilogue also emits GroupNorm statistics
if gn:
    from test_gpu_ops import ACT, _p, _rc
    xin = x.permute(0, 2, 3, 1).contiguous(); outb = torch.empty((n, h, w, cout), device="cuda")
    gam, bet = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
    sc, sh = torch.empty((n, cout), device="cuda"), torch.empty((n, cout), device="cuda")
    wt_h, b_h = wt.cpu().contiguous(), b.cpu().contiguous()
    def run_conv(*a, **k):
        _rc(lib, lib.keepop_conv2d_gn(mode, _p(xin), n, h, w, cin, _p(wt_h), _p(b_h), cout, ksz, ksz, 1, pads[0], pads[1], pads[2], pads[3], 1,
                                      _p(pre[0]) if act != "none" else None, _p(pre[1]) if act != "none" else None, ACT[act], ACT["none"], None,
                                      _p(outb), _p(gam), _p(bet), _p(sc), _p(sh), None))
    run_conv()
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
lib.keepop_tc_trace(ctypes.c_void_p(buf.data_ptr()))
run_conv(lib, x, wt, b, 1, pads, 1, pre if act != "none" else None, act, "none", None, use_tc=mode)
lib.keepop_tc_trace(None)
t = buf.cpu().reshape(16, 16); t0 = int(t[t > 0].min())
names = ["prod:loads_issued", "prod:got_A_EMPTY", "prod:A_FULL_arrive", "mma:got_ACC_EMPTY", "mma:got_A_FULL", "mma:issued_all", "epi:got_ACC_FULL", "epi:done", "load:first_tap", "cta:entry/setup/pdl/done",
         "prod:tile_top*", "prod:located*", "prod:ldg_issued*", "prod:converted*"]   # * = KEEP_TC_TRACE_FINE builds only
for i, nm in enumerate(names):
    print("%-20s" % nm, " ".join("%7d" % (int(v) - t0 if v > 0 else -1) for v in t[i][:10]))
