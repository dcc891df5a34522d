"""Per-role timeline (clock64) of CTA 0 of the fused attention kernel: GMFlow window shape by default."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, keep_b200
lib = keep_b200.keep_net.load_library()
lib.keepop_attn_trace.argtypes = [ctypes.c_void_p]
nimg, W, wsz, dh, shift = 8, 64, 32, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 16
q, k, v = (torch.randn((nimg, W * W, dh), device="cuda") for _ in range(3))
out = torch.empty_like(q)
reg = torch.zeros((4, 1024), dtype=torch.uint8, device="cuda")
P = lambda t: ctypes.c_void_p(t.data_ptr())
run = lambda: lib.keepop_attention_window(P(q), P(k), P(v), nimg, W, wsz, shift, dh, dh ** -0.5, P(reg) if shift else None, P(out), None)
run(); run()
buf = torch.zeros(320, dtype=torch.int64, device="cuda")
lib.keepop_attn_trace(P(buf)); run(); lib.keepop_attn_trace(None)
t = buf.cpu().reshape(8, 40); t0 = int(t[7, 0])
names = ["K stage produced", "V stage produced", "S issued", "PV issued", "passA consumed", "P handed over", "-", "entry/Q/done"]
for i, nm in enumerate(names):
    vals = [int(x) - t0 for x in t[i] if x > 0]
    print("%-18s" % nm, " ".join("%6d" % x for x in vals[:34]))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); [run() for _ in range(5)]; b.record(); torch.cuda.synchronize()
print("avg us per call (incl. sync overhead):", a.elapsed_time(b) * 200)
