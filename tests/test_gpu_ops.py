"""GPU tier, op level: every hand-written kernel family vs the plain PyTorch fp32 op it replaces,
called through the C-ABI (ctypes) on seeded tensors at shapes taken from the KEEP path."""
import ctypes
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ACT = {"none": 0, "swish": 1, "relu": 2, "lrelu": 3, "gelu": 4, "sigmoid": 5}


def _act(x, name):
    return {"none": lambda t: t, "swish": lambda t: t * torch.sigmoid(t), "relu": F.relu,
            "lrelu": lambda t: F.leaky_relu(t, 0.2), "gelu": F.gelu, "sigmoid": torch.sigmoid}[name](x)


@pytest.fixture(autouse=True)
def _exact_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _rc(lib, rc):
    assert rc == 0, lib.keep_last_error().decode()


def run_conv(lib, x_nchw, w, b, stride=1, pads=(0, 0, 0, 0), up=1, pre=None, pre_act="none", act="none", res=None, use_tc=0):
    """x NCHW cuda fp32 -> NCHW result from our NHWC kernel."""
    n, cin, h, wd = x_nchw.shape
    cout, _, kh, kw = w.shape
    x = x_nchw.permute(0, 2, 3, 1).contiguous()
    ho = (h * up + pads[0] + pads[2] - kh) // stride + 1
    wo = (wd * up + pads[1] + pads[3] - kw) // stride + 1
    out = torch.empty((n, ho, wo, cout), device="cuda", dtype=torch.float32)
    wh = w.detach().cpu().contiguous()
    bh = b.detach().cpu().contiguous() if b is not None else None
    sc = sh = None
    if pre is not None:
        sc, sh = pre[0].contiguous(), pre[1].contiguous()
    r = res.permute(0, 2, 3, 1).contiguous() if res is not None else None
    _rc(lib, lib.keepop_conv2d(use_tc, _p(x), n, h, wd, cin, _p(wh), _p(bh), cout, kh, kw, stride, pads[0], pads[1], pads[2],
                               pads[3], up, _p(sc), _p(sh), ACT[pre_act], ACT[act], _p(r), _p(out), None))
    return out.permute(0, 3, 1, 2).contiguous()


def ref_conv(x, w, b, stride=1, pads=(0, 0, 0, 0), up=1, pre=None, pre_act="none", act="none", res=None):
    if pre is not None:
        x = x * pre[0][:, :, None, None] + pre[1][:, :, None, None]
    x = _act(x, pre_act)
    if up > 1:
        x = F.interpolate(x, scale_factor=float(up), mode="nearest")
    x = F.pad(x, (pads[1], pads[3], pads[0], pads[2]))
    y = _act(F.conv2d(x, w, b, stride=stride), act)
    return y + res if res is not None else y


CONV_CASES = [
    # name, n, cin, h, w, cout, k, stride, pads, up, pre, pre_act, act, res, bias
    ("3x3_64_64", 2, 64, 32, 32, 64, 3, 1, (1, 1, 1, 1), 1, False, "none", "none", False, True),
    ("3x3_gn_swish_res", 1, 64, 40, 24, 128, 3, 1, (1, 1, 1, 1), 1, True, "swish", "none", True, True),
    ("down_s2_asym", 2, 64, 32, 32, 64, 3, 2, (0, 0, 1, 1), 1, False, "none", "none", False, True),
    ("up2_128", 1, 128, 16, 16, 128, 3, 1, (1, 1, 1, 1), 2, False, "none", "none", False, True),
    ("stem_3_64", 2, 3, 64, 64, 64, 3, 1, (1, 1, 1, 1), 1, False, "none", "none", False, True),
    ("head_64_3_gn", 1, 64, 48, 48, 3, 3, 1, (1, 1, 1, 1), 1, True, "none", "none", False, True),
    ("gm_stem_7x7_s2", 2, 3, 64, 64, 64, 7, 2, (3, 3, 3, 3), 1, False, "none", "none", False, False),
    ("gm_s2_p1_96", 2, 64, 32, 32, 96, 3, 2, (1, 1, 1, 1), 1, True, "relu", "none", False, False),
    ("linear_splitk", 1, 512, 256, 1, 1024, 1, 1, (0, 0, 0, 0), 1, False, "none", "gelu", False, True),
    ("c16_512_splitk_res", 1, 512, 16, 16, 512, 3, 1, (1, 1, 1, 1), 1, True, "swish", "none", True, True),
    ("gain_head_1", 3, 256, 16, 16, 1, 1, 1, (0, 0, 0, 0), 1, False, "none", "sigmoid", False, True),
    ("lrelu_256", 1, 256, 32, 32, 256, 3, 1, (1, 1, 1, 1), 1, False, "none", "lrelu", False, True),
    ("ragged_m", 1, 64, 37, 29, 64, 3, 1, (1, 1, 1, 1), 1, False, "none", "none", False, True),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_simt_matches_torch(lib, case):
    name, n, cin, h, w, cout, k, stride, pads, up, pre, pre_act, act, res, bias = case
    g = torch.Generator(device="cpu").manual_seed(hash(name) & 0xFFFF)
    x = torch.randn((n, cin, h, w), generator=g).cuda()
    wt = (torch.randn((cout, cin, k, k), generator=g) / math.sqrt(cin * k * k)).cuda()
    b = torch.randn((cout,), generator=g).cuda() if bias else None
    prep = None
    if pre:
        prep = (1.0 + 0.2 * torch.randn((n, cin), generator=g)).cuda(), (0.2 * torch.randn((n, cin), generator=g)).cuda()
    want0 = ref_conv(x, wt, b, stride, pads, up, prep, pre_act, act, None)
    r = torch.randn(want0.shape, generator=g).cuda() if res else None
    want = want0 + r if res else want0
    got = run_conv(lib, x, wt, b, stride, pads, up, prep, pre_act, act, r)
    assert got.shape == want.shape
    err = float((got - want).abs().max())
    assert err <= 2e-4 * max(1.0, float(want.abs().max())), "%s: max abs err %g" % (name, err)


@pytest.mark.parametrize("n,hw,c,groups,eps,affine", [(2, 64 * 64, 64, 32, 1e-6, True), (1, 16 * 16, 512, 32, 1e-6, True),
                                                      (3, 32 * 32, 96, 96, 1e-5, False), (1, 512 * 512, 64, 32, 1e-6, True),
                                                      (2, 37 * 11, 256, 32, 1e-6, True)])
def test_groupnorm_affine_matches_torch(lib, n, hw, c, groups, eps, affine):
    g = torch.Generator(device="cpu").manual_seed(7)
    x = (torch.randn((n, hw, c), generator=g) * 2.0 + 3.0).cuda()   # non-zero mean: exercises the variance path
    gamma = (1 + 0.1 * torch.randn((c,), generator=g)).cuda() if affine else None
    beta = (0.1 * torch.randn((c,), generator=g)).cuda() if affine else None
    scale = torch.empty((n, c), device="cuda")
    shift = torch.empty((n, c), device="cuda")
    _rc(lib, lib.keepop_groupnorm_affine(_p(x), n, hw, c, groups, eps, _p(gamma), _p(beta), _p(scale), _p(shift), None))
    got = x * scale[:, None, :] + shift[:, None, :]
    want = F.group_norm(x.permute(0, 2, 1).contiguous(), groups, gamma, beta, eps).permute(0, 2, 1)
    assert float((got - want).abs().max()) < 5e-5


@pytest.mark.parametrize("rows,c", [(256, 512), (20 * 256, 256), (8192, 128), (7, 1024)])
def test_layernorm_matches_torch(lib, rows, c):
    g = torch.Generator(device="cpu").manual_seed(3)
    x = (torch.randn((rows, c), generator=g) + 0.5).cuda()
    w = (1 + 0.1 * torch.randn((c,), generator=g)).cuda()
    b = (0.1 * torch.randn((c,), generator=g)).cuda()
    out = torch.empty_like(x)
    _rc(lib, lib.keepop_layernorm(_p(x), rows, c, _p(w), _p(b), 1e-5, _p(out), None))
    assert float((out - F.layer_norm(x, (c,), w, b, 1e-5)).abs().max()) < 2e-5


# (nb, Lq, Lk, heads, dh): AttnBlock, code transformer, sparse-causal, temporal, CFA-16, CFA-32 (small), GMFlow window (small)
@pytest.mark.parametrize("nb,Lq,Lk,heads,dh", [(2, 256, 256, 1, 512), (1, 256, 256, 8, 64), (3, 256, 512, 8, 48),
                                               (64, 20, 20, 8, 48), (1, 256, 256, 4, 256), (1, 1024, 1024, 4, 256),
                                               (4, 1024, 1024, 1, 128), (5, 3, 3, 8, 48)])
def test_attention_matches_torch(lib, nb, Lq, Lk, heads, dh):
    g = torch.Generator(device="cpu").manual_seed(11)
    D = heads * dh
    q = torch.randn((nb, Lq, D), generator=g).cuda()
    k = torch.randn((nb, Lk, D), generator=g).cuda()
    v = torch.randn((nb, Lk, D), generator=g).cuda()
    scale = dh ** -0.5
    out = torch.empty((nb, Lq, D), device="cuda")
    _rc(lib, lib.keepop_attention(_p(q), _p(k), _p(v), nb, Lq, Lk, heads, dh, scale, _p(out), None))
    qh, kh, vh = (t.reshape(nb, -1, heads, dh).transpose(1, 2) for t in (q, k, v))
    want = (torch.softmax(qh @ kh.transpose(-1, -2) * scale, -1) @ vh).transpose(1, 2).reshape(nb, Lq, D)
    assert float((out - want).abs().max()) < 2e-5


def test_flow_warp_matches_grid_sample(lib):
    g = torch.Generator(device="cpu").manual_seed(5)
    n, h, w, c = 2, 96, 80, 3
    img = torch.randn((n, c, h, w), generator=g).cuda()
    flow = (torch.randn((n, 2, h, w), generator=g) * 6.0).cuda()
    flow[0, :, :8] *= 40.0  # far out of bounds -> zero padding
    gy, gx = torch.meshgrid(torch.arange(h, device="cuda", dtype=torch.float32),
                            torch.arange(w, device="cuda", dtype=torch.float32), indexing="ij")
    vx, vy = gx[None] + flow[:, 0], gy[None] + flow[:, 1]
    grid = torch.stack((2 * vx / (w - 1) - 1, 2 * vy / (h - 1) - 1), dim=3)
    want = F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    out = torch.empty((n, h, w, c), device="cuda")
    _rc(lib, lib.keepop_flow_warp(_p(img.permute(0, 2, 3, 1).contiguous()), _p(flow.permute(0, 2, 3, 1).contiguous()), _p(out),
                                  n, h, w, c, None))
    assert float((out.permute(0, 3, 1, 2) - want).abs().max()) < 1e-3


def test_convex_upsample_matches_torch(lib):
    g = torch.Generator(device="cpu").manual_seed(9)
    n, h, w = 2, 16, 12
    mask = torch.randn((n, 576, h, w), generator=g).cuda()
    flow = torch.randn((n, 2, h, w), generator=g).cuda()
    m = torch.softmax(mask.view(n, 1, 9, 8, 8, h, w), dim=2)
    uf = F.unfold(8 * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
    want = torch.sum(m * uf, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(n, 2, 8 * h, 8 * w)
    out = torch.empty((n, 8 * h, 8 * w, 2), device="cuda")
    _rc(lib, lib.keepop_convex_upsample8(_p(mask.permute(0, 2, 3, 1).contiguous()), _p(flow.permute(0, 2, 3, 1).contiguous()),
                                         _p(out), n, h, w, None))
    assert float((out.permute(0, 3, 1, 2) - want).abs().max()) < 1e-4


def test_window_sine_pos_matches_oracle(lib):
    from oracle import keep_oracle
    n, h, w, c = 2, 64, 64, 128
    x = torch.zeros((n, h, w, c), device="cuda")
    _rc(lib, lib.keepop_window_sine_pos(_p(x), n, h, w, c, 2, None))
    pos = keep_oracle._sine_pos(h // 2, w // 2, c // 2, torch.float32).repeat(1, 2, 2)  # (c, h, w)
    assert float((x[0].permute(2, 0, 1).cpu() - pos).abs().max()) < 1e-5
    assert torch.equal(x[0], x[1])


def test_argmax_gather_matches_torch(lib):
    g = torch.Generator(device="cpu").manual_seed(13)
    logits = torch.randn((256, 1024), generator=g).cuda()
    logits[5, 100] = logits[5, 900] = 50.0  # exact tie -> lowest index
    cb = torch.randn((1024, 256), generator=g).cuda()
    idx = torch.empty((256,), dtype=torch.int32, device="cuda")
    quant = torch.empty((256, 256), device="cuda")
    _rc(lib, lib.keepop_argmax_gather(_p(logits), 256, 1024, _p(cb), 256, _p(idx), _p(quant), None))
    want = logits.argmax(dim=1)
    assert int(idx[5]) == 100
    mask = torch.ones(256, dtype=torch.bool, device="cuda"); mask[5] = False
    assert torch.equal(idx.long()[mask], want[mask])
    assert torch.equal(quant, cb[idx.long()])


SMALL_CASES = [
    # name, n, cin, h, w, cout, k, stride, pad, pre
    ("stem_3x3", 2, 3, 96, 80, 64, 3, 1, 1, False),
    ("gm_stem_7x7_s2", 2, 3, 96, 64, 64, 7, 2, 3, False),
    ("head_64_3_gn", 2, 64, 72, 50, 3, 3, 1, 1, True),
    ("head_64_3_full", 1, 64, 512, 512, 3, 3, 1, 1, True),
]


@pytest.mark.parametrize("case", SMALL_CASES, ids=[c[0] for c in SMALL_CASES])
def test_conv_small_kernels_match_torch(lib, case):
    name, n, cin, h, w, cout, k, stride, pad, pre = case
    g = torch.Generator(device="cpu").manual_seed(hash(name) & 0xFFFF)
    x = torch.randn((n, cin, h, w), generator=g).cuda()
    wt = (torch.randn((cout, cin, k, k), generator=g) / math.sqrt(cin * k * k)).cuda()
    b = torch.randn((cout,), generator=g).cuda() if name != "gm_stem_7x7_s2" else None
    prep = None
    if pre:
        prep = (1.0 + 0.2 * torch.randn((n, cin), generator=g)).cuda(), (0.2 * torch.randn((n, cin), generator=g)).cuda()
    pads = (pad, pad, pad, pad)
    want = ref_conv(x, wt, b, stride, pads, 1, prep, "none", "none", None)
    got = run_conv(lib, x, wt, b, stride, pads, 1, prep, "none", "none", None, use_tc=4)
    err = float((got - want).abs().max())
    assert got.shape == want.shape and err <= 2e-4 * max(1.0, float(want.abs().max())), "%s: %g" % (name, err)
