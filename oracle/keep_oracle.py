"""TEST INFRASTRUCTURE ONLY — CPU/torch fp32 restatement of the reference KEEP forward.

This file is the *checker* for the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it; the product path
(`comfyui-keep_b200/`) never does and fails loudly when its CUDA library is missing.

It restates — functionally, over a plain state dict, with no nn.Module — what the reference
computes in `keep_net(x, need_upscale=False)`:

  KEEP.forward                    modules/deps/wm_basicsr/archs/keep_arch.py:1008-1145
  KEEP.get_flow / FlowGenerator   keep_arch.py:976-986 ; archs/gmflow_arch.py:40-66
  GMFlow.forward (1 scale, swin splits=2, global corr + global propagation)
                                  archs/gmflow/gmflow/gmflow.py:92-170, upsample_flow :67-90
  CNNEncoder / ResidualBlock      gmflow/backbone.py:8-117
  FeatureTransformer              gmflow/transformer.py:8-105,108-185,244-322
  FeatureFlowAttention            gmflow/transformer.py:343-374
  global_correlation_softmax      gmflow/matching.py:7-36
  sine position / window split    gmflow/position.py:26-46 ; gmflow/utils.py:5-86
  Encoder / Generator / ResBlock / AttnBlock / Down / Up / codebook gather
                                  archs/vqgan_arch.py:16-22,78-91,129-343
  KalmanFilter (gain, predict, update), BasicTransformerBlock, SparseCausalAttention
                                  keep_arch.py:640-821 ; CrossAttention :137-241
  TransformerSALayer              keep_arch.py:423-439
  Fuse_sft_block (CFT)            keep_arch.py:465-472
  CrossFrameFusionLayer (CFA)     keep_arch.py:519-541
  flow_warp                       archs/arch_util.py:113-144
  diffusers FeedForward (GEGLU)   third-party, absent from the tree; restated from the 0.11-era
                                  definition (see oracle/ref_loader.py) — "parity unpinned" there.

Pinning: `oracle/make_golden.py` runs the *real* reference (imported from /root/reference) and
this restatement on the same seeded weights/inputs in the build container and checks they agree
to fp32 round-off (result recorded in tests/golden/pin_report.json); the reference's outputs are
committed as fixtures under tests/golden/.  The reference itself ships no golden vectors / KATs
for this path, so the live reference run is the only pin.
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------


class _W:
    """Prefix view over a flat state dict."""

    def __init__(self, sd, prefix=""):
        self.sd, self.p = sd, prefix

    def __call__(self, name):
        return self.sd[self.p + name]

    def has(self, name):
        return (self.p + name) in self.sd

    def sub(self, name):
        return _W(self.sd, self.p + name + ".")


def _gn(x, w, name, swish):
    y = F.group_norm(x, 32, w(name + ".weight"), w(name + ".bias"), eps=1e-6)
    return y * torch.sigmoid(y) if swish else y


def _conv(x, w, name, stride=1, padding=0):
    b = w(name + ".bias") if w.has(name + ".bias") else None
    return F.conv2d(x, w(name + ".weight"), b, stride=stride, padding=padding)


def _lin(x, w, name):
    b = w(name + ".bias") if w.has(name + ".bias") else None
    return F.linear(x, w(name + ".weight"), b)


def _ln(x, w, name):
    wt = w(name + ".weight")
    return F.layer_norm(x, (wt.shape[0],), wt, w(name + ".bias"), eps=1e-5)


# ----------------------------------------------------------------------------------------------
# VQGAN blocks (vqgan_arch.py)
# ----------------------------------------------------------------------------------------------

def res_block(x, w):
    """vqgan_arch.py:170-181"""
    h = _conv(_gn(x, w, "norm1", True), w, "conv1", padding=1)
    h = _conv(_gn(h, w, "norm2", True), w, "conv2", padding=1)
    skip = _conv(x, w, "conv_out") if w.has("conv_out.weight") else x
    return h + skip


def attn_block(x, w):
    """vqgan_arch.py:219-243 — single head over H*W tokens, d = C, scale C^-0.5."""
    n, c, hh, ww = x.shape
    hn = _gn(x, w, "norm", False)
    q = _conv(hn, w, "q").reshape(n, c, hh * ww).transpose(1, 2)   # (n, L, c)
    k = _conv(hn, w, "k").reshape(n, c, hh * ww)                   # (n, c, L)
    v = _conv(hn, w, "v").reshape(n, c, hh * ww).transpose(1, 2)   # (n, L, c)
    p = torch.softmax(torch.bmm(q, k) * (c ** -0.5), dim=2)
    o = torch.bmm(p, v).transpose(1, 2).reshape(n, c, hh, ww)
    return x + _conv(o, w, "proj_out")


# block programme of Encoder (vqgan_arch.py:246-292) for nf=64, ch_mult [1,2,2,4,4,8], 2 res blocks,
# attention at 16: kinds in order of `blocks`
ENCODER_PROGRAM = (
    ["conv_in"] + ["res", "res", "down"] * 5 + ["res", "attn", "res", "attn"] + ["res", "attn", "res"]
    + ["norm_out", "conv_out"])
# Generator (vqgan_arch.py:295-343)
GENERATOR_PROGRAM = (
    ["conv_in", "res", "attn", "res"] + ["res", "attn", "res", "attn", "up"] + ["res", "res", "up"] * 4
    + ["res", "res"] + ["norm_out", "conv_out"])
assert len(ENCODER_PROGRAM) == 25 and len(GENERATOR_PROGRAM) == 25


def vq_block(kind, x, w):
    if kind in ("conv_in", "conv_out"):
        return F.conv2d(x, w("weight"), w("bias"), padding=1)
    if kind == "res":
        return res_block(x, w)
    if kind == "attn":
        return attn_block(x, w)
    if kind == "down":  # vqgan_arch.py:135-139: pad right/bottom by one, 3x3 stride 2
        return _conv(F.pad(x, (0, 1, 0, 1)), w, "conv", stride=2)
    if kind == "up":    # vqgan_arch.py:148-152
        return _conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, "conv", padding=1)
    if kind == "norm_out":
        return F.group_norm(x, 32, w("weight"), w("bias"), eps=1e-6)
    raise ValueError(kind)


def encoder_forward(x, w, taps=()):
    """Encoder over NCHW batch; returns (z, {block_index: activation})."""
    out = {}
    for i, kind in enumerate(ENCODER_PROGRAM):
        x = vq_block(kind, x, w.sub("blocks.%d" % i))
        if i in taps:
            out[i] = x
    return x, out


# ----------------------------------------------------------------------------------------------
# attention helpers (keep_arch.py CrossAttention)
# ----------------------------------------------------------------------------------------------

def _mh_attention(q, k, v, heads, scale):
    """q (B, Lq, h*d), k/v (B, Lk, h*d) -> (B, Lq, h*d)   keep_arch.py:103-115,200-241"""
    B, Lq, D = q.shape
    d = D // heads
    qh = q.reshape(B, Lq, heads, d).transpose(1, 2)
    kh = k.reshape(B, -1, heads, d).transpose(1, 2)
    vh = v.reshape(B, -1, heads, d).transpose(1, 2)
    p = torch.softmax(torch.matmul(qh, kh.transpose(-1, -2)) * scale, dim=-1)
    return torch.matmul(p, vh).transpose(1, 2).reshape(B, Lq, D)


def _geglu_ff(x, w):
    """diffusers FeedForward(geglu): net.0.proj -> h * gelu(g) -> net.2"""
    h, g = _lin(x, w, "net.0.proj").chunk(2, dim=-1)
    return _lin(h * F.gelu(g), w, "net.2")


def kalman_block(h, w, T, heads=8):
    """BasicTransformerBlock.forward, keep_arch.py:640-682.  h: (b*T, L, C)."""
    BT, L, C = h.shape
    b = BT // T
    # sparse-causal attention: keys/values = [frame 0 || frame i-1]  (:704-716)
    hn = _ln(h, w, "norm1")
    a = w.sub("attn1")
    q = _lin(hn, a, "to_q")
    k = _lin(hn, a, "to_k").reshape(b, T, L, -1)
    v = _lin(hn, a, "to_v").reshape(b, T, L, -1)
    prev = [max(i - 1, 0) for i in range(T)]
    k = torch.cat([k[:, [0] * T], k[:, prev]], dim=2).reshape(BT, 2 * L, -1)
    v = torch.cat([v[:, [0] * T], v[:, prev]], dim=2).reshape(BT, 2 * L, -1)
    dh = q.shape[-1] // heads
    h = _lin(_mh_attention(q, k, v, heads, dh ** -0.5), a, "to_out.0") + h
    # feed-forward
    h = _geglu_ff(_ln(h, w, "norm3"), w.sub("ff")) + h
    # temporal attention over frames for each spatial token (:672-680)
    ht = h.reshape(b, T, L, C).permute(0, 2, 1, 3).reshape(b * L, T, C)
    hn = _ln(ht, w, "norm_temp")
    a = w.sub("attn_temp")
    o = _mh_attention(_lin(hn, a, "to_q"), _lin(hn, a, "to_k"), _lin(hn, a, "to_v"), heads, dh ** -0.5)
    ht = _lin(o, a, "to_out.0") + ht
    return ht.reshape(b, L, T, C).permute(0, 2, 1, 3).reshape(BT, L, C)


def kalman_gains(z_codes, w, n_layers=3):
    """KalmanFilter.calc_gain, keep_arch.py:801-821.  z_codes (b,T,256,16,16) -> (b,T,1,16,16)."""
    b, T, C, hh, ww = z_codes.shape
    h = z_codes.reshape(b * T, C, hh * ww).transpose(1, 2)
    for i in range(n_layers):
        h = kalman_block(h, w.sub("uncertainty_estimator.%d" % i), T)
    x = h.transpose(1, 2).reshape(b * T, C, hh, ww)
    g = w.sub("kalman_gain_calculator")
    for i in range(3):
        x = res_block(x, g.sub(str(i)))
    x = torch.sigmoid(F.conv2d(x, g("3.weight"), g("3.bias")))
    return x.reshape(b, T, 1, hh, ww)


def flow_warp(x, flow_nchw):
    """arch_util.py:113-144: bilinear grid_sample, zero padding, align_corners=True.
    flow (n,2,h,w) in pixels, channel 0 = x displacement."""
    n, _, h, w = x.shape
    gy, gx = torch.meshgrid(torch.arange(h, dtype=x.dtype), torch.arange(w, dtype=x.dtype), indexing="ij")
    vx = gx[None] + flow_nchw[:, 0]
    vy = gy[None] + flow_nchw[:, 1]
    grid = torch.stack((2.0 * vx / max(w - 1, 1) - 1.0, 2.0 * vy / max(h - 1, 1) - 1.0), dim=3)
    return F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


def sa_layer(t, pos, w, heads=8):
    """TransformerSALayer.forward, keep_arch.py:423-439.  t: (L, b, E) sequence-first."""
    L, b, E = t.shape
    tn = _ln(t, w, "norm1")
    qk_in = tn + pos
    wi, bi = w("self_attn.in_proj_weight"), w("self_attn.in_proj_bias")
    q = F.linear(qk_in, wi[:E], bi[:E])
    k = F.linear(qk_in, wi[E:2 * E], bi[E:2 * E])
    v = F.linear(tn, wi[2 * E:], bi[2 * E:])
    o = _mh_attention(q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1), heads, (E // heads) ** -0.5)
    t = t + _lin(o.transpose(0, 1), w, "self_attn.out_proj")
    tn = _ln(t, w, "norm2")
    return t + _lin(F.gelu(_lin(tn, w, "linear1")), w, "linear2")


def code_logits(z_hat, W):
    """feat_emb -> 9 SA layers -> LN -> Linear(512->1024): keep_arch.py:1073-1081. returns (b,256,1024)"""
    b = z_hat.shape[0]
    pos = W("position_emb").unsqueeze(1).repeat(1, b, 1)
    t = _lin(z_hat.flatten(2).permute(2, 0, 1), W, "feat_emb")
    for i in range(9):
        t = sa_layer(t, pos, W.sub("ft_layers.%d" % i))
    t = _ln(t, W, "idx_pred_layer.0")
    return F.linear(t, W("idx_pred_layer.1.weight")).permute(1, 0, 2)


def cft_block(enc_feat, dec_feat, w, cond=1.0):
    """Fuse_sft_block.forward, keep_arch.py:465-472"""
    f = res_block(torch.cat([enc_feat, dec_feat], dim=1), w.sub("encode_enc"))

    def branch(name):
        y = F.leaky_relu(_conv(f, w, name + ".0", padding=1), 0.2)
        return _conv(y, w, name + ".2", padding=1)

    return dec_feat + cond * (dec_feat * branch("scale") + branch("shift"))


def cfa_block(cur, prev, w, heads=4):
    """CrossFrameFusionLayer.forward (residual=True), keep_arch.py:519-541"""
    n, c, hh, ww = cur.shape
    x = cur.flatten(2).transpose(1, 2)
    p = prev.flatten(2).transpose(1, 2)
    a = w.sub("attn")
    q, k, v = _lin(x, a, "to_q"), _lin(p, a, "to_k"), _lin(p, a, "to_v")
    dh = q.shape[-1] // heads
    y = _lin(_mh_attention(q, k, v, heads, dh ** -0.5), a, "to_out.0")
    x = _ln(y, w, "norm1") + x
    x = _ln(_geglu_ff(x, w.sub("ff")), w, "norm2") + x
    return x.transpose(1, 2).reshape(n, c, hh, ww)


# ----------------------------------------------------------------------------------------------
# GMFlow (gmflow/*.py), the configuration KEEP uses: 1 scale, attn_splits=2, global correlation,
# global flow propagation, convex x8 upsampling.
# ----------------------------------------------------------------------------------------------

def _inorm(x):
    return F.instance_norm(x, eps=1e-5)


def _gm_resblock(x, w, stride):
    """backbone.py:8-36 (InstanceNorm, affine-free)"""
    y = F.relu(_inorm(F.conv2d(x, w("conv1.weight"), None, stride=stride, padding=1)))
    y = F.relu(_inorm(F.conv2d(y, w("conv2.weight"), None, padding=1)))
    if w.has("downsample.0.weight"):
        x = _inorm(F.conv2d(x, w("downsample.0.weight"), w("downsample.0.bias"), stride=stride))
    return F.relu(x + y)


def gm_backbone(img, w):
    """CNNEncoder.forward, backbone.py:101-117 -> (n,128,H/8,W/8)"""
    x = F.relu(_inorm(F.conv2d(img, w("conv1.weight"), None, stride=2, padding=3)))
    for name, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
        x = _gm_resblock(x, w.sub(name + ".0"), stride)
        x = _gm_resblock(x, w.sub(name + ".1"), 1)
    return F.conv2d(x, w("conv2.weight"), w("conv2.bias"))


def _sine_pos(hh, ww, feats, dtype):
    """position.py:26-46 with normalize=True, scale 2*pi, temperature 1e4 -> (2*feats, hh, ww)"""
    y = torch.arange(1, hh + 1, dtype=torch.float32)[:, None].expand(hh, ww)
    x = torch.arange(1, ww + 1, dtype=torch.float32)[None, :].expand(hh, ww)
    eps, scale = 1e-6, 2 * math.pi
    y = y / (hh + eps) * scale
    x = x / (ww + eps) * scale
    i = torch.arange(feats, dtype=torch.float32)
    dim_t = 10000.0 ** (2 * torch.div(i, 2, rounding_mode="floor") / feats)
    px = x[:, :, None] / dim_t
    py = y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).permute(2, 0, 1).to(dtype)


def _split_windows(x, k):
    """(B,H,W,C) -> (B*k*k, H/k, W/k, C)   utils.py:5-30 channel_last"""
    B, H, W, C = x.shape
    return x.reshape(B, k, H // k, k, W // k, C).permute(0, 1, 3, 2, 4, 5).reshape(B * k * k, H // k, W // k, C)


def _merge_windows(x, k):
    Bk, h, w, C = x.shape
    B = Bk // (k * k)
    return x.reshape(B, k, k, h, w, C).permute(0, 1, 3, 2, 4, 5).reshape(B, k * h, k * w, C)


def _shift_mask(H, W, wh, ww_, sh, sw):
    """transformer.py:19-43 -> (k*k, wh*ww, wh*ww) with entries in {0, -100}"""
    m = torch.zeros((1, H, W, 1))
    cnt = 0
    for hs in (slice(0, -wh), slice(-wh, -sh), slice(-sh, None)):
        for ws in (slice(0, -ww_), slice(-ww_, -sw), slice(-sw, None)):
            m[:, hs, ws, :] = cnt
            cnt += 1
    mw = _split_windows(m, W // ww_).reshape(-1, wh * ww_)
    d = mw.unsqueeze(1) - mw.unsqueeze(2)
    return torch.where(d != 0, torch.full_like(d, -100.0), torch.zeros_like(d))


def _swin_attention(q, k, v, H, W, splits, shift, mask):
    """transformer.py:46-105: single head, window split `splits`, optional half-window shift."""
    B, L, C = q.shape
    wh, ww_ = H // splits, W // splits
    q, k, v = (t.reshape(B, H, W, C) for t in (q, k, v))
    if shift:
        q, k, v = (torch.roll(t, shifts=(-(wh // 2), -(ww_ // 2)), dims=(1, 2)) for t in (q, k, v))
    q, k, v = (_split_windows(t, splits).reshape(B * splits * splits, wh * ww_, C) for t in (q, k, v))
    s = torch.matmul(q, k.transpose(1, 2)) / (C ** 0.5)
    if shift:
        s = s + mask.repeat(B, 1, 1)
    o = torch.matmul(torch.softmax(s, dim=-1), v)
    o = _merge_windows(o.reshape(B * splits * splits, wh, ww_, C), splits)
    if shift:
        o = torch.roll(o, shifts=(wh // 2, ww_ // 2), dims=(1, 2))
    return o.reshape(B, L, C)


def _gm_layer(src, tgt, w, H, W, splits, shift, mask, ffn):
    """TransformerLayer.forward, transformer.py:147-185"""
    q, k, v = _lin(src, w, "q_proj"), _lin(tgt, w, "k_proj"), _lin(tgt, w, "v_proj")
    m = _ln(_lin(_swin_attention(q, k, v, H, W, splits, shift, mask), w, "merge"), w, "norm1")
    if ffn:
        m = _lin(F.gelu(_lin(torch.cat([src, m], dim=-1), w, "mlp.0")), w, "mlp.2")
        m = _ln(m, w, "norm2")
    return src + m


def gm_transformer(f0, f1, w, splits=2):
    """FeatureTransformer.forward, transformer.py:273-322.  f0,f1 (B,C,H,W) -> same."""
    B, C, H, W = f0.shape
    mask = _shift_mask(H, W, H // splits, W // splits, H // splits // 2, W // splits // 2).to(f0.dtype)
    a = f0.flatten(2).transpose(1, 2)
    b_ = f1.flatten(2).transpose(1, 2)
    c0 = torch.cat((a, b_), dim=0)
    c1 = torch.cat((b_, a), dim=0)
    for i in range(6):
        lw = w.sub("layers.%d" % i)
        shift = (i % 2 == 1)
        c0 = _gm_layer(c0, c0, lw.sub("self_attn"), H, W, splits, shift, mask, ffn=False)
        c0 = _gm_layer(c0, c1, lw.sub("cross_attn_ffn"), H, W, splits, shift, mask, ffn=True)
        c1 = torch.cat((c0[B:], c0[:B]), dim=0)
    o0, o1 = c0[:B], c0[B:]
    return (o0.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous(),
            o1.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous())


def _coords(H, W):
    y, x = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    return torch.stack((x, y), dim=0).float()  # (2,H,W): x first


def gmflow_forward(img0, img1, w):
    """FlowGenerator.forward + GMFlow.forward.  img in [-1,1], (n,3,H,W) -> flow (n,2,H,W)."""
    n, _, H, W = img0.shape
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1).to(img0.dtype)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1).to(img0.dtype)
    i0 = ((img0 + 1) / 2 * 255 / 255. - mean) / std
    i1 = ((img1 + 1) / 2 * 255 / 255. - mean) / std
    feat = gm_backbone(torch.cat((i0, i1), dim=0), w.sub("backbone"))
    f0, f1 = feat[:n], feat[n:]
    C, h, ww = f0.shape[1], f0.shape[2], f0.shape[3]
    # position added per 2x2 window (utils.py:66-86)
    pos = _sine_pos(h // 2, ww // 2, C // 2, f0.dtype).repeat(1, 2, 2)[None]
    f0, f1 = f0 + pos, f1 + pos
    f0, f1 = gm_transformer(f0, f1, w.sub("transformer"), splits=2)
    # global correlation softmax (matching.py:7-36)
    a = f0.flatten(2).transpose(1, 2)
    corr = torch.matmul(a, f1.flatten(2)) / (C ** 0.5)
    grid = _coords(h, ww).reshape(2, -1).t()[None].expand(n, -1, -1).to(f0.dtype)
    flow = torch.matmul(torch.softmax(corr, dim=-1), grid) - grid          # (n, L, 2)
    # flow propagation by feature self-similarity (transformer.py:343-374; note k = k_proj(q_proj(x)))
    fw = w.sub("feature_flow_attn")
    q = _lin(a, fw, "q_proj")
    k = _lin(q, fw, "k_proj")
    flow = torch.matmul(torch.softmax(torch.matmul(q, k.transpose(1, 2)) / (C ** 0.5), dim=-1), flow)
    flow = flow.transpose(1, 2).reshape(n, 2, h, ww)
    # convex x8 upsampling (gmflow.py:74-88)
    m = F.relu(F.conv2d(torch.cat((flow, f0), dim=1), w("upsampler.0.weight"), w("upsampler.0.bias"), padding=1))
    m = F.conv2d(m, w("upsampler.2.weight"), w("upsampler.2.bias"))
    m = torch.softmax(m.reshape(n, 1, 9, 8, 8, h, ww), dim=2)
    uf = F.unfold(8 * flow, [3, 3], padding=1).reshape(n, 2, 9, 1, 1, h, ww)
    up = torch.sum(m * uf, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(n, 2, 8 * h, 8 * ww)
    return up


# ----------------------------------------------------------------------------------------------
# nearest-neighbour quantiser (the VQGAN's own `quantize` module; KEEP inference uses get_codebook_feat instead)
# ----------------------------------------------------------------------------------------------

@torch.no_grad()
def vq_nearest(z, codebook):
    """VectorQuantizer.forward, inference values only (vqgan_arch.py:37-76).
    z (n, C, h, w), codebook (K, C) -> z_q (n, C, h, w) [the forward value of z + (z_q - z).detach(), :61],
    indices (n*h*w,) int64 [torch.argmin, :47], d (n*h*w, K) the distance matrix [:43-44]."""
    zp = z.permute(0, 2, 3, 1).contiguous()
    zf = zp.view(-1, codebook.shape[1])
    d = (zf ** 2).sum(dim=1, keepdim=True) + (codebook ** 2).sum(1) - 2 * torch.matmul(zf, codebook.t())
    idx = torch.argmin(d, dim=1)
    z_q = codebook[idx].view(zp.shape)          # one_hot @ codebook (:50-55) is an exact row gather
    z_q = zp + (z_q - zp)
    return z_q.permute(0, 3, 1, 2).contiguous(), idx, d


# ----------------------------------------------------------------------------------------------
# full forward
# ----------------------------------------------------------------------------------------------

FUSE_ENCODER_BLOCK = {"512": 2, "256": 5, "128": 8, "64": 11, "32": 14, "16": 18}    # keep_arch.py:950-951
FUSE_GENERATOR_BLOCK = {"16": 6, "32": 9, "64": 12, "128": 15, "256": 18, "512": 21}  # keep_arch.py:953-954


def fusion_lists(sd):
    """(cft_list, cfa_list) of the checkpoint: the feature sizes that own `cft.<s>.*` / `cfa.<s>.*` tensors.
    'KEEP' general: cft 16/32/64 (modules/utils.py:46); 'Asian': cft 32/64/128/256 (:62); cfa 16/32 in both."""
    cft = [s for s in FUSE_GENERATOR_BLOCK if "cft.%s.scale.0.weight" % s in sd]
    cfa = [s for s in FUSE_GENERATOR_BLOCK if "cfa.%s.attn.to_q.weight" % s in sd]
    return cft, cfa


@torch.no_grad()
def keep_forward(sd, x, force_codes=None, force_flows=None, force_prev=None, capture=False):
    """x (b,T,3,512,512) fp32 in [-1,1] -> (b,T,3,512,512) unclamped.   keep_arch.py:1008-1145.

    Teacher forcing for stage-level parity:
      force_codes (b,T,256) long  — use these code indices instead of this run's argmax
      force_flows (b,T-1,2,H,W)   — use these flows instead of running GMFlow
      force_prev  (b,T,3,H,W)     — use these as `prev_out` (frame i uses force_prev[:, i-1])
    capture=True additionally returns a dict of intermediates.
    """
    W = _W(sd)
    b, T, c, H, Wd = x.shape
    cap = {}
    cft_list, cfa_list = fusion_lists(sd)
    ENC_TAPS = {FUSE_ENCODER_BLOCK[s]: s for s in cft_list}       # keep_arch.py:1030-1037
    GEN_CFT = {FUSE_GENERATOR_BLOCK[s]: s for s in cft_list}      # keep_arch.py:1053-1054
    GEN_CFA = {FUSE_GENERATOR_BLOCK[s]: s for s in cfa_list}      # keep_arch.py:1056-1057
    if force_flows is None:
        flows = gmflow_forward(x[:, 1:].reshape(-1, c, H, Wd), x[:, :-1].reshape(-1, c, H, Wd),
                               W.sub("flownet.model")).reshape(b, T - 1, 2, H, Wd)
    else:
        flows = force_flows
    z, taps = encoder_forward(x.reshape(-1, c, H, Wd), W.sub("encoder"), taps=tuple(ENC_TAPS))
    enc_feat = {ENC_TAPS[i]: t.reshape(b, T, *t.shape[1:]) for i, t in taps.items()}
    z_codes = z.reshape(b, T, *z.shape[1:])
    gains = kalman_gains(z_codes, W.sub("kalman_filter"))
    codebook = W("quantize.embedding.weight")
    outs, all_logits, all_codes, all_zhat = [], [], [], []
    cfa_prev = {}
    prev_out = None
    for i in range(T):
        if i == 0:
            z_hat = z_codes[:, 0]
        else:
            src = force_prev[:, i - 1] if force_prev is not None else prev_out
            z_prime, _ = encoder_forward(flow_warp(src, flows[:, i - 1]), W.sub("hq_encoder"))
            g = gains[:, i]
            z_hat = (1 - g) * z_codes[:, i] + g * z_prime
        logits = code_logits(z_hat, W)
        codes = logits.argmax(dim=2) if force_codes is None else force_codes[:, i]
        xg = codebook[codes].reshape(b, 16, 16, 256).permute(0, 3, 1, 2).contiguous()
        for j, kind in enumerate(GENERATOR_PROGRAM):
            xg = vq_block(kind, xg, W.sub("generator.blocks.%d" % j))
            if j in GEN_CFT:
                s = GEN_CFT[j]
                xg = cft_block(enc_feat[s][:, i], xg, W.sub("cft.%s" % s), 1.0)
            if j in GEN_CFA:
                s = GEN_CFA[j]
                if i > 0:
                    xg = cfa_block(xg, cfa_prev[s], W.sub("cfa.%s" % s))
                cfa_prev[s] = xg
        prev_out = xg
        outs.append(xg)
        if capture:
            all_logits.append(logits)
            all_codes.append(logits.argmax(dim=2))
            all_zhat.append(z_hat)
    out = torch.stack(outs, dim=1)
    if capture:
        cap.update(flows=flows, z_codes=z_codes, gains=gains, enc_feat=enc_feat,
                   logits=torch.stack(all_logits, 1), codes=torch.stack(all_codes, 1),
                   z_hat=torch.stack(all_zhat, 1))
        return out, cap
    return out
