"""Seeded synthetic weights and clips for the KEEP network, 'KEEP' general or 'Asian' config (benchmarks and tests; not the oracle).

No checkpoint is available offline (the reference downloads `KEEP-b76feb75.pth` at run time,
modules/utils.py:55), so parity and benchmarks use a seeded synthetic state dict with the
reference's exact key set and shapes (`keep_state_shapes.json`, dumped from the reference's own
`KEEP(**cfg).state_dict()` by oracle/dump_shapes.py — general: 896 tensors, 158.49 M parameters; `keep_state_shapes_asian.json`:
914 tensors, 143.57 M: no `cft.16`, `cft.128` and `cft.256` added, modules/utils.py:58-73).

The reference's default init makes several paths no-ops (CFT convs, CFA linears,
`position_emb`, a ±1/1024 codebook), so this generator draws *every* tensor from a live
distribution instead:
  * conv / linear weights  ~ U(±sqrt(3 / fan_in))   (variance-preserving)
  * 1-D `*.weight` (GroupNorm / LayerNorm gains)   ~ 1 + 0.1 N(0,1)
  * all biases                                     ~ 0.05 N(0,1)
  * position_emb                                   ~ 0.1 N(0,1)
  * quantize.embedding.weight (codebook)           ~ 0.5 N(0,1)
Generation uses a CPU `torch.Generator` seeded per tensor (seed, key-index), so the dict is
reproducible on any box with the same torch build, independent of iteration order.
"""
import json
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


SHAPE_TABLES = {"KEEP": "keep_state_shapes.json", "Asian": "keep_state_shapes_asian.json"}


def load_shapes(config="KEEP"):
    with open(os.path.join(_HERE, SHAPE_TABLES[config])) as f:
        return json.load(f)  # insertion-ordered: reference state_dict order


def make_state_dict(seed=0, dtype=torch.float32, config="KEEP"):
    shapes = load_shapes(config)
    sd = {}
    for idx, (key, shape) in enumerate(shapes.items()):
        g = torch.Generator(device="cpu")
        g.manual_seed((seed * 1000003 + idx * 7919 + 12345) & 0x7FFFFFFF)
        if key == "position_emb":
            t = 0.1 * torch.randn(shape, generator=g)
        elif key == "quantize.embedding.weight":
            t = 0.5 * torch.randn(shape, generator=g)
        elif key.endswith("bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        elif len(shape) == 1:  # norm gains
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:  # conv (O, I, kh, kw) / linear (O, I)
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            bound = math.sqrt(3.0 / fan_in)
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
        sd[key] = t.to(dtype).contiguous()
    return sd


def make_clip(T, seed=1234, coherent=True, b=1):
    """Synthetic aligned clip (b, T, 3, 512, 512) fp32 in [-1, 1].

    coherent=False: i.i.d. uniform noise.
    coherent=True : one smooth random image translated by a sub-pixel drift per frame plus a
                    little noise, so flow / warp / cross-frame attention see realistic motion.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    if not coherent:
        return torch.rand((b, T, 3, 512, 512), generator=g) * 2.0 - 1.0
    low = torch.randn((b, 3, 24, 24), generator=g)
    base = torch.nn.functional.interpolate(low, size=(560, 560), mode="bicubic", align_corners=False)
    mid = torch.randn((b, 3, 96, 96), generator=g) * 0.35
    base = base + torch.nn.functional.interpolate(mid, size=(560, 560), mode="bicubic", align_corners=False)
    base = base / base.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-6)
    frames = []
    for t in range(T):
        dx = 24 + int(round(3.0 * t))
        dy = 24 + int(round(-2.0 * t + 0.5 * t * t / max(T, 1)))
        dx = max(0, min(48, dx))
        dy = max(0, min(48, dy))
        f = base[:, :, dy:dy + 512, dx:dx + 512]
        f = f + 0.02 * torch.randn(f.shape, generator=g)
        frames.append(f.clamp(-1.0, 1.0))
    return torch.stack(frames, dim=1).contiguous()


def make_latents(codebook, n=2, seed=99, hw=16):
    """Synthetic encoder latents (n, C, hw, hw) for the nearest-neighbour quantiser: half of the tokens sit near a
    codebook entry (entry + small noise, as a trained encoder's outputs do), a quarter are far from everything (pure noise,
    near-tied distances), an eighth equal an entry exactly (distance ~ 0) and an eighth are midpoints of two entries
    (a constructed near-tie).  Deterministic in (codebook, n, seed)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    K, C = codebook.shape
    L = n * hw * hw
    pick = torch.randint(0, K, (L,), generator=g)
    pick2 = torch.randint(0, K, (L,), generator=g)
    kind = torch.randint(0, 8, (L,), generator=g)
    noise = torch.randn((L, C), generator=g)
    z = codebook[pick] + 0.1 * codebook.std() * noise                      # kinds 0-3: near an entry
    z = torch.where((kind >= 4)[:, None] & (kind < 6)[:, None], codebook.std() * noise, z)
    z = torch.where((kind == 6)[:, None], codebook[pick], z)
    z = torch.where((kind == 7)[:, None], 0.5 * (codebook[pick] + codebook[pick2]) + 1e-3 * noise, z)
    return z.view(n, hw, hw, C).permute(0, 3, 1, 2).contiguous()


def make_vq_case(codebook, n=2, seed=99):
    """(codebook', z) for the quantiser tests: entry 700 is made an exact duplicate of entry 3 and the first eight tokens
    sit on / next to that pair, so exact ties occur and must resolve to the lowest index (torch.argmin, vqgan_arch.py:47)."""
    cb = codebook.clone()
    cb[700] = cb[3]
    z = make_latents(cb, n=n, seed=seed)
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + 1)
    zt = z.permute(0, 2, 3, 1).contiguous()
    flat = zt.view(-1, cb.shape[1])
    flat[0:4] = cb[700]
    flat[4:8] = cb[700] + 0.05 * cb.std() * torch.randn((4, cb.shape[1]), generator=g)
    return cb, zt.permute(0, 3, 1, 2).contiguous()
