"""TEST INFRASTRUCTURE ONLY — loads the *real* reference KEEP network from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
`oracle/make_golden.py` to pin the oracle restatement (`oracle/keep_oracle.py`) and to
generate the committed fixtures under `tests/golden/`.

Two shims are needed:
  1. `wm_basicsr/__init__.py` star-imports training-only code needing ffmpeg/av/comfy, so we
     pre-seed an empty namespace package whose __path__ points at the reference tree.
  2. `keep_arch.py:21` imports `FeedForward`/`AdaLayerNorm` from `diffusers`, which is not
     installed (nor declared by the reference).  We restate the diffusers-0.11-era GEGLU
     FeedForward (proj: Linear(dim, 8*dim); h, g = chunk(2); h * gelu_erf(g); Linear(4*dim, dim))
     with the key names the checkpoint uses (`net.0.proj`, `net.2`).  Parity at this one
     boundary is therefore pinned by this shim, not by diffusers itself ("parity unpinned").
"""
import os
import sys
import types

import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("KEEP_REFERENCE_ROOT", "/root/reference")
REF_DEPS = os.path.join(REF_ROOT, "modules", "deps")

KEEP_GENERAL_CFG = dict(  # modules/utils.py:42-57 merged with defaults :76-90
    img_size=512, emb_dim=256, dim_embd=512, n_head=8, n_layers=9, codebook_size=1024,
    cft_list=['16', '32', '64'], kalman_attn_head_dim=48, num_uncertainty_layers=3,
    cfa_list=['16', '32'], cfa_nhead=4, cfa_dim=256, cond=1, nf=64, ch_mult=[1, 2, 2, 4, 4, 8],
    attn_resolutions=[16], res_blocks=2, quantizer_type='nearest', beta=0.25, temp_reg_list=['32'],
    gumbel_straight_through=False, gumbel_kl_weight=1e-8, vqgan_path=None, latent_size=256,
    fix_modules=['quantize', 'generator'], flownet_path=None, cfa_nlayers=4, cross_residual=True,
    mask_ratio=0.)


KEEP_ASIAN_CFG = dict(KEEP_GENERAL_CFG, cft_list=['32', '64', '128', '256'], temp_reg_list=[])  # modules/utils.py:58-73

CONFIGS = {"KEEP": KEEP_GENERAL_CFG, "Asian": KEEP_ASIAN_CFG}


def available():
    return os.path.isdir(os.path.join(REF_DEPS, "wm_basicsr"))


def _install_shims():
    if "wm_basicsr" not in sys.modules:
        pkg = types.ModuleType("wm_basicsr")
        pkg.__path__ = [os.path.join(REF_DEPS, "wm_basicsr")]
        sys.modules["wm_basicsr"] = pkg
    if "diffusers.models.attention" not in sys.modules:
        class GEGLU(nn.Module):
            def __init__(self, dim_in, dim_out):
                super().__init__()
                self.proj = nn.Linear(dim_in, dim_out * 2)

            def forward(self, x):
                h, g = self.proj(x).chunk(2, dim=-1)
                return h * F.gelu(g)

        class FeedForward(nn.Module):
            def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu"):
                super().__init__()
                assert activation_fn == "geglu"
                self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(dropout),
                                          nn.Linear(dim * mult, dim_out or dim)])

            def forward(self, x):
                for m in self.net:
                    x = m(x)
                return x

        class AdaLayerNorm(nn.Module):  # never constructed by KEEP (num_embeds_ada_norm is None)
            pass

        d = types.ModuleType("diffusers")
        dm = types.ModuleType("diffusers.models")
        da = types.ModuleType("diffusers.models.attention")
        da.FeedForward, da.AdaLayerNorm = FeedForward, AdaLayerNorm
        sys.modules.update({"diffusers": d, "diffusers.models": dm, "diffusers.models.attention": da})
    if REF_DEPS not in sys.path:
        sys.path.insert(0, REF_DEPS)


def load_reference_keep(state_dict=None, config="KEEP"):
    """Build reference KEEP ('KEEP' general or 'Asian' config), optionally load a state dict (strict)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_shims()
    from wm_basicsr.archs.keep_arch import KEEP  # noqa
    net = KEEP(**CONFIGS[config]).eval()
    if state_dict is not None:
        net.load_state_dict(state_dict, strict=True)
    return net
