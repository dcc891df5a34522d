"""TEST INFRASTRUCTURE ONLY — imports the reference's own HOST classes (`KEEPModelLoader`, `KEEPModelPack`,
`KEEPFaceProcessor`; /root/reference/modules/keep_model_loader.py, keep_processor.py) in this container, so the CPU test
tier can drive the real load -> cache-hit -> load_device -> process -> offload lifecycle with the B200 engine swapped in.

Only usable where /root/reference exists (the build container; not the GPU box).  What is stubbed, and nothing else:
  * `comfy.model_management` / `comfy.utils` / `folder_paths` — the ComfyUI host, absent here;
  * the checkpoint download (`load_file_from_url_comfy`): returns a seeded synthetic checkpoint written to a temp dir, with
    the LEGACY key names (`cross_fuse`, `fuse_convs_dict`) so the loader's rename (keep_model_loader.py:110-118) runs;
  * `ffmpeg` (ffmpeg-python, imported by wm_basicsr/utils/video_util.py for the VideoReader the nodes never use; the
    reference would otherwise try to pip-install it): an empty module;
  * `FaceRestoreHelper` (needs detector / parser weights from the network): an object with the attributes the pack touches.
The reference modules use package-relative imports (`from .. import logger`), so they are imported under a synthetic
parent package `_refkeep` whose `__path__` is the reference root.
"""
import importlib
import logging
import os
import sys
import types

import torch

from . import ref_loader

PKG = "_refkeep"


class _FakeFaceHelper:
    """Stands in for wm_facelib's FaceRestoreHelper(upscale_factor=1, face_size=512, ...) (keep_model_loader.py:131-135)."""

    def __init__(self, **kw):
        self.kw = kw
        self.device = kw.get("device")
        self.upscale_factor = kw.get("upscale_factor", 1)
        self.cropped_faces, self.restored_faces, self.affine_matrices, self.all_landmarks_5 = [], [], [], []
        self.is_gray = False

    def clean_all(self):
        self.cropped_faces, self.restored_faces, self.affine_matrices, self.all_landmarks_5 = [], [], [], []


def install(device="cpu", models_dir="/tmp/keep_ref_models"):
    """Import the reference's `modules` package with the ComfyUI host stubbed; returns (loader_module, processor_module)."""
    if not ref_loader.available():
        raise RuntimeError("reference tree not present at %s" % ref_loader.REF_ROOT)
    ref_loader._install_shims()
    sys.modules.setdefault("ffmpeg", types.ModuleType("ffmpeg"))
    import wm_basicsr.archs.keep_arch  # noqa: F401  (registers 'KEEP' in ARCH_REGISTRY, as wm_basicsr/__init__.py would)
    if "comfy" not in sys.modules:
        comfy = types.ModuleType("comfy")
        mm = types.ModuleType("comfy.model_management")
        cu = types.ModuleType("comfy.utils")

        class ProgressBar:
            def __init__(self, total):
                self.total, self.n = total, 0

            def update(self, n):
                self.n += n

        cu.ProgressBar = ProgressBar
        cu.tiled_scale = lambda x, fn, **kw: fn(x)
        mm.soft_empty_cache = lambda: None
        comfy.model_management, comfy.utils = mm, cu
        fp = types.ModuleType("folder_paths")
        fp.models_dir = models_dir
        sys.modules.update({"comfy": comfy, "comfy.model_management": mm, "comfy.utils": cu, "folder_paths": fp})
    mm = sys.modules["comfy.model_management"]
    mm.get_torch_device = lambda: torch.device(device)
    mm.unet_offload_device = lambda: torch.device("cpu")
    if PKG not in sys.modules:
        root = types.ModuleType(PKG)
        root.__path__ = [ref_loader.REF_ROOT]
        root.logger = logging.getLogger("ComfyUI-KEEP")
        mods = types.ModuleType(PKG + ".modules")
        mods.__path__ = [os.path.join(ref_loader.REF_ROOT, "modules")]
        sys.modules[PKG], sys.modules[PKG + ".modules"] = root, mods
    loader_mod = importlib.import_module(PKG + ".modules.keep_model_loader")
    proc_mod = importlib.import_module(PKG + ".modules.keep_processor")
    loader_mod.FaceRestoreHelper = _FakeFaceHelper
    return loader_mod, proc_mod


def fake_checkpoint(state_dict, path, legacy_names=True):
    """Write `state_dict` the way the released checkpoint stores it: under 'params_ema', with the pre-rename key names."""
    sd = {}
    for k, v in state_dict.items():
        if legacy_names:
            k = k.replace("cfa.", "cross_fuse.", 1) if k.startswith("cfa.") else k
            k = k.replace("cft.", "fuse_convs_dict.", 1) if k.startswith("cft.") else k
        sd[k] = v
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({"params_ema": sd}, path)
    return path
