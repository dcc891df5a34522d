"""CPU tier: the reference's OWN host classes (KEEPModelLoader / KEEPModelPack / KEEPFaceProcessor, imported from
/root/reference with the ComfyUI host stubbed -- oracle/ref_host.py) driven through load -> cache hit -> load_device ->
offload with the B200 engine swapped in the two ways INTEGRATION.md documents, and the real `process_image_sequence`
clip loop (keep_processor.py:258-273) pinned against the host mirror in comfyui-keep_b200/sharding.py.

No device: the engine is built with KEEP_FLAG_PLAN_ONLY (`KeepNetB200(plan_only=True)`: strict key / shape check and the
workspace plan run in libkeep_b200.so, nothing is uploaded).  Skipped where /root/reference is absent (the GPU box)."""
import importlib.util
import os
import sys
import types

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="needs the reference tree (/root/reference)")


@pytest.fixture(scope="module")
def ref_mods(tmp_path_factory, state_dict):
    from oracle import ref_host
    d = tmp_path_factory.mktemp("keep_models")
    loader_mod, proc_mod = ref_host.install(device="cuda:0", models_dir=str(d))
    ckpt = ref_host.fake_checkpoint(state_dict, str(d / "keep_models" / "KEEP" / "KEEP-b76feb75.pth"))
    other = str(d / "facelib.pth")
    torch.save({}, other)
    loader_mod.load_file_from_url_comfy = lambda url, model_dir_name, file_name, sha256=None: ckpt if "KEEP-" in url else other
    return loader_mod, proc_mod


def _lifecycle(loader, keep_mod):
    """first load, second load (cache hit), then load_device / offload twice, as two node executions do (nodes.py:119-136)"""
    pack1 = loader.load_keep_model_pack("KEEP", "retinaface_resnet50")
    net = pack1.keep_net
    assert isinstance(net, keep_mod.KeepNetB200), "first load must hand out the B200 engine"
    assert net._flags & keep_mod.DEFAULT_FLAGS == keep_mod.DEFAULT_FLAGS, "the default engine is the measured tc3 + CUDA-graph mode"
    assert all(p.keep_net is net for p in loader.loaded_models.values()), "the loader's cache must hold the engine, not the torch module"
    pack2 = loader.load_keep_model_pack("KEEP", "retinaface_resnet50")          # cache hit (keep_model_loader.py:76-86)
    assert pack2 is not pack1 and pack2.keep_net is net, "second load (cache hit) must hand out the same engine"
    for pack in (pack1, pack2):
        assert net._engine is None
        pack.load_device()                                                     # keep_net.to(device) (:28-31)
        assert net._engine is not None and net._device.type == "cuda"
        lib = keep_mod.keep_net.load_library()
        assert lib.keep_workspace_bytes(net._engine, 1, 20) > (1 << 30)         # the C++ side accepted all 896 tensors and planned T=20
        pack.offload()                                                         # keep_net.to(offload_device) (:45-48)
        assert net._engine is None and net._device.type == "cpu"
    return net


def test_install_into_loader_lifecycle(ref_mods, keep_mod, state_dict):
    loader_mod, _ = ref_mods
    loader = keep_mod.install_into_loader(loader_mod.KEEPModelLoader(), plan_only=True)
    assert keep_mod.install_into_loader(loader) is loader                      # idempotent
    net = _lifecycle(loader, keep_mod)
    # the engine carries the checkpoint's tensors under the reference's (renamed) key set
    sd = net.state_dict()
    assert set(sd) == set(state_dict)
    for k in ("cfa.16.attn.to_q.weight", "cft.32.scale.0.weight", "encoder.blocks.0.weight"):
        assert torch.equal(sd[k], state_dict[k])


def test_documented_source_patch_of_the_loader(ref_mods, keep_mod):
    """INTEGRATION.md's one-line edit, applied literally to the reference's loader source at test time: `net` itself is
    replaced between `net.eval()` (keep_model_loader.py:121) and the pack construction (:140), so pack AND cache get it."""
    loader_mod, _ = ref_mods
    src = open(loader_mod.__file__).read()
    anchor = "        net.eval()\n"
    assert src.count(anchor) == 1
    patched = src.replace(anchor, anchor + "        import keep_b200; net = keep_b200.from_reference(net, plan_only=True)\n")
    name = loader_mod.__name__ + "_patched"
    mod = types.ModuleType(name)
    mod.__package__ = loader_mod.__package__
    mod.__file__ = loader_mod.__file__
    sys.modules[name] = mod
    exec(compile(patched, loader_mod.__file__, "exec"), mod.__dict__)
    mod.FaceRestoreHelper = loader_mod.FaceRestoreHelper
    mod.load_file_from_url_comfy = loader_mod.load_file_from_url_comfy
    _lifecycle(mod.KEEPModelLoader(), keep_mod)


def test_old_one_pack_swap_misses_the_cache(ref_mods, keep_mod):
    """Why `install_into_model_pack` alone is not the documented swap: the cache keeps the torch module (round-1 finding)."""
    loader_mod, _ = ref_mods
    loader = loader_mod.KEEPModelLoader()
    pack1 = loader.load_keep_model_pack("KEEP", "retinaface_resnet50")
    keep_mod.install_into_model_pack(pack1, plan_only=True)
    pack2 = loader.load_keep_model_pack("KEEP", "retinaface_resnet50")
    assert isinstance(pack1.keep_net, keep_mod.KeepNetB200) and not isinstance(pack2.keep_net, keep_mod.KeepNetB200)


class _RecordingNet:
    """keep_net stand-in: records every call the reference's processor makes and returns the clip unchanged"""

    def __init__(self):
        self.calls = []

    def to(self, *_a, **_k):
        return self

    def __call__(self, x, need_upscale=True):
        assert need_upscale is False and x.dtype == torch.float32 and x.dim() == 5 and x.shape[0] == 1
        assert float(x.min()) >= -1.0 - 1e-6 and float(x.max()) <= 1.0 + 1e-6
        self.calls.append(tuple(x.shape))
        return x.clone()


@pytest.mark.parametrize("n_frames,max_clip", [(41, 20), (7, 3), (20, 20), (5, 100)])
def test_real_process_image_sequence_clip_loop_matches_host_mirror(ref_mods, keep_mod, n_frames, max_clip):
    """The REAL KEEPFaceProcessor.process_image_sequence (aligned frames) with a recording keep_net: the calls it makes are
    exactly `sharding.split_clips` (1-frame tail duplicated to T=2, keep_processor.py:266-268), and `sharding.run_clips` --
    what bench.py --config 3 / 5 time on the GPU -- reproduces its output tensor."""
    loader_mod, proc_mod = ref_mods
    from oracle import ref_host
    net = _RecordingNet()
    pack = loader_mod.KEEPModelPack(net, ref_host._FakeFaceHelper(device="cpu"), None, None, "KEEP")
    pack.device = torch.device("cpu")
    proc = proc_mod.KEEPFaceProcessor(pack)
    g = torch.Generator().manual_seed(5)
    images = torch.rand((n_frames, 64, 64, 3), generator=g)                    # ComfyUI IMAGE: (N, H, W, 3) fp32 RGB in [0, 1]
    seen = {}
    real_cat = torch.cat

    def spy_cat(tensors, dim=0):
        out = real_cat(tensors, dim=dim)
        if dim == 1 and out.dim() == 5 and out.shape[1] == n_frames:
            seen["restored"] = out
        return out

    proc_mod.torch.cat = spy_cat
    try:
        out = proc.process_image_sequence(images, 1.0, True, True, False, max_clip_length=max_clip)
    finally:
        proc_mod.torch.cat = real_cat
    assert tuple(out.shape) == (n_frames, 64, 64, 3)                            # aligned path returns the background
    clips = keep_mod.sharding.split_clips(n_frames, max_clip)
    assert net.calls == [(1, max(2, e - s), 3, 512, 512) for s, e, _ in clips]
    # the mirror, fed the same crops the processor built, makes the same calls and the same (1, N, 3, 512, 512) tensor
    mirror = _RecordingNet()
    import cv2
    import numpy as np
    crops = []
    for i in range(n_frames):
        bgr = proc_mod.comfy_image_to_cv2(images[i].unsqueeze(0))
        face = cv2.resize(bgr, (512, 512), interpolation=cv2.INTER_LINEAR)
        t = proc_mod.img2tensor(face / 255., bgr2rgb=True, float32=True)
        crops.append((t - 0.5) / 0.5)
    frames = torch.stack(crops, 0).unsqueeze(0)
    got = keep_mod.sharding.run_clips(mirror, frames, max_clip)
    assert mirror.calls == net.calls
    assert torch.equal(got, seen["restored"])
    # and the uint8 path's host conversions are the reference's: tensor2img(min_max=(-1, 1)) on what the net returned
    ref_u8 = proc_mod.tensor2img(got[0, 0], rgb2bgr=True, min_max=(-1, 1))
    x = got[0, 0].clamp(-1, 1)
    mine = ((x + 1) / 2 * 255.0).round().permute(1, 2, 0).flip(-1).to(torch.uint8).numpy()
    assert np.array_equal(ref_u8, mine)


# ---- the two reference-side edits INTEGRATION.md documents for the caller, applied to the real source ------
_LOOP_OLD = """            for start_idx in tqdm(range(0, num_total_faces, max_clip_length), desc="Restoring faces with KEEP"):
                end_idx = min(start_idx + max_clip_length, num_total_faces)
                current_clip = batched_cropped_faces[:, start_idx:end_idx, ...]
                if current_clip.shape[1] == 1:
                    current_clip = torch.cat([current_clip, current_clip], dim=1)
                    temp_restored_tensors.append(self.keep_net(current_clip, need_upscale=False)[:, 0:1, ...])
                elif current_clip.shape[1] > 1:
                    temp_restored_tensors.append(self.keep_net(current_clip, need_upscale=False))
"""
_LOOP_NEW = """            import keep_b200   # all clips of equal length in ONE call: the engine overlaps / lock-steps them
            temp_restored_tensors.append(keep_b200.sharding.run_clips_batched(
                self.keep_net, batched_cropped_faces, max_clip_length, clips_per_call=4))
"""
_DROP_OLD = """            if num_faces_this_frame == 0 or has_aligned_frames:
                output_frames_cv2.append(bg_img_final) # a little simplified, aligned case could be handled better
                continue
"""
_DROP_NEW = """            if has_aligned_frames and num_faces_this_frame:   # aligned input: the restored crop IS the frame (was: dropped)
                face = all_restored_faces_cv2[restored_face_idx_counter].astype('uint8')
                restored_face_idx_counter += num_faces_this_frame
                output_frames_cv2.append(cv2.resize(face, (target_w, target_h), interpolation=cv2.INTER_LANCZOS4))
                continue
            if num_faces_this_frame == 0:
                output_frames_cv2.append(bg_img_final)
                continue
"""


def test_documented_processor_patch_batches_clips_and_returns_the_restored_frames(ref_mods, keep_mod):
    """INTEGRATION.md's caller-side edits (keep_processor.py:263-270 and :289-291), applied literally to the reference's source
    at test time: the clip loop becomes one batched call per group of equal-length clips (what `batch_clips` / `concurrent_clips`
    engines overlap), and an aligned sequence returns the restored frames instead of the untouched background."""
    loader_mod, proc_mod = ref_mods
    from oracle import ref_host
    src = open(proc_mod.__file__).read()
    assert src.count(_LOOP_OLD) == 1 and src.count(_DROP_OLD) == 1, "the reference's clip loop / paste-back changed: update INTEGRATION.md"
    name = proc_mod.__name__ + "_patched"
    mod = types.ModuleType(name)
    mod.__package__ = proc_mod.__package__
    mod.__file__ = proc_mod.__file__
    sys.modules[name] = mod
    exec(compile(src.replace(_LOOP_OLD, _LOOP_NEW).replace(_DROP_OLD, _DROP_NEW), proc_mod.__file__, "exec"), mod.__dict__)
    net = _RecordingNet.__new__(_RecordingNet)
    net.calls = []
    net.__class__ = type("BatchNet", (_RecordingNet,), {"__call__": lambda self, x, need_upscale=True: (self.calls.append(tuple(x.shape)), x.clone())[1]})
    pack = loader_mod.KEEPModelPack(net, ref_host._FakeFaceHelper(device="cpu"), None, None, "KEEP")
    pack.device = torch.device("cpu")
    g = torch.Generator().manual_seed(11)
    images = torch.rand((9, 64, 64, 3), generator=g)
    out = mod.KEEPFaceProcessor(pack).process_image_sequence(images, 1.0, True, True, False, max_clip_length=2)
    # 9 frames, clips of 2: four full clips go in as ONE (4, 2, 3, 512, 512) call, the 1-frame tail duplicated to T = 2
    assert net.calls == [(4, 2, 3, 512, 512), (1, 2, 3, 512, 512)]
    assert tuple(out.shape) == (9, 64, 64, 3)
    # the identity "network" makes the restored frame the input itself (up to the 64 -> 512 -> 64 resampling and uint8 rounding)
    assert float((out - images).abs().mean()) < 0.05 and not torch.equal(out, images)   # (unpatched: out IS the background = images)
