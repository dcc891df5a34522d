"""GPU tier: lockstep clip batching (KEEP_FLAG_BATCH_CLIPS / KeepNetB200(batch_clips=N)).

A batch of b > 1 clips handed to ONE engine walks the per-frame recurrence in lockstep (one batched hq_encoder / code
transformer / generator pass per frame index).  Clips are independent in the reference (keep_processor.py:263-270), so the
result must equal the clip-by-clip loop -- up to fp32 summation order (a batch of two picks different K-splits), i.e. to
the pixel bar (max-abs <= 1e-2 on clamped pixels) until a code index flips at a near-tie of the per-clip run's own logits."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NEAR_TIE = 2e-2


def _report(tag, **kw):
    from conftest import parity_report
    parity_report(tag, **kw)


def _make(keep_mod, state_dict, mode, **kw):
    kn = keep_mod.keep_net
    flags = (0 if mode == "fp32" else kn.TC3_FLAGS) | kw.pop("extra_flags", 0)
    n = keep_mod.KeepNetB200(flags=flags, **kw)
    n.load_state_dict(state_dict, strict=True)
    return n.eval().to("cuda")


def _clips(b, T):
    from oracle import weights
    return torch.cat([weights.make_clip(T, seed=4321 + 17 * c, coherent=True) for c in range(b)], 0)


def _compare(out, ref, margins, tag):
    """out / ref (b, T, 3, H, W); margins[c] (T, 256) top1-top2 logit margins of the per-clip run."""
    b, T = out.shape[:2]
    for c in range(b):
        for i in range(T):
            e = float((out[c, i].clamp(-1, 1) - ref[c, i].clamp(-1, 1)).abs().max())
            if e <= 1e-2:
                continue
            worst = float(margins[c][i].min())
            assert i > 0 or worst < 1e-3, "%s clip %d frame 0 differs by %g (min margin %g)" % (tag, c, e, worst)
            assert worst < NEAR_TIE, "%s clip %d frame %d differs by %g at a non-tie (min margin %g)" % (tag, c, i, e, worst)
            break   # later frames of this clip diverged legitimately through the recurrence


@pytest.mark.parametrize("mode", ["tc3", "fp32"])
def test_lockstep_pair_matches_clip_by_clip(keep_mod, state_dict, mode):
    b, T = 2, 2
    x = _clips(b, T).cuda()
    single = _make(keep_mod, state_dict, mode)
    single.debug_capture(True)
    ref, margins = [], []
    for c in range(b):
        ref.append(single(x[c:c + 1], need_upscale=False).cpu())
        top2 = single.debug_read("logits", (T, 256, 1024)).topk(2, dim=2).values
        margins.append(top2[..., 0] - top2[..., 1])
    ref = torch.cat(ref, 0)
    single.to("cpu")
    net = _make(keep_mod, state_dict, mode, batch_clips=2)
    out = net(x, need_upscale=False).cpu()
    assert out.shape == ref.shape and bool(torch.isfinite(out).all())
    _report("lockstep_pair[%s]" % mode,
            maxabs_per_clip_frame=[["%.2e" % float((out[c, i].clamp(-1, 1) - ref[c, i].clamp(-1, 1)).abs().max()) for i in range(T)]
                                   for c in range(b)],
            min_margin=[["%.2e" % float(margins[c][i].min()) for i in range(T)] for c in range(b)])
    _compare(out, ref, margins, "lockstep[%s]" % mode)
    # frame 0 has no recurrence behind it and the LQ encoder runs clip by clip on both paths: tight agreement
    assert float((out[:, 0] - ref[:, 0]).abs().max()) < 5e-3
    net.to("cpu")


def test_lockstep_clip_matches_the_reference_fixture(keep_mod, state_dict):
    """The lockstep path against the REAL reference (tests/golden/ref_T3_coherent.npz, written by oracle/make_golden.py
    from /root/reference's KEEP.forward), not just against this engine's own clip-by-clip run: clip 0 of the batch is the
    fixture's clip.  Frame 0 (no recurrence behind it) must meet the pixel bar outright; later frames meet it until the
    reference's own logits have a near-tie (the same rule as test_gpu_parity.check_free_running)."""
    from oracle import weights
    from test_gpu_parity import psnr
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_T3_coherent.npz"))
    T = 3
    x = torch.cat([weights.make_clip(T, seed=1234, coherent=True), weights.make_clip(T, seed=4338, coherent=True)], 0).cuda()
    net = _make(keep_mod, state_dict, "tc3", batch_clips=2)
    out = net(x, need_upscale=False).cpu()
    net.to("cpu")
    ref_sub = torch.from_numpy(g["out_sub4"])                     # (1, T, 3, 128, 128)
    top2 = torch.from_numpy(g["logit_top2"])[0]
    margin = top2[..., 0] - top2[..., 1]
    sub = out[:1, :, :, ::4, ::4]
    e0 = float((sub[:, 0].clamp(-1, 1) - ref_sub[:, 0].clamp(-1, 1)).abs().max())
    p0 = psnr(sub[:, 0], ref_sub[:, 0])
    _report("lockstep_vs_reference_fixture", frame0_maxabs=e0, frame0_psnr=p0,
            maxabs=["%.2e" % float((sub[:, i].clamp(-1, 1) - ref_sub[:, i].clamp(-1, 1)).abs().max()) for i in range(T)],
            min_margin=["%.2e" % float(margin[i].min()) for i in range(T)])
    assert e0 <= 1e-2 and p0 >= 50.0
    _compare(sub, ref_sub, [margin], "lockstep-vs-fixture")


def test_lockstep_odd_clip_graph_replay_and_u8(keep_mod, state_dict):
    """b = 3 with groups of 2: one lockstep pair + the odd clip on the per-clip path; CUDA-graph replay equals the eager
    call bit for bit; the uint8 call goes through the same lockstep path."""
    kn = keep_mod.keep_net
    b, T = 3, 3
    x = _clips(b, T).cuda()
    net = _make(keep_mod, state_dict, "tc3", batch_clips=2, extra_flags=kn.FLAG_CUDA_GRAPH)
    first = net(x, need_upscale=False).clone()       # eager (packs weights, warms the allocations)
    second = net(x, need_upscale=False).clone()      # graph capture + launch
    third = net(x, need_upscale=False).clone()       # graph replay
    assert torch.equal(first, second) and torch.equal(second, third)
    single = _make(keep_mod, state_dict, "tc3")
    last = single(x[2:3], need_upscale=False)
    assert torch.equal(first[2:3], last)             # the odd clip took the ordinary per-clip path
    u8 = ((x.permute(0, 1, 3, 4, 2).flip(-1) * 0.5 + 0.5) * 255.0).round().clamp(0, 255).to(torch.uint8).contiguous()
    got = net.forward_u8(u8)
    want = torch.cat([single.forward_u8(u8[c:c + 1]) for c in range(b)], 0)
    assert got.shape == want.shape and got.dtype == torch.uint8
    close = ((got.int() - want.int()).abs() <= 1).float().mean()
    _report("lockstep_u8", within_1_lsb=float(close), exact=float((got == want).float().mean()))
    assert float(close) > 0.98, "uint8 lockstep output differs from the clip-by-clip loop on %.3f of the bytes" % (1 - float(close))
    net.to("cpu")
    single.to("cpu")
