"""GPU tier: BASELINE.json configs[1] at full size -- one 20-frame aligned 512x512 clip -- against the REAL reference's
outputs (tests/golden/ref_T20_coherent.npz, written by oracle/make_golden.py; the oracle restatement reproduces that run
with 100 % code agreement over all 5120 indices, tests/golden/pin_report.json).

What full size adds over the T = 2 / 3 cases: the LQ encoder in two passes of 10 frames, GMFlow in 5 chunks on the side
stream, the gain estimator's temporal attention over 20 frames, and 19 steps of the CFA recurrence.  The free-running
pixel comparison stops at the first near-tie flip (chaotic synthetic weights); with only the discrete code
indices teacher-forced, every one of the 20 frames must meet the pixel bar (max-abs <= 1e-2 clamped, PSNR >= 50 dB)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
T = 20


def psnr(a, b):
    a, b = a.double().clamp(-1, 1), b.double().clamp(-1, 1)
    mse = float(((a - b) ** 2).mean()) / 4.0
    return 999.0 if mse == 0 else 10 * np.log10(1.0 / mse)


def _report(tag, **kw):
    from conftest import parity_report
    parity_report(tag, **kw)


@pytest.fixture(scope="module")
def net(keep_mod, state_dict):
    kn = keep_mod.keep_net
    n = keep_mod.KeepNetB200(flags=kn.TC3_FLAGS)
    n.load_state_dict(state_dict, strict=True)
    n.eval().to("cuda")
    n.debug_capture(True)
    yield n
    n.to("cpu")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "ref_T20_coherent.npz"))


@pytest.fixture(scope="module")
def clip():
    from oracle import weights
    return weights.make_clip(T, seed=1234, coherent=True)


def test_T20_batched_stages_match_reference(net, gold, clip):
    """Free-running call: the stages that are batched over all 20 frames (flows, LQ-encoder latents, Kalman gains) against
    the reference, then pixels frame by frame until the first near-tie flip."""
    out = net(clip.cuda(), need_upscale=False).cpu()
    assert out.shape == clip.shape and bool(torch.isfinite(out).all())
    z = net.debug_read("z_codes", (T, 16, 16, 256)).permute(0, 3, 1, 2)
    gains = net.debug_read("gains", (T, 16, 16))
    flows = net.debug_read("flows", (T - 1, 512, 512, 2)).permute(0, 3, 1, 2)[None]
    codes = net.debug_read("codes", (T, 256), torch.int32).long()
    e_z0 = float((z[0] - torch.from_numpy(gold["z_first"])).abs().max())
    e_z19 = float((z[T - 1] - torch.from_numpy(gold["z_last"])).abs().max())
    e_zm = float((z.double().mean(dim=(1, 2, 3)) - torch.from_numpy(gold["z_mean"])).abs().max())
    e_g = float((gains - torch.from_numpy(gold["gains"]).reshape(T, 16, 16)).abs().max())
    fmax = float(np.abs(gold["flows_sub16"]).max())
    e_f = float((flows[:, :, :, ::16, ::16] - torch.from_numpy(gold["flows_sub16"])).abs().max())
    ref_codes = torch.from_numpy(gold["codes"].astype(np.int64))[0]
    top2 = torch.from_numpy(gold["logit_top2"])[0]
    margin = top2[..., 0] - top2[..., 1]
    agree = [float((codes[i] == ref_codes[i]).float().mean()) for i in range(T)]
    _report("T20_free[tc3]", z_first=e_z0, z_last=e_z19, z_mean=e_zm, gain=e_g, flow=e_f, flow_max=fmax, agree=agree)
    assert e_z0 < 2e-3 and e_z19 < 2e-3 and e_zm < 1e-4 and e_g < 2e-4 and e_f < 1e-3 * fmax
    ref_sub = torch.from_numpy(gold["out_sub8"])
    compared = 0
    for i in range(T):
        flips = codes[i] != ref_codes[i]
        if bool(flips.any()):
            worst = float(margin[i][flips].max())
            _report("T20_free[tc3].first_flip", frame=i, flips=int(flips.sum()), worst_margin=worst)
            # one near-tie threshold for every frame (the reference's own fp32-vs-fp64 logit spread is 5.8e-3 after
            # 3 frames; the engine's measured first flip on this clip sits at a 2.7e-3 margin)
            assert worst < 2e-2, "frame %d: code index flipped at a non-tie (margin %g)" % (i, worst)
            break
        a, b = out[:, i, :, ::8, ::8], ref_sub[:, i]
        e = float((a.clamp(-1, 1) - b.clamp(-1, 1)).abs().max())
        assert e <= 1e-2 and psnr(a, b) >= 50.0, "frame %d: max-abs %g, PSNR %g" % (i, e, psnr(a, b))
        compared += 1
    _report("T20_free[tc3].frames_compared", n=compared)
    assert compared >= 1


def test_T20_codes_forced_every_frame_meets_the_pixel_bar(net, gold, clip):
    """Only the discrete decisions are teacher-forced (the reference's 20 x 256 code indices); LQ encoder taps, generator,
    CFT and 19 steps of the CFA recurrence run free on the engine's own intermediates."""
    try:
        net.debug_force("codes", torch.from_numpy(gold["codes"].astype(np.int32))[0])
        out = net(clip.cuda(), need_upscale=False).cpu()
    finally:
        net.debug_force("codes", None)
    ref_sub = torch.from_numpy(gold["out_sub8"])
    worst_e, worst_p = 0.0, 999.0
    for i in range(T):
        a, b = out[:, i, :, ::8, ::8], ref_sub[:, i]
        e = float((a.clamp(-1, 1) - b.clamp(-1, 1)).abs().max())
        worst_e, worst_p = max(worst_e, e), min(worst_p, psnr(a, b))
    crop = torch.from_numpy(gold["out_crop"])
    e_crop = float((out[:, [0, T // 2, T - 1], :, 192:320, 192:320].clamp(-1, 1) - crop.clamp(-1, 1)).abs().max())
    e_mean = float((out.double().mean(dim=(2, 3, 4)) - torch.from_numpy(gold["out_mean"])).abs().max())
    _report("T20_codes_forced[tc3]", out_sub8=worst_e, psnr_min=worst_p, crop=e_crop, frame_mean=e_mean)
    assert worst_e <= 1e-2 and worst_p >= 50.0 and e_crop <= 1e-2 and e_mean < 1e-3
