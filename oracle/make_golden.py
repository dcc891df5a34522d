"""TEST INFRASTRUCTURE ONLY — pin the oracle against the real reference and write fixtures.

Run in the build container (needs /root/reference):   python oracle/make_golden.py

1. Builds the real reference `KEEP` (oracle/ref_loader.py), loads the seeded synthetic weights
   (oracle/weights.py, strict=True) and runs `keep_net(x, need_upscale=False)` on seeded clips.
2. Runs the restatement `oracle/keep_oracle.py` on the same weights/inputs and records how far it
   is from the reference (stage by stage) -> tests/golden/pin_report.json.
3. Also runs the reference in float64 to record the reference's own fp32 round-off floor
   (code-index agreement, top1-top2 logit margins).
4. Commits small fixtures (reference outputs, subsampled where large) to tests/golden/*.npz.
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import keep_oracle, ref_loader, weights  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def run_reference(net, x):
    cap = {"logits": []}
    hooks = [
        net.flownet.register_forward_hook(lambda m, i, o: cap.__setitem__("flows", o.detach().clone())),
        net.encoder.blocks[-1].register_forward_hook(lambda m, i, o: cap.__setitem__("z_codes", o.detach().clone())),
        net.kalman_filter.kalman_gain_calculator.register_forward_hook(
            lambda m, i, o: cap.__setitem__("gains", o.detach().clone())),
        net.idx_pred_layer.register_forward_hook(lambda m, i, o: cap["logits"].append(o.detach().clone())),
    ]
    with torch.no_grad():
        out = net(x, need_upscale=False)
    for h in hooks:
        h.remove()
    cap["logits"] = torch.stack([l.permute(1, 0, 2) for l in cap["logits"]], dim=1)  # (b,T,256,1024)
    cap["codes"] = cap["logits"].argmax(dim=3)
    return out, cap


def maxabs(a, b):
    return float((a.double() - b.double()).abs().max())


def psnr(a, b):
    a = a.double().clamp(-1, 1)
    b = b.double().clamp(-1, 1)
    mse = float(((a - b) ** 2).mean()) / 4.0  # [-1,1] -> [0,1] scale
    return 999.0 if mse == 0 else 10 * np.log10(1.0 / mse)


CASES = (  # fixture name, model config, T, coherent clip, clip seed
    ("T3_coherent", "KEEP", 3, True, 1234),
    ("T2_noise", "KEEP", 2, False, 1234),
    ("asian_T2_coherent", "Asian", 2, True, 1234),   # CFT at 32/64/128/256, none at 16
    ("T20_coherent", "KEEP", 20, True, 1234),        # BASELINE.json configs[1] at full size (compact fixture, see below)
    ("T12_coherent", "KEEP", 12, True, 1236),        # the 12-frame tail clip of BASELINE.json configs[4] (512 frames = 25 x 20 + 12)
)


def pin_vq(report):
    """Pin oracle vq_nearest against the real VectorQuantizer.forward (vqgan_arch.py:37-76) -> ref_vq.npz."""
    ref_loader._install_shims()
    from wm_basicsr.archs.vqgan_arch import VectorQuantizer
    sd = weights.make_state_dict(seed=0)
    cb, z = weights.make_vq_case(sd["quantize.embedding.weight"], n=2, seed=99)   # entry 700 duplicates entry 3: exact ties
    vq = VectorQuantizer(1024, 256, 0.25).eval()
    vq.embedding.weight.data.copy_(cb)
    with torch.no_grad():
        zq_ref, _, info = vq(z)
    idx_ref = info["min_encoding_indices"].view(-1)
    zq, idx, d = keep_oracle.vq_nearest(z, cb)
    top2 = d.topk(2, dim=1, largest=False).values
    case = {"tokens": int(idx.numel()), "index_agreement": float((idx == idx_ref).float().mean()),
            "zq_bit_equal": bool(torch.equal(zq, zq_ref)), "mean_distance_ref": float(info["mean_distance"]),
            "mean_distance_oracle": float(d.mean()), "unique_codes": int(idx_ref.unique().numel()),
            "margin_min": float((top2[:, 1] - top2[:, 0]).min()), "hits_duplicate_low": int((idx_ref == 3).sum()),
            "hits_duplicate_high": int((idx_ref == 700).sum())}
    report["cases"]["vq_nearest"] = case
    print("vq_nearest", json.dumps(case, indent=1))
    np.savez_compressed(os.path.join(GOLD, "ref_vq.npz"), idx=idx_ref.numpy().astype(np.int16),
                        top2=top2.numpy().astype(np.float32), zq_sum=zq_ref.double().sum(dim=(2, 3)).numpy(),
                        zq_crop=zq_ref[:, :8, :4, :4].numpy())


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    only = set(sys.argv[1:])
    report = {"torch": torch.__version__, "weights_seed": 0, "cases": {}}
    rp = os.path.join(GOLD, "pin_report.json")
    if only and os.path.exists(rp):
        with open(rp) as f:
            report = json.load(f)
    nets = {}
    for name, config, T, coherent, seed in CASES:
        if only and name not in only:
            continue
        if config not in nets:
            sd_c = weights.make_state_dict(seed=0, config=config)
            nets[config] = (sd_c, ref_loader.load_reference_keep(sd_c, config=config))
        sd, net = nets[config]
        x = weights.make_clip(T, seed=seed, coherent=coherent)
        t0 = time.time()
        ref_out, ref_cap = run_reference(net, x)
        t_ref = time.time() - t0
        t0 = time.time()
        ora_out, ora_cap = keep_oracle.keep_forward(sd, x, capture=True)
        t_ora = time.time() - t0
        case = {
            "config": config, "T": T, "coherent": coherent, "clip_seed": seed, "ref_seconds": t_ref, "oracle_seconds": t_ora,
            "oracle_vs_ref_maxabs": {
                "flows": maxabs(ora_cap["flows"].reshape(-1), ref_cap["flows"].reshape(-1)),
                "z_codes": maxabs(ora_cap["z_codes"].reshape(-1), ref_cap["z_codes"].reshape(-1)),
                "gains": maxabs(ora_cap["gains"].reshape(-1), ref_cap["gains"].reshape(-1)),
                "logits": maxabs(ora_cap["logits"], ref_cap["logits"]),
                "out": maxabs(ora_out, ref_out),
                "out_clamped": maxabs(ora_out.clamp(-1, 1), ref_out.clamp(-1, 1)),   # what tensor2img sees (img_util.py:56)
            },
            "oracle_vs_ref_code_agreement": float((ora_cap["codes"] == ref_cap["codes"]).float().mean()),
            "oracle_vs_ref_psnr_db": psnr(ora_out, ref_out),
            "ref_out_absmax": float(ref_out.abs().max()),
            "ref_out_std": float(ref_out.std()),
            "ref_flow_absmax": float(ref_cap["flows"].abs().max()),
            "ref_gain_range": [float(ref_cap["gains"].min()), float(ref_cap["gains"].max())],
        }
        top2 = ref_cap["logits"].topk(2, dim=3).values
        margin = (top2[..., 0] - top2[..., 1])
        case["ref_logit_margin"] = {"min": float(margin.min()), "median": float(margin.median()),
                                    "p01": float(margin.flatten().kthvalue(max(1, margin.numel() // 100)).values)}
        case["ref_unique_codes"] = int(ref_cap["codes"].unique().numel())
        # reference's own fp32 noise floor: same net in float64
        if name == "T20_coherent":
            # the same floor at full size: where do the fp32 reference and the fp64 restatement part ways, and at what margin
            sd64 = {k: v.double() for k, v in sd.items()}
            _, cap64 = keep_oracle.keep_forward(sd64, x.double(), capture=True)
            c64, rc = cap64["codes"][0], ref_cap["codes"][0]
            case["ref_fp32_vs_fp64"] = {
                "code_agreement_per_frame": [float((c64[i] == rc[i]).float().mean()) for i in range(T)],
                "worst_flipped_margin_per_frame": [float(margin[0, i][c64[i] != rc[i]].max()) if bool((c64[i] != rc[i]).any()) else 0.0
                                                   for i in range(T)],
            }
        if name == "T3_coherent":
            # (the reference itself cannot run in float64: matching.py:31 mixes a float32 grid in;
            #  the restatement, pinned above, is run in float64 instead)
            sd64 = {k: v.double() for k, v in sd.items()}
            out64, cap64 = keep_oracle.keep_forward(sd64, x.double(), capture=True)
            case["ref_fp32_vs_fp64"] = {
                "code_agreement_per_frame": [float((cap64["codes"][:, i] == ref_cap["codes"][:, i]).float().mean())
                                             for i in range(T)],
                "out_maxabs": maxabs(out64, ref_out), "psnr_db": psnr(out64, ref_out),
                "logits_maxabs": maxabs(cap64["logits"], ref_cap["logits"]),
            }
        report["cases"][name] = case
        print(name, json.dumps(case, indent=1))
        if T > 8:
            # full-size clip: compact fixture -- stride-8 pixels of every frame, a full-resolution crop of three frames, the
            # discrete decisions with their margins, the gains (temporal attention over all T frames) and per-frame
            # latent statistics plus the first / last latents in full
            zc = ref_cap["z_codes"].reshape(T, 256, 16, 16)
            np.savez_compressed(
                os.path.join(GOLD, "ref_%s.npz" % name),
                out_sub8=ref_out[:, :, :, ::8, ::8].numpy().astype(np.float32),
                out_crop=ref_out[:, [0, T // 2, T - 1], :, 192:320, 192:320].numpy().astype(np.float32),
                out_mean=ref_out.double().mean(dim=(2, 3, 4)).numpy(),
                gains=ref_cap["gains"].numpy().astype(np.float32),
                codes=ref_cap["codes"].numpy().astype(np.int16),
                logit_top2=top2.numpy().astype(np.float32),
                z_mean=zc.double().mean(dim=(1, 2, 3)).numpy(), z_sqmean=(zc.double() ** 2).mean(dim=(1, 2, 3)).numpy(),
                z_first=zc[0].numpy().astype(np.float32), z_last=zc[T - 1].numpy().astype(np.float32),
                flows_sub16=ref_cap["flows"].reshape(1, T - 1, 2, 512, 512)[:, :, :, ::16, ::16].numpy().astype(np.float32),
            )
            continue
        # fixtures: full-res output is 3 MB/frame fp32 -> keep a stride-4 subsample + fp16 copy of frame 0 crop
        np.savez_compressed(
            os.path.join(GOLD, "ref_%s.npz" % name),
            out_sub4=ref_out[:, :, :, ::4, ::4].numpy().astype(np.float32),
            out_crop=ref_out[:, :, :, 192:320, 192:320].numpy().astype(np.float32),
            out_mean=ref_out.double().mean(dim=(2, 3, 4)).numpy(),
            out_sqmean=(ref_out.double() ** 2).mean(dim=(2, 3, 4)).numpy(),
            flows_sub8=ref_cap["flows"].reshape(1, T - 1, 2, 512, 512)[:, :, :, ::8, ::8].numpy().astype(np.float32),
            z_codes=ref_cap["z_codes"].numpy().astype(np.float32),
            gains=ref_cap["gains"].numpy().astype(np.float32),
            codes=ref_cap["codes"].numpy().astype(np.int16),
            logit_top2=top2.numpy().astype(np.float32),
        )
    if not only or "vq_nearest" in only:
        pin_vq(report)
    with open(rp, "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
