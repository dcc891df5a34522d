"""CPU tier: oracle vs golden fixtures, host logic, C-ABI surface. No compute calls on the library."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_pin_report_says_oracle_matches_reference():
    rep = json.load(open(os.path.join(GOLD, "pin_report.json")))
    vq = rep["cases"]["vq_nearest"]
    assert vq["index_agreement"] == 1.0 and vq["zq_bit_equal"] and vq["hits_duplicate_low"] == 8 and vq["hits_duplicate_high"] == 0
    for name, case in rep["cases"].items():
        if name == "vq_nearest":
            continue
        assert case["oracle_vs_ref_code_agreement"] == 1.0, name
        # unclamped outputs of the synthetic-weight nets reach |70| (Asian: CFT at 256^2); the pixel bar is on the clamped range
        assert case["oracle_vs_ref_maxabs"]["out"] < 1e-3 * max(1.0, case["ref_out_absmax"] / 16), name
        assert case["oracle_vs_ref_maxabs"].get("out_clamped", case["oracle_vs_ref_maxabs"]["out"]) < 1e-3, name
        assert case["oracle_vs_ref_maxabs"]["flows"] < 1e-3, name
        assert case["oracle_vs_ref_psnr_db"] > 90.0, name


def test_oracle_reproduces_reference_golden_T2(state_dict):
    """The committed fixture was produced by the REAL reference (oracle/make_golden.py); the restatement must
    reproduce it from the seeded weights/inputs alone."""
    from oracle import keep_oracle, weights
    g = np.load(os.path.join(GOLD, "ref_T2_noise.npz"))
    x = weights.make_clip(2, seed=1234, coherent=False)
    torch.set_num_threads(os.cpu_count() or 1)
    out, cap = keep_oracle.keep_forward(state_dict, x, capture=True)
    assert np.array_equal(cap["codes"].numpy().astype(np.int16), g["codes"])
    np.testing.assert_allclose(cap["z_codes"].reshape(g["z_codes"].shape).numpy(), g["z_codes"], atol=2e-4)
    np.testing.assert_allclose(cap["gains"].reshape(g["gains"].shape).numpy(), g["gains"], atol=2e-5)
    np.testing.assert_allclose(cap["flows"][:, :, :, ::8, ::8].numpy(), g["flows_sub8"], atol=2e-2)
    np.testing.assert_allclose(out[:, :, :, ::4, ::4].numpy(), g["out_sub4"], atol=2e-3)


def test_weights_are_reproducible_and_complete(state_dict):
    from oracle import weights
    shapes = weights.load_shapes()
    assert len(shapes) == 896
    assert sum(int(np.prod(s)) for s in shapes.values()) == 158485860          # 158.49 M, as KEEP(**cfg).state_dict()
    asian = weights.load_shapes("Asian")
    assert len(asian) == 914 and sum(int(np.prod(s)) for s in asian.values()) == 143573092
    sd2 = weights.make_state_dict(seed=0)
    for k in ("encoder.blocks.0.weight", "cft.16.scale.0.weight", "position_emb"):
        assert torch.equal(state_dict[k], sd2[k])
    # every path is live: no zero-initialised tensor
    assert all(float(v.abs().max()) > 0 for v in state_dict.values())


def _header_functions():
    src = open(os.path.join(ROOT, "include", "keep_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(keep(?:op)?_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = _header_functions()
    assert "keep_create" in names and "keep_forward" in names and "keep_destroy" in names
    for n in names:
        assert hasattr(lib, n), "libkeep_b200.so does not export %s" % n


def test_plan_only_engine_checks_keys_and_sizes_workspace(keep_mod, lib, state_dict):
    net = keep_mod.KeepNetB200()
    net.load_state_dict(state_dict, strict=True)
    h = net._make_engine(flags=256)  # KEEP_FLAG_PLAN_ONLY: no device needed
    w2, w3, w20 = (lib.keep_workspace_bytes(h, 1, T) for T in (2, 3, 20))
    assert 0 < w2 <= w3 <= w20 < 8 << 30
    assert lib.keep_workspace_bytes(h, 1, 1) == 0 and b"T" in lib.keep_last_error()
    # forward on a plan-only engine must fail loudly, not fall back
    rc = lib.keep_forward(h, ctypes.c_void_p(256), 1, 2, ctypes.c_void_p(256), 0, None, 0, None)
    assert rc != 0 and b"PLAN_ONLY" in lib.keep_last_error()
    net._drop_engine()


def test_default_flags_and_rejected_flags(keep_mod, lib, state_dict):
    """The default-constructed module is the measured engine (tcgen05 split precision + CUDA graph), 'Asian' adds the wide
    operand range; the reserved fp16-feature flag is refused by the C side instead of failing inside the first forward."""
    kn = keep_mod.keep_net
    assert kn.DEFAULT_FLAGS == kn.FLAG_TCGEN05 | kn.FLAG_TC_SPLIT3 | kn.FLAG_TC_WIDE | kn.FLAG_CUDA_GRAPH
    assert keep_mod.KeepNetB200()._flags == kn.DEFAULT_FLAGS
    assert keep_mod.KeepNetB200(flags=0)._flags == 0                                  # exact-fp32 CUDA-core engine, on request
    assert keep_mod.KeepNetB200(**kn.KEEP_ASIAN_CFG)._flags == kn.DEFAULT_FLAGS
    net = keep_mod.KeepNetB200()
    net.load_state_dict(state_dict, strict=True)
    for bad in (kn.FLAG_FP16_FEATURES, kn.FLAG_FP16_FEATURES | kn.FLAG_TCGEN05 | kn.FLAG_TC_SPLIT3):
        with pytest.raises(RuntimeError, match="FP16_FEATURES"):
            net._make_engine(flags=bad | kn.FLAG_PLAN_ONLY)
    # a plan-only module walks the residency calls without a device (what the reference-host lifecycle test relies on)
    po = keep_mod.KeepNetB200(plan_only=True)
    po.load_state_dict(state_dict, strict=True)
    po.to("cuda")
    assert po._engine is not None and lib.keep_workspace_bytes(po._engine, 1, 2) > 0
    st = ctypes.c_int(-1)
    assert lib.keep_status(po._engine, 1, ctypes.byref(st)) == 0 and st.value == 0
    with pytest.raises(RuntimeError):
        po(torch.zeros(1, 2, 3, 512, 512), need_upscale=False)
    po.to("cpu")
    assert po._engine is None


def test_strict_state_dict_and_no_cpu_fallback(keep_mod, state_dict):
    net = keep_mod.KeepNetB200()
    bad = dict(state_dict)
    bad.pop("cfa.16.norm1.weight")
    with pytest.raises(RuntimeError, match="Missing key"):
        net.load_state_dict(bad, strict=True)
    bad = dict(state_dict)
    bad["feat_emb.weight"] = torch.zeros(3, 3)
    with pytest.raises(RuntimeError, match="size mismatch"):
        net.load_state_dict(bad, strict=True)
    net.load_state_dict(state_dict, strict=True)
    assert set(net.state_dict().keys()) == set(state_dict.keys())
    with pytest.raises(RuntimeError, match="no device engine|no CPU fallback"):
        net(torch.zeros(1, 2, 3, 512, 512), need_upscale=False)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device"):
            net.to("cuda")
    with pytest.raises(ValueError):
        keep_mod.KeepNetB200(n_layers=4)


def test_missing_key_is_rejected_by_the_c_side(keep_mod, lib, state_dict):
    net = keep_mod.KeepNetB200()
    net.load_state_dict(state_dict, strict=True)
    del net._weights["quantize.embedding.weight"]
    with pytest.raises(RuntimeError, match="quantize.embedding.weight"):
        net._make_engine(flags=256)


# ---- 'Asian' model config (modules/utils.py:58-73) ----------------------------------------------------

def test_oracle_reproduces_reference_golden_asian_T2(state_dict_asian):
    """Fixture produced by the REAL reference built with the 'Asian' config (CFT after the 32^2..256^2 generator levels)."""
    from oracle import keep_oracle, weights
    g = np.load(os.path.join(GOLD, "ref_asian_T2_coherent.npz"))
    assert keep_oracle.fusion_lists(state_dict_asian) == (["32", "64", "128", "256"], ["16", "32"])
    x = weights.make_clip(2, seed=1234, coherent=True)
    torch.set_num_threads(os.cpu_count() or 1)
    out, cap = keep_oracle.keep_forward(state_dict_asian, x, capture=True)
    assert sorted(cap["enc_feat"]) == ["128", "256", "32", "64"]
    assert np.array_equal(cap["codes"].numpy().astype(np.int16), g["codes"])
    np.testing.assert_allclose(cap["z_codes"].reshape(g["z_codes"].shape).numpy(), g["z_codes"], atol=2e-4)
    np.testing.assert_allclose(cap["gains"].reshape(g["gains"].shape).numpy(), g["gains"], atol=2e-5)
    np.testing.assert_allclose(out[:, :, :, ::4, ::4].clamp(-1, 1).numpy(), np.clip(g["out_sub4"], -1, 1), atol=2e-3)


def test_asian_config_host_mirror_and_plan_only_engine(keep_mod, lib, state_dict, state_dict_asian):
    kn = keep_mod.keep_net
    assert kn.config_name(dict(kn.KEEP_GENERAL_CFG)) == "KEEP"
    assert kn.config_name(dict(kn.KEEP_ASIAN_CFG, temp_reg_list=[])) == "Asian"
    with pytest.raises(ValueError):
        kn.config_name(dict(cft_list=["16", "64"]))
    assert len(state_dict_asian) == 914 and "cft.16.scale.0.weight" not in state_dict_asian
    net = keep_mod.KeepNetB200(cft_list=["32", "64", "128", "256"])
    assert net.config == "Asian"
    with pytest.raises(RuntimeError, match="Missing key|Unexpected key"):
        net.load_state_dict(state_dict, strict=True)          # general weights into the Asian programme
    net.load_state_dict(state_dict_asian, strict=True)
    h = net._make_engine(flags=256)  # KEEP_FLAG_PLAN_ONLY: dry run of the Asian programme (taps at 32..256, CFA without CFT at 16)
    gen = keep_mod.KeepNetB200()
    gen.load_state_dict(state_dict, strict=True)
    hg = gen._make_engine(flags=256)
    wa, wg = lib.keep_workspace_bytes(h, 1, 20), lib.keep_workspace_bytes(hg, 1, 20)
    # the Asian programme keeps (T,128,128,128) and (T,256,256,128) encoder taps resident: 0.84 GB more fp32 at T = 20
    assert wa > wg and wa - wg > 20 * (128 * 128 * 128 + 256 * 256 * 128 - 16 * 16 * 512) * 4 * 0.9
    assert wa < 12 << 30
    net._drop_engine()
    gen._drop_engine()
    # an incomplete CFT block is rejected by the C side by name
    bad = keep_mod.KeepNetB200(cft_list=["32", "64", "128", "256"])
    bad.load_state_dict(state_dict_asian, strict=True)
    del bad._weights["cft.256.shift.2.weight"]
    with pytest.raises(RuntimeError, match="cft.256.shift.2.weight"):
        bad._make_engine(flags=256)


# ---- nearest-neighbour quantiser (vqgan_arch.py:37-76) -------------------------------------------------

def test_oracle_vq_nearest_reproduces_reference_fixture(state_dict):
    """ref_vq.npz holds what the REAL VectorQuantizer.forward returned (oracle/make_golden.py::pin_vq)."""
    from oracle import keep_oracle, weights
    g = np.load(os.path.join(GOLD, "ref_vq.npz"))
    cb, z = weights.make_vq_case(state_dict["quantize.embedding.weight"], n=2, seed=99)
    zq, idx, d = keep_oracle.vq_nearest(z, cb)
    assert np.array_equal(idx.numpy().astype(np.int16), g["idx"])
    assert int((idx[:8] == 3).sum()) == 8 and int((idx == 700).sum()) == 0      # exact ties -> lowest index
    np.testing.assert_allclose(d.topk(2, dim=1, largest=False).values.numpy(), g["top2"], rtol=0, atol=1e-4)
    assert np.array_equal(zq[:, :8, :4, :4].numpy(), g["zq_crop"])
    np.testing.assert_allclose(zq.double().sum(dim=(2, 3)).numpy(), g["zq_sum"], atol=1e-9)


def test_vector_quantize_has_no_cpu_fallback(keep_mod, state_dict):
    cb = state_dict["quantize.embedding.weight"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        keep_mod.vector_quantize(torch.zeros(1, 256, 16, 16), cb)


# ---- lockstep clip batching: host-side plan ------------------------------------------------------------

def test_lockstep_plan_only_engine(keep_mod, lib, state_dict):
    kn = keep_mod.keep_net
    net = keep_mod.KeepNetB200(batch_clips=2)
    assert net._flags & kn.FLAG_BATCH_CLIPS
    net.load_state_dict(state_dict, strict=True)
    h = net._make_engine(flags=256 | kn.FLAG_BATCH_CLIPS | kn.FLAG_TCGEN05 | kn.FLAG_TC_SPLIT3)
    w1, w2, w3 = (lib.keep_workspace_bytes(h, b, 20) for b in (1, 2, 3))
    assert 0 < w1 < w2 == w3 < 2 * w1          # groups of 2: resident taps / latents double, the encoder passes do not
    assert lib.keep_set_batch_clips(h, 4) == 0
    assert lib.keep_workspace_bytes(h, 4, 20) > w2
    assert lib.keep_set_batch_clips(h, 0) != 0 and b"max_clips" in lib.keep_last_error()
    assert lib.keep_set_batch_clips(None, 2) != 0
    net._drop_engine()


# ---- host-side plan (keep_plan_dump): control flow of both configs and of the lockstep path, no device --------------------

def _plan(keep_mod, lib, sd, clips, T, flags, tmp_path, **cfg):
    import collections
    net = keep_mod.KeepNetB200(flags=flags, **cfg)
    net.load_state_dict(sd, strict=True)
    h = net._make_engine(flags=256 | net._flags)
    path = os.path.join(str(tmp_path), "plan_%d_%d_%d.txt" % (clips, T, flags))
    assert lib.keep_plan_dump(h, clips, T, path.encode()) == 0, lib.keep_last_error()
    net._drop_engine()
    lines = open(path).read().splitlines()
    kinds = collections.Counter(l.split()[0] for l in lines)
    return lines, kinds


def test_plan_per_frame_chain_and_kernel_choice(keep_mod, lib, state_dict, tmp_path):
    kn = keep_mod.keep_net
    tc3 = kn.FLAG_TCGEN05 | kn.FLAG_TC_SPLIT3
    l2, k2 = _plan(keep_mod, lib, state_dict, 1, 2, tc3, tmp_path)
    l3, k3 = _plan(keep_mod, lib, state_dict, 1, 3, tc3, tmp_path)
    l4, k4 = _plan(keep_mod, lib, state_dict, 1, 4, tc3, tmp_path)
    _, k5 = _plan(keep_mod, lib, state_dict, 1, 5, tc3, tmp_path)
    # one more frame = one more pass of the serial chain (hq_encoder + code transformer + generator with CFT / CFA); GMFlow runs
    # in chunks of 2 pairs (KEEP_FLOW_CHUNK), so T = 2 and 3 share one chunk, T = 4 and 5 need a second one
    per_frame = {k: k3[k] - k2[k] for k in k3}
    per_chunk = {k: k4[k] - k3[k] - per_frame[k] for k in k4}
    assert {k: k5[k] - k4[k] for k in k5} == per_frame and per_chunk["attention"] == 12 and per_chunk["conv"] > 0
    assert per_frame["conv"] == 168 and per_frame["attention"] == 17 and per_frame["layernorm"] == 23
    # every GEMM-shaped layer runs on the tcgen05 kernel in the tensor-core modes; the stems / heads on their own kernels
    for l in l2:
        if l.startswith("conv"):
            f = dict(kv.split("=") for kv in l.split()[1:])
            strided_1x1 = f["k"] == "1" and f["stride"] != "1"      # GMFlow's three downsample shortcuts stay on CUDA cores
            if int(f["c0"]) + int(f["c1"]) >= 32 and int(f["cout"]) % 16 == 0 and not strided_1x1:
                assert f["kernel"] == "tcgen05", l
            assert f["wide"] == "0", l                              # general config: fp16 pairs everywhere
    lf, kf = _plan(keep_mod, lib, state_dict, 1, 2, 0, tmp_path)
    # fp32 engine mode: same programme on CUDA-core kernels (its batched attention GEMMs are not traced as "gemm" lines)
    # (tc3 also traces GMFlow's 12 window attentions per chunk of pairs: they run on the fused tcgen05 attention kernel there)
    fused = [l for l in l2 if l.startswith("attention") and "kernel=tcgen05_fused" in l]
    gm = [l for l in fused if "Lq=1024 Lk=1024 heads=1 dh=128" in l]
    ft = [l for l in fused if "Lq=256 Lk=256 heads=8 dh=64" in l]          # code transformer: 9 layers per frame, same kernel
    assert len(gm) == 12 and len(ft) == 2 * 9 and len(fused) == len(gm) + len(ft)
    assert {k: v - (12 if k == "attention" else 0) for k, v in k2.items() if k != "gemm"} == dict(kf) and not any("tcgen05" in l for l in lf)


def test_plan_norms_ride_on_the_producing_kernels(keep_mod, lib, state_dict, tmp_path):
    """GroupNorm statistics come from the producing conv / split-K reduce ('gnstats=1'; with KEEP_GN_REDUCE_FINAL=1 split-K layers
    with few slots also finalize them inside the reduce: 'gnstats=2' + 'groupnorm ... fused=2', no finalize launch), and every LayerNorm of the code
    transformer is written by the preceding linear's reduce kernel ('layernorm ... fused=1') -- per additional frame of the chain."""
    kn = keep_mod.keep_net
    tc3 = kn.FLAG_TCGEN05 | kn.FLAG_TC_SPLIT3
    l2, _ = _plan(keep_mod, lib, state_dict, 1, 2, tc3, tmp_path)
    l3, _ = _plan(keep_mod, lib, state_dict, 1, 3, tc3, tmp_path)

    def count(lines, pred):
        return sum(1 for l in lines if pred(l))

    per_frame = lambda pred: count(l3, pred) - count(l2, pred)
    gn_all = per_frame(lambda l: l.startswith("groupnorm"))
    gn_fin_in_reduce = per_frame(lambda l: l.startswith("groupnorm") and l.endswith("fused=2"))
    gn_fin_launch = per_frame(lambda l: l.startswith("groupnorm") and l.endswith("fused=1"))
    assert per_frame(lambda l: l.startswith("conv") and l.endswith("gnstats=2")) == gn_fin_in_reduce
    assert per_frame(lambda l: l.startswith("conv") and l.endswith("gnstats=1")) == gn_fin_launch
    # (finalize inside the reduce is opt-in -- KEEP_GN_REDUCE_FINAL=1 -- since the separate tiny launch measured 0.6 % faster)
    assert gn_fin_in_reduce == (30 if os.environ.get("KEEP_GN_REDUCE_FINAL") == "1" else 0)
    assert gn_fin_in_reduce + gn_fin_launch >= 0.85 * gn_all
    # code transformer: feat_emb -> norm1, 9 x norm2, 8 x the next layer's norm1, idx_pred_layer.0
    assert per_frame(lambda l: l.startswith("layernorm") and "rows=256 c=512" in l and l.endswith("fused=1")) == 19
    # (the only stand-alone LayerNorms of that shape left are the two post-norms of the 16^2 cross-frame attention block)
    assert per_frame(lambda l: l.startswith("layernorm") and "rows=256 c=512" in l and not l.endswith("fused=1")) == 2


def test_plan_lockstep_shares_the_chain_and_asian_adds_a_cft(keep_mod, lib, state_dict, state_dict_asian, tmp_path):
    kn = keep_mod.keep_net
    tc3 = kn.FLAG_TCGEN05 | kn.FLAG_TC_SPLIT3
    _, s2 = _plan(keep_mod, lib, state_dict, 1, 2, tc3, tmp_path)
    _, s3 = _plan(keep_mod, lib, state_dict, 1, 3, tc3, tmp_path)
    lk2, b2 = _plan(keep_mod, lib, state_dict, 2, 2, tc3, tmp_path, batch_clips=2)
    _, b3 = _plan(keep_mod, lib, state_dict, 2, 3, tc3, tmp_path, batch_clips=2)
    # two clips in lockstep: the per-frame chain runs once per frame index (at batch 2), not once per clip
    assert {k: b3[k] - b2[k] for k in b3} == {k: s3[k] - s2[k] for k in s3}
    assert b2["conv"] < 2 * s2["conv"]
    assert any(l.startswith("conv n=2 h=512 w=512 c0=3 ") for l in lk2)                 # hq_encoder stem at batch 2
    assert any(l.startswith("attention nb=2 Lq=256 Lk=256 heads=8 dh=64") for l in lk2)   # code transformer at batch 2
    # 'Asian': four CFT blocks instead of three (7 convs + 2 GroupNorms each per frame), wide operands on raw-input layers
    la, a2 = _plan(keep_mod, lib, state_dict_asian, 1, 2, tc3, tmp_path, **kn.KEEP_ASIAN_CFG)
    assert a2["conv"] - s2["conv"] == 2 * 7 and a2["groupnorm"] - s2["groupnorm"] == 2 * 2
    assert any("c0=128 c1=128 cout=128 k=3" in l and "h=256" in l for l in la)        # cat[enc, dec] at 256^2
    wide = [l for l in la if "wide=1" in l]
    assert wide and all(" pre=0 " in l and "kernel=tcgen05" in l for l in wide)
    assert any("up=2" in l for l in wide) and not any(" w=1 " in l for l in wide)


def _plan_gflop(lines):
    """Algorithmic GFLOP (2 x MAC) of the conv / linear / GEMM / attention ops of a plan."""
    tot = 0.0
    for l in lines:
        kind = l.split()[0]
        f = dict(kv.split("=") for kv in l.split()[1:])
        if kind == "conv":
            n, h, w, c0, c1, co, k, st, up = (int(f[x]) for x in ("n", "h", "w", "c0", "c1", "cout", "k", "stride", "up"))
            tot += 2.0 * n * (h * up // st) * (w * up // st) * co * k * k * (c0 + c1)
        elif kind == "gemm":
            tot += 2.0 * int(f["nb"]) * int(f["M"]) * int(f["K"]) * int(f["N"])
        elif kind == "attention":
            tot += 4.0 * int(f["nb"]) * int(f["heads"]) * int(f["Lq"]) * int(f["Lk"]) * int(f["dh"])
    return tot / 1e9


def test_engine_executes_the_reference_flop_count(keep_mod, lib, state_dict, tmp_path):
    """No work skipped: the engine's planned conv / GEMM / attention FLOPs equal what torch's FlopCounterMode counts for the
    reference forward -- FLOPs(T) = 1058.97 T - 410.13 GFLOP -- to 0.1 %.  The remainder is the one-hot @
    codebook matmul the engine replaces by a row gather (0.13 GFLOP / frame), the two V-in-R^2 softmax expectations of GMFlow
    and the gain estimator's T x T temporal attention, which run outside the traced GEMM family."""
    kn = keep_mod.keep_net
    for T in (2, 5, 20):
        lines, _ = _plan(keep_mod, lib, state_dict, 1, T, kn.FLAG_TCGEN05 | kn.FLAG_TC_SPLIT3, tmp_path)
        want = 1058.97 * T - 410.13
        got = _plan_gflop(lines)
        assert abs(got / want - 1.0) < 1e-3, (T, got, want)
        assert got <= want                       # never more than the reference either (nothing recomputed)


def test_gmflow_fused_projection_plan_keeps_the_flops(tmp_path):
    """KEEP_GM_FUSE_QKV=1 (opt-in): q|k|v (self-attention) and k|v (cross-attention) projections of GMFlow as one GEMM over fused
    weights -- fewer, wider GEMMs, the same FLOPs.  The knob is read once per process, hence the subprocess."""
    import subprocess
    import sys
    code = (
        "import sys, os, collections; sys.path.insert(0, %r)\n"
        "import keep_b200\n"
        "sys.path.insert(0, os.path.join(%r, 'tests'))\n"
        "lib = keep_b200.keep_net.load_library()\n"
        "net = keep_b200.KeepNetB200(flags=6); net.load_state_dict(keep_b200.synth.make_state_dict(seed=0), strict=True)\n"
        "h = net._make_engine(flags=256 | 6)\n"
        "p = os.path.join(%r, 'plan_fuse.txt')\n"
        "assert lib.keep_plan_dump(h, 1, 3, p.encode()) == 0\n"
        "L = open(p).read().splitlines()\n"
        "import test_cpu_host as t\n"
        "print(len([l for l in L if l.startswith('conv')]), sum(' c0=128 c1=0 cout=384 k=1 ' in l and ' h=4096 ' in l for l in L), sum(' c0=128 c1=0 cout=256 k=1 ' in l and ' h=4096 ' in l for l in L), '%%.3f' %% t._plan_gflop(L))\n"
    ) % (ROOT, ROOT, str(tmp_path))
    out = {}
    for knob in ("0", "1"):
        env = dict(os.environ, KEEP_GM_FUSE_QKV=knob)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[knob] = r.stdout.strip().splitlines()[-1].split()
    n0, q0, kv0, g0 = int(out["0"][0]), int(out["0"][1]), int(out["0"][2]), float(out["0"][3])
    n1, q1, kv1, g1 = int(out["1"][0]), int(out["1"][1]), int(out["1"][2]), float(out["1"][3])
    assert q0 == 0 and kv0 == 0 and q1 == 6 and kv1 == 6   # T = 3: one GMFlow pass, 6 self-attention + 6 cross-attention layers
    assert n1 == n0 - 6 * 2 - 6 * 1                    # three projections -> one (self), two -> one (cross)
    assert abs(g1 - g0) < 1e-6 * g0


def test_maximum_clip_length_fits_the_gpu(keep_mod, lib, state_dict, state_dict_asian):
    """T = 100 is the node's upper bound for max_clip_length (nodes.py): the workspace plan must fit a 180 GB B200 many times
    over, in both configs and for the largest lockstep group; T = 101 and T = 1 are rejected by name."""
    kn = keep_mod.keep_net
    for sd, cfg in ((state_dict, {}), (state_dict_asian, kn.KEEP_ASIAN_CFG)):
        net = keep_mod.KeepNetB200(flags=kn.FLAG_TCGEN05 | kn.FLAG_TC_SPLIT3, batch_clips=4, **cfg)
        net.load_state_dict(sd, strict=True)
        h = net._make_engine(flags=256 | net._flags)
        w1, w4 = lib.keep_workspace_bytes(h, 1, 100), lib.keep_workspace_bytes(h, 4, 100)
        assert 0 < w1 < w4 < 32 << 30
        assert lib.keep_workspace_bytes(h, 1, 101) == 0 and b"T <= 100" in lib.keep_last_error()
        assert lib.keep_workspace_bytes(h, 1, 1) == 0
        net._drop_engine()
