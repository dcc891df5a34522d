#!/usr/bin/env python
"""bench.py — KEEP hot path (`keep_net(clip, need_upscale=False)`) on N B200s of one node.

Metric (BASELINE.json): aligned 512x512 face frames/sec on synthetic 20-frame clips (configs[1]), with
  value        frames/s with the clip already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e          frames/s through the plugin call with HOST (pinned) buffers: H2D of the clip + D2H of the
               decoded frames inside the timed region
  roofline     dominant kernel family (conv / linear implicit GEMM): algorithmic FLOPs / per-launch CUDA-event
               time on the launching stream, vs the measured tensor peak in MEASURED_PEAKS.json
  cpu_baseline the oracle port (fp32 torch restatement of the reference forward) on the host cores (N=1 only)

One step = one 20-frame clip per rank (weak scaling: clips are independent, keep_processor.py:263-270);
for N>1 the decoded frames of every rank are gathered to rank 0 over NCCL inside the timed step.

`--impl reference` times the reference's own CPU implementation of the path (the oracle port — the reference
is Python and /root/reference does not exist on the GPU box) on a bounded sample (one T=2 clip per step).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T_CLIP = 20
# Engine mode benchmarked by default: the fastest one that meets the parity bar (tests/test_gpu_parity.py):
#   fp32 = exact CUDA-core kernels; tc3 = tcgen05 with split-precision operands (fp32-grade); tc = tcgen05 fp16 operands
DEFAULT_MODE = "tc3"
DTYPE_OF = {"fp32": "f32", "tc3": "f16x2 split operands, fp32 accumulate (fp32-grade)", "tc": "f16 operands, fp32 accumulate"}


def clip_flops(T):
    """Algorithmic FLOPs of the reference forward on one clip (torch FlopCounterMode): 2*MAC."""
    return (1058.97 * T - 410.13) * 1e9


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops_sustained": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"], "hbm_gbs": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": busy[len(busy) // 2] if busy else 0.0, "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_sample(T=2, seed=1234):
    """Oracle port on the host cores: one T-frame clip (bounded sample of the 20-frame workload)."""
    import torch
    import keep_b200
    from oracle import keep_oracle
    torch.set_num_threads(os.cpu_count() or 1)
    sd = keep_b200.synth.make_state_dict(seed=0)
    x = keep_b200.synth.make_clip(T, seed=seed, coherent=True)
    return sd, x, keep_oracle


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port), rank 0 only."""
    if rank != 0:
        return
    import torch
    sd, x, keep_oracle = cpu_baseline_sample(T=2)
    for _ in range(max(1, min(args.warmup, 1))):
        keep_oracle.keep_forward(sd, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        keep_oracle.keep_forward(sd, x)
    dt = time.perf_counter() - t0
    fps = 2 * args.steps / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "512x512 aligned-face frames/sec through keep_net (20-frame clips)", "value": fps,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "KEEP general model, aligned 512x512 clip, T=2 sample of the 20-frame clip per step (CPU)",
                   "weights": "seeded synthetic (no checkpoint offline)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "T=2 clip per step, oracle port of keep_arch.py:1008-1145 in torch fp32 on the host cores"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=T_CLIP)
    ap.add_argument("--mode", default=os.environ.get("KEEP_BENCH_MODE", "auto"), choices=["auto", "fp32", "tc", "tc3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clips-per-step", type=int, default=1,
                    help="clips per GPU per step (default 1 = BASELINE configs[1]); > 1 is the stream-of-clips workload (configs[4]): "
                         "independent clips, two in flight per GPU on engine replicas (KeepNetB200(concurrent_clips=2))")
    ap.add_argument("--batch-clips", type=int, default=int(os.environ.get("KEEP_BENCH_BATCH", "1")),
                    help="with --clips-per-step > 1: clips per lockstep group inside one engine (KeepNetB200(batch_clips=N)) instead of "
                         "engine replicas on separate streams")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying the per-clip CUDA graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import keep_b200
    keep_b200.build()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    T = args.frames
    flags = 0
    mode = args.mode
    if mode == "auto":
        mode = os.environ.get("KEEP_DEFAULT_MODE", DEFAULT_MODE)
    if mode == "tc":
        flags |= keep_b200.keep_net.FLAG_TCGEN05
    elif mode == "tc3":
        flags |= keep_b200.keep_net.FLAG_TCGEN05 | keep_b200.keep_net.FLAG_TC_SPLIT3
    if not args.no_graph:
        flags |= keep_b200.keep_net.FLAG_CUDA_GRAPH
    B = max(1, args.clips_per_step)
    batch = min(B, max(1, args.batch_clips))
    net = keep_b200.KeepNetB200(flags=flags, batch_clips=batch,
                                concurrent_clips=min(B, int(os.environ.get("KEEP_BENCH_REPLICAS", "2"))) if (B > 1 and B > batch) else 1)
    net.load_state_dict(keep_b200.synth.make_state_dict(seed=0), strict=True)
    net.eval().to(dev)
    x_host = torch.cat([keep_b200.synth.make_clip(T, seed=1234 + rank + 100 * i, coherent=True) for i in range(B)], 0).pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    x = x_host.to(dev)
    gather_buf = None
    if world > 1 and rank == 0:
        gather_buf = [torch.empty((B, T, 3, 512, 512), dtype=torch.float16, device=dev) for _ in range(world)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(inp):
        out = net(inp, need_upscale=False)
        if world > 1:  # the one collective of the path: decoded frames to rank 0 (fp16, 31.5 MB per clip)
            dist.gather(out.to(torch.float16), gather_buf, dst=0)
        return out

    for _ in range(args.warmup):
        step(x)
    torch.cuda.synchronize()

    # ---- timed region: K steps, device-timed per step, L2 flushed between steps (outside the timed spans)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = net.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(x)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = net.launch_count() - l0
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * T * args.steps / (ms * 1e-3)

    # ---- e2e through the plugin call with host buffers (H2D + D2H inside the timed span)
    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        out = step(xd)
        out_host.copy_(out, non_blocking=True)

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    k2 = max(2, min(args.steps, 5))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k2):
        e2e_step()
    b.record()
    torch.cuda.synchronize()
    t2 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * T * k2 / (float(t2.item()) * 1e-3)
    nbytes = x_host.numel() * 4

    # ---- the same call with uint8 BGR crops in / out (keep_forward_u8): the host-side img2tensor /
    # normalize / tensor2img of keep_processor.py folded into the device path, a quarter of the host<->device bytes
    u8_host = ((x_host.permute(0, 1, 3, 4, 2).flip(-1) * 0.5 + 0.5) * 255.0).round().clamp(0, 255).to(torch.uint8).contiguous().pin_memory()
    u8_out_host = torch.empty_like(u8_host).pin_memory()

    def e2e_u8_step():
        out = net.forward_u8(u8_host.to(dev, non_blocking=True))
        u8_out_host.copy_(out, non_blocking=True)

    e2e_u8_step(); e2e_u8_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k2):
        e2e_u8_step()
    b.record()
    torch.cuda.synchronize()
    t3 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    e2e_u8_value = world * B * T * k2 / (float(t3.item()) * 1e-3)

    # ---- roofline leg: per-launch CUDA events around every conv/GEMM launch (one extra clip, rank 0's view)
    pk = peaks()
    net.profile(True)
    step(x)
    prof = net.profile_read()
    dump = os.path.join(tempfile.gettempdir(), "keep_layers_%d.csv" % os.getpid())
    net.profile_dump(dump)
    net.profile(False)
    # the single most expensive layer shape of the family (per-launch numbers, same CUDA-event timing)
    import collections
    import csv as _csv
    shapes = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in _csv.DictReader(open(dump)):
        k = (int(r["tag"]), int(r["M"]), int(r["K"]), int(r["N"]))
        shapes[k][0] += 1; shapes[k][1] += float(r["ms"]); shapes[k][2] += float(r["gflop"])
    os.unlink(dump)
    dom_k, dom_v = max(shapes.items(), key=lambda kv: kv[1][1])
    ncu_path = os.path.join(ROOT, "profiles", "r1_ncu_conv_tc3_v19_512x512_64_64.json")
    traffic, traffic_note = None, None
    if os.path.exists(ncu_path):
        l0 = json.load(open(ncu_path))["launches"][0]
        traffic = (float(l0["dram__bytes_read.sum"].split()[0]) + float(l0["dram__bytes_write.sum"].split()[0])) * 1e6
        traffic_note = ("dram read+write bytes of ONE conv_tc_kernel<3,false> launch at M=524288 K=576 N=64 (two 512x512 frames, 64->64 3x3; "
                        "algorithmic 268.6 MB) from profiles/r1_ncu_conv_tc3_v19_512x512_64_64.json")
    fam = "tcgen05" if prof["tcgen05"]["gflop"] > prof["cuda_core"]["gflop"] else "cuda_core"
    pf = prof[fam]
    achieved = pf["gflop"] / max(pf["ms"], 1e-9)  # GFLOP / ms == TFLOP/s
    roofline = {
        "bound": "tensor", "kernel": "conv/linear implicit GEMM (%s path)" % fam, "achieved": achieved,
        "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tflops_sustained"], "traffic": traffic,
        "traffic_note": traffic_note,
        "dominant_shape": {"path": "tcgen05" if dom_k[0] else "cuda_core", "M": dom_k[1], "K": dom_k[2], "N": dom_k[3],
                           "launches_per_clip": dom_v[0], "avg_us": 1e3 * dom_v[1] / dom_v[0],
                           "tflops": dom_v[2] / max(dom_v[1], 1e-9), "share_of_family_time": dom_v[1] / max(pf["ms"], 1e-9)},
        "peak_source": pk["source"] + ", bf16 sustained", "launches_per_clip": pf["launches"], "ms_per_clip": pf["ms"],
        "gflop_per_clip": pf["gflop"], "algorithmic_gb_per_clip": pf["gbytes"],
        "hbm_achieved_gbs": pf["gbytes"] / max(pf["ms"], 1e-9) * 1e3, "hbm_peak_gbs": pk["hbm_gbs"],
        "other_family": prof["cuda_core" if fam == "tcgen05" else "tcgen05"],
        "whole_path_tflops": clip_flops(T) * B * world * args.steps / (ms * 1e-3) / 1e12 / world,
        "whole_path_frac_of_tensor_peak": clip_flops(T) * B * args.steps / (ms * 1e-3) / 1e12 / pk["tflops_sustained"],
    }

    if world > 1:
        dist.barrier()
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sd, xc, keep_oracle = cpu_baseline_sample(T=2)
            keep_oracle.keep_forward(sd, xc)  # warm-up
            t0 = time.perf_counter()
            keep_oracle.keep_forward(sd, xc)
            dt = time.perf_counter() - t0
            cpu = {"value": 2 / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": "one T=2 clip (after one warm-up) through the oracle port, torch fp32, %d host threads" % torch.get_num_threads()}
        line = {
            "metric": "512x512 aligned-face frames/sec through keep_net (20-frame clips)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_OF[mode], "data": "synthetic",
            "config": {"workload": "KEEP general model, %d-frame aligned 512x512 synthetic clip, %s per GPU per step" % (
                           T, "one clip" if B == 1 else ("%d independent clips (lockstep groups of %d%s)" % (B, batch, ", groups on engine replicas" if B > batch else " inside one engine") if batch > 1
                                                     else "%d independent clips (two in flight on engine replicas)" % B)),
                       "weights": "seeded synthetic (no checkpoint offline)", "engine_mode": mode,
                       "cuda_graph": not args.no_graph,
                       "l2": "256 MiB buffer zeroed between timed steps; per-step working set >> 126 MB L2",
                       "collective": "NCCL gather of fp16 decoded frames to rank 0" if world > 1 else "none"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes},
            "e2e_u8": {"value": e2e_u8_value, "unit": "frames/s", "h2d_bytes_per_step": u8_host.numel(), "d2h_bytes_per_step": u8_host.numel(),
                       "note": "keep_forward_u8: uint8 BGR crops in/out, host-side img2tensor/normalize/tensor2img folded in (no gather leg)"},
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
