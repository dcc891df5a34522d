#!/usr/bin/env python
"""bench.py — KEEP hot path (`keep_net(clip, need_upscale=False)`) on N B200s of one node.

Metric (BASELINE.json): aligned 512x512 face frames/sec, with
  value        frames/s with the frames already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e          frames/s through the plugin call with HOST (pinned) buffers: H2D of the frames + D2H of the
               decoded frames inside the timed region (copies pipelined on side streams around the call)
  roofline     dominant kernel (conv / linear implicit GEMM) at its most expensive layer shape: algorithmic FLOPs of those
               launches / their CUDA-event time on the launching stream, vs the measured tensor peak in MEASURED_PEAKS.json
               (`family`: the same over every launch of the kernel family)
  cpu_baseline the oracle port (fp32 torch restatement of the reference forward) on the host cores (N=1 only):
               BASELINE.json configs[0] -- one aligned face duplicated to T=2 (keep_processor.py:173-175)
  gpu_eager_baseline (N=1 only) the same restatement of the reference's eager PyTorch forward on the B200 itself
               (cuDNN / cuBLAS dispatch), fp32 with TF32 off and with TF32 on -- the reference's own Blackwell path

Workloads (`--config`, BASELINE.json configs[i-1]):
  2 (default)  one 20-frame aligned clip per rank per step (weak scaling; N>1: + NCCL gather of fp16 frames to rank 0
               = configs[3]); `--clips-per-step B` puts B independent clips per rank in one call
  3            a 100-frame aligned sequence through the reference caller's clip loop (keep_processor.py:258-273,
               max_clip_length=20 -> 5 clips; detection / alignment / paste-back are the reference's CPU code, not timed)
  5            a 512-frame stream: 25 clips of 20 + one of 12, round-robin over the ranks (strong scaling), each rank's
               clips handed to its engine in one call (lockstep groups with --batch-clips), one gather at the end

`--impl reference` times the reference's own CPU implementation of the path (the oracle port — the reference is
Python and /root/reference does not exist on the GPU box) on the SAME config: one full clip per step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T_CLIP = 20
METRIC = "512x512 aligned-face frames/sec through keep_net (20-frame clips)"
# Engine mode benchmarked by default = the product default (keep_net.DEFAULT_FLAGS): the fastest one that meets the parity bar
#   fp32 = exact CUDA-core kernels; tc3 = tcgen05 with split-precision operands (fp32-grade); tc = tcgen05 fp16 operands
DEFAULT_MODE = "tc3"
DTYPE_OF = {"fp32": "f32", "tc3": "f16x2 split operands, fp32 accumulate (fp32-grade)", "tc": "f16 operands, fp32 accumulate"}


def clip_flops(T):
    """Algorithmic FLOPs of the reference forward on one clip (torch FlopCounterMode): 2*MAC."""
    return (1058.97 * T - 410.13) * 1e9


def workload_config(cfg, T, B):
    """The `config` object both arms print (identical for `--impl reference` and the engine)."""
    if cfg == 3:
        w = "KEEP general model, 100-frame aligned 512x512 synthetic sequence through the caller's clip loop, max_clip_length=20 (5 clips)"
    elif cfg == 5:
        w = "KEEP general model, 512-frame aligned 512x512 synthetic stream, max_clip_length=20 (25 clips of 20 + 1 of 12), clips round-robin over the GPUs"
    else:
        w = "KEEP general model, %d-frame aligned 512x512 synthetic clip, %s per GPU per step" % (
            T, "one clip" if B == 1 else "%d independent clips" % B)
    return {"workload": w, "weights": "seeded synthetic (no checkpoint offline)", "frames_per_clip": T, "baseline_config": cfg}


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


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def oracle_cpu():
    """The oracle port on the host cores (checker / baseline only -- never on the product path)."""
    import torch
    import keep_b200
    from oracle import keep_oracle
    torch.set_num_threads(os.cpu_count() or 1)
    return keep_b200.synth.make_state_dict(seed=0), keep_oracle


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port), rank 0 only, on the SAME
    config as the engine arm -- one full clip (config 2) per step; configs 3 / 5 run one of their 20-frame clips per step
    (a bounded sample: the stream is a loop over independent clips, keep_processor.py:263-270)."""
    if rank != 0:
        return
    import torch
    import keep_b200
    sd, keep_oracle = oracle_cpu()
    T = args.frames
    x = keep_b200.synth.make_clip(T, seed=1234, coherent=True)
    keep_oracle.keep_forward(sd, x[:, :2])            # warm-up (thread pool, oneDNN primitives): one T=2 clip
    t0 = time.perf_counter()
    for _ in range(args.steps):
        keep_oracle.keep_forward(sd, x)
    dt = time.perf_counter() - t0
    fps = T * args.steps / dt
    cores = torch.get_num_threads()
    sample = ("one full %d-frame clip per step (the engine arm's clip, seed 1234), oracle port of keep_arch.py:1008-1145 in torch "
              "fp32 on %d host threads (%s); warm-up = one T=2 clip" % (T, cores, cpu_model()))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak" if args.config != 5 else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.config, T, max(1, args.clips_per_step)),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(T, dev, calls=2):
    """The reference's own Blackwell path: its eager fp32 PyTorch forward on the B200 -- cuDNN / cuBLAS /
    ATen dispatch, no custom kernels -- restated functionally by the oracle port (pinned against the real reference,
    tests/golden/pin_report.json; /root/reference itself is absent on the GPU box).  Timed with TF32 off (true fp32, the
    parity oracle's arithmetic) and with TF32 on for cuDNN + cuBLAS; the code-index flip rate between the two is the
    reference's own device-side noise floor.  Baseline leg only: nothing here is on the product path."""
    import torch
    import keep_b200
    from oracle import keep_oracle
    sd = {k: v.to(dev) for k, v in keep_b200.synth.make_state_dict(seed=0).items()}
    x = keep_b200.synth.make_clip(T, seed=1234, coherent=True).to(dev)
    out = {"kind": "port", "what": "oracle/keep_oracle.py (functional restatement of the reference's eager PyTorch forward) on cuda:0, "
                                    "cudnn.benchmark on, %d timed calls after 1 warm-up" % calls}
    codes = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        torch.backends.cudnn.benchmark = True
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.device(dev), torch.no_grad():
                _, cap = keep_oracle.keep_forward(sd, x, capture=True)       # warm-up (cuDNN autotune) + the code indices
                codes[name] = cap["codes"][0].cpu()
                del cap
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(calls):
                    keep_oracle.keep_forward(sd, x)
                b.record()
                torch.cuda.synchronize()
            ms = a.elapsed_time(b) / calls
            out[name] = {"value": T / (ms * 1e-3), "unit": "frames/s", "ms_per_clip": ms}
        flips = (codes["fp32"] != codes["tf32"]).float().mean(dim=1)
        out["tf32"]["code_flip_rate_vs_fp32_per_frame"] = [round(float(v), 4) for v in flips]
        out["tf32"]["code_flip_rate_vs_fp32"] = float(flips.mean())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
        del sd
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 5], help="BASELINE.json workload (see the module docstring)")
    ap.add_argument("--frames", type=int, default=T_CLIP)
    ap.add_argument("--mode", default=os.environ.get("KEEP_BENCH_MODE", "auto"), choices=["auto", "fp32", "tc", "tc3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--clips-per-step", type=int, default=1,
                    help="config 2: clips per GPU per step (default 1 = BASELINE configs[1]); > 1 = independent clips in ONE call, "
                         "two in flight per GPU on engine replicas unless --batch-clips groups them")
    ap.add_argument("--batch-clips", type=int, default=int(os.environ.get("KEEP_BENCH_BATCH", "1")),
                    help="clips per lockstep group inside one engine (KeepNetB200(batch_clips=N)); config 5 default 2")
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
    kn = keep_b200.keep_net
    mode = args.mode
    if mode == "auto":
        mode = os.environ.get("KEEP_DEFAULT_MODE", DEFAULT_MODE)
    flags = {"fp32": 0, "tc": kn.FLAG_TCGEN05, "tc3": kn.TC3_FLAGS}[mode]
    if not args.no_graph:
        flags |= kn.FLAG_CUDA_GRAPH
    cfg = args.config
    sh = keep_b200.sharding
    if cfg == 2:
        B = max(1, args.clips_per_step)
        batch = min(B, max(1, args.batch_clips))
        n_frames = B * T
    else:
        n_frames = 100 if cfg == 3 else 512
        n_clips = len(sh.split_clips(n_frames, T))
        B = (n_clips + world - 1) // world if cfg == 5 else 1
        batch = max(1, args.batch_clips) if cfg == 5 else 1
        if cfg == 5 and args.batch_clips <= 1:
            batch = 2
    replicas = min(B, int(os.environ.get("KEEP_BENCH_REPLICAS", "2"))) if (B > 1 and B > batch) else 1
    net = keep_b200.KeepNetB200(flags=flags, batch_clips=batch, concurrent_clips=replicas)
    assert flags == kn.DEFAULT_FLAGS or args.mode != "auto" or args.no_graph, "bench default must be the product default"
    net.load_state_dict(keep_b200.synth.make_state_dict(seed=0), strict=True)
    net.eval().to(dev)
    if cfg == 2:
        x_host = torch.cat([keep_b200.synth.make_clip(T, seed=1234 + rank + 100 * i, coherent=True) for i in range(B)], 0).pin_memory()
    else:   # one long sequence, every rank holds it (config 5 shards its clips by rank; config 3 is single-GPU per rank)
        x_host = torch.cat([keep_b200.synth.make_clip(min(T, n_frames - s), seed=1234 + s, coherent=True)
                            for s in range(0, n_frames, T)], 1).pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    x = x_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    gather_buf, gather_host = None, None
    if cfg == 2 and world > 1 and rank == 0:
        gather_buf = [torch.empty((B, T, 3, 512, 512), dtype=torch.float16, device=dev) for _ in range(world)]
        gather_host = torch.empty((world, B, T, 3, 512, 512), dtype=torch.float16).pin_memory()
    side = torch.cuda.Stream(device=dev)       # gather / D2H of step k overlaps compute of step k+1
    copy_in = torch.cuda.Stream(device=dev)    # H2D of step k+1 overlaps compute of step k

    def compute(inp):
        """one step of the hot path on resident frames; returns the decoded frames (rank 0: of the whole job for config 5)"""
        if cfg == 2:
            return net(inp, need_upscale=False, out_dtype=torch.float16 if world > 1 else torch.float32)
        if cfg == 5 and world > 1:
            return sh.run_clips_sharded_batched(net, inp, T)
        if batch > 1:
            return sh.run_clips_batched(net, inp, T, clips_per_call=B)
        return sh.run_clips(net, inp, T)

    def collect(out, to_host):
        """the one collective of the path (config 2, N>1: fp16 frames to rank 0) and, for e2e, the D2H read of the result --
        on the side stream, so it overlaps the next step's compute"""
        cur = torch.cuda.current_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            if cfg == 2 and world > 1:
                dist.gather(out, gather_buf, dst=0)
                if to_host and rank == 0:
                    for r in range(world):
                        gather_host[r].copy_(gather_buf[r], non_blocking=True)
            elif to_host and out is not None:
                (out_host if out.dtype == out_host.dtype else out_host_h).copy_(out, non_blocking=True)
            if out is not None:
                out.record_stream(side)

    out_host_h = torch.empty(x_host.shape, dtype=torch.float16).pin_memory() if (cfg == 5 and world > 1) else None

    def step(inp, to_host=False):
        out = compute(inp)
        if (cfg == 2 and world > 1) or to_host:
            collect(out, to_host)
        return out

    for _ in range(args.warmup):
        step(x)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()

    # ---- timed region: K steps, device-timed per step, L2 flushed between steps (outside the timed spans)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = net.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = []
    cur = torch.cuda.current_stream(dev)
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(x)
        cur.wait_stream(side)          # the step's gather is part of the step
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = net.launch_count() - l0
    clocks = sampler.stop()
    status_bits = net.status()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    job_frames = (world * n_frames) if cfg != 5 else n_frames          # config 5: the ranks share ONE 512-frame stream
    value = job_frames * args.steps / (ms * 1e-3)

    # ---- e2e through the plugin call with host buffers: H2D of every step's frames from pinned memory and D2H of its decoded
    # frames inside the timed span; the copies run on side streams around the call (step k+1's upload and step k-1's
    # download overlap step k's kernels), K steps back to back, one span
    xin_bufs = [torch.empty_like(x_dev0) for x_dev0 in (x, x)]      # double-buffered device inputs: no allocation inside the span

    def e2e_run(k):
        with torch.cuda.stream(copy_in):
            xin_bufs[0].copy_(x_host, non_blocking=True)
        for i in range(k):
            cur.wait_stream(copy_in)                 # upload of step i done
            xin = xin_bufs[i & 1]
            if i + 1 < k:
                copy_in.wait_stream(cur)             # (the buffer being refilled was read by step i-1, already enqueued on `cur`)
                with torch.cuda.stream(copy_in):
                    xin_bufs[(i + 1) & 1].copy_(x_host, non_blocking=True)
            step(xin, to_host=True)
        cur.wait_stream(side)

    e2e_run(3)                                       # warm the caching allocator's pool for the side-stream outputs
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    k2 = max(3, min(args.steps, 8))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e2e_run(k2)
    b.record()
    torch.cuda.synchronize()
    t2 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = job_frames * k2 / (float(t2.item()) * 1e-3)
    h2d = x_host.numel() * 4                                          # per rank per step
    if cfg == 2 and world > 1:
        d2h = world * B * T * 3 * 512 * 512 * 2                       # rank 0 reads the gathered fp16 frames of the whole job
    elif cfg == 5 and world > 1:
        d2h = n_frames * 3 * 512 * 512 * 2
    else:
        d2h = x_host.numel() * 4

    # ---- the same call with uint8 BGR crops in / out (keep_forward_u8): the host-side img2tensor /
    # normalize / tensor2img of keep_processor.py folded into the device path, a quarter of the host<->device bytes
    e2e_u8 = None
    if cfg == 2:
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
        e2e_u8 = {"value": world * B * T * k2 / (float(t3.item()) * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": u8_host.numel(),
                  "d2h_bytes_per_step": u8_host.numel(),
                  "note": "keep_forward_u8: uint8 BGR crops in/out, host-side img2tensor/normalize/tensor2img folded in (no gather leg)"}
        del u8_host, u8_out_host

    # ---- roofline leg: per-launch CUDA events around every conv/GEMM launch (one extra 20-frame clip, rank 0's view)
    pk = peaks()
    net.profile(True)
    net(x[:1, :T].contiguous(), need_upscale=False)
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
    traffic, traffic_note = None, None
    for cand in ("r2_ncu_conv_tc3_512x512_64_64.json", "r1_ncu_conv_tc3_v19_512x512_64_64.json"):
        ncu_path = os.path.join(ROOT, "profiles", cand)
        if os.path.exists(ncu_path):
            ln = json.load(open(ncu_path))["launches"][0]
            traffic = (float(ln["dram__bytes_read.sum"].split()[0]) + float(ln["dram__bytes_write.sum"].split()[0])) * 1e6
            traffic_note = ("dram read+write bytes of ONE conv_tc_kernel<3,false> launch at M=524288 K=576 N=64 (two 512x512 frames, 64->64 "
                            "3x3; algorithmic 268.6 MB) from profiles/" + cand)
            break
    fam = "tcgen05" if prof["tcgen05"]["gflop"] > prof["cuda_core"]["gflop"] else "cuda_core"
    pf = prof[fam]
    achieved = pf["gflop"] / max(pf["ms"], 1e-9)  # GFLOP / ms == TFLOP/s
    flops_per_step = sum(clip_flops(e - s) for s, e, _ in sh.split_clips(n_frames, T)) if cfg != 2 else clip_flops(T) * B
    job_flops = flops_per_step * (world if cfg != 5 else 1)
    # achieved / frac: the DOMINANT kernel instance -- the conv / linear kernel's most expensive layer shape (algorithmic FLOPs of
    # its launches / their CUDA-event time); `family` = the same over every launch of the kernel family
    dom_tflops = dom_v[2] / max(dom_v[1], 1e-9)
    roofline = {
        "bound": "tensor", "kernel": "conv_tc_kernel (conv / linear implicit GEMM, %s path), shape M=%d K=%d N=%d" % (fam, dom_k[1], dom_k[2], dom_k[3]),
        "achieved": dom_tflops, "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": dom_tflops / pk["tflops_sustained"],
        "traffic": traffic / 2 if traffic else None,
        "traffic_note": (traffic_note + "; halved: the chain's launches of this shape process ONE frame (M=262144, algorithmic 134.3 MB)") if traffic_note else None,
        "dominant_shape": {"path": "tcgen05" if dom_k[0] else "cuda_core", "M": dom_k[1], "K": dom_k[2], "N": dom_k[3],
                           "launches_per_clip": dom_v[0], "avg_us": 1e3 * dom_v[1] / dom_v[0],
                           "tflops": dom_tflops, "share_of_family_time": dom_v[1] / max(pf["ms"], 1e-9)},
        "family": {"achieved": achieved, "frac": achieved / pk["tflops_sustained"], "what": "every conv / linear launch of the clip"},
        "peak_source": pk["source"] + ", bf16 sustained", "launches_per_clip": pf["launches"], "ms_per_clip": pf["ms"],
        "gflop_per_clip": pf["gflop"], "algorithmic_gb_per_clip": pf["gbytes"],
        "hbm_achieved_gbs": pf["gbytes"] / max(pf["ms"], 1e-9) * 1e3, "hbm_peak_gbs": pk["hbm_gbs"],
        "other_family": prof["cuda_core" if fam == "tcgen05" else "tcgen05"],
        "whole_path_tflops_per_gpu": job_flops * args.steps / (ms * 1e-3) / 1e12 / world,
        "whole_path_frac_of_tensor_peak": job_flops * args.steps / (ms * 1e-3) / 1e12 / world / pk["tflops_sustained"],
        "note": "per-launch events of ONE profiled 20-frame clip (eager, serialised by the events, GMFlow inline on the main stream so that every launch is timed alone); the headline value is the graph replay with the GMFlow side branch",
    }

    if world > 1:
        dist.barrier()
    if rank == 0:
        cpu, eager = None, None
        if world == 1 and not args.no_eager_baseline:
            net.to("cpu")                      # free the engine's workspace for the eager PyTorch run
            xin_bufs.clear()
            del x
            torch.cuda.empty_cache()
            eager = gpu_eager_baseline(T, dev)
        if world == 1 and not args.no_cpu_baseline:
            sd, keep_oracle = oracle_cpu()
            g = torch.Generator().manual_seed(1234)
            face = torch.rand((1, 1, 3, 512, 512), generator=g) * 2 - 1          # BASELINE.json configs[0]
            xc = torch.cat([face, face], dim=1)                                  # keep_processor.py:173-175
            keep_oracle.keep_forward(sd, xc)  # warm-up
            t0 = time.perf_counter()
            n_cpu = 3
            for _ in range(n_cpu):
                keep_oracle.keep_forward(sd, xc)
            dt = (time.perf_counter() - t0) / n_cpu
            cpu = {"value": 2 / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port", "faces_per_s": 1 / dt,
                   "seconds_per_call": dt, "cpu_model": cpu_model(),
                   "sample": "BASELINE configs[0]: one aligned face duplicated to T=2 (keep_processor.py:173-175), %d calls after one "
                             "warm-up through the oracle port, torch fp32, %d host threads; the full 20-frame clip is timed by "
                             "`--impl reference`" % (n_cpu, torch.get_num_threads())}
        conf = workload_config(cfg, T, B if cfg == 2 else 1)
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if cfg == 5 else "weak", "vs_baseline": None, "dtype": DTYPE_OF[mode], "data": "synthetic",
            "config": conf,
            "engine": {"engine_mode": mode, "cuda_graph": not args.no_graph, "flags": int(net._flags),
                       "clips_per_call": B, "lockstep_group": batch, "engine_replicas": replicas,
                       "l2": "256 MiB buffer zeroed between timed steps; per-step working set >> 126 MB L2",
                       "collective": ("NCCL gather of fp16 decoded frames to rank 0 on a side stream" if world > 1 else "none"),
                       "nonfinite_status_bits": int(status_bits)},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "pinned host buffers; per step H2D of the input frames and D2H of the decoded frames, pipelined on side "
                            "streams around the call; %d steps in one span" % k2},
            "roofline": roofline,
        }
        if e2e_u8 is not None:
            line["e2e_u8"] = e2e_u8
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if eager is not None:
            line["gpu_eager_baseline"] = eager
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
