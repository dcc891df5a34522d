"""Clip-level data parallelism for `keep_net`: one process per GPU, clips round-robin over ranks, one gather.

The reference cuts the flattened face list into independent `keep_net` calls with no carried state
(modules/keep_processor.py:263-270); inside a clip the frame recurrence is serial.  So the only sharding that
pays is by clip: clip k -> rank k mod world, every rank holds a full weight replica, and the
decoded frames are gathered to rank 0 once per round (fp16, 31.5 MB per 20-frame clip) over NCCL / NVLink.
No collective touches the data path inside a clip.
"""
import torch
import torch.distributed as dist


def split_clips(n_frames, max_clip_length):
    """[(start, end, keep)] exactly as the reference loop: a 1-frame tail clip is run duplicated to T=2 and only its
    first output frame is kept (keep_processor.py:266-268)."""
    if max_clip_length < 1:
        raise ValueError("max_clip_length must be >= 1")
    clips = []
    for s in range(0, n_frames, max_clip_length):
        e = min(s + max_clip_length, n_frames)
        clips.append((s, e, e - s))
    return clips


def assign_round_robin(n_clips, world):
    """clip index -> rank (26 clips over 8 ranks -> 4,4,3,3,3,3,3,3)."""
    return [k % world for k in range(n_clips)]


def run_clips(net, frames, max_clip_length):
    """Single-process reference semantics: frames (1, N, 3, 512, 512) -> (1, N, 3, 512, 512)."""
    outs = []
    for s, e, keep in split_clips(frames.shape[1], max_clip_length):
        clip = frames[:, s:e]
        if clip.shape[1] == 1:
            clip = torch.cat([clip, clip], dim=1)
        outs.append(net(clip, need_upscale=False)[:, :keep])
    return torch.cat(outs, dim=1)


def run_clips_batched(net, frames, max_clip_length, clips_per_call=4):
    """Same result as `run_clips`, but the full-length clips go to the engine `clips_per_call` at a time as one
    (b, T, 3, 512, 512) call, so a `KeepNetB200(concurrent_clips=...)` / `KeepNetB200(batch_clips=...)` engine can overlap
    them; the shorter tail clip is its own call, a 1-frame tail is duplicated as in the reference."""
    clips = split_clips(frames.shape[1], max_clip_length)
    full = [c for c in clips if c[2] == max_clip_length and max_clip_length >= 2]
    outs = {}
    for g0 in range(0, len(full), max(1, clips_per_call)):
        group = full[g0:g0 + max(1, clips_per_call)]
        x = torch.cat([frames[:, s:e] for s, e, _ in group], dim=0)          # (b, T, 3, 512, 512)
        y = net(x, need_upscale=False)
        for j, (s, e, keep) in enumerate(group):
            outs[s] = y[j:j + 1, :keep]
    for s, e, keep in clips:
        if s in outs:
            continue
        clip = frames[:, s:e]
        if clip.shape[1] == 1:
            clip = torch.cat([clip, clip], dim=1)
        outs[s] = net(clip, need_upscale=False)[:, :keep]
    return torch.cat([outs[s] for s, _, _ in clips], dim=1) if clips else frames[:, :0]


def run_clips_sharded(net, frames, max_clip_length, group=None, gather_dtype=torch.float16):
    """Every rank holds `frames`; rank r runs clips k with k % world == r; rank 0 returns the reassembled sequence
    (other ranks return None).  One gather per round of `world` clips."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    clips = split_clips(frames.shape[1], max_clip_length)
    owner = assign_round_robin(len(clips), world)
    result = [None] * len(clips) if rank == 0 else None
    for r0 in range(0, len(clips), world):
        batch = list(range(r0, min(r0 + world, len(clips))))
        mine = [k for k in batch if owner[k] == rank]
        T = max_clip_length if max_clip_length >= 2 else 2
        buf = torch.zeros((1, T, 3, 512, 512), dtype=gather_dtype, device=frames.device)
        if mine:
            s, e, keep = clips[mine[0]]
            clip = frames[:, s:e]
            if clip.shape[1] == 1:
                clip = torch.cat([clip, clip], dim=1)
            out = net(clip, need_upscale=False)[:, :keep]
            buf[:, :keep] = out.to(gather_dtype)
        gathered = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, gathered, dst=0, group=group)
        if rank == 0:
            for k in batch:
                keep = clips[k][2]
                result[k] = gathered[owner[k]][:, :keep]
    if rank == 0:
        return torch.cat(result, dim=1)
    return None


def run_clips_sharded_batched(net, frames, max_clip_length, group=None, gather_dtype=torch.float16):
    """Config 5 of BASELINE.json (a long stream, clips round-robin over the ranks) with each rank's clips handed to its
    engine in ONE call, so `KeepNetB200(batch_clips=...)` / `(concurrent_clips=...)` overlap them, and ONE
    gather of all decoded frames at the end instead of one per round.  Every rank holds `frames`; rank 0 returns the
    reassembled (1, N, 3, 512, 512) sequence, other ranks None.  Same result as `run_clips`."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    clips = split_clips(frames.shape[1], max_clip_length)
    owner = assign_round_robin(len(clips), world)
    slots = [[k for k in range(len(clips)) if owner[k] == r] for r in range(world)]   # rank r's clips, in order
    mine = slots[rank]
    outs = {}
    full = [k for k in mine if clips[k][2] == max_clip_length and max_clip_length >= 2]
    if full:
        y = net(torch.cat([frames[:, clips[k][0]:clips[k][1]] for k in full], dim=0), need_upscale=False)
        for j, k in enumerate(full):
            outs[k] = y[j:j + 1]
    for k in mine:
        if k in outs:
            continue
        s, e, keep = clips[k]
        clip = frames[:, s:e]
        if clip.shape[1] == 1:
            clip = torch.cat([clip, clip], dim=1)
        outs[k] = net(clip, need_upscale=False)[:, :keep]
    per_rank = max(1, max(len(sl) for sl in slots))
    T = max(2, max_clip_length)
    buf = torch.zeros((per_rank, T) + tuple(frames.shape[2:]), dtype=gather_dtype, device=frames.device)
    for j, k in enumerate(mine):
        buf[j, :clips[k][2]] = outs[k][0].to(gather_dtype)
    gathered = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0, group=group)
    if rank != 0:
        return None
    if not clips:
        return frames[:, :0].to(gather_dtype)
    return torch.cat([gathered[owner[k]][slots[owner[k]].index(k), :clips[k][2]].unsqueeze(0) for k in range(len(clips))], dim=1)
