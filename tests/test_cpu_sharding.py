"""CPU tier: clip splitting / round-robin assignment / gather order of the N>1 path, world_size-2 gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_split_matches_reference_loop(keep_mod):
    sh = keep_mod.sharding
    assert sh.split_clips(100, 20) == [(0, 20, 20), (20, 40, 20), (40, 60, 20), (60, 80, 20), (80, 100, 20)]
    c = sh.split_clips(512, 20)            # BASELINE.json config 5: 25 x 20 + 1 x 12
    assert len(c) == 26 and c[-1] == (500, 512, 12)
    assert sh.split_clips(21, 20)[-1] == (20, 21, 1)   # 1-frame tail (duplicated to T=2 by the runner)
    assert sh.split_clips(0, 20) == []
    with pytest.raises(ValueError):
        sh.split_clips(5, 0)
    owner = sh.assign_round_robin(26, 8)
    assert [owner.count(r) for r in range(8)] == [4, 4, 3, 3, 3, 3, 3, 3]


class _FakeNet:
    """Stands in for keep_net on CPU: marks every frame with (clip length, frame position) so order/ownership show."""

    def __call__(self, x, need_upscale=False):
        assert x.shape[1] >= 2 and not need_upscale
        t = torch.arange(x.shape[1], dtype=x.dtype).view(1, -1, 1, 1, 1)
        return x * 2.0 + t * 1e-3


@pytest.mark.parametrize("n_frames,clip_len,per_call", [(7, 3, 2), (9, 2, 4), (4, 4, 4), (5, 1, 3), (41, 20, 2)])
def test_batched_clip_runner_equals_the_reference_loop(keep_mod, n_frames, clip_len, per_call):
    sh = keep_mod.sharding
    g = torch.Generator().manual_seed(1)
    frames = torch.rand((1, n_frames, 3, 8, 8), generator=g)
    calls = []

    class Net(_FakeNet):
        def __call__(self, x, need_upscale=False):
            calls.append(tuple(x.shape[:2]))
            return super().__call__(x, need_upscale)

    out = sh.run_clips_batched(Net(), frames, clip_len, clips_per_call=per_call)
    ref = sh.run_clips(_FakeNet(), frames, clip_len)
    assert torch.equal(out, ref)
    assert max(b for b, _ in calls) <= per_call
    if clip_len >= 2 and n_frames >= 2 * clip_len and per_call >= 2:
        assert any(b > 1 for b, _ in calls)        # full-length clips really went in together


def _worker(rank, world, port, n_frames, clip_len, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import keep_b200
    g = torch.Generator().manual_seed(0)
    frames = torch.rand((1, n_frames, 3, 512, 512), generator=g)[:, :, :, :8, :8].repeat(1, 1, 1, 64, 64)
    out = keep_b200.sharding.run_clips_sharded(_FakeNet(), frames, clip_len, gather_dtype=torch.float32)
    calls = []

    class Net(_FakeNet):
        def __call__(self, x, need_upscale=False):
            calls.append(int(x.shape[0]))
            return super().__call__(x, need_upscale)

    out2 = keep_b200.sharding.run_clips_sharded_batched(Net(), frames, clip_len, gather_dtype=torch.float32)
    if rank == 0:
        ref = keep_b200.sharding.run_clips(_FakeNet(), frames, clip_len)
        n_full_mine = len([k for k, c in enumerate(keep_b200.sharding.split_clips(n_frames, clip_len))
                           if k % world == 0 and c[2] == clip_len and clip_len >= 2])
        batched_ok = (max(calls) if calls else 0) == max(1, n_full_mine)       # rank 0's full clips went in as one call
        q.put((tuple(out.shape), bool(torch.equal(out, ref)) and bool(torch.equal(out2, ref)) and batched_ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,clip_len", [(7, 3), (5, 2), (4, 4), (13, 2)])
def test_sharded_equals_single_process_gloo_world2(n_frames, clip_len):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, clip_len, q)) for r in range(2)]
    for p in procs:
        p.start()
    shape, same = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert shape == (1, n_frames, 3, 512, 512) and same
