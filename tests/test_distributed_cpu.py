"""CPU, world_size 2 over gloo: the sample-DP host logic of the N>1 path (clip->rank assignment, rank-independent
seeds, one all-gather of the decoded clips in rank order)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from t2v_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    mine = D.clips_for_rank(5, rank, ws)
    seeds = [D.clip_seed(123, i) for i in mine]
    # stand-in for a rendered clip: deterministic function of the seed
    clip = torch.full((3, 4, 4, 3), seeds[0] % 251, dtype=torch.uint8)
    got = D.gather_clips(clip)
    q.put((rank, mine, seeds, [int(g[0, 0, 0, 0]) for g in got]))
    dist.destroy_process_group()


def test_sample_dp_gather_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    assert res[0][2] == [123, 125, 127] and res[1][2] == [124, 126]
    assert res[0][3] == res[1][3] == [123 % 251, 124 % 251]          # rank order, identical on every rank


def test_single_process_paths():
    assert D.world() == (0, 1)
    x = torch.zeros(2, 2, 2, 3, dtype=torch.uint8)
    assert D.gather_clips(x)[0] is x
    assert D.clips_for_rank(3, 0, 1) == [0, 1, 2]


def _cfg_worker(rank, ws, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['T2V_CFG_SPLIT'] = '1'
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    from t2v_b200 import samplers
    calls = []

    def model(x, t, c):                     # stand-in denoiser: records which branch this rank evaluated
        calls.append(float(c.mean()))
        return x * 0.5 + c.mean()
    x = torch.arange(24, dtype=torch.float32).reshape(1, 4, 3, 2, 1)
    c, uc = torch.full((1, 7, 8), 0.25), torch.full((1, 7, 8), -0.5)
    e_c, e_u = samplers._eval_pair(model, x, torch.tensor([5]), c, uc)
    q.put((rank, D.units(), calls, float((e_c - (x * 0.5 + 0.25)).abs().max()), float((e_u - (x * 0.5 - 0.5)).abs().max())))
    dist.destroy_process_group()


def test_cfg_pair_split_world2():
    """CFG-split mode: each rank of a pair runs ONE branch, one all-gather gives both ranks (eps_cond, eps_uncond)."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cfg_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == (0, 1) and res[1][1] == (0, 1)              # one clip-rendering unit (the pair)
    assert res[0][2] == [0.25] and res[1][2] == [-0.5]              # even rank: conditional, odd rank: unconditional
    assert all(r[3] == 0.0 and r[4] == 0.0 for r in res)


def _frame_shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from t2v_b200 import distributed as D

    class FakeVAE:
        def decode_video(self, z, z_scale, as_uint8=True):            # [1, 4, n, h, w] -> [n, 2, 2, 3] "frames" tagged by content
            n = z.shape[2]
            return (z[0, 0, :, 0, 0].view(n, 1, 1, 1).expand(n, 2, 2, 3) * z_scale).to(torch.uint8)

    class FakeUNet:
        _shard = ('g', rank, world)

        def set_clip_frames(self, F):
            self.F = F
    fs = D.FrameShardedClip(FakeUNet(), FakeVAE())
    F = 7
    full = torch.arange(F, dtype=torch.float32).view(1, 1, F, 1, 1).expand(1, 4, F, 3, 2).contiguous()
    fs.begin(F, seed=None)
    try:
        mine = fs.local(full)
        assert mine.shape[2] == D.frame_bounds(F, world)[rank + 1] - D.frame_bounds(F, world)[rank]
        torch.manual_seed(5)
        n1 = D.step_noise(mine)
        torch.manual_seed(5)
        ref = torch.randn(full.shape)
        assert torch.equal(n1, ref[:, :, fs.f0:fs.f1])                # every rank draws the full clip's noise, keeps its frames
        back = fs.gather_latent(mine)
        frames = fs.decode(back, 2.0)
    finally:
        fs.end()
    q.put((rank, torch.equal(back, full), frames[:, 0, 0, 0].tolist()))
    dist.destroy_process_group()


def test_frame_shard_host_logic_world3():
    """Ragged frame ranges (7 frames over 3 ranks), the latent all-gather before the VAE and the frame-sharded decode +
    frame gather, on gloo."""
    import torch.multiprocessing as mp
    from t2v_b200 import distributed as D
    assert D.frame_bounds(125, 8) == [0, 16, 32, 48, 64, 80, 95, 110, 125]
    assert D.frame_bounds(7, 3) == [0, 3, 5, 7]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_frame_shard_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, ok, frames in res:
        assert ok and frames == [0, 2, 4, 6, 8, 10, 12], (rank, ok, frames)


def _role_group_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), T2V_CFG_SPLIT='1')
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from t2v_b200 import distributed as D
    pair, role, pgrp = D.cfg_pair()
    rgrp = D.cfg_role_group()
    t = torch.tensor([float(rank)])
    dist.all_reduce(t, group=pgrp)                   # pair (2i, 2i+1): sum = 4i + 1
    u = torch.tensor([float(rank)])
    dist.all_reduce(u, group=rgrp)                   # role group: ranks of the same parity
    e = torch.full((1, 4, 2, 2, 2), float(rank))
    ec, eu = D.exchange_eps(e, pgrp)
    q.put((rank, pair, role, t.item(), u.item(), dist.get_rank(rgrp), dist.get_world_size(rgrp), ec[0, 0, 0, 0, 0].item(), eu[0, 0, 0, 0, 0].item()))
    dist.destroy_process_group()


def test_cfg_split_times_frame_shard_groups_world4():
    """2 x 2 layout: pairs (0,1), (2,3) exchange eps; role groups {0,2} (cond) and {1,3} (uncond) are the frame-shard groups."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_role_group_worker, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    for rank, pair, role, tsum, usum, rg_rank, rg_ws, ec, eu in res:
        assert (pair, role) == (rank // 2, rank % 2)
        assert tsum == 4 * pair + 1 and usum == (2.0 if role == 0 else 4.0)
        assert (rg_rank, rg_ws) == (rank // 2, 2)
        assert (ec, eu) == (2.0 * pair, 2.0 * pair + 1)


def test_exchange_kernel_magic_division_is_exact():
    """csrc/shard.cu indexes 16-byte vectors with 32-bit round-up magic numbers (Granlund-Montgomery) instead of integer divisions:
    q = (t + ((n - t) >> s1)) >> s2, t = umulhi(m, n), m = floor(2^32 (2^l - d) / d) + 1, l = ceil(log2 d).  The same arithmetic in
    Python must equal n // d for every 32-bit n, for the divisors the kernel uses (channels / 8, pixels and frames of a block)."""
    import random

    def magic(d):
        l = 0
        while (1 << l) < d:
            l += 1
        m = (((1 << 32) * ((1 << l) - d)) // d + 1) & 0xffffffff
        s1 = min(l, 1)
        return m, s1, l - s1

    def fdiv(n, mg):
        m, s1, s2 = mg
        t = (m * n) >> 32
        return ((t + (((n - t) & 0xffffffff) >> s1)) & 0xffffffff) >> s2
    rng = random.Random(3)
    divisors = [1, 2, 3, 5, 7, 8, 15, 16, 40, 63, 80, 125, 128, 160, 512, 1024, 9216] + [rng.randrange(1, 1 << 20) for _ in range(200)]
    for d in divisors:
        mg = magic(d)
        for n in [0, 1, d - 1, d, d + 1, 2 * d - 1, (1 << 31) - 1, (1 << 32) - 1] + [rng.randrange(0, 1 << 32) for _ in range(100)]:
            assert fdiv(n, mg) == n // d, (n, d)


def test_exchange_kernel_row_algebra_emulated():
    """csrc/shard.cu moves the (own frames x peer's pixels) block of every rank with the row formulas
         FS -> PS:  src row (b * nf_me + fl) * P + pb[s] + pl          -> dst row on rank s: (b * F + fb[me] + fl) * np_s + pl
         PS -> FS:  src row (b * F + fb[s] + fl) * np_me + pl          -> dst row on rank s: (b * nf_s + fl) * P + pb[me] + pl
    Emulated here in numpy for ragged frame / pixel partitions: FS -> PS must give every rank all F frames of its pixel range,
    PS -> FS must be its inverse.  (The same formulas, with the same partition function, as the kernel -- this pins the algebra
    on the CPU; the kernel itself is checked on GPUs by tests/test_frame_shard_gpu.py.)"""
    import numpy as np
    from t2v_b200.distributed import frame_bounds
    for (B, F, P, nr) in [(2, 9, 12, 2), (1, 7, 10, 3), (2, 13, 8, 8), (1, 125, 16, 8)]:
        fb, pb = frame_bounds(F, nr), frame_bounds(P, nr)           # the kernel partitions pixels with the same function
        full = np.arange(B * F * P, dtype=np.int64).reshape(B, F, P)         # token id = (b, f, p)
        fs = [full[:, fb[r]:fb[r + 1], :].reshape(-1).copy() for r in range(nr)]          # rows (b, own f, all p)
        ps = [np.full(B * F * (pb[r + 1] - pb[r]), -1, dtype=np.int64) for r in range(nr)]
        for me in range(nr):
            nf_me = fb[me + 1] - fb[me]
            for s in range(nr):
                np_s = pb[s + 1] - pb[s]
                for b in range(B):
                    for fl in range(nf_me):
                        for pl in range(np_s):
                            ps[s][(b * F + fb[me] + fl) * np_s + pl] = fs[me][(b * nf_me + fl) * P + pb[s] + pl]
        for r in range(nr):
            assert np.array_equal(ps[r], full[:, :, pb[r]:pb[r + 1]].reshape(-1)), (B, F, P, nr, r)
        back = [np.full(B * (fb[r + 1] - fb[r]) * P, -1, dtype=np.int64) for r in range(nr)]
        for me in range(nr):
            np_me = pb[me + 1] - pb[me]
            for s in range(nr):
                nf_s = fb[s + 1] - fb[s]
                for b in range(B):
                    for fl in range(nf_s):
                        for pl in range(np_me):
                            back[s][(b * nf_s + fl) * P + pb[me] + pl] = ps[me][(b * F + fb[s] + fl) * np_me + pl]
        for r in range(nr):
            assert np.array_equal(back[r], fs[r]), (B, F, P, nr, r)
