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
