"""Per-launch-group device time of one FRAME-SHARDED forward (every rank profiles simultaneously: the exchange kernels wait for
their peers).  torchrun --nproc-per-node N scripts/profile_shard.py [F h w]; rank 0 prints the groups incl. the FS<->PS exchanges
and the cross-rank GroupNorm statistics."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sd-webui-text2video_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch                       # noqa: E402
import torch.distributed as dist   # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from t2v_b200.modules import UNetSD          # noqa: E402
from t2v_b200.synthetic import randomize_    # noqa: E402
F, h, w = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (125, 32, 32)))
B = 2
with torch.device('cuda'):
    net = UNetSD()
net = randomize_(net.half().cuda().eval(), seed=0)
net.shard_setup()
net.set_clip_frames(F)
f0, f1 = net.frame_range(F)
x = torch.randn(B, 4, f1 - f0, h, w, device='cuda')
y = torch.randn(B, 77, 1024, device='cuda')
t = torch.full((B,), 500.0, device='cuda')
for _ in range(3):
    net(x, t, y)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    net(x, t, y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
out = f'gpurun_out/shard_steps_rank{rank}.tsv'
os.environ['T2V_PROFILE_DUMP'] = out
dist.barrier()
p = net.profile(B, F, h, w)
dist.barrier()
if rank == 0:
    print(f'sharded forward B{B} F{F} ({f1 - f0} local) {h}x{w} over {world} ranks: {ms:.2f} ms graphed, launches {net.num_launches()}')
    print({k: v for k, v in p.items()})
    rows = [l.rstrip('\n').split('\t') for l in open(out)]
    agg = {}
    for i, kind, ms_, fl_, label in rows:
        key = label.split(' rows=')[0] if label.startswith(('exchange', 'gn_stats', 'gn_apply', 'gn_fused', 'ln_rowstats')) else ('gemm' if label.startswith('gemm') else (label or 'other'))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(ms_)
    for key, (n, ms_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
        print(f'{ms_:8.3f} ms  x{n:4d}  {key}')
dist.destroy_process_group()
