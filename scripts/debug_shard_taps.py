"""Frame-shard bring-up: per-module comparison of the sharded denoiser (this rank's taps, reassembled over the ranks) with the
unsharded one on the same GPU.  torchrun --nproc-per-node N scripts/debug_shard_taps.py ; prints the first modules whose
output diverges.  FS taps hold (b, own frames, all pixels), PS taps (b, all frames, own pixels) -- t2v_unet_tap_info."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sd-webui-text2video_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch                       # noqa: E402
import torch.distributed as dist   # noqa: E402


def rel_rms(a, b):
    a, b = a.float(), b.float()
    return ((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-9)).item()


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from oracle import unet_oracle as UO
    from t2v_b200.modules import UNetSD
    from t2v_b200.distributed import frame_bounds
    dim = int(os.environ.get('DIM', '64'))
    cfg = UO.UNetConfig(dim=dim)
    W = UO.make_weights(UO.param_specs(cfg), seed=1)

    def make():
        with torch.device('cuda'):
            net = UNetSD(dim=dim)
        net = net.half()
        net.load_state_dict(W, strict=True)
        return net.cuda().eval()
    plain, shard = make(), make()
    shard.shard_setup()
    plain.enable_taps(True)
    shard.enable_taps(True)
    B, F, h, w = [int(v) for v in os.environ.get('SHAPE', '2,9,32,16').split(',')]
    g = torch.Generator().manual_seed(B * 100 + F)
    x = torch.randn(B, 4, F, h, w, generator=g)
    y = torch.randn(B, 77, 1024, generator=g).half().float()
    t = torch.randint(0, 1000, (B,), generator=g)
    ref = plain(x.cuda(), t.cuda(), y.cuda())
    fb = frame_bounds(F, world)
    f0, f1 = fb[rank], fb[rank + 1]
    shard.set_clip_frames(F)
    mine = shard(x[:, :, f0:f1].contiguous().cuda(), t.cuda(), y.cuda())
    names = []
    for n, _ in plain.named_modules():
        parts = n.split('.')
        if (parts[0] in ('input_blocks', 'output_blocks') and len(parts) == 3) or (parts[0] == 'middle_block' and len(parts) == 2):
            names.append(n)
    names.append('out')
    for n in names:
        try:
            a = plain.read_tap_auto(n)          # [(B F), C, hh, ww]
        except RuntimeError:
            continue
        s = shard.read_tap_auto(n)              # FS: [(B Fl), C, hh, ww]   PS: [(B F), C, 1, P_own]
        C_, hh, ww = a.shape[1], a.shape[2], a.shape[3]
        full = torch.zeros_like(a)
        ps = s.shape[2] == 1 and hh != 1
        shp = torch.tensor([s.shape[0], s.shape[3]], device='cuda')
        shapes = [torch.zeros_like(shp) for _ in range(world)]
        dist.all_gather(shapes, shp)
        # pad to a common size for the all-gather
        nmax = max(int(q[0]) * int(q[1]) for q in shapes) * C_ * (1 if ps else hh)
        buf = torch.zeros(nmax, device='cuda', dtype=torch.float16)
        buf[:s.numel()] = s.reshape(-1)
        bufs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(bufs, buf)
        if ps:
            off = 0
            fullv = full.view(B * F, C_, hh * ww)
            for r in range(world):
                pr = int(shapes[r][1])
                fullv[:, :, off:off + pr] = bufs[r][:B * F * C_ * pr].view(B * F, C_, pr)
                off += pr
        else:
            for r in range(world):
                fl = fb[r + 1] - fb[r]
                part = bufs[r][:B * fl * C_ * hh * ww].view(B, fl, C_, hh, ww)
                full.view(B, F, C_, hh, ww)[:, fb[r]:fb[r + 1]] = part
        if rank == 0:
            e = rel_rms(full, a)
            per_f = [(full.view(B, F, -1)[:, f] - a.view(B, F, -1)[:, f]).float().pow(2).mean().sqrt().item() for f in range(F)]
            print(f'TAP {n:28s} {"PS" if ps else "FS"} C={C_} {hh}x{ww} rel_rms {e:.5f}' + ('   <-- per-frame abs rms ' + ' '.join(f'{v:.3f}' for v in per_f) if e > 5e-3 else ''), flush=True)
    gathered = [torch.zeros((B, 4, F, h, w), dtype=torch.float16, device='cuda') for _ in range(world)]
    buf = torch.zeros((B, 4, F, h, w), dtype=torch.float16, device='cuda')
    buf[:, :, f0:f1] = mine
    dist.all_gather(gathered, buf)
    if rank == 0:
        print('FINAL rel_rms', rel_rms(sum(gathered), ref), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
