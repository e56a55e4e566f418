"""Worker of tests/test_frame_shard_gpu.py: one rank of a frame-sharded clip (launched by torch.distributed.run, one process
per GPU).  Compares the sharded denoiser / pipeline with the unsharded one on the same GPU and with the CPU oracle; rank 0
prints one JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sd-webui-text2video_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch                       # noqa: E402
import torch.distributed as dist   # noqa: E402


def rel_rms(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-9)).item()


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from oracle import unet_oracle as UO, vae_oracle as VO, samplers_oracle as SO
    from t2v_b200.modules import UNetSD
    from t2v_b200.pipeline import TextToVideoSynthesis
    from t2v_b200.distributed import frame_bounds
    full_model = os.environ.get('T2V_SHARD_FULL') == '1'
    dim = 320 if full_model else 64
    cfg = UO.UNetConfig(dim=dim)
    W = UO.make_weights(UO.param_specs(cfg), seed=1)
    Wh = {k: v.half().float() for k, v in W.items()}

    def make():
        with torch.device('cuda'):
            net = UNetSD(dim=dim)
        net = net.half()
        net.load_state_dict(W, strict=True)
        return net.cuda().eval()
    plain, shard = make(), make()
    shard.shard_setup()
    out = []
    # the deepest level (3 downsamples) must keep >= world pixels: 32 x 16 latent -> 4 x 2
    shapes = [(2, 24, 32, 32)] if full_model else [(2, 9, 32, 16), (1, 11, 16, 32), (2, 13, 32, 16), (1, world, 32, 16)]
    for (B, F, h, w) in shapes:
        g = torch.Generator().manual_seed(B * 100 + F)
        x = torch.randn(B, 4, F, h, w, generator=g)
        y = torch.randn(B, 77, 1024, generator=g).half().float()
        t = torch.randint(0, 1000, (B,), generator=g)
        ref = plain(x.cuda(), t.cuda(), y.cuda())
        f0, f1 = shard.frame_range(F)
        assert (f0, f1) == tuple(frame_bounds(F, world)[rank:rank + 2])
        shard.set_clip_frames(F)
        for rep in range(3):                 # eager, graph capture, graph replay
            mine = shard(x[:, :, f0:f1].contiguous().cuda(), t.cuda(), y.cuda())
        gathered = [torch.zeros((B, 4, F, h, w), dtype=torch.float16, device='cuda') for _ in range(world)]
        buf = torch.zeros((B, 4, F, h, w), dtype=torch.float16, device='cuda')
        buf[:, :, f0:f1] = mine
        dist.all_gather(gathered, buf)
        full = sum(gathered)
        rec = {'case': f'forward B{B} F{F} {h}x{w}', 'world': world, 'sharded_vs_unsharded_rms': rel_rms(full, ref),
               'equal_bits_frac': (full == ref).float().mean().item()}
        if not full_model:
            orc = UO.unet_forward(Wh, cfg, x, t, y)
            rec['sharded_vs_oracle_rms'] = rel_rms(full, orc)
            rec['unsharded_vs_oracle_rms'] = rel_rms(ref, orc)
        out.append(rec)
    if not full_model:
        # whole pipeline: TextToVideoSynthesis.infer sharded == unsharded (x_T from the same CPU seed, eta 0 and eta > 0)
        Wv = UO.make_weights(VO.decoder_param_specs(VO.VAEConfig()), seed=3)
        pp = TextToVideoSynthesis(None, model_cfg={'unet_dim': 64}, unet_state=W, vae_state=Wv)
        ps = TextToVideoSynthesis(None, model_cfg={'unet_dim': 64}, unet_state=W, vae_state=Wv)
        ps.enable_frame_shard()
        g = torch.Generator().manual_seed(2)
        c = torch.randn(1, 77, 1024, generator=g).half()
        uc = torch.randn(1, 77, 1024, generator=g).half()
        for sampler, eta in (('DDIM_Gaussian', 0.0), ('DDIM', 0.0), ('UniPC', 0.0), ('DDIM', 0.5)):
            S = 6
            torch.cuda.manual_seed(77)
            fr_p, lat_p, _ = pp.infer(c, uc, S, 9, 77, 5.0, 256, 128, eta, 'GPU (half precision)', torch.device('cuda'), None, 0, 0.0,
                                      None, False, sampler)
            fr_s, lat_s, _ = ps.infer(c, uc, S, 9, 77, 5.0, 256, 128, eta, 'GPU (half precision)', torch.device('cuda'), None, 0, 0.0,
                                      None, False, sampler)
            import numpy as np
            d = np.abs(np.stack(fr_p).astype(int) - np.stack(fr_s).astype(int))
            out.append({'case': f'infer {sampler} eta {eta}', 'latent_rms': rel_rms(lat_s, lat_p), 'frames': len(fr_s),
                        'u8_mean_abs_diff': float(d.mean()), 'u8_max_abs_diff': int(d.max())})
    dist.barrier()
    if rank == 0:
        for r in out:
            print('SHARD ' + json.dumps(r), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
