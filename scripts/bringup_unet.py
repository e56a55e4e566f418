"""GPU bring-up of the model-level calls against the CPU oracle (tiny UNet with taps, VAE)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch
from oracle import unet_oracle as UO, vae_oracle as VO
from t2v_b200.modules import UNetSD, AutoencoderKL

def rel(a, b):
    a = a.float().cpu(); b = b.float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-6)).item(), ((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-9)).item()

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
if which in ('all', 'tiny'):
    cfg = UO.UNetConfig(dim=64)
    W = UO.make_weights(UO.param_specs(cfg), seed=1)
    Wh = {k: v.half().float() for k, v in W.items()}      # oracle sees the fp16-rounded weights the GPU gets
    net = UNetSD(dim=64).half()
    net.load_state_dict(W, strict=True)
    net = net.cuda().eval()
    for (B, F, h, w) in ((1, 3, 16, 8), (2, 4, 8, 8)):
        g = torch.Generator().manual_seed(123)
        x = torch.randn(B, 4, F, h, w, generator=g); y = torch.randn(B, 77, 1024, generator=g)
        t = torch.tensor([981] * B)
        taps = {}
        ref = UO.unet_forward(Wh, cfg, x, t, y.half().float(), taps)
        net.enable_taps(True)
        out = net(x.cuda(), t.cuda(), y.cuda()); torch.cuda.synchronize()
        for name, v in taps.items():
            try:
                got = net.read_tap(name, tuple(v.shape))
                e = rel(got, v)
                print(f'  tap {name:28s} {tuple(v.shape)} max-rel {e[0]:.3e} rms-rel {e[1]:.3e}' + ('   <<<<' if e[1] > 2e-2 else ''), flush=True)
            except Exception as ex:
                print('  tap', name, 'ERR', str(ex)[:100])
        e = rel(out, ref)
        print(f'tiny B{B} F{F} {h}x{w}: eps max-rel {e[0]:.3e} rms-rel {e[1]:.3e}  launches {net.num_launches()}', flush=True)
        net.enable_taps(False)
        out2 = net(x.cuda(), t.cuda(), y.cuda()); torch.cuda.synchronize()
        print('   arena-reuse run equals no-reuse run:', torch.equal(out, out2), rel(out2, ref))

if which in ('all', 'vae'):
    vcfg = VO.VAEConfig()
    W = UO.make_weights(VO.decoder_param_specs(vcfg), seed=3)
    Wh = {k: v.half().float() for k, v in W.items()}
    dd = {'double_z': True, 'z_channels': 4, 'resolution': 256, 'in_channels': 3, 'out_ch': 3, 'ch': 128,
          'ch_mult': [1, 2, 4, 4], 'num_res_blocks': 2, 'attn_resolutions': [], 'dropout': 0.0}
    vae = AutoencoderKL(dd, 4).half()
    vae.load_state_dict(W, strict=False)
    vae = vae.cuda().eval()
    g = torch.Generator().manual_seed(5)
    z = torch.randn(2, 4, 8, 16, generator=g) / 0.18215 * 0.8
    ref = VO.vae_decode(Wh, vcfg, z)
    out = vae.decode(z.cuda()); torch.cuda.synchronize()
    print('vae decode:', rel(out, ref), tuple(out.shape), flush=True)

if which in ('all', 'full'):
    cfg = UO.UNetConfig()
    t0 = time.time()
    W = UO.make_weights(UO.param_specs(cfg), seed=0)
    net = UNetSD().half()
    net.load_state_dict(W, strict=True); del W
    net = net.cuda().eval()
    print('full model built in', time.time() - t0, flush=True)
    gold = torch.load(os.path.join(ROOT, 'tests/golden/unet_cfg1.pt'))
    g = torch.Generator('cpu').manual_seed(123); x = torch.randn((1, 4, 4, 16, 16), generator=g)
    g2 = torch.Generator('cpu').manual_seed(2); c = torch.randn(1, 77, 1024, generator=g2); uc = torch.randn(1, 77, 1024, generator=g2)
    t = torch.tensor([981])
    out = net(x.cuda(), t.cuda(), c.cuda()); torch.cuda.synchronize()
    print('cfg1 eps_cond vs reference fp32 golden:', rel(out, gold['eps_cond']), flush=True)
    xb = torch.cat([x, x]).cuda(); yb = torch.cat([c, uc]).cuda()
    outb = net(xb, torch.tensor([981, 981]).cuda(), yb); torch.cuda.synchronize()
    print('cfg1 batched cond:', rel(outb[:1], gold['eps_cond']), 'uncond:', rel(outb[1:], gold['eps_uncond']), flush=True)
    # timing at config 2 shape
    for (B, F, h, w) in ((1, 24, 32, 32), (2, 24, 32, 32)):
        x = torch.randn(B, 4, F, h, w, device='cuda'); y = torch.randn(B, 77, 1024, device='cuda'); t = torch.full((B,), 500.0, device='cuda')
        for _ in range(2): net(x, t, y)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): net(x, t, y)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = net.flops(B, F, h, w)
        print(f'forward B{B} F{F} {h}x{w}: {ms:.2f} ms, {fl/1e12:.3f} TFLOP -> {fl/ms/1e9:.1f} TFLOP/s, launches {net.num_launches()}', flush=True)
