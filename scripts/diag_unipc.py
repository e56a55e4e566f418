"""Diagnostic: where do the product UniPC sampler and the oracle UniPC sampler diverge when both are driven by the SAME GPU
denoiser?  Prints, per model call, the relative RMS difference of the latents handed to the model and of the eps returned."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'sd-webui-text2video_b200')):
    sys.path.insert(0, p)
import torch                                                      # noqa: E402
from oracle import unet_oracle as UO, samplers_oracle as SO       # noqa: E402
from oracle.make_golden import synth_inputs                       # noqa: E402
from t2v_b200.modules import UNetSD                               # noqa: E402
from t2v_b200 import samplers                                     # noqa: E402


def rr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-12)).item()


def main():
    dim = int(sys.argv[1]) if len(sys.argv) > 1 else 320
    F, h, w = (24, 32, 32) if dim == 320 else (3, 16, 8)
    cfg = UO.UNetConfig(dim=dim)
    W = UO.make_weights(UO.param_specs(cfg), seed=0)
    with torch.device('cuda'):
        net = UNetSD(dim=dim)
    net = net.half()
    net.load_state_dict(W, strict=True)
    net = net.cuda().eval()
    betas = SO.linear_sd_betas()
    net.register_schedule(given_betas=betas.numpy())
    x, c, uc = synth_inputs(F, h, w)
    log = {'ours': [], 'orc': []}

    class Stop(Exception):
        pass

    class Rec(object):
        def __init__(self, tag, n):
            self.tag, self.n = tag, n

        def __getattr__(self, k):
            return getattr(net, k)

        def __call__(self, xx, tt, cc):
            e = net(xx.cuda(), torch.as_tensor(tt).cuda().float(), cc.cuda())
            log[self.tag].append((xx.detach().float().cpu().clone(), torch.as_tensor(tt).float().cpu().clone(), e.float().cpu().clone()))
            if len(log[self.tag]) == self.n:
                raise Stop()
            return e if xx.is_cuda else e.cpu()
    N = 8
    entry = [s for s in samplers.available_samplers if s.name == 'UniPC'][0]
    m = Rec('ours', N)
    try:
        entry.init_sampler(m, betas=betas, device=torch.device('cuda')).sample(
            S=30, conditioning=c.cuda(), unconditional_conditioning=uc.cuda(), unconditional_guidance_scale=17.0, x_T=x.cuda(),
            shape=tuple(x.shape), eta=0.0, batch_size=1)
    except Stop:
        pass
    try:
        SO.unipc_sample(Rec('orc', N), betas, x, 30, c, uc, 17.0)
    except Stop:
        pass
    for i, (a, b) in enumerate(zip(log['ours'], log['orc'])):
        print(f'call {i + 1}: t ours {a[1].flatten()[0].item():.4f} orc {b[1].flatten()[0].item():.4f}  x diff {rr(a[0], b[0]):.3e}  '
              f'eps diff {rr(a[2], b[2]):.3e}  |x| {b[0].abs().max().item():.3f} |eps| {b[2].abs().max().item():.3f}')


if __name__ == '__main__':
    main()
