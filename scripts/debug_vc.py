"""Per-block error of the VideoCrafter UNet (GPU) vs the CPU oracle at the full configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch
from oracle import vc_oracle as VC, unet_oracle as UO
from t2v_b200.modules import UNetModel
cfg = VC.VCConfig()
W = UO.make_weights(VC.vc_param_specs(cfg), seed=0)
net = UNetModel().half(); net.load_state_dict(W, strict=True); net = net.cuda().eval()
Wh = {k: v.half().float() for k, v in W.items()}
T, h, w = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (16, 16, 16)))
g = torch.Generator().manual_seed(1)
x = torch.randn(1, 4, T, h, w, generator=g); ctx = torch.randn(1, 77, 768, generator=g).half().float(); t = torch.tensor([500])
if os.environ.get('FIXTURE'):
    gg = torch.load(os.path.join(ROOT, 'tests/golden/vc_unet_full.pt'))
    x = torch.randn(gg['shape'], generator=torch.Generator('cpu').manual_seed(gg['x_seed']))
    ctx = torch.randn((1, gg['L'], 768), generator=torch.Generator('cpu').manual_seed(gg['ctx_seed'])); t = gg['t']
if os.environ.get('TVAL'):
    t = torch.tensor([int(os.environ['TVAL'])])
taps = {}
ref = VC.vc_unet_forward(Wh, cfg, x, t, ctx, taps)
net.enable_taps(True)
out = net(x.cuda(), t.cuda(), context=ctx.cuda())
def err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-9)).item()
for k, v in taps.items():
    B, C, TT, hh, ww = v.shape
    r = v.permute(0, 2, 1, 3, 4).reshape(B * TT, C, hh, ww)
    try:
        o = net.read_tap(k, tuple(r.shape))
        print(f'{k:28s} C{C:5d} {hh}x{ww}  rel rms {err(o, r):.3e}')
    except Exception as e:
        print(k, 'tap failed', str(e)[:80])
print('final', err(out, ref))
