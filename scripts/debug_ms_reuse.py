"""ModelScope UNetSD: arena-reuse vs no-reuse outputs (must be identical)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch
from t2v_b200.modules import UNetSD
from t2v_b200.synthetic import randomize_
def err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-9)).item()
with torch.device('cuda'):
    net = UNetSD()
net = randomize_(net.half().cuda().eval(), seed=0)
for (B, F, h, w) in ((1, 4, 16, 16), (2, 8, 16, 16), (2, 24, 32, 32)):
    x = torch.randn(B, 4, F, h, w, device='cuda'); y = torch.randn(B, 77, 1024, device='cuda'); t = torch.full((B,), 500.0, device='cuda')
    outs = {}
    for mode in ('1', ''):
        if mode: os.environ['T2V_ARENA_NO_REUSE'] = '1'
        else: os.environ.pop('T2V_ARENA_NO_REUSE', None)
        p = getattr(net.time_embed, '0').bias; p.data = p.data.clone(); net.mark_dirty()
        outs[mode] = net(x, t, y).clone()
    print(f'UNetSD B{B} F{F} {h}x{w}: reuse-vs-noreuse rel rms {err(outs[""], outs["1"]):.3e}', flush=True)
