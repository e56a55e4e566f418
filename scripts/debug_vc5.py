"""reuse vs no-reuse outputs for several VC configs in one process (plan rebuilt by re-shipping one parameter)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch
from oracle import vc_oracle as VC, unet_oracle as UO
from t2v_b200.modules import UNetModel
def err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-9)).item()
for mc, cd, tl, T, h, w, L in ((64, 48, 4, 4, 8, 8, 7), (64, 48, 16, 16, 16, 16, 77), (128, 768, 16, 16, 16, 16, 77), (320, 768, 16, 16, 16, 16, 77), (320, 768, 16, 4, 16, 16, 77), (320, 768, 16, 16, 8, 8, 77), (320, 768, 16, 16, 16, 16, 7)):
    cfg = VC.VCConfig(model_channels=mc, context_dim=cd, temporal_length=tl)
    W = UO.make_weights(VC.vc_param_specs(cfg), seed=3)
    net = UNetModel(model_channels=mc, context_dim=cd, temporal_length=tl).half(); net.load_state_dict(W, strict=True); net = net.cuda().eval()
    x = torch.randn(1, 4, T, h, w).cuda(); ctx = torch.randn(1, L, cd).cuda(); t = torch.tensor([981]).cuda()
    outs = {}
    for mode in ('1', ''):
        if mode: os.environ['T2V_ARENA_NO_REUSE'] = '1'
        else: os.environ.pop('T2V_ARENA_NO_REUSE', None)
        p = getattr(net.time_embed, '0').bias; p.data = p.data.clone(); net.mark_dirty()
        outs[mode] = net(x, t, context=ctx).clone()
    print(f'mc{mc} ctx{cd} L{tl} T{T} {h}x{w} tokens{L}: reuse-vs-noreuse rel rms {err(outs[""], outs["1"]):.3e}', flush=True)
    del net
