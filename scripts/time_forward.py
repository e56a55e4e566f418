"""Device time of the CUDA-graphed UNet forward (B=2 CFG pair) -- A/B harness for library variants (T2V_LIB_PATH) and env switches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch
from t2v_b200.modules import UNetSD
from t2v_b200.synthetic import randomize_
B, F, h, w = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (2, 24, 32, 32)))
with torch.device('cuda'):
    net = UNetSD()
net = randomize_(net.half().cuda().eval(), seed=0)
x = torch.randn(B, 4, F, h, w, device='cuda'); y = torch.randn(B, 77, 1024, device='cuda'); t = torch.full((B,), 500.0, device='cuda')
for _ in range(3): net(x, t, y)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): out = net(x, t, y)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
fl = net.flops(B, F, h, w)
print(f'{os.environ.get("TAG", "")} forward B{B} F{F} {h}x{w}: {best:.3f} ms  {fl/best/1e9:.1f} TFLOP/s  checksum {out.float().abs().sum().item():.4f}')
