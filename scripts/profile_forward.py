"""Per-family and per-launch device time of one UNet forward at the benchmark shape (run under gpurun)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch
from t2v_b200.modules import UNetSD
from t2v_b200.synthetic import randomize_
B, F, h, w = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (2, 24, 32, 32)))
out = sys.argv[5] if len(sys.argv) > 5 else 'gpurun_out/forward_steps.tsv'
with torch.device('cuda'):
    net = UNetSD()
net = randomize_(net.half().cuda().eval(), seed=0)
x = torch.randn(B, 4, F, h, w, device='cuda'); y = torch.randn(B, 77, 1024, device='cuda'); t = torch.full((B,), 500.0, device='cuda')
for _ in range(3): net(x, t, y)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): net(x, t, y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
fl = net.flops(B, F, h, w)
print(f'forward B{B} F{F} {h}x{w}: {ms:.2f} ms  {fl/1e12:.3f} TFLOP  {fl/ms/1e9:.1f} TFLOP/s  launches {net.num_launches()}')
os.environ['T2V_PROFILE_DUMP'] = out
p = net.profile(B, F, h, w)
for k, v in p.items():
    print(k, v)
rows = [l.rstrip('\n').split('\t') for l in open(out)]
agg = {}
for i, kind, ms_, fl_, label in rows:
    a = agg.setdefault(label, [0, 0.0, 0.0]); a[0] += 1; a[1] += float(ms_); a[2] += float(fl_)
print('--- top launch groups by time')
for key, (n, ms_, fl_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f'{ms_:8.3f} ms  x{n:3d}  {fl_/max(ms_,1e-9)/1e9:8.1f} TF/s  {key}')
