"""Final-state GEMM cases for `ncu --set full` (one CTA per tile): K-heavy linear, level-0 QKV, level-0 GEGLU FF, level-0 3x3 conv."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch
from t2v_b200 import ops
torch.manual_seed(0)
dev = 'cuda'
def lin(M, K, N, bn, flags=0, res=False):
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(1, N, K, device=dev) / K ** 0.5).half()
    b = torch.randn(N, device=dev).half(); r = torch.randn(M, N, device=dev).half() if res else None
    out = torch.empty(M, N // 2 if flags & 1 else N, device=dev, dtype=torch.half)
    for _ in range(2): ops.gemm(a, w, N, bias=b, residual=r, out=out, force_bn=bn, flags=flags)
def conv(NF, h, w_, C, Co, bn):
    x = torch.randn(NF * h * w_, C, device=dev).half(); wt = (torch.randn(Co, C, 3, 3, device=dev) / (9 * C) ** 0.5).half()
    wp = ops.pack_conv_weight(wt); b = torch.randn(Co, device=dev).half(); out = torch.empty(NF * h * w_, Co, device=dev, dtype=torch.half)
    for _ in range(2): ops.gemm(x, wp, Co, dims=[w_, h, NF], taps=ops.conv_taps_2d(), bias=b, out=out, force_bn=bn)
lin(24576, 1280, 2560, 256)
lin(49152, 320, 960, 256)
lin(49152, 320, 2560, 256, flags=1)
conv(48, 32, 32, 320, 320, 160)
torch.cuda.synchronize()
print('done')
