"""Round-2 final-state GEMM cases for `ncu --set full` (tile width chosen by gemm_plan): K-heavy 3x3 conv (level 1), level-1 QKV
(N = 1920 -> 9 x 224), level-0 QKV (K = 320), level-0 attention out-projection with residual (K = 320), level-0 GEGLU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch
from t2v_b200 import ops
torch.manual_seed(0)
dev = 'cuda'
def lin(M, K, N, flags=0, res=False):
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(1, N, K, device=dev) / K ** 0.5).half()
    b = torch.randn(N, device=dev).half(); r = torch.randn(M, N, device=dev).half() if res else None
    out = torch.empty(M, N // 2 if flags & 1 else N, device=dev, dtype=torch.half)
    for _ in range(2): ops.gemm(a, w, N, bias=b, residual=r, out=out, flags=flags)
def conv(NF, h, w_, C, Co):
    x = torch.randn(NF * h * w_, C, device=dev).half(); wt = (torch.randn(Co, C, 3, 3, device=dev) / (9 * C) ** 0.5).half()
    wp = ops.pack_conv_weight(wt); b = torch.randn(Co, device=dev).half(); out = torch.empty(NF * h * w_, Co, device=dev, dtype=torch.half)
    for _ in range(2): ops.gemm(x, wp, Co, dims=[w_, h, NF], taps=ops.conv_taps_2d(), bias=b, out=out)
conv(48, 16, 16, 640, 640)          # launches 0,1
lin(12288, 640, 1920)               # 2,3
lin(49152, 320, 960)                # 4,5
lin(49152, 320, 320, res=True)      # 6,7
lin(49152, 320, 2560, flags=1)      # 8,9
torch.cuda.synchronize()
print('done')
