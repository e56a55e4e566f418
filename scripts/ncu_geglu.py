"""One GEGLU GEMM launch at the level-0 FF shape (for ncu --set full)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'sd-webui-text2video_b200'))
import torch
from t2v_b200 import ops
M, K, N = 49152, 320, 2560
a = torch.randn(M, K, device='cuda').half()
w = (torch.randn(1, N, K, device='cuda') / K ** 0.5).half()
b = torch.randn(N, device='cuda').half()
out = torch.empty(M, N // 2, device='cuda', dtype=torch.half)
for _ in range(3):
    ops.gemm(a, w, N, bias=b, out=out, force_bn=256, flags=1)
torch.cuda.synchronize()
