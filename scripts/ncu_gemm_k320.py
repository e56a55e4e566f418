"""ncu target: the level-0 K = 320 GEMMs that sit at 300-600 TFLOP/s (to_out + residual, qkv).  3 warm-up + 2 profiled launches each."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch                       # noqa: E402
from t2v_b200 import ops           # noqa: E402

dev = 'cuda'
M, K = 49152, 320
a = torch.randn(M, K, device=dev).half()
for N, res in ((320, True), (960, False)):
    w = (torch.randn(1, N, K, device=dev) / K ** 0.5).half()
    b = torch.randn(N, device=dev).half()
    r = torch.randn(M, N, device=dev).half() if res else None
    out = torch.empty(M, N, device=dev, dtype=torch.half)
    for _ in range(5):
        ops.gemm(a, w, N, bias=b, residual=r, out=out)
    torch.cuda.synchronize()
