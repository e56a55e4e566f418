"""A/B of the GEMM epilogue variants (T2V_LIB_PATH selects the build): warm per-launch time of the layers that are bound by
the epilogue / operand stream rather than by the tensor pipe.  Run under gpurun, once per library build."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch                       # noqa: E402
from t2v_b200 import ops           # noqa: E402

dev = 'cuda'
TAG = os.environ.get('TAG', '')


def bench(M, K, N, res, flags=0, taps=None, dims=None, iters=30):
    a = torch.randn(M, K, device=dev).half()
    nt = 1 if taps is None else len(taps)
    w = (torch.randn(nt, N, K, device=dev) / (K * nt) ** 0.5).half()
    b = torch.randn(N, device=dev).half()
    r = torch.randn(M, N, device=dev).half() if res else None
    out = torch.empty(M, N, device=dev, dtype=torch.half)

    def run():
        ops.gemm(a, w, N, bias=b, residual=r, out=out, flags=flags, taps=taps, dims=dims)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    return us, 2.0 * M * N * K * nt / us / 1e6


cases = [('L0 to_out  49152x320x320 +res', 49152, 320, 320, True, None, None),
         ('L0 proj    49152x320x320', 49152, 320, 320, False, None, None),
         ('L0 qkv     49152x960x320', 49152, 320, 960, False, None, None),
         ('L0 ff2     49152x320x1280 +res', 49152, 1280, 320, True, None, None),
         ('L1 to_out  12288x640x640 +res', 12288, 640, 640, True, None, None),
         ('L1 qkv     12288x1920x640', 12288, 640, 1920, False, None, None),
         ('L2 to_out  3072x1280x1280 +res', 3072, 1280, 1280, True, None, None),
         ('L0 conv3x3 48f 32x32 320->320 +res', 49152, 320, 320, True, ops.conv_taps_2d(), [32, 32, 48]),
         ('L0 tconv   2x24x1024 320->320', 49152, 320, 320, False, ops.conv_taps_temporal(), [1024, 24, 2])]
for name, M, K, N, res, taps, dims in cases:
    us, tf = bench(M, K, N, res, taps=taps, dims=dims)
    us2, tf2 = bench(M, K, N, res, flags=ops.GEMM_NO_BS, taps=taps, dims=dims)
    print(f'{TAG:6s} {name:38s} auto {us:7.1f} us {tf:7.1f} TF/s | streaming {us2:7.1f} us {tf2:7.1f} TF/s', flush=True)
