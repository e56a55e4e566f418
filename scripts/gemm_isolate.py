"""Isolation sweep for the small-K GEMM: which stage bounds it? (run under gpurun)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch
from t2v_b200 import ops
dev = 'cuda'
def bench(M, K, N, bn, flags=0, res=False, bias=True, cg=1, taps=None, dims=None, iters=20):
    a = torch.randn(M, K, device=dev).half()
    nt = 1 if taps is None else len(taps)
    w = (torch.randn(nt, N, K, device=dev) / K ** 0.5).half()
    b = torch.randn(N, device=dev).half() if bias else None
    r = torch.randn(M, N, device=dev).half() if res else None
    out = torch.empty(M, N, device=dev, dtype=torch.half)
    big = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    def run():
        ops.gemm(a, w, N, bias=b, residual=r, out=out, force_bn=bn, force_cg=cg, flags=flags, taps=taps, dims=dims)
    for _ in range(3): run()
    torch.cuda.synchronize()
    t = 0.0
    for _ in range(iters):
        big.zero_()                      # flush L2
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); t += e0.elapsed_time(e1)
    cold = t / iters
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / iters
    fl = 2.0 * M * N * K * nt
    return cold * 1e3, warm * 1e3, fl / warm / 1e9
NS, NE, NM = 256, 512, 1024
GEGLU = 1
cases = [(49152, 320, 320, 160, 0), (49152, 320, 2560, 256, GEGLU), (12288, 640, 5120, 256, GEGLU), (49152, 320, 960, 256, 0)]
if len(sys.argv) > 1:
    cases = [c for c in cases if str(c[2]) in sys.argv[1:]]
for (M, K, N, bn, base) in cases:
    for name, fl, res in (('full+res', 0, True), ('full', 0, False), ('no_store', NS, False), ('no_epilogue', NE, False), ('no_mma', NM, False), ('no_mma_no_epi', NM | NE, False)):
        if base and res:
            continue
        c, w_, tf = bench(M, K, N, bn, flags=fl | base, res=res)
        print(f'M{M} K{K} N{N} bn{bn} {"geglu " if base else ""}{name:14s}: cold {c:7.1f} us  warm {w_:7.1f} us  ({tf:6.1f} TF/s warm)', flush=True)
