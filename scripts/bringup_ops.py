"""GPU bring-up checks of the kernel-level entry points (run under gpurun). Prints max errors; exit 1 on failure."""
import os, sys, time
os.environ.setdefault('T2V_BRINGUP', '1')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch
import torch.nn.functional as F
from t2v_b200 import ops, _lib

torch.manual_seed(0)
dev = 'cuda'
fails = []

def rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6)).item()

def report(name, got, ref, tol=2e-3):
    e = rel(got, ref)
    ok = e < tol and torch.isfinite(got.float()).all().item()
    print(f'{"OK  " if ok else "FAIL"} {name}: rel-max err {e:.3e}', flush=True)
    if not ok:
        fails.append(name)

def t_linear(M, K, N, bn=0, bias=True, res=False):
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    b = torch.randn(N, device=dev).half() if bias else None
    r = torch.randn(M, N, device=dev).half() if res else None
    out = ops.gemm(a, w.view(1, N, K), N, bias=b, residual=r, force_bn=bn)
    ref = a.float() @ w.float().t()
    if bias: ref = ref + b.float()
    if res: ref = ref + r.float()
    report(f'linear M{M} K{K} N{N} bn{bn} bias{int(bias)} res{int(res)}', out, ref)

def t_conv2d(NF, h, w, Cin, Cout, bn=0):
    x = torch.randn(NF, h, w, Cin, device=dev).half()
    wt = (torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5).half()
    b = torch.randn(Cout, device=dev).half()
    n_alloc = max(Cout, 16)
    wp = ops.pack_conv_weight(wt, n_alloc=n_alloc)
    out = ops.gemm(x.view(-1, Cin), wp, Cout, dims=[w, h, NF], taps=ops.conv_taps_2d(), n_alloc=n_alloc, bias=b, force_bn=bn)
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), wt.float(), b.float(), padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    report(f'conv3x3 NF{NF} {h}x{w} {Cin}->{Cout} bn{bn}', out, ref)

def t_tconv(B, Fr, P, C, bn=0):
    x = torch.randn(B, Fr, P, C, device=dev).half()
    wt = (torch.randn(C, C, 3, 1, 1, device=dev) / (3 * C) ** 0.5).half()
    b = torch.randn(C, device=dev).half()
    wp = ops.pack_conv_weight(wt)
    res = x.view(-1, C)
    out = ops.gemm(x.view(-1, C), wp, C, dims=[P, Fr, B], taps=ops.conv_taps_temporal(), bias=b, residual=res, force_bn=bn)
    x5 = x.permute(0, 3, 1, 2).reshape(B, C, Fr, P, 1).float()
    ref = F.conv3d(x5, wt.float(), b.float(), padding=(1, 0, 0)) + x5
    ref = ref.reshape(B, C, Fr, P).permute(0, 2, 3, 1).reshape(-1, C)
    report(f'tconv B{B} F{Fr} P{P} C{C} bn{bn}', out, ref)

def t_geglu(M, K, H, bn):
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(2 * H, K, device=dev) / K ** 0.5).half()
    b = torch.randn(2 * H, device=dev).half()
    wp, bp = ops.pack_geglu_weight(w, b, bn)
    out = ops.gemm(a, wp, 2 * H, bias=bp, flags=ops.GEMM_GEGLU, force_bn=bn)
    h = (a.float() @ w.float().t() + b.float()).half()
    xa, gate = h.chunk(2, dim=-1)
    ref = xa * F.gelu(gate)
    report(f'geglu M{M} K{K} H{H} bn{bn}', out, ref, tol=4e-3)

def t_batched(nb, S, C):
    q = torch.randn(nb, S, C, device=dev).half()
    k = torch.randn(nb, S, C, device=dev).half()
    out = ops.gemm(q.view(-1, C), k, S, dims=[S, nb], taps=[[0, 0]], n_alloc=S, b_batch_dim=1, alpha=C ** -0.5)
    ref = torch.bmm(q.float(), k.float().transpose(1, 2)).reshape(-1, S) * C ** -0.5
    report(f'batched nb{nb} S{S} C{C}', out, ref)

def t_bias_rows(B, rows, K, N):
    a = torch.randn(B * rows, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    b = torch.randn(B, N, device=dev).half()
    out = ops.gemm(a, w.view(1, N, K), N, bias=b, bias_rows=rows, bias_stride=N)
    ref = (a.float() @ w.float().t()).view(B, rows, N) + b.float()[:, None, :]
    report(f'bias_rows B{B} rows{rows}', out, ref.view(-1, N))

print('device', torch.cuda.get_device_name(0), 'sms', _lib.lib().t2v_num_sms(), flush=True)
t_linear(128, 64, 64, bn=64)
t_linear(128, 64, 64, bn=64, bias=False)
t_linear(256, 128, 128, bn=128)
t_linear(1000, 320, 320)
t_linear(1000, 320, 320, bn=160, res=True)
t_linear(4096, 512, 256, bn=256)
t_linear(24576, 320, 2560)
t_linear(384, 1280, 1280, res=True)
t_linear(77, 1024, 640)
t_linear(512, 320, 4, bn=16)
t_conv2d(2, 16, 16, 64, 64)
t_conv2d(3, 32, 32, 320, 320)
t_conv2d(4, 8, 8, 128, 256)
t_conv2d(4, 4, 4, 256, 128)
t_conv2d(5, 2, 2, 64, 64)
t_conv2d(2, 16, 8, 64, 128)
t_conv2d(2, 18, 32, 64, 64)
t_conv2d(2, 9, 16, 64, 64)
t_conv2d(3, 16, 16, 8, 64)
t_conv2d(3, 16, 16, 320, 4)
t_tconv(1, 24, 256, 320)
t_tconv(2, 4, 16, 128)
t_tconv(2, 5, 4, 64)
t_tconv(1, 3, 128, 64)
t_geglu(1024, 320, 1280, 256)
t_geglu(512, 64, 256, 128)
t_geglu(300, 64, 256, 64)
t_batched(3, 256, 512)
t_batched(2, 128, 512)
t_bias_rows(2, 640, 320, 320)

# timing of a big GEMM
M, K, N = 24576, 1280, 2560
a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) / K ** 0.5).half().view(1, N, K)
out = torch.empty(M, N, device=dev, dtype=torch.half)
for bn in (256, 128):
    for _ in range(3): ops.gemm(a, w, N, out=out, force_bn=bn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm(a, w, N, out=out, force_bn=bn)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'gemm {M}x{N}x{K} bn{bn}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s', flush=True)
for _ in range(3): torch.matmul(a, w[0].t())
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): torch.matmul(a, w[0].t())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f'cublas same: {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TFLOP/s')
print('FAILS:', fails)
sys.exit(1 if fails else 0)
