"""In-kernel timeline of gemm_tc_kernel (needs the trace build: T2V_BUILD_OUT=.../libt2v_b200_trace.so T2V_BUILD_DIR=build_trace
csrc/build.sh -DT2V_GEMM_TRACE=1, then T2V_LIB_PATH=<that .so> python scripts/gemm_trace.py).  Two CTAs (first, middle of the grid)
stamp clock64 where each role stops waiting; this prints, per case, how a tile's time splits between waiting for operands,
for the accumulator to drain, and the epilogue's own phases -- i.e. WHICH role is the bottleneck of the layers that run
far below the tensor peak.  Numbers are SM cycles (1.9 GHz under load)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch                       # noqa: E402
from t2v_b200 import ops, _lib     # noqa: E402

CAP = 4096
dev = 'cuda'


def run_case(name, M, K, N, res, flags=0, taps=None, dims=None, geglu=False, force_bn=0):
    a = torch.randn(M, K, device=dev).half()
    nt = 1 if taps is None else len(taps)
    w = (torch.randn(nt, N, K, device=dev) / (K * nt) ** 0.5).half()
    b = torch.randn(N, device=dev).half()
    r = torch.randn(M, N, device=dev).half() if res else None
    out = torch.empty(M, N // 2 if geglu else N, device=dev, dtype=torch.half)
    fl = flags | (ops.GEMM_GEGLU if geglu else 0)

    def run():
        ops.gemm(a, w, N, bias=b, residual=r, out=out, flags=fl, taps=taps, dims=dims, force_bn=force_bn)
    l = _lib.lib()
    l.t2v_debug_gemm_trace.argtypes = [C.c_void_p]
    l.t2v_debug_gemm_trace.restype = C.c_int
    l.t2v_debug_gemm_trace(None)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    buf = torch.zeros(2 * 4 * CAP, dtype=torch.int64, device=dev)
    assert l.t2v_debug_gemm_trace(C.c_void_p(buf.data_ptr())) == 0
    run()
    torch.cuda.synchronize()
    l.t2v_debug_gemm_trace(None)
    t = buf.cpu().view(2, 4, CAP)
    print(f'=== {name}: {us:.1f} us, {2.0 * M * N * K * nt / us / 1e6:.0f} TF/s')
    for cta in range(2):
        ev = {}
        for role in range(4):
            row = t[cta, role]
            n = int((row != 0).sum())
            ev[role] = [(int(v) >> 8, int(v) & 255) for v in row[:n].tolist()]
        allt = [x[0] for role in ev for x in ev[role]]
        if not allt:
            print(f'  cta slot {cta}: no events')
            continue
        t0 = min(allt)
        span = max(allt) - t0
        prod = [x[0] - t0 for x in ev[0] if x[1] == 1]
        mma_full = [x[0] - t0 for x in ev[1] if x[1] == 4]
        # per-k-iteration: TMA issue -> operands landed (same index: the ring is FIFO)
        lat = [m - p for p, m in zip(prod, mma_full)]
        # MMA: time blocked on the accumulator (tag 2 -> 3) per tile; time from tile start (3) to last operands landed
        mma = ev[1]
        acc_wait, tiles_mma = [], []
        for i, (ts, tag) in enumerate(mma):
            if tag == 2 and i + 1 < len(mma) and mma[i + 1][1] == 3:
                acc_wait.append(mma[i + 1][0] - ts)
                tiles_mma.append(mma[i + 1][0] - t0)
        print(f'  cta slot {cta}: span {span} clk, {len(prod)} k-iters, {len(tiles_mma)} tiles -> {span / max(1, len(tiles_mma)):.0f} clk/tile')
        if lat:
            lat_s = sorted(lat)
            print(f'    TMA issue -> operands seen by MMA: median {lat_s[len(lat_s) // 2]} clk, p10 {lat_s[len(lat_s) // 10]}, p90 {lat_s[len(lat_s) * 9 // 10]}')
            gaps = [b_ - a_ for a_, b_ in zip(prod[:-1], prod[1:])]
            gs = sorted(gaps)
            print(f'    producer issue gaps: median {gs[len(gs) // 2]} clk, mean {sum(gaps) / len(gaps):.0f}')
            mg = [b_ - a_ for a_, b_ in zip(mma_full[:-1], mma_full[1:])]
            ms = sorted(mg)
            print(f'    MMA k-iter gaps (operands landed): median {ms[len(ms) // 2]} clk, mean {sum(mg) / len(mg):.0f}')
        if acc_wait:
            print(f'    MMA blocked on accumulator drain per tile: mean {sum(acc_wait) / len(acc_wait):.0f} clk (sum {sum(acc_wait)} = {100.0 * sum(acc_wait) / span:.0f} % of span)')
        for role in (2, 3):
            e = ev[role]
            if not e:
                continue
            wait_acc, chunk_ld, stage_wait, pack, tail, tile_len = [], [], [], [], [], []
            last = None
            t5 = None
            for ts, tag in e:
                if tag == 5:
                    t5 = ts
                    last = ts
                elif tag == 6:
                    wait_acc.append(ts - last); last = ts
                elif tag == 7:
                    chunk_ld.append(ts - last); last = ts
                elif tag == 8:
                    stage_wait.append(ts - last); last = ts
                elif tag == 9:
                    pack.append(ts - last); last = ts
                elif tag == 10:
                    tail.append(ts - last); last = ts
                    if t5 is not None:
                        tile_len.append(ts - t5)

            def m(x):
                return sum(x) / len(x) if x else 0.0
            print(f'    epilogue warp role {role}: tiles {len(tile_len)}, per tile {m(tile_len):.0f} clk = wait-for-accumulator {m(wait_acc):.0f} + '
                  f'chunks x [tmem ld {m(chunk_ld):.0f} + staging-free wait {m(stage_wait):.0f} + math/st.shared/fence {m(pack):.0f}] '
                  f'(chunks/tile {len(chunk_ld) / max(1, len(tile_len)):.1f}) + tail {m(tail):.0f}')
        if cta == 0 and os.environ.get('TRACE_DUMP'):
            for role in range(4):
                print('    raw role', role, ' '.join(f'{x[0] - t0}:{x[1]}' for x in ev[role][:int(os.environ.get('TRACE_DUMP'))]))


cases = [('L0 to_out 49152x320x320 +res', 49152, 320, 320, True, 0, None, None, False),
         ('L0 proj 49152x320x320', 49152, 320, 320, False, 0, None, None, False),
         ('L0 qkv 49152x960x320', 49152, 320, 960, False, 0, None, None, False),
         ('L0 qkv 49152x960x320 streaming', 49152, 320, 960, False, ops.GEMM_NO_BS, None, None, False),
         ('L0 ff2 49152x320x1280 +res', 49152, 1280, 320, True, 0, None, None, False),
         ('L1 to_out 12288x640x640 +res', 12288, 640, 640, True, 0, None, None, False),
         ('L2 to_out 3072x1280x1280 +res', 3072, 1280, 1280, True, 0, None, None, False),
         ('L1 conv3x3 12288 640->640', 12288, 640, 640, False, 0, ops.conv_taps_2d(), [16, 16, 48], False),
         ('L0 geglu 49152x2560x320', 49152, 320, 2560, False, 0, None, None, True)]
sel = os.environ.get('CASES')
for c in cases:
    if sel and not any(k in c[0] for k in sel.split(',')):
        continue
    try:
        run_case(c[0], c[1], c[2], c[3], c[4], flags=c[5], taps=c[6], dims=c[7], geglu=c[8])
    except Exception as ex:
        print('case failed', c[0], ex)
