"""CPU: the C-ABI shared library loads and exports every symbol include/t2v_b200.h declares; the ctypes table
mirrors the header; host-only entry points (create / param_info / flops) work without a GPU."""
import ctypes as C
import os
import re

import pytest

from t2v_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 't2v_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(t2v_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), 'run `python -c "import __graft_entry__ as g; g.build()"` first'
    lib = C.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/t2v_b200.h but not exported'


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_host_only_calls_work_without_gpu():
    l = _lib.load_library()
    assert b't2v_b200' in l.t2v_version()
    from t2v_b200.modules import UNetSD
    net = UNetSD(dim=64)
    assert len(net.state_dict()) == 1480
    fl = net.flops(1, 24, 32, 32, 77)
    assert fl > 0
    big = UNetSD.__new__(UNetSD)      # flop model at the public config without allocating 1.4 B parameters
    cfg = _lib.UNetConfigC()
    cfg.in_dim, cfg.dim, cfg.context_dim, cfg.out_dim = 4, 320, 1024, 4
    for i, m in enumerate((1, 2, 4, 4)):
        cfg.dim_mult[i] = m
    cfg.n_mult, cfg.num_heads, cfg.head_dim, cfg.num_res_blocks = 4, 8, 64, 2
    for i, s in enumerate((1.0, 0.5, 0.25)):
        cfg.attn_scales[i] = s
    cfg.n_attn_scales = 3
    h = C.c_void_p()
    assert l.t2v_unet_create(C.byref(cfg), C.byref(h)) == 0
    fl = l.t2v_unet_flops(h, 1, 24, 32, 32, 77)
    l.t2v_unet_destroy(h)
    # 7.336 TFLOP per forward at 24f x 256^2 (torch flop counter on the reference); ours counts the
    # cross-attention K/V projections once per sample instead of once per frame (-0.09 TF) -> within 3 %
    assert abs(fl / 1e12 - 7.336) / 7.336 < 0.03, fl / 1e12


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        _lib.lib()
