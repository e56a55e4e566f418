"""Thin Python wrappers over the kernel-level C entry points (t2v_op_*).  Used by the parity tests and by
small host-side utilities; the model-level path (t2v_unet_forward / t2v_vae_decode) does not go through here.

All tensors are CUDA fp16, channels-last token matrices [rows, C] unless stated otherwise.
"""
import ctypes as C

import torch

from . import _lib

GEMM_GEGLU = 1
GEMM_FORCE_BS, GEMM_NO_BS = 2048, 4096      # bring-up switches of t2v_op_gemm: B-stationary variant on / off
GEMM_OUT_F32 = 2


def _ia(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def gemm(a, w_packed, N, *, K=None, lda=None, dims=None, taps=None, n_alloc=None, b_batch_dim=-1, flags=0, out=None,
         ldo=None, bias=None, bias_rows=0, bias_stride=0, residual=None, ldr=None, alpha=1.0, force_bn=0, force_cg=0):
    """out[row, n] = alpha * sum_tap sum_k a[row + tap, k] w[tap, n, k] (+bias) (+residual)."""
    l = _lib.lib()
    rows = a.shape[0]
    K = K if K is not None else a.shape[1]
    lda = lda if lda is not None else a.stride(0)
    dims = list(dims) if dims is not None else [rows]
    nd = len(dims)
    taps = taps if taps is not None else [[0] * nd]
    flat = [o for t in taps for o in t]
    n_alloc = n_alloc if n_alloc is not None else w_packed.shape[-2]
    ncols = N // 2 if (flags & GEMM_GEGLU) else N
    if out is None:
        out = torch.empty((rows, ncols), device=a.device, dtype=torch.float32 if (flags & GEMM_OUT_F32) else torch.float16)
    ldo = ldo if ldo is not None else out.stride(0)
    ldr = ldr if ldr is not None else (residual.stride(0) if residual is not None else 0)
    rc = l.t2v_op_gemm(_lib.ptr(a), lda, K, nd, _ia(dims), len(taps), _ia(flat), _lib.ptr(w_packed), n_alloc, N,
                       b_batch_dim, flags, _lib.ptr(out), ldo, _lib.ptr(bias), bias_rows, bias_stride,
                       _lib.ptr(residual), ldr, alpha, force_bn, force_cg, _lib.stream_ptr())
    _lib.check(rc, 'op_gemm')
    return out


def pack_conv_weight(w, n_alloc=None, k_alloc=None):
    """w [Cout, Cin, *k] (fp16/fp32, CUDA) -> [taps, n_alloc, k_alloc] fp16."""
    l = _lib.lib()
    w = w.contiguous()
    cout, cin = w.shape[0], w.shape[1]
    taps = 1
    for s in w.shape[2:]:
        taps *= s
    n_alloc = n_alloc or cout
    k_alloc = k_alloc or cin
    dst = torch.empty((taps, n_alloc, k_alloc), device=w.device, dtype=torch.float16)
    rc = l.t2v_op_pack_conv_weight(_lib.ptr(w), int(w.dtype == torch.float32), _lib.ptr(dst), cout, cin, taps, n_alloc,
                                   k_alloc, _lib.stream_ptr())
    _lib.check(rc, 'pack_conv_weight')
    return dst


def pack_geglu_weight(w, b, bn):
    l = _lib.lib()
    w = w.contiguous()
    H = w.shape[0] // 2
    K = w.shape[1]
    wd = torch.empty((1, 2 * H, K), device=w.device, dtype=torch.float16)
    bd = torch.empty((2 * H,), device=w.device, dtype=torch.float16)
    rc = l.t2v_op_pack_geglu_weight(_lib.ptr(w), _lib.ptr(b.contiguous()), int(w.dtype == torch.float32), _lib.ptr(wd),
                                    _lib.ptr(bd), H, K, bn, _lib.stream_ptr())
    _lib.check(rc, 'pack_geglu_weight')
    return wd, bd


def conv_taps_2d():
    """tap = ky*3+kx over row dims (w, h, frames)."""
    return [[kx - 1, ky - 1, 0] for ky in range(3) for kx in range(3)]


def conv_taps_temporal():
    """tap = kt over row dims (pixels, frames, samples)."""
    return [[0, kt - 1, 0] for kt in range(3)]


def groupnorm(x, gamma, beta, rows_per_inst, eps, silu):
    l = _lib.lib()
    y = torch.empty_like(x)
    rc = l.t2v_op_groupnorm(_lib.ptr(x), x.stride(0), _lib.ptr(y), y.stride(0), x.shape[0], x.shape[1], rows_per_inst,
                            _lib.ptr(gamma), _lib.ptr(beta), eps, int(silu), _lib.stream_ptr())
    _lib.check(rc, 'op_groupnorm')
    return y


def layernorm(x, gamma, beta, eps=1e-5):
    l = _lib.lib()
    y = torch.empty_like(x)
    rc = l.t2v_op_layernorm(_lib.ptr(x), x.stride(0), _lib.ptr(y), y.stride(0), x.shape[0], x.shape[1], _lib.ptr(gamma),
                            _lib.ptr(beta), eps, _lib.stream_ptr())
    _lib.check(rc, 'op_layernorm')
    return y


def attention(q, k, v, o, q_bs, q_ss, k_bs, k_ss, v_bs, v_ss, o_bs, o_ss, batch, heads, sq, skv, kv_batch_div=1,
              scale=0.125):
    l = _lib.lib()
    rc = l.t2v_op_attention(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), q_bs, q_ss, k_bs, k_ss, v_bs, v_ss, o_bs,
                            o_ss, batch, heads, sq, skv, kv_batch_div, scale, _lib.stream_ptr())
    _lib.check(rc, 'op_attention')
    return o


def attention_hd(q, k, v, o, q_bs, q_ss, k_bs, k_ss, v_bs, v_ss, o_bs, o_ss, batch, heads, head_dim, sq, skv, kv_batch_div=1,
                 scale=None):
    """softmax(QK^T * scale)V for any supported head_dim (head h at column h*head_dim of the token matrices)."""
    l = _lib.lib()
    scale = head_dim ** -0.5 if scale is None else scale
    rc = l.t2v_op_attention_hd(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), q_bs, q_ss, k_bs, k_ss, v_bs, v_ss, o_bs,
                               o_ss, batch, heads, head_dim, sq, skv, kv_batch_div, scale, _lib.stream_ptr())
    _lib.check(rc, 'op_attention_hd')
    return o


def attention_relpos(q, k, v, o, table_k, table_v, n_seq, seq_inner, bs_outer, bs_inner, ss, o_bs_outer, o_bs_inner, o_ss,
                     heads, head_dim, T, max_rel, scale=None):
    """Temporal self-attention with relative-position key / value tables (VideoCrafter TemporalCrossAttention)."""
    l = _lib.lib()
    scale = head_dim ** -0.5 if scale is None else scale
    rc = l.t2v_op_attention_relpos(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), _lib.ptr(table_k), _lib.ptr(table_v),
                                   n_seq, seq_inner, bs_outer, bs_inner, ss, o_bs_outer, o_bs_inner, o_ss, heads, head_dim, T,
                                   max_rel, scale, _lib.stream_ptr())
    _lib.check(rc, 'op_attention_relpos')
    return o
