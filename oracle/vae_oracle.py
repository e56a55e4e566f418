"""TEST INFRASTRUCTURE ONLY -- CPU oracle for `AutoencoderKL.decode` and `AutoencoderKL.encode`.

Functional restatement of t2v_model.py:1646-1649 (post_quant_conv -> Decoder) where Decoder is the
un-vendored `ldm.modules.diffusionmodules.model.Decoder` (third-party package "stablediffusion",
version unpinned by the reference).  The reference vendors a twin of the same upstream code under
scripts/videocrafter/lvdm/models/modules/autoencoder_modules.py; citations below are into that
file.  Pinned against that module executed in-process (oracle/make_golden.py).
"""
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


@dataclass
class VAEConfig:
    """ddconfig hard-coded at t2v_pipeline.py:117-128."""
    ch: int = 128
    ch_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    z_channels: int = 4
    out_ch: int = 3
    embed_dim: int = 4


def decoder_param_specs(cfg: VAEConfig) -> Dict[str, Tuple[int, ...]]:
    """Keys as in autoencoder_modules.py:509-555 with the `decoder.` prefix, plus post_quant_conv
    (t2v_model.py:1604)."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(p, o, i, k):
        s[p + '.weight'] = (o, i, k, k)
        s[p + '.bias'] = (o,)

    def norm(p, c):
        s[p + '.weight'] = (c,)
        s[p + '.bias'] = (c,)

    def resnet(p, ci, co):
        norm(p + '.norm1', ci)
        conv(p + '.conv1', co, ci, 3)
        norm(p + '.norm2', co)
        conv(p + '.conv2', co, co, 3)
        if ci != co:
            conv(p + '.nin_shortcut', co, ci, 1)

    conv('post_quant_conv', cfg.z_channels, cfg.embed_dim, 1)
    nres = len(cfg.ch_mult)
    block_in = cfg.ch * cfg.ch_mult[-1]
    conv('decoder.conv_in', block_in, cfg.z_channels, 3)
    resnet('decoder.mid.block_1', block_in, block_in)
    norm('decoder.mid.attn_1.norm', block_in)
    for n in ('q', 'k', 'v', 'proj_out'):
        conv(f'decoder.mid.attn_1.{n}', block_in, block_in, 1)
    resnet('decoder.mid.block_2', block_in, block_in)
    for lvl in reversed(range(nres)):
        block_out = cfg.ch * cfg.ch_mult[lvl]
        for j in range(cfg.num_res_blocks + 1):
            resnet(f'decoder.up.{lvl}.block.{j}', block_in, block_out)
            block_in = block_out
        if lvl != 0:
            conv(f'decoder.up.{lvl}.upsample.conv', block_in, block_in, 3)
    norm('decoder.norm_out', block_in)
    conv('decoder.conv_out', cfg.out_ch, block_in, 3)
    return s


def _gn(W, p, x):
    return F.group_norm(x, 32, W[p + '.weight'], W[p + '.bias'], 1e-6)     # Normalize :33-34


def _swish(x):
    return x * torch.sigmoid(x)                                            # nonlinearity :29-31


def _resnet(W, p, x):
    """ResnetBlock.forward :207-228 with temb None."""
    h = F.conv2d(_swish(_gn(W, p + '.norm1', x)), W[p + '.conv1.weight'], W[p + '.conv1.bias'], padding=1)
    h = F.conv2d(_swish(_gn(W, p + '.norm2', h)), W[p + '.conv2.weight'], W[p + '.conv2.bias'], padding=1)
    if (p + '.nin_shortcut.weight') in W:
        x = F.conv2d(x, W[p + '.nin_shortcut.weight'], W[p + '.nin_shortcut.bias'])
    return x + h


def _attn(W, p, x):
    """AttnBlock.forward :91-116 -- single head, d = C, softmax over keys."""
    h = _gn(W, p + '.norm', x)
    q = F.conv2d(h, W[p + '.q.weight'], W[p + '.q.bias'])
    k = F.conv2d(h, W[p + '.k.weight'], W[p + '.k.bias'])
    v = F.conv2d(h, W[p + '.v.weight'], W[p + '.v.bias'])
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    a = torch.bmm(q, k) * (int(c) ** -0.5)
    a = F.softmax(a, dim=2)
    v = v.reshape(b, c, hh * ww)
    o = torch.bmm(v, a.permute(0, 2, 1)).reshape(b, c, hh, ww)
    o = F.conv2d(o, W[p + '.proj_out.weight'], W[p + '.proj_out.bias'])
    return x + o


@torch.no_grad()
def vae_decode(W: Dict[str, torch.Tensor], cfg: VAEConfig, z, taps=None):
    """AutoencoderKL.decode t2v_model.py:1646-1649; Decoder.forward autoencoder_modules.py:557-596.
    z [N,4,h,w] (already divided by 0.18215 by the caller, t2v_pipeline.py:348) -> [N,3,8h,8w]."""
    z = z.to(W['post_quant_conv.weight'].dtype)
    z = F.conv2d(z, W['post_quant_conv.weight'], W['post_quant_conv.bias'])
    h = F.conv2d(z, W['decoder.conv_in.weight'], W['decoder.conv_in.bias'], padding=1)
    h = _resnet(W, 'decoder.mid.block_1', h)
    h = _attn(W, 'decoder.mid.attn_1', h)
    h = _resnet(W, 'decoder.mid.block_2', h)
    if taps is not None:
        taps['mid'] = h
    for lvl in reversed(range(len(cfg.ch_mult))):
        for j in range(cfg.num_res_blocks + 1):
            h = _resnet(W, f'decoder.up.{lvl}.block.{j}', h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode='nearest')       # Upsample :161-165
            h = F.conv2d(h, W[f'decoder.up.{lvl}.upsample.conv.weight'],
                         W[f'decoder.up.{lvl}.upsample.conv.bias'], padding=1)
        if taps is not None:
            taps[f'up{lvl}'] = h
    h = _swish(_gn(W, 'decoder.norm_out', h))
    return F.conv2d(h, W['decoder.conv_out.weight'], W['decoder.conv_out.bias'], padding=1)


def encoder_param_specs(cfg: VAEConfig, in_channels: int = 3) -> Dict[str, Tuple[int, ...]]:
    """Keys of the ldm Encoder (autoencoder_modules.py:382-446) with the `encoder.` prefix, plus quant_conv
    (t2v_model.py:1603: Conv2d(2*z_channels, 2*embed_dim, 1))."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(p, o, i, k):
        s[p + '.weight'] = (o, i, k, k)
        s[p + '.bias'] = (o,)

    def norm(p, c):
        s[p + '.weight'] = (c,)
        s[p + '.bias'] = (c,)

    def resnet(p, ci, co):
        norm(p + '.norm1', ci)
        conv(p + '.conv1', co, ci, 3)
        norm(p + '.norm2', co)
        conv(p + '.conv2', co, co, 3)
        if ci != co:
            conv(p + '.nin_shortcut', co, ci, 1)

    conv('encoder.conv_in', cfg.ch, in_channels, 3)
    nres = len(cfg.ch_mult)
    in_mult = (1,) + tuple(cfg.ch_mult)
    block_in = cfg.ch
    for lvl in range(nres):
        block_in = cfg.ch * in_mult[lvl]
        block_out = cfg.ch * cfg.ch_mult[lvl]
        for j in range(cfg.num_res_blocks):
            resnet(f'encoder.down.{lvl}.block.{j}', block_in, block_out)
            block_in = block_out
        if lvl != nres - 1:
            conv(f'encoder.down.{lvl}.downsample.conv', block_in, block_in, 3)
    resnet('encoder.mid.block_1', block_in, block_in)
    norm('encoder.mid.attn_1.norm', block_in)
    for n in ('q', 'k', 'v', 'proj_out'):
        conv(f'encoder.mid.attn_1.{n}', block_in, block_in, 1)
    resnet('encoder.mid.block_2', block_in, block_in)
    norm('encoder.norm_out', block_in)
    conv('encoder.conv_out', 2 * cfg.z_channels, block_in, 3)
    conv('quant_conv', 2 * cfg.embed_dim, 2 * cfg.z_channels, 1)
    return s


@torch.no_grad()
def vae_encode_moments(W: Dict[str, torch.Tensor], cfg: VAEConfig, x):
    """AutoencoderKL.encode t2v_model.py:1640-1644 up to the moments: Encoder.forward autoencoder_modules.py:448-482, then
    quant_conv.  x [N,3,H,W] in [-1,1] -> moments [N, 2*embed_dim, H/8, W/8] = (mean | logvar); the latent the pipeline uses
    is `mean * 0.18215` (t2v_pipeline.py:181-183).  Downsample (autoencoder_modules.py:183-195): pad (0,1,0,1), 3x3 stride 2."""
    x = x.to(W['encoder.conv_in.weight'].dtype)
    h = F.conv2d(x, W['encoder.conv_in.weight'], W['encoder.conv_in.bias'], padding=1)
    nres = len(cfg.ch_mult)
    for lvl in range(nres):
        for j in range(cfg.num_res_blocks):
            h = _resnet(W, f'encoder.down.{lvl}.block.{j}', h)
        if lvl != nres - 1:
            h = F.pad(h, (0, 1, 0, 1), mode='constant', value=0)
            h = F.conv2d(h, W[f'encoder.down.{lvl}.downsample.conv.weight'], W[f'encoder.down.{lvl}.downsample.conv.bias'],
                         stride=2, padding=0)
    h = _resnet(W, 'encoder.mid.block_1', h)
    h = _attn(W, 'encoder.mid.attn_1', h)
    h = _resnet(W, 'encoder.mid.block_2', h)
    h = _swish(_gn(W, 'encoder.norm_out', h))
    h = F.conv2d(h, W['encoder.conv_out.weight'], W['encoder.conv_out.bias'], padding=1)
    return F.conv2d(h, W['quant_conv.weight'], W['quant_conv.bias'])


def tensor2vid_u8(video):
    """t2v_pipeline.py:447-460: [1,3,F,H,W] float in [-1,1] -> uint8 [F,H,W,3] (RGB), truncating cast."""
    v = video.float() * 0.5 + 0.5
    v = v.clamp(0, 1)
    v = v[0].permute(1, 2, 3, 0)          # f h w c   ('i c f h w -> f h (i w) c' with i = 1)
    return (v.numpy() * 255).astype('uint8')
