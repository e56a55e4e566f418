"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the VideoCrafter denoiser.

Functional (weights-dict) restatement of `UNetModel.forward`
(/root/reference/scripts/videocrafter/lvdm/models/modules/openaimodel3d.py:632-670) as configured by
base_t2v/model_config.yaml:21-46, plus the pieces it is built from:

  * `ResBlock._forward`                    openaimodel3d.py:244-271  (Conv3d (1,3,3); GroupNorm32 = fp32 stats over T*H*W, util.py:271-273)
  * `Downsample` / `Upsample`              openaimodel3d.py:110-140 / :57-92 (stride (1,2,2) conv / nearest x2 in H,W then conv)
  * `SpatialTemporalTransformer.forward`   attention_temporal.py:386-399
  * `BasicTransformerBlockST._forward`     attention_temporal.py:301-335 (spatial self, temporal self, spatial cross, temporal "cross"
                                           with context None = self, GEGLU feed-forward)
  * `CrossAttention.forward`               attention_temporal.py:167-190 (einsum + softmax)
  * `TemporalCrossAttention.forward`       attention_temporal.py:107-144 with `RelativePosition` :46-65
  * `timestep_embedding`                   util.py:142-162

The sampler of this path (lvdm/samplers/ddim.py:135-279) is arithmetically the ldm DDIM already restated in
samplers_oracle.ddim_sample (same `make_ddim_timesteps` / `make_ddim_sampling_parameters`, util.py:36-63; CFG
`e_u + g (e_c - e_u)`, ddim.py:229; x0 / direction / x_prev ddim.py:262-277).  `vc_ddim_sample` below differs from it only
where the VideoCrafter copy does: 5-D-aware coefficient tensors (ddim.py:253-260) and per-step noise from the sampler's own
CPU generator (util.py:321-325) instead of the global RNG.  Quirk kept out of the harness: ddim.py:148-149 draws x_T from
the GLOBAL RNG (the UI seed does not determine it), so x_T is always passed explicitly.

Pinned against the unmodified reference modules executed in this container by oracle/make_golden.py (max |diff| printed
there and asserted in tests/test_oracle_golden.py); only tests/, smoke() and bench.py's CPU leg may import this file.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

import numpy as np

from .samplers_oracle import ddim_schedule


@dataclass
class VCConfig:
    """base_t2v/model_config.yaml:21-46"""
    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    attention_resolutions: Tuple[int, ...] = (4, 2, 1)
    num_res_blocks: int = 2
    channel_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_heads: int = 8
    context_dim: int = 768
    temporal_length: int = 16
    use_relative_position: bool = True


@dataclass
class VCBlock:
    kind: str                 # 'conv' | 'res' | 'st' | 'down' | 'up'
    prefix: str
    cin: int = 0
    cout: int = 0
    heads: int = 0
    dim_head: int = 0


@dataclass
class VCLayout:
    input_blocks: List[List[VCBlock]] = field(default_factory=list)
    middle: List[VCBlock] = field(default_factory=list)
    output_blocks: List[List[VCBlock]] = field(default_factory=list)


def vc_enumerate(cfg: VCConfig) -> VCLayout:
    """Mirrors UNetModel.__init__ (openaimodel3d.py:407-617): legacy=False, num_head_channels=-1 -> dim_head = ch // num_heads."""
    L = VCLayout()
    mc = cfg.model_channels
    L.input_blocks.append([VCBlock('conv', 'input_blocks.0.0', cfg.in_channels, mc)])
    chans = [mc]
    ch, ds, idx = mc, 1, 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [VCBlock('res', f'input_blocks.{idx}.0', ch, mult * mc)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers.append(VCBlock('st', f'input_blocks.{idx}.1', ch, ch, cfg.num_heads, ch // cfg.num_heads))
            L.input_blocks.append(layers)
            chans.append(ch)
            idx += 1
        if level != len(cfg.channel_mult) - 1:
            L.input_blocks.append([VCBlock('down', f'input_blocks.{idx}.0', ch, ch)])
            chans.append(ch)
            idx += 1
            ds *= 2
    L.middle = [VCBlock('res', 'middle_block.0', ch, ch),
                VCBlock('st', 'middle_block.1', ch, ch, cfg.num_heads, ch // cfg.num_heads),
                VCBlock('res', 'middle_block.2', ch, ch)]
    oidx = 0
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            layers = [VCBlock('res', f'output_blocks.{oidx}.0', ch + ich, mc * mult)]
            ch = mc * mult
            if ds in cfg.attention_resolutions:
                layers.append(VCBlock('st', f'output_blocks.{oidx}.{len(layers)}', ch, ch, cfg.num_heads, ch // cfg.num_heads))
            if level and i == cfg.num_res_blocks:
                layers.append(VCBlock('up', f'output_blocks.{oidx}.{len(layers)}', ch, ch))
                ds //= 2
            L.output_blocks.append(layers)
            oidx += 1
    return L


def vc_param_specs(cfg: VCConfig) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys / shapes of the reference UNetModel (the on-disk format `model.diffusion_model.*` minus that prefix)."""
    S: Dict[str, Tuple[int, ...]] = {}
    mc = cfg.model_channels
    ted = 4 * mc
    S['time_embed.0.weight'] = (ted, mc)
    S['time_embed.0.bias'] = (ted,)
    S['time_embed.2.weight'] = (ted, ted)
    S['time_embed.2.bias'] = (ted,)

    def add(b: VCBlock):
        p = b.prefix
        if b.kind == 'conv':
            S[p + '.weight'] = (b.cout, b.cin, 1, 3, 3)
            S[p + '.bias'] = (b.cout,)
        elif b.kind == 'down':
            S[p + '.op.weight'] = (b.cout, b.cin, 1, 3, 3)
            S[p + '.op.bias'] = (b.cout,)
        elif b.kind == 'up':
            S[p + '.conv.weight'] = (b.cout, b.cin, 1, 3, 3)
            S[p + '.conv.bias'] = (b.cout,)
        elif b.kind == 'res':
            S[p + '.in_layers.0.weight'] = (b.cin,)
            S[p + '.in_layers.0.bias'] = (b.cin,)
            S[p + '.in_layers.2.weight'] = (b.cout, b.cin, 1, 3, 3)
            S[p + '.in_layers.2.bias'] = (b.cout,)
            S[p + '.emb_layers.1.weight'] = (b.cout, ted)
            S[p + '.emb_layers.1.bias'] = (b.cout,)
            S[p + '.out_layers.0.weight'] = (b.cout,)
            S[p + '.out_layers.0.bias'] = (b.cout,)
            S[p + '.out_layers.3.weight'] = (b.cout, b.cout, 1, 3, 3)
            S[p + '.out_layers.3.bias'] = (b.cout,)
            if b.cin != b.cout:
                S[p + '.skip_connection.weight'] = (b.cout, b.cin, 1, 1, 1)
                S[p + '.skip_connection.bias'] = (b.cout,)
        elif b.kind == 'st':
            C, d = b.cin, b.dim_head
            inner = b.heads * d
            S[p + '.norm.weight'] = (C,)
            S[p + '.norm.bias'] = (C,)
            S[p + '.proj_in.weight'] = (inner, C, 1, 1, 1)
            S[p + '.proj_in.bias'] = (inner,)
            S[p + '.proj_out.weight'] = (C, inner, 1, 1, 1)
            S[p + '.proj_out.bias'] = (C,)
            t = p + '.transformer_blocks.0'
            for a, kdim in (('attn1', inner), ('attn2', cfg.context_dim), ('attn1_tmp', inner), ('attn2_tmp', inner)):
                S[f'{t}.{a}.to_q.weight'] = (inner, inner)
                S[f'{t}.{a}.to_k.weight'] = (inner, kdim)
                S[f'{t}.{a}.to_v.weight'] = (inner, kdim)
                S[f'{t}.{a}.to_out.0.weight'] = (inner, inner)
                S[f'{t}.{a}.to_out.0.bias'] = (inner,)
                if a.endswith('_tmp') and cfg.use_relative_position:
                    S[f'{t}.{a}.relative_position_k.embeddings_table'] = (2 * cfg.temporal_length + 1, d)
                    S[f'{t}.{a}.relative_position_v.embeddings_table'] = (2 * cfg.temporal_length + 1, d)
            S[f'{t}.ff.net.0.proj.weight'] = (8 * inner, inner)
            S[f'{t}.ff.net.0.proj.bias'] = (8 * inner,)
            S[f'{t}.ff.net.2.weight'] = (inner, 4 * inner)
            S[f'{t}.ff.net.2.bias'] = (inner,)
            for n in range(1, 6):
                S[f'{t}.norm{n}.weight'] = (inner,)
                S[f'{t}.norm{n}.bias'] = (inner,)

    L = vc_enumerate(cfg)
    for blk in L.input_blocks:
        for b in blk:
            add(b)
    for b in L.middle:
        add(b)
    for blk in L.output_blocks:
        for b in blk:
            add(b)
    S['out.0.weight'] = (mc,)
    S['out.0.bias'] = (mc,)
    S['out.2.weight'] = (cfg.out_channels, mc, 1, 3, 3)
    S['out.2.bias'] = (cfg.out_channels,)
    return S


def vc_timestep_embedding(t, dim, max_period=10000):
    """util.py:142-162 (repeat_only=False): [cos | sin], freqs = exp(-ln(max_period) * i / half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(W, p, x, eps):
    # GroupNorm32 (util.py:271-273): fp32 statistics over (C/32, T, H, W); Normalize (util.py:198-199) the same with eps 1e-6
    return F.group_norm(x.float(), 32, W[p + '.weight'].float(), W[p + '.bias'].float(), eps).type(x.dtype)


def _split_heads(t, h):
    b, n, _ = t.shape
    return t.reshape(b, n, h, -1).permute(0, 2, 1, 3).reshape(b * h, n, -1)


def _merge_heads(t, h):
    bh, n, d = t.shape
    return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, h * d)


def _cross_attention(W, p, x, ctx, heads):
    """attention_temporal.py:167-190"""
    ctx = x if ctx is None else ctx
    q = F.linear(x, W[p + '.to_q.weight'])
    k = F.linear(ctx, W[p + '.to_k.weight'])
    v = F.linear(ctx, W[p + '.to_v.weight'])
    d = q.shape[-1] // heads
    q, k, v = (_split_heads(t, heads) for t in (q, k, v))
    sim = torch.einsum('bid,bjd->bij', q, k) * d ** -0.5
    attn = sim.softmax(dim=-1)
    out = _merge_heads(torch.einsum('bij,bjd->bid', attn, v), heads)
    return F.linear(out, W[p + '.to_out.0.weight'], W[p + '.to_out.0.bias'])


def _relative_position(table, lq, lk, max_rel):
    """attention_temporal.py:56-65: table[clamp(k - q, -L, L) + L] -> [lq, lk, d]"""
    rq = torch.arange(lq, device=table.device)
    rk = torch.arange(lk, device=table.device)
    idx = torch.clamp(rk[None, :] - rq[:, None], -max_rel, max_rel) + max_rel
    return table[idx.long()]


def _temporal_attention(W, p, x, heads, cfg: VCConfig):
    """attention_temporal.py:107-144 with context = x, mask = None"""
    q = F.linear(x, W[p + '.to_q.weight'])
    k = F.linear(x, W[p + '.to_k.weight'])
    v = F.linear(x, W[p + '.to_v.weight'])
    d = q.shape[-1] // heads
    scale = d ** -0.5
    q, k, v = (_split_heads(t, heads) for t in (q, k, v))
    sim = torch.einsum('bid,bjd->bij', q, k) * scale
    if cfg.use_relative_position:
        k2 = _relative_position(W[p + '.relative_position_k.embeddings_table'], q.shape[1], k.shape[1], cfg.temporal_length)
        sim = sim + torch.einsum('btd,tsd->bts', q, k2) * scale
    attn = sim.softmax(dim=-1)
    out = torch.einsum('bij,bjd->bid', attn, v)
    if cfg.use_relative_position:
        v2 = _relative_position(W[p + '.relative_position_v.embeddings_table'], q.shape[1], v.shape[1], cfg.temporal_length)
        out = out + torch.einsum('bts,tsd->btd', attn, v2)
    return F.linear(_merge_heads(out, heads), W[p + '.to_out.0.weight'], W[p + '.to_out.0.bias'])


def _ln(W, p, x):
    return F.layer_norm(x, (x.shape[-1],), W[p + '.weight'], W[p + '.bias'], 1e-5)


def _st_block(W, p, x, ctx, heads, cfg: VCConfig):
    """attention_temporal.py:301-335 on x [b, c, t, h, w]"""
    b, c, t, h, w = x.shape

    def to_spatial(z):
        return z.permute(0, 2, 3, 4, 1).reshape(b * t, h * w, c)

    def from_spatial(z):
        return z.reshape(b, t, h, w, c).permute(0, 4, 1, 2, 3)

    def to_temporal(z):
        return z.permute(0, 3, 4, 2, 1).reshape(b * h * w, t, c)

    def from_temporal(z):
        return z.reshape(b, h, w, t, c).permute(0, 4, 3, 1, 2)

    s = to_spatial(x)
    s = _cross_attention(W, p + '.attn1', _ln(W, p + '.norm1', s), None, heads) + s
    x = from_spatial(s)
    m = to_temporal(x)
    m = _temporal_attention(W, p + '.attn1_tmp', _ln(W, p + '.norm4', m), heads, cfg) + m
    x = from_temporal(m)
    s = to_spatial(x)
    ctx_rep = None if ctx is None else ctx.repeat_interleave(t, dim=0)          # :321-325
    s = _cross_attention(W, p + '.attn2', _ln(W, p + '.norm2', s), ctx_rep, heads) + s
    x = from_spatial(s)
    m = to_temporal(x)
    m = _temporal_attention(W, p + '.attn2_tmp', _ln(W, p + '.norm5', m), heads, cfg) + m
    hgl = F.linear(_ln(W, p + '.norm3', m), W[p + '.ff.net.0.proj.weight'], W[p + '.ff.net.0.proj.bias'])
    a, gate = hgl.chunk(2, dim=-1)
    m = F.linear(a * F.gelu(gate), W[p + '.ff.net.2.weight'], W[p + '.ff.net.2.bias']) + m
    return from_temporal(m)


def _st(W, b: VCBlock, x, ctx, cfg: VCConfig):
    """attention_temporal.py:386-399"""
    p = b.prefix
    h = _gn(W, p + '.norm', x, 1e-6)
    h = F.conv3d(h, W[p + '.proj_in.weight'], W[p + '.proj_in.bias'])
    h = _st_block(W, p + '.transformer_blocks.0', h, ctx, b.heads, cfg)
    h = F.conv3d(h, W[p + '.proj_out.weight'], W[p + '.proj_out.bias'])
    return h + x


def _res(W, b: VCBlock, x, emb):
    """openaimodel3d.py:244-271 (use_scale_shift_norm False, no up/down)"""
    p = b.prefix
    h = F.silu(_gn(W, p + '.in_layers.0', x, 1e-5))
    h = F.conv3d(h, W[p + '.in_layers.2.weight'], W[p + '.in_layers.2.bias'], padding=(0, 1, 1))
    e = F.linear(F.silu(emb), W[p + '.emb_layers.1.weight'], W[p + '.emb_layers.1.bias']).type(h.dtype)
    h = h + e[:, :, None, None, None]
    h = F.silu(_gn(W, p + '.out_layers.0', h, 1e-5))
    h = F.conv3d(h, W[p + '.out_layers.3.weight'], W[p + '.out_layers.3.bias'], padding=(0, 1, 1))
    if b.cin != b.cout:
        x = F.conv3d(x, W[p + '.skip_connection.weight'], W[p + '.skip_connection.bias'])
    return x + h


def _run(W, blk: List[VCBlock], h, emb, ctx, cfg: VCConfig, taps=None):
    for b in blk:
        if b.kind == 'conv':
            h = F.conv3d(h, W[b.prefix + '.weight'], W[b.prefix + '.bias'], padding=(0, 1, 1))
        elif b.kind == 'res':
            h = _res(W, b, h, emb)
        elif b.kind == 'st':
            h = _st(W, b, h, ctx, cfg)
        elif b.kind == 'down':
            h = F.conv3d(h, W[b.prefix + '.op.weight'], W[b.prefix + '.op.bias'], stride=(1, 2, 2), padding=(0, 1, 1))
        elif b.kind == 'up':
            h = F.interpolate(h, (h.shape[2], h.shape[3] * 2, h.shape[4] * 2), mode='nearest')
            h = F.conv3d(h, W[b.prefix + '.conv.weight'], W[b.prefix + '.conv.bias'], padding=(0, 1, 1))
        if taps is not None:
            taps[b.prefix] = h
    return h


def vc_unet_forward(W: Dict[str, torch.Tensor], cfg: VCConfig, x, t, ctx, taps=None):
    """UNetModel.forward (openaimodel3d.py:632-670): x [B,4,T,h,w], t [B], ctx [B,77,context_dim] -> eps [B,4,T,h,w]."""
    L = vc_enumerate(cfg)
    emb = vc_timestep_embedding(t, cfg.model_channels)
    emb = F.linear(emb, W['time_embed.0.weight'], W['time_embed.0.bias'])
    emb = F.linear(F.silu(emb), W['time_embed.2.weight'], W['time_embed.2.bias'])
    hs = []
    h = x
    for blk in L.input_blocks:
        h = _run(W, blk, h, emb, ctx, cfg, taps)
        hs.append(h)
    h = _run(W, L.middle, h, emb, ctx, cfg, taps)
    for blk in L.output_blocks:
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run(W, blk, h, emb, ctx, cfg, taps)
    h = F.silu(_gn(W, 'out.0', h, 1e-5))
    return F.conv3d(h, W['out.2.weight'], W['out.2.bias'], padding=(0, 1, 1))


@torch.no_grad()
def vc_ddim_sample(model, betas, x_T, S, cond, uncond, guide_scale, eta=0.0, noise_gen=None, trace=None):
    """DDIMSampler.sample / ddim_sampling / p_sample_ddim (lvdm/samplers/ddim.py:62-279); `model(x, t, c)` = apply_model."""
    acp = torch.cumprod(1 - betas, dim=0)
    ts, alphas, alphas_prev, sigmas = ddim_schedule(acp, S, eta)
    sqrt_1m = np.sqrt(1.0 - alphas)
    img = x_T
    b = img.shape[0]
    size = (b,) + (1,) * (img.dim() - 1)
    total = ts.shape[0]
    if noise_gen is None:
        noise_gen = torch.Generator(device='cpu')
    for i, step in enumerate(np.flip(ts)):
        index = total - i - 1
        t = torch.full((b,), int(step), dtype=torch.long, device=img.device)
        if uncond is None or guide_scale == 1.0:
            e_t = model(img, t, cond)
        else:
            e_c = model(img, t, cond)
            e_u = model(img, t, uncond)
            e_t = e_u + guide_scale * (e_c - e_u)
        dev = img.device
        a_t = torch.full(size, float(alphas[index]), device=dev)
        a_prev = torch.full(size, float(alphas_prev[index]), device=dev)
        sigma_t = torch.full(size, float(sigmas[index]), device=dev)
        s1m = torch.full(size, float(sqrt_1m[index]), device=dev)
        pred_x0 = (img - s1m * e_t) / a_t.sqrt()
        dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
        noise = sigma_t * torch.randn(img.shape, generator=noise_gen).to(dev) * 1.0
        img = a_prev.sqrt() * pred_x0 + dir_xt + noise
        if trace is not None:
            trace.append(img.clone())
    return img
