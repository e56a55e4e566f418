"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the OpenCLIP ViT-H-14 text tower as FrozenOpenCLIPEmbedder drives it
(/root/reference/scripts/modelscope/clip_hardcode.py:112-119 encode_with_transformer, :269-274 text_transformer_forward with
layer = 'penultimate', :397-422 process_tokens).

The transformer itself lives in the third-party package `open_clip` (open_clip_torch; the reference pins no version --
requirements.txt lists only ffmpeg / audio helpers -- and the package is not installed in this container): PARITY UNPINNED
against open_clip itself.  What is restated is open_clip's published TextTransformer: pre-LN ResidualAttentionBlocks
    x = x + attn(ln_1(x)),  attn = nn.MultiheadAttention(width, heads) with the additive causal mask (upper triangle -inf)
    x = x + c_proj(gelu(c_fc(ln_2(x))))                        (nn.GELU, erf form; LayerScale is identity for ViT-H-14)
on x = token_embedding(tokens) + positional_embedding, then ln_final (no text projection on this path).  tests/ pin this
restatement against a module built from torch's own nn.MultiheadAttention / nn.LayerNorm / nn.GELU with the same weights.
"""
from dataclasses import dataclass
from typing import Dict, Tuple
import math

import torch
import torch.nn.functional as F


@dataclass
class ClipConfig:
    width: int = 1024
    heads: int = 16
    layers: int = 24
    layers_run: int = 23          # layer='penultimate': the loop breaks at i == layers - 1 (clip_hardcode.py:269-274)
    context: int = 77
    vocab: int = 49408


def clip_param_specs(cfg: ClipConfig) -> Dict[str, Tuple[int, ...]]:
    W = cfg.width
    s = {'token_embedding.weight': (cfg.vocab, W), 'positional_embedding': (cfg.context, W)}
    for i in range(cfg.layers_run):
        p = f'transformer.resblocks.{i}'
        s.update({p + '.ln_1.weight': (W,), p + '.ln_1.bias': (W,), p + '.attn.in_proj_weight': (3 * W, W),
                  p + '.attn.in_proj_bias': (3 * W,), p + '.attn.out_proj.weight': (W, W), p + '.attn.out_proj.bias': (W,),
                  p + '.ln_2.weight': (W,), p + '.ln_2.bias': (W,), p + '.mlp.c_fc.weight': (4 * W, W),
                  p + '.mlp.c_fc.bias': (4 * W,), p + '.mlp.c_proj.weight': (W, 4 * W), p + '.mlp.c_proj.bias': (W,)})
    s.update({'ln_final.weight': (W,), 'ln_final.bias': (W,)})
    return s


@torch.no_grad()
def clip_text_forward(Wt: Dict[str, torch.Tensor], cfg: ClipConfig, tokens):
    """tokens [B, context] int -> [B, context, width] = ln_final(resblocks[:layers_run](tok_emb + pos_emb))."""
    x = F.embedding(tokens.long(), Wt['token_embedding.weight']) + Wt['positional_embedding']
    B, L, W = x.shape
    H, d = cfg.heads, W // cfg.heads
    mask = torch.full((L, L), float('-inf'), dtype=x.dtype, device=x.device).triu_(1)       # open_clip build_attention_mask
    for i in range(cfg.layers_run):
        p = f'transformer.resblocks.{i}'
        h = F.layer_norm(x, (W,), Wt[p + '.ln_1.weight'], Wt[p + '.ln_1.bias'], 1e-5)
        qkv = F.linear(h, Wt[p + '.attn.in_proj_weight'], Wt[p + '.attn.in_proj_bias'])
        q, k, v = (t.reshape(B, L, H, d).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
        a = torch.softmax((q * (d ** -0.5)) @ k.transpose(-1, -2) + mask, dim=-1) @ v
        a = a.permute(0, 2, 1, 3).reshape(B, L, W)
        x = x + F.linear(a, Wt[p + '.attn.out_proj.weight'], Wt[p + '.attn.out_proj.bias'])
        h = F.layer_norm(x, (W,), Wt[p + '.ln_2.weight'], Wt[p + '.ln_2.bias'], 1e-5)
        h = F.gelu(F.linear(h, Wt[p + '.mlp.c_fc.weight'], Wt[p + '.mlp.c_fc.bias']))
        x = x + F.linear(h, Wt[p + '.mlp.c_proj.weight'], Wt[p + '.mlp.c_proj.bias'])
    return F.layer_norm(x, (W,), Wt['ln_final.weight'], Wt['ln_final.bias'], 1e-5)


def process_tokens(z, multipliers):
    """clip_hardcode.py:416-420: per-token emphasis weights, then the original mean is restored."""
    m = torch.as_tensor(multipliers, dtype=z.dtype, device=z.device)
    original_mean = z.mean()
    z = z * m.reshape(m.shape + (1,)).expand(z.shape)
    return z * (original_mean / z.mean())


class TorchTextTower(torch.nn.Module):
    """The same tower out of torch's own modules (nn.MultiheadAttention etc.) -- what open_clip's ResidualAttentionBlock is
    made of; used by the tests to pin clip_text_forward."""

    def __init__(self, cfg: ClipConfig):
        super().__init__()
        W = cfg.width
        self.cfg = cfg
        self.token_embedding = torch.nn.Embedding(cfg.vocab, W)
        self.positional_embedding = torch.nn.Parameter(torch.zeros(cfg.context, W))
        blocks = []
        for _ in range(cfg.layers_run):
            b = torch.nn.Module()
            b.ln_1 = torch.nn.LayerNorm(W)
            b.attn = torch.nn.MultiheadAttention(W, cfg.heads)
            b.ln_2 = torch.nn.LayerNorm(W)
            b.mlp = torch.nn.Sequential()
            b.mlp.add_module('c_fc', torch.nn.Linear(W, 4 * W))
            b.mlp.add_module('gelu', torch.nn.GELU())
            b.mlp.add_module('c_proj', torch.nn.Linear(4 * W, W))
            blocks.append(b)
        self.transformer = torch.nn.Module()
        self.transformer.resblocks = torch.nn.ModuleList(blocks)
        self.ln_final = torch.nn.LayerNorm(W)
        self.register_buffer('attn_mask', torch.full((cfg.context, cfg.context), float('-inf')).triu_(1), persistent=False)

    @torch.no_grad()
    def forward(self, tokens):
        x = self.token_embedding(tokens.long()) + self.positional_embedding
        x = x.permute(1, 0, 2)                                     # NLD -> LND (clip_hardcode.py:115)
        for r in self.transformer.resblocks:
            h = r.ln_1(x)
            x = x + r.attn(h, h, h, need_weights=False, attn_mask=self.attn_mask)[0]
            x = x + r.mlp(r.ln_2(x))
        return self.ln_final(x.permute(1, 0, 2))
