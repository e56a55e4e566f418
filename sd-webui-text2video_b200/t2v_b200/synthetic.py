"""Seeded random-init weights for benchmarking (no checkpoints can be downloaded in this environment).

Every residual branch of the reference is zero-initialised (t2v_model.py:326,:631-636,:708-713,:955-956,:1214-1216),
which would make a random-init network skip most of its arithmetic numerically (not in time); we draw every tensor
non-zero with variance-preserving scales so activations stay O(1..10) in fp16 through all 28 blocks."""
import math

import torch


@torch.no_grad()
def randomize_(module, seed=0, gain=1.0):
    """In-place, on whatever device the parameters live on (use after .cuda() for speed)."""
    params = sorted(module.named_parameters(), key=lambda kv: kv[0])
    dev = params[0][1].device
    g = torch.Generator(device=dev).manual_seed(seed)
    for name, p in params:
        if p.dim() == 1:
            v = torch.randn(p.shape, generator=g, device=dev)
            if name.endswith('.weight'):
                v = 1.0 + 0.1 * v
            else:
                v = 0.05 * v
        else:
            fan_in = p[0].numel()
            v = torch.randn(p.shape, generator=g, device=dev) * (gain / math.sqrt(fan_in))
        p.copy_(v.to(p.dtype))
    if hasattr(module, 'mark_dirty'):
        module.mark_dirty()
    return module
