"""One launch of the tcgen05 attention kernel at the cfg-3 L0 shape (for ncu --set full)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'sd-webui-text2video_b200'))
from t2v_b200 import ops  # noqa: E402

batch, heads, S = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (8, 5, 9216)))
C = heads * 64
qkv = torch.randn(batch * S, 3 * C, device='cuda').half()
o = torch.zeros(batch * S, C, device='cuda', dtype=torch.half)
ld = 3 * C
for _ in range(2):
    ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], o, S * ld, ld, S * ld, ld, S * ld, ld, S * C, C, batch, heads, S, S)
torch.cuda.synchronize()
