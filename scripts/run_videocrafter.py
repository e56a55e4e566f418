"""BASELINE config 5: VideoCrafter base_t2v, 16 frames x 256^2, 50 DDIM steps, through the public entry point
(synthetic weights).  Prints one JSON line: clip seconds, frames/s, achieved TFLOP/s, forward breakdown."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch
from t2v_b200 import videocrafter as vcm
from t2v_b200.synthetic import randomize_
m = vcm.LatentDiffusion().half()
randomize_(m.model.diffusion_model, seed=0); randomize_(m.first_stage_model, seed=3)
m = m.cuda().eval()
g = torch.Generator().manual_seed(2)
c = torch.randn(1, 77, 768, generator=g).half().cuda(); uc = torch.randn(1, 77, 768, generator=g).half().cuda()
x_T = torch.randn(1, 4, 16, 32, 32, generator=torch.Generator().manual_seed(123)).cuda()
steps, frames = 50, 16
args = dict(prompt_embeds=c, n_prompt_embeds=uc, steps=steps, frames=frames, seed=123, cfg_scale=15.0, eta=1.0, batch_count=1, x_T=x_T)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = vcm.process_videocrafter(args, model=m)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
unet = m.model.diffusion_model
fl = steps * unet.flops(2, frames, 32, 32, 77) + m.first_stage_model.flops(frames, 32, 32)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
xb = torch.randn(2, 4, 16, 32, 32, device='cuda'); tb = torch.full((2,), 500, device='cuda'); cb = torch.cat([c, uc])
for _ in range(3): unet(xb, tb, context=cb)
e0.record()
for _ in range(10): unet(xb, tb, context=cb)
e1.record(); torch.cuda.synchronize()
fwd = e0.elapsed_time(e1) / 10
prof = unet.profile(2, frames, 32, 32, 77)
print(json.dumps({'config': 'cfg5 VideoCrafter base_t2v 16f x 256^2, 50-step DDIM (eta 1.0, cfg 15), B=2 cond+uncond forward',
                  'clip_s': round(dt, 3), 'frames_per_s': round(frames / dt, 2), 'tflop_per_clip': round(fl / 1e12, 1),
                  'achieved_tflops': round(fl / dt / 1e12, 1), 'forward_ms_B2': round(fwd, 2),
                  'forward_tflops': round(unet.flops(2, frames, 32, 32, 77) / fwd / 1e9, 1), 'launches': unet.num_launches(),
                  'out_shape': list(out[0].shape), 'out_mean': float(out[0].mean()),
                  'families_ms': {k: round(v['ms'], 2) for k, v in prof.items() if isinstance(v, dict)}}), flush=True)
