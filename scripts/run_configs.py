"""Runs the other BASELINE.json configurations once through the public pipeline (synthetic weights) and prints
frames/s + per-family forward breakdown.  Usage: python scripts/run_configs.py [cfg3] [cfg4] [cfg2]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch
from t2v_b200.pipeline import TextToVideoSynthesis
from t2v_b200.synthetic import randomize_
CFG = {'cfg2': dict(frames=24, H=256, W=256, steps=50, sampler='DDIM_Gaussian'),
       'cfg3': dict(frames=24, H=576, W=1024, steps=30, sampler='UniPC'),
       'cfg4': dict(frames=125, H=256, W=256, steps=50, sampler='DDIM_Gaussian'),
       'ddim': dict(frames=24, H=256, W=256, steps=50, sampler='DDIM')}
which = sys.argv[1:] or ['cfg3', 'cfg4']
pipe = TextToVideoSynthesis(None)
randomize_(pipe.sd_model, seed=0); randomize_(pipe.autoencoder, seed=3)
g = torch.Generator().manual_seed(2)
c = torch.randn(1, 77, 1024, generator=g).half().pin_memory(); uc = torch.randn(1, 77, 1024, generator=g).half().pin_memory()
for name in which:
    k = CFG[name]
    torch.cuda.reset_peak_memory_stats()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        frames, lat, info = pipe.infer(c, uc, k['steps'], k['frames'], 123, 17.0, k['W'], k['H'], 0.0, 'GPU (half precision)',
                                       torch.device('cuda'), None, 0, 0.0, None, False, k['sampler'])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    h, w = k['H'] // 8, k['W'] // 8
    fl = k['steps'] * pipe.sd_model.flops(2, k['frames'], h, w, 77) + pipe.autoencoder.flops(k['frames'], h, w)
    import numpy as np
    fr = np.stack(frames)
    free, total = torch.cuda.mem_get_info()
    print(json.dumps({'config': name, **k, 'clip_s': round(dt, 3), 'frames_per_s': round(k['frames'] / dt, 2),
                      'tflop_per_clip': round(fl / 1e12, 1), 'achieved_tflops': round(fl / dt / 1e12, 1),
                      'out_shape': list(fr.shape), 'out_mean': float(fr.mean()), 'finite_latent': bool(torch.isfinite(lat).all()),
                      'gpu_mem_used_gb': round((total - free) / 2 ** 30, 1)}), flush=True)
