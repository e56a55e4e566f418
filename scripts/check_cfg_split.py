"""torchrun --nproc-per-node 2: CFG-pair split (one guidance branch per GPU) vs the batched B=2 forward on one GPU:
latent agreement and per-clip latency.  Prints one JSON line on rank 0."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'sd-webui-text2video_b200'))
import torch, torch.distributed as dist
rank = int(os.environ['RANK']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
from t2v_b200.pipeline import TextToVideoSynthesis
from t2v_b200.synthetic import randomize_
from t2v_b200 import samplers
pipe = TextToVideoSynthesis(None, device=dev)
randomize_(pipe.sd_model, seed=0); randomize_(pipe.autoencoder, seed=3)
g = torch.Generator().manual_seed(2)
c = torch.randn(1, 77, 1024, generator=g).half().to(dev); uc = torch.randn(1, 77, 1024, generator=g).half().to(dev)
entry = [s for s in samplers.available_samplers if s.name == 'DDIM_Gaussian'][0]
F, h, w, S = 24, 32, 32, 50
def clip(seed):
    x_T = torch.randn((1, 4, F, h, w), device=dev, generator=torch.Generator(device=dev).manual_seed(seed))
    smp = entry.init_sampler(pipe.sd_model, betas=pipe.diffusion.betas, device=dev)
    return smp.sample(S=S, conditioning=c, unconditional_conditioning=uc, unconditional_guidance_scale=17.0, x_T=x_T,
                      shape=tuple(x_T.shape), eta=0.0, batch_size=1)
def timed(n):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): out = clip(123)
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / n, out
os.environ['T2V_CFG_SPLIT'] = '1'
clip(1); ts, lat_split = timed(2)
os.environ.pop('T2V_CFG_SPLIT')
clip(1); tb, lat_b2 = timed(2)
both = [torch.empty_like(lat_split) for _ in range(2)]
dist.all_gather(both, lat_split)
if rank == 0:
    rel = ((lat_split - lat_b2).abs().max() / lat_b2.abs().max()).item()
    print(json.dumps({'check': 'cfg_pair_split', 'clip_s_split_2gpu': round(ts, 4), 'clip_s_batched_1gpu': round(tb, 4),
                      'latency_speedup': round(tb / ts, 3), 'latent_max_rel_diff_vs_batched': rel,
                      'replicas_identical': bool(torch.equal(both[0], both[1]))}), flush=True)
dist.destroy_process_group()
