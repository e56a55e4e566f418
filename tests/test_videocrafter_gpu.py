"""GPU: the VideoCrafter text2video path end to end on a reduced configuration:
`LatentDiffusion.apply_model` / `DDIMSampler.sample` / `decode_first_stage` / `sample_text2video` / `process_videocrafter`
against the CPU oracle (oracle/vc_oracle.py, bit-exact vs the reference modules; oracle/vae_oracle.py)."""
import numpy as np
import pytest
import torch

from oracle import unet_oracle as UO, vae_oracle as VO, vc_oracle as VC, samplers_oracle as SO

from parity_util import report  # noqa: E402

pytestmark = pytest.mark.gpu

CFG = VC.VCConfig(model_channels=64, context_dim=48, temporal_length=4)


@pytest.fixture(scope='module')
def ldm():
    from t2v_b200.videocrafter import LatentDiffusion
    W = UO.make_weights(VC.vc_param_specs(CFG), seed=4)
    Wv = UO.make_weights(VO.decoder_param_specs(VO.VAEConfig()), seed=3)
    m = LatentDiffusion(unet_config=dict(model_channels=64, context_dim=48, temporal_length=4), image_size=[8, 8],
                        video_length=4).half()
    m.model.diffusion_model.load_state_dict(W, strict=True)
    m.first_stage_model.load_state_dict(Wv, strict=False)
    return m.cuda().eval(), W, Wv


def conds(B=1):
    g = torch.Generator('cpu').manual_seed(2)
    return torch.randn(B, 9, 48, generator=g).half().float(), torch.randn(B, 9, 48, generator=g).half().float()


def test_state_dict_uses_the_checkpoint_key_layout(ldm):
    m = ldm[0]
    keys = set(m.state_dict())
    assert 'model.diffusion_model.input_blocks.1.1.transformer_blocks.0.attn1_tmp.relative_position_k.embeddings_table' in keys
    assert 'first_stage_model.decoder.conv_in.weight' in keys and 'alphas_cumprod' in keys


@pytest.mark.parametrize('S,scale,eta', [(5, 7.5, 0.0), (4, 3.0, 0.5), (2, 1.0, 0.0)])
def test_ddim_trajectory_vs_oracle(ldm, S, scale, eta):
    from t2v_b200.videocrafter import DDIMSampler
    m, W, _ = ldm
    Wh = {k: v.half().float() for k, v in W.items()}
    c, uc = conds()
    x_T = torch.randn((1, 4, 4, 8, 8), generator=torch.Generator('cpu').manual_seed(5))
    smp = DDIMSampler(m)
    smp.noise_gen.manual_seed(11)
    out, inter = smp.sample(S=S, batch_size=1, shape=(4, 4, 8, 8), conditioning=c.cuda(), x_T=x_T.cuda(), eta=eta,
                            unconditional_guidance_scale=scale, unconditional_conditioning=uc.cuda(), verbose=False)
    ref = VC.vc_ddim_sample(lambda a, b, d: VC.vc_unet_forward(Wh, CFG, a, b, d), SO.linear_sd_betas(), x_T, S, c, uc, scale,
                            eta=eta, noise_gen=torch.Generator('cpu').manual_seed(11))
    err = (out.cpu() - ref).abs().max() / ref.abs().max()
    report(f'vc_ddim_tiny:S{S}_g{scale}_eta{eta}', max=float(err))
    assert err < 5e-3, err                      # measured <= 3.2e-3
    assert 'x_inter' in inter


def test_batched_samples_match_single(ldm):
    """B = 2 latents with different prompts == two B = 1 runs (per-sample GroupNorm / attention)."""
    from t2v_b200.videocrafter import DDIMSampler
    m = ldm[0]
    c, uc = conds(2)
    x_T = torch.randn((2, 4, 4, 8, 8), generator=torch.Generator('cpu').manual_seed(6)).cuda()
    smp = DDIMSampler(m)
    both, _ = smp.sample(S=4, batch_size=2, shape=(4, 4, 8, 8), conditioning=c.cuda(), x_T=x_T, eta=0.0,
                         unconditional_guidance_scale=5.0, unconditional_conditioning=uc.cuda())
    one, _ = smp.sample(S=4, batch_size=1, shape=(4, 4, 8, 8), conditioning=c[1:].cuda(), x_T=x_T[1:], eta=0.0,
                        unconditional_guidance_scale=5.0, unconditional_conditioning=uc[1:].cuda())
    assert (both[1:] - one).abs().max() / one.abs().max() < 5e-3


def test_sample_text2video_and_entry_point(ldm):
    from t2v_b200 import videocrafter as vcm
    m, W, Wv = ldm
    c, uc = conds()
    x_T = torch.randn((1, 4, 4, 8, 8), generator=torch.Generator('cpu').manual_seed(9)).cuda()
    vids = vcm.sample_text2video(m, c.cuda(), uc.cuda(), 1, 1, ddim_steps=4, eta=0.0, cfg_scale=4.0, num_frames=4, x_T=x_T)
    assert vids.shape == (1, 3, 4, 64, 64) and vids.min() >= 0 and vids.max() <= 255
    # decode parity: latent -> frames vs the VAE oracle
    smp = vcm.DDIMSampler(m)
    lat, _ = smp.sample(S=4, batch_size=1, shape=(4, 4, 8, 8), conditioning=c.cuda(), x_T=x_T, eta=0.0,
                        unconditional_guidance_scale=4.0, unconditional_conditioning=uc.cuda())
    dec = m.decode_first_stage(lat, return_cpu=True)                               # [1, 3, 4, 64, 64]
    ref = VO.vae_decode({k: v.half().float() for k, v in Wv.items()}, VO.VAEConfig(),
                        lat[0].float().cpu().permute(1, 0, 2, 3) / 0.18215)        # [4, 3, 64, 64]
    assert (dec[0].permute(1, 0, 2, 3) - ref).abs().max() / ref.abs().max() < 2e-2
    out = vcm.process_videocrafter(dict(prompt_embeds=c.cuda(), n_prompt_embeds=uc.cuda(), steps=4, frames=4, seed=3,
                                        cfg_scale=4.0, eta=0.0, batch_count=2, x_T=x_T), model=m)
    assert len(out) == 2 and out[0].shape == (1, 3, 4, 64, 64)
    assert np.array_equal(out[0], vids)                                            # eta = 0, same x_T: deterministic


def test_errors_are_loud(ldm):
    from t2v_b200 import videocrafter as vcm
    m = ldm[0]
    with pytest.raises(RuntimeError):
        m.get_learned_conditioning(['a prompt'])                                   # no text tower attached
    with pytest.raises(RuntimeError):
        m.model.diffusion_model(torch.randn(1, 4, 33, 8, 8).cuda(), torch.tensor([5]).cuda(),
                                context=torch.randn(1, 9, 48).cuda())              # more than 32 frames per clip
