"""GPU: the public entry points end to end (tiny UNet config, full-size VAE, synthetic weights): `process_modelscope`,
`TextToVideoSynthesis.infer` for the three UI samplers, determinism in the seed, the per-step callback contract, and
agreement of a short DDIM_Gaussian trajectory + decode with the CPU oracle driven by the same weights."""
import numpy as np
import pytest
import torch

from oracle import unet_oracle as UO, vae_oracle as VO, samplers_oracle as SO

from parity_util import report  # noqa: E402

pytestmark = pytest.mark.gpu

TINY = {'unet_dim': 64}


@pytest.fixture(scope='module')
def pipe():
    from t2v_b200.pipeline import TextToVideoSynthesis
    cfg = UO.UNetConfig(dim=64)
    W = UO.make_weights(UO.param_specs(cfg), seed=1)
    Wv = UO.make_weights(VO.decoder_param_specs(VO.VAEConfig()), seed=3)
    p = TextToVideoSynthesis(None, model_cfg=TINY, unet_state=W, vae_state=Wv)
    return p, cfg, W, Wv


def conds():
    g = torch.Generator().manual_seed(2)
    return torch.randn(1, 77, 1024, generator=g).half(), torch.randn(1, 77, 1024, generator=g).half()


@pytest.mark.parametrize('sampler,steps', [('DDIM_Gaussian', 6), ('DDIM', 5), ('UniPC', 5)])
def test_infer_all_samplers(pipe, sampler, steps):
    p = pipe[0]
    c, uc = conds()
    frames, latent, info = p.infer(c, uc, steps, 3, 123, 7.5, 64, 64, 0.0, 'GPU (half precision)', torch.device('cuda'),
                                   None, 0, 0.0, None, False, sampler)
    assert len(frames) == 3 and frames[0].shape == (64, 64, 3) and frames[0].dtype == np.uint8
    assert latent.shape == (1, 4, 3, 8, 8) and torch.isfinite(latent).all()
    frames2, latent2, _ = p.infer(c, uc, steps, 3, 123, 7.5, 64, 64, 0.0, 'GPU (half precision)', torch.device('cuda'),
                                  None, 0, 0.0, None, False, sampler)
    assert torch.equal(latent, latent2) and all(np.array_equal(a, b) for a, b in zip(frames, frames2))
    _, latent3, _ = p.infer(c, uc, steps, 3, 124, 7.5, 64, 64, 0.0, 'GPU (half precision)', torch.device('cuda'), None, 0,
                            0.0, None, False, sampler)
    assert not torch.equal(latent, latent3)
    assert 'steps' in info and sampler in info


def test_short_trajectory_and_decode_vs_oracle(pipe):
    p, cfg, W, Wv = pipe
    c, uc = conds()
    S, F = 4, 2
    frames, latent, _ = p.infer(c, uc, S, F, 77, 5.0, 64, 64, 0.0, 'GPU (half precision)', torch.device('cuda'), None, 0,
                                0.0, None, False, 'DDIM_Gaussian')
    Wh = {k: v.half().float() for k, v in W.items()}
    x_T = torch.randn((1, 4, F, 8, 8), generator=torch.Generator('cpu').manual_seed(77))
    ref = SO.ddim_gaussian_sample(lambda a, b, d: UO.unet_forward(Wh, cfg, a, b, d), SO.linear_sd_betas(), x_T, S,
                                  c.float(), uc.float(), 5.0)
    err = (latent.cpu() - ref).abs().max() / ref.abs().max()
    report('pipeline:ddim_gaussian_4steps_tiny', max=float(err))
    assert err < 6e-3, err                      # 4 steps x 2 fp16 forwards each, vs fp32 oracle (measured 3.8e-3)
    dec = VO.vae_decode({k: v.half().float() for k, v in Wv.items()}, VO.VAEConfig(),
                        (ref[0].permute(1, 0, 2, 3) / 0.18215))
    ref_u8 = VO.tensor2vid_u8(dec.permute(1, 0, 2, 3).unsqueeze(0))          # [F, H, W, 3] RGB
    got = np.stack([f[:, :, ::-1] for f in frames])                             # BGR -> RGB
    assert np.abs(got.astype(int) - ref_u8.astype(int)).mean() < 3.0


@pytest.mark.parametrize('sampler', ['DDIM_Gaussian', 'DDIM'])
def test_eta_positive_matches_oracle_on_the_same_noise(pipe, sampler, monkeypatch):
    """eta > 0: the fused step kernel's sigma / direction terms against the oracle (gaussian_sampler.py:269-283, ddim/sampler.py:
    199-218).  The GPU path draws its per-step noise from the CUDA generator and the oracle from the CPU one, so both sides are fed
    the SAME noise tape here: the product through its single noise hook (`distributed.step_noise`), the oracle through torch.randn /
    randn_like.  DDIM_Gaussian's reference draws a second, unused randn per step (the inpaint hook, :285-291): the tape skips it."""
    p, cfg, W, Wv = pipe
    c, uc = conds()
    S, F, eta = 4, 2, 0.7
    tape = [torch.randn((1, 4, F, 8, 8), generator=torch.Generator('cpu').manual_seed(500 + i)) for i in range(S)]
    from t2v_b200 import distributed as D
    it = iter(tape)
    monkeypatch.setattr(D, 'step_noise', lambda like: next(it).to(device=like.device, dtype=like.dtype))
    _, latent, _ = p.infer(c, uc, S, F, 77, 5.0, 64, 64, eta, 'GPU (half precision)', torch.device('cuda'), None, 0,
                           0.0, None, False, sampler)
    monkeypatch.undo()
    Wh = {k: v.half().float() for k, v in W.items()}
    x_T = torch.randn((1, 4, F, 8, 8), generator=torch.Generator('cpu').manual_seed(77))
    calls = {'n': 0}
    real_randn, real_randn_like = torch.randn, torch.randn_like

    def tape_randn_like(x, *a, **k):                      # DDIM_Gaussian: draw 0, 2, 4, ... are the step noises
        i = calls['n']
        calls['n'] += 1
        return tape[i // 2].to(x) if i % 2 == 0 else real_randn_like(x, *a, **k)

    def tape_randn(*a, **k):                              # DDIM: one draw per step
        i = calls['n']
        calls['n'] += 1
        return tape[i]
    model = lambda a, b, d: UO.unet_forward(Wh, cfg, a, b, d)     # noqa: E731
    if sampler == 'DDIM_Gaussian':
        monkeypatch.setattr(torch, 'randn_like', tape_randn_like)
        ref = SO.ddim_gaussian_sample(model, SO.linear_sd_betas(), x_T, S, c.float(), uc.float(), 5.0, eta=eta)
    else:
        monkeypatch.setattr(torch, 'randn', tape_randn)
        ref = SO.ddim_sample(model, SO.linear_sd_betas(), x_T, S, c.float(), uc.float(), 5.0, eta=eta)
    monkeypatch.undo()
    assert calls['n'] == (2 * S if sampler == 'DDIM_Gaussian' else S)
    err = float((latent.cpu() - ref).abs().max() / ref.abs().max())
    report(f'pipeline:{sampler}_eta{eta}_4steps_tiny', max=err)
    assert err < 6e-3, err                      # same budget as the eta = 0 trajectory (4 steps x 2 fp16 forwards vs fp32 oracle)


def test_callback_and_interrupt(pipe):
    from t2v_b200 import samplers
    p = pipe[0]
    c, uc = conds()
    samplers.state.interrupted = False
    calls = []
    smp = [s for s in samplers.available_samplers if s.name == 'DDIM_Gaussian'][0].init_sampler(
        p.sd_model, betas=p.diffusion.betas, device=torch.device('cuda'))
    x = torch.randn(1, 4, 2, 8, 8, device='cuda')
    smp.sample(S=5, conditioning=c.cuda(), unconditional_conditioning=uc.cuda(), unconditional_guidance_scale=3.0, x_T=x,
               callback=lambda step: calls.append(step))
    assert calls == [0, 1, 2, 3, 4]

    class Boom(Exception):
        pass

    def bad(step):
        if step == 2:
            raise Boom()
    with pytest.raises(Boom):
        smp.sample(S=5, conditioning=c.cuda(), unconditional_conditioning=uc.cuda(), unconditional_guidance_scale=3.0, x_T=x,
                   callback=bad)
    samplers.state.interrupted = True
    try:
        with pytest.raises(samplers.InterruptedException):
            p.infer(c, uc, 4, 2, 1, 3.0, 64, 64, 0.0, 'GPU (half precision)', torch.device('cuda'), None, 0, 0.0, None, False,
                    'DDIM')
    finally:
        samplers.state.interrupted = False


def test_process_modelscope_entry_point(pipe):
    from t2v_b200 import process_modelscope as pm
    pm.pipe = pipe[0]
    c, uc = conds()
    outs = pm.process_modelscope({'prompt_embeds': c, 'n_prompt_embeds': uc, 'steps': 3, 'frames': 2, 'seed': 5, 'cfg_scale': 4.0,
                                  'width': 64, 'height': 64, 'batch_count': 2, 'sampler': 'DDIM_Gaussian', 'return_frames': True})
    assert len(outs) == 2 and len(outs[0]) == 2 and outs[0][0].shape == (64, 64, 3)
    assert not np.array_equal(outs[0][0], outs[1][0])           # batch i uses seed + i
    # default return type = the reference's: a list of data-URL videos (process_modelscope.py:34, :256-262)
    urls = pm.process_modelscope({'prompt_embeds': c, 'n_prompt_embeds': uc, 'steps': 3, 'frames': 2, 'seed': 5, 'cfg_scale': 4.0,
                                  'width': 64, 'height': 64, 'batch_count': 1, 'sampler': 'DDIM_Gaussian'})
    assert len(urls) == 1 and isinstance(urls[0], str) and urls[0].startswith('data:video/') and ';base64,' in urls[0]
    if urls[0].startswith('data:video/avi'):                   # no ffmpeg on the box: uncompressed AVI, frames recoverable bit for bit
        import base64
        raw = base64.b64decode(urls[0].split(',', 1)[1])
        assert raw[:4] == b'RIFF' and raw[8:12] == b'AVI '
        first = raw.index(b'00db') + 8
        got = np.frombuffer(raw[first:first + 64 * 64 * 3], dtype=np.uint8).reshape(64, 64, 3)[::-1]
        assert np.array_equal(got, outs[0][0])
    with pytest.raises(RuntimeError):
        pipe[0].infer('a cat', '', 3, 2, 1, 3.0, 64, 64)        # string prompts need a clip_encoder
    with pytest.raises(RuntimeError):
        pipe[0].infer(c, uc, 3, 2, 1, 3.0, 64, 64, 0.0, 'CPU (Low VRAM)')


@pytest.mark.parametrize('sampler', ['DDIM_Gaussian', 'DDIM', 'UniPC'])
def test_vid2vid_through_the_entry_point(pipe, sampler):
    """vid2vid: input frames -> compute_latents (AutoencoderKL.encode on the library) -> encode_latent -> denoise -> decode
    (process_modelscope.py:118-221 minus the file reading)."""
    from t2v_b200 import process_modelscope as pm
    p = pipe[0]
    Wenc = UO.make_weights(VO.encoder_param_specs(VO.VAEConfig()), seed=5)
    p.autoencoder.load_state_dict(Wenc, strict=False)
    p.autoencoder.cuda()
    pm.pipe = p
    c, uc = conds()
    vid = torch.rand((1, 3, 3, 64, 64), generator=torch.Generator().manual_seed(4)) * 2 - 1
    lat = p.compute_latents(vid, 'GPU (half precision)', torch.device('cuda'))
    assert lat.shape == (1, 4, 3, 8, 8) and lat.dtype == torch.float32 and not lat.is_cuda and torch.isfinite(lat).all()
    # 8 steps at strength 0.5 -> skip 4, denoise 4 (the reference's 1000 // steps grid needs a divisor of 1000: with 3 steps its
    # make_ddim_timesteps yields index 1000 and raises, ldm util.py:36-50 -- reproduced, not papered over)
    base = dict(prompt_embeds=c, n_prompt_embeds=uc, steps=8, frames=3, seed=11, cfg_scale=5.0, width=64, height=64,
                sampler=sampler, return_frames=True)
    out = pm.process_modelscope(dict(base, do_vid2vid=True, vid2vid_frames_tensor=vid, strength=0.5))
    txt = pm.process_modelscope(dict(base))
    assert len(out) == 1 and len(out[0]) == 3 and out[0][0].shape == (64, 64, 3)
    assert any((a != b).any() for a, b in zip(out[0], txt[0]))          # the input video steers the result
    with pytest.raises(NotImplementedError):
        pm.process_modelscope(dict(base, do_vid2vid=True))                # no frames given: file reading is webui plumbing
    pm.pipe = None


def test_img2vid_inpainting_latents_and_entry_point(pipe):
    """img2vid (process_modelscope.py:170-219): per-frame weights from the key-frame schedule, the fp64 blend
    image_latents * (1 - mask) + noise * mask on the device vs numpy, and the run through the entry point."""
    from t2v_b200 import process_modelscope as pm
    p = pipe[0]
    Wenc = UO.make_weights(VO.encoder_param_specs(VO.VAEConfig()), seed=5)
    p.autoencoder.load_state_dict(Wenc, strict=False)
    p.autoencoder.cuda()
    pm.pipe = p
    c, uc = conds()
    img = torch.rand((3, 64, 64), generator=torch.Generator().manual_seed(8)) * 2 - 1
    F_, n_i = 5, 3
    noise = np.random.RandomState(3).normal(size=(1, 4, F_, 8, 8))
    lat, mask = pm.inpainting_latents(p, img, F_, 64, 64, n_i, '0:(t/max_i_f), "max_i_f":(1)', 7, 'GPU (half precision)', noise)
    il = p.compute_latents(img.view(1, 3, 1, 64, 64), 'GPU (half precision)', torch.device('cuda')).numpy()      # [1,4,1,8,8]
    w = np.array([0.0, 0.5, 1.0, 1.0, 1.0]).reshape(1, 1, F_, 1, 1)
    ref = il * (1 - w) + noise * w                                           # numpy float64, as the reference
    assert lat.dtype == torch.float64 and np.array_equal(lat.cpu().numpy(), ref)
    assert np.array_equal(mask.cpu().numpy(), np.broadcast_to(w, ref.shape))
    base = dict(prompt_embeds=c, n_prompt_embeds=uc, steps=4, frames=F_, seed=11, cfg_scale=5.0, width=64, height=64,
                sampler='DDIM_Gaussian', return_frames=True)
    out = pm.process_modelscope(dict(base, inpainting_frames=n_i, inpainting_image_tensor=img, inpainting_noise=noise))
    txt = pm.process_modelscope(dict(base))
    assert len(out[0]) == F_ and any((a != b).any() for a, b in zip(out[0], txt[0]))
    pm.pipe = None
