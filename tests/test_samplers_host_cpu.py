"""CPU: the HOST side of the product schedulers (coefficient algebra, call order, history handling of t2v_b200/samplers.py)
against the reference-pinned oracle schedulers, driven by the same fp16 denoiser (the tiny oracle UNet rounded to fp16 -- the
dtype, magnitudes and CFG scale of the real path).  The three device kernels the schedulers launch (t2v_ddim_step, t2v_cfg_x0,
t2v_lincomb; their own parity is tests/test_model_gpu.py) are replaced by torch statements of the same formulas, so this runs
without a GPU and isolates what the GPU single-step gates cannot: on the GPU a 3.5e-8 change of the latent re-rolls the fp16
rounding noise of the whole denoiser (eps moves by 2.4e-3, scripts/diag_unipc.py), which UniPC's update amplifies."""
import ctypes as C

import pytest
import torch

from oracle import unet_oracle as UO, samplers_oracle as SO
from t2v_b200 import samplers as S, _lib


class _TorchKernels(object):
    """csrc/elementwise.cu ddim_step_kernel / cfg_x0_kernel / lincomb_kernel restated with torch ops on registered tensors."""

    def __init__(self, reg):
        self.reg = reg

    @staticmethod
    def _cfg(ec, eu, g, fp16):
        if fp16:
            return (eu + g * (ec - eu)).float()                 # fp16 tensors: op-by-op fp16 rounding, as the kernel
        return eu.float() + g * (ec.float() - eu.float())

    def t2v_cfg_x0(self, x, ec, eu, is32, x0, n, g, alpha, sigma, fp16, stream):
        R = self.reg
        e = self._cfg(R[ec.value], R[eu.value], g, fp16) if eu.value else R[ec.value].float()
        R[x0.value].copy_((R[x.value] - sigma * e) / alpha)
        return 0

    def t2v_lincomb(self, out, srcs, coefs, k, n, stream):
        acc = torch.zeros_like(self.reg[out.value])
        for i in range(k):
            acc = acc + coefs[i] * self.reg[srcs[i]]
        self.reg[out.value].copy_(acc)
        return 0

    def t2v_ddim_step(self, x, ec, eu, is32, xo, n, chan_stride, Cc, gch, g, mode, a0, a1, a2, a3, a4, noise, fp16, stream):
        R = self.reg
        X, EC = R[x.value], R[ec.value]
        e = EC.float().clone()
        if eu.value:
            e[:, :gch] = self._cfg(EC[:, :gch], R[eu.value][:, :gch], g, fp16)
        nz = a4 * R[noise.value] if (noise is not None and getattr(noise, 'value', None) and a4 != 0.0) else 0.0
        if mode == 0:
            ax = a0 * X
            x0 = ax - a1 * e
            eps = (ax - x0) / a1
            R[xo.value].copy_(a2 * x0 + a3 * eps + nz)
        else:
            x0 = (X - a0 * e) / a1
            R[xo.value].copy_(a2 * x0 + a3 * e + nz)
        return 0


@pytest.fixture()
def cpu_samplers(monkeypatch):
    reg = {}

    def ptr(t):
        if t is None:
            return C.c_void_p(0)
        reg[t.data_ptr()] = t
        return C.c_void_p(t.data_ptr())
    fake = _TorchKernels(reg)
    monkeypatch.setattr(_lib, 'lib', lambda: fake)
    monkeypatch.setattr(_lib, 'ptr', ptr)
    monkeypatch.setattr(_lib, 'stream_ptr', lambda: None)
    monkeypatch.setattr(S, '_need_cuda', lambda x: None)
    orig = torch.empty_like

    def empty_like(t, *a, **k):
        o = orig(t, *a, **k)
        reg[o.data_ptr()] = o
        return o
    monkeypatch.setattr(torch, 'empty_like', empty_like)
    return reg


@pytest.mark.parametrize('name,oracle_fn,S_,stops', [('UniPC', SO.unipc_sample, 30, (3, 5, 7)), ('DDIM_Gaussian', SO.ddim_gaussian_sample, 50, (3, 5)),
                                                    ('DDIM', SO.ddim_sample, 50, (3, 5))])
def test_scheduler_host_algebra_matches_oracle_with_fp16_eps(cpu_samplers, name, oracle_fn, S_, stops):
    reg = cpu_samplers
    cfg = UO.UNetConfig(dim=64)
    W = UO.make_weights(UO.param_specs(cfg), seed=1)
    betas = SO.linear_sd_betas()
    g = torch.Generator().manual_seed(123)
    x = torch.randn(1, 4, 3, 8, 8, generator=g)
    c = torch.randn(1, 77, 1024, generator=g)
    uc = torch.randn(1, 77, 1024, generator=g)

    class Stop(Exception):
        pass

    def latent_at_call(run, stop):
        calls = []

        class M(object):
            device = torch.device('cpu')
            alphas_cumprod = torch.cumprod(1 - betas, 0)
            num_timesteps = 1000

            def __call__(self, a, b, d):
                calls.append(a.clone())
                reg[a.data_ptr()] = a
                if len(calls) == stop:
                    raise Stop()
                return UO.unet_forward(W, cfg, a, b, d).half()           # fp16 eps, as the GPU denoiser returns
        try:
            run(M())
        except Stop:
            pass
        return calls[-1]
    entry = [s for s in S.available_samplers if s.name == name][0]
    for stop in stops:
        torch.manual_seed(0)
        ours = latent_at_call(lambda m: entry.init_sampler(m, betas=betas, device=torch.device('cpu')).sample(
            S=S_, conditioning=c, unconditional_conditioning=uc, unconditional_guidance_scale=17.0, x_T=x, shape=tuple(x.shape), eta=0.0,
            batch_size=1), stop)
        torch.manual_seed(0)
        orc = latent_at_call(lambda m: oracle_fn(m, betas, x, S_, c, uc, 17.0), stop)
        rel = ((ours - orc).pow(2).mean().sqrt() / orc.pow(2).mean().sqrt()).item()
        assert rel < 1e-4, (name, stop, rel)          # fp32 re-association only (measured 0 .. 2e-5); an algebra slip would be >= 1e-2
