"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the three schedulers behind `Txt2VideoSampler`.

Restates (compactly, as plain functions over a `model(x, t, cond) -> eps` callable) the per-step
latent update of
  * "DDIM_Gaussian"  scripts/samplers/ddim/gaussian_sampler.py:214-296 (+ :73-86, :93-95, :103-108,
                     :125-136, :138-183, :199-211) and `_i` modelscope/t2v_model.py:1232-1237
  * "DDIM"           scripts/samplers/ddim/sampler.py:24-53, :110-166, :169-220 and the ldm helpers
                     (vendored twin: videocrafter/lvdm/models/modules/util.py:36-66)
  * "UniPC"          scripts/samplers/uni_pc/sampler.py:32-89, uni_pc.py:8-153 (NoiseScheduleVP),
                     :238-311 (model_wrapper), :551-677 (bh update), :683-743 (sample), :750-795.
Arithmetic order and dtypes follow the reference so that on CPU fp32 the trajectories agree with
the reference classes bit for bit (pinned in oracle/make_golden.py -> tests/golden/samplers_*.pt).
"""
import math

import numpy as np
import torch


def linear_sd_betas(n=1000, init_beta=0.00085, last_beta=0.0120):
    """beta_schedule('linear_sd') t2v_model.py:1240-1249; called at t2v_pipeline.py:107-111."""
    return torch.linspace(init_beta ** 0.5, last_beta ** 0.5, n, dtype=torch.float64) ** 2


def _i(tab, t, x):
    """t2v_model.py:1232-1237: gather tab[t], reshape to (B,1,1,...), cast to x's dtype."""
    return tab[t.cpu()].view((x.size(0),) + (1,) * (x.ndim - 1)).to(x)


# ------------------------------------------------------------------------------------------
# DDIM_Gaussian (UI default)
# ------------------------------------------------------------------------------------------
def gaussian_timesteps(num_timesteps, S):
    """get_time_steps :73-86 -> e.g. S=50: 981, 961, ..., 1."""
    stride = num_timesteps // S
    steps = (1 + torch.arange(0, num_timesteps, stride)).clamp(0, num_timesteps - 1)
    return steps.flip(0), stride


def gaussian_cfg(y_out, u_out, scale):
    """do_classifier_guidance :125-136 with var_type 'learned_range' (ctor default :15, never
    overridden, samplers_common.py:87): get_dim :93-95 returns C//2, so only the first half of the
    channels is guided; the rest is the *conditional* output unguided."""
    dim = y_out.size(1) // 2
    a = u_out[:, :dim]
    b = scale * (y_out[:, :dim] - u_out[:, :dim])
    return torch.cat([a + b, y_out[:, dim:]], dim=1)


@torch.no_grad()
def ddim_gaussian_sample(model, betas, x_T, S, cond, uncond, guide_scale, eta=0.0, callback=None,
                         trace=None):
    """GaussianDiffusion.sample :214-296."""
    acp = torch.cumprod(1 - betas, dim=0)                       # :33 (fp64 when betas are)
    sqrt_recip = torch.sqrt(1.0 / acp)                          # :40
    sqrt_recipm1 = torch.sqrt(1.0 / acp - 1)                    # :41
    ts, stride = gaussian_timesteps(len(betas), S)
    xt = x_T.clone()
    for step in range(S):
        t = torch.full((xt.shape[0],), int(ts[step]), dtype=torch.long)
        if guide_scale is None or guide_scale == 1:             # :121-123, :152-153
            out = model(xt, t, cond)
        else:
            y_out = model(xt, t, cond)                          # two sequential forwards :161-162
            u_out = model(xt, t, uncond)
            out = gaussian_cfg(y_out, u_out, guide_scale)
        x0 = _i(sqrt_recip, t, xt) * xt - _i(sqrt_recipm1, t, xt) * out        # mean_x0 :103-105
        alphas = _i(acp, t, xt)
        alphas_prev = _i(acp, (t - stride).clamp(0), xt)                          # :269-270
        eps = (_i(sqrt_recip, t, xt) * xt - x0) / _i(sqrt_recipm1, t, xt)         # get_eps :201-202
        a = (1 - alphas_prev) / (1 - alphas)
        b = (1 - alphas / alphas_prev)
        sigmas = eta * torch.sqrt(a * b)
        noise = torch.randn_like(xt)                                                # drawn even at eta 0 (:279)
        direction = torch.sqrt(1 - alphas_prev - sigmas ** 2) * eps
        mask = t.ne(0).float().view(-1, *((1,) * (xt.ndim - 1))).to(xt.device)
        xt = torch.sqrt(alphas_prev) * x0 + direction + mask * sigmas * noise
        torch.randn_like(xt)     # the inpaint-mask hook draws a second randn every step (:285-291); result unused
        if trace is not None:
            trace.append(xt.clone())
        if callback is not None:
            callback(step)
    return xt


# ------------------------------------------------------------------------------------------
# DDIM (ldm-style)
# ------------------------------------------------------------------------------------------
def ddim_schedule(alphas_cumprod, S, eta=0.0):
    """make_ddim_timesteps ('uniform') + make_ddim_sampling_parameters (util.py:36-66)."""
    n = alphas_cumprod.shape[0]
    c = n // S
    ts = np.asarray(list(range(0, n, c))) + 1
    acp = alphas_cumprod.double().cpu().numpy()
    alphas = acp[ts]
    alphas_prev = np.asarray([acp[0]] + acp[ts[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return ts, alphas, alphas_prev, sigmas


@torch.no_grad()
def ddim_sample(model, betas, x_T, S, cond, uncond, guide_scale, eta=0.0, callback=None, trace=None):
    """DDIMSampler.sample/ddim_sampling/p_sample_ddim (sampler.py:57-220); all 4 channels guided (:181);
    coefficient tensors are torch.full((b,1,1,1), ...) in the default dtype fp32 (:194-197)."""
    acp = torch.cumprod(1 - betas, dim=0)
    ts, alphas, alphas_prev, sigmas = ddim_schedule(acp, S, eta)
    sqrt_1m = np.sqrt(1.0 - alphas)
    img = x_T
    b = img.shape[0]
    total = ts.shape[0]
    for i, step in enumerate(np.flip(ts)):
        index = total - i - 1
        t = torch.full((b,), int(step), dtype=torch.long)
        if uncond is None or guide_scale == 1.0:
            e_t = model(img, t, cond)
        else:
            e_c = model(img, t, cond)
            e_u = model(img, t, uncond)
            e_t = e_u + guide_scale * (e_c - e_u)
        dev = img.device
        a_t = torch.full((b, 1, 1, 1), float(alphas[index]), device=dev)
        a_prev = torch.full((b, 1, 1, 1), float(alphas_prev[index]), device=dev)
        sigma_t = torch.full((b, 1, 1, 1), float(sigmas[index]), device=dev)
        s1m = torch.full((b, 1, 1, 1), float(sqrt_1m[index]), device=dev)
        pred_x0 = (img - s1m * e_t) / a_t.sqrt()
        dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
        noise = sigma_t * torch.randn(img.shape, device=dev) * 1.0
        img = a_prev.sqrt() * pred_x0 + dir_xt + noise
        if trace is not None:
            trace.append(img.clone())
        if callback is not None:
            callback(i)
    return img


# ------------------------------------------------------------------------------------------
# UniPC (bh1, multistep order 3, predict_x0, time_uniform, lower_order_final)
# ------------------------------------------------------------------------------------------
def _interp(x, xp, yp):
    """interpolate_fn uni_pc.py:750-795 specialised to C = 1 (piecewise-linear, linear extrapolation
    from the two outermost key points); same arithmetic order as the reference's `cand`."""
    N, K = x.shape[0], xp.shape[1]
    all_x = torch.cat([x.unsqueeze(2), xp.unsqueeze(0).repeat((N, 1, 1))], dim=2)
    sorted_all_x, x_indices = torch.sort(all_x, dim=2)
    x_idx = torch.argmin(x_indices, dim=2)
    cand_start = x_idx - 1
    start_idx = torch.where(x_idx == 0, torch.tensor(1), torch.where(x_idx == K, torch.tensor(K - 2), cand_start))
    end_idx = torch.where(start_idx == cand_start, start_idx + 2, start_idx + 1)
    start_x = torch.gather(sorted_all_x, 2, start_idx.unsqueeze(2)).squeeze(2)
    end_x = torch.gather(sorted_all_x, 2, end_idx.unsqueeze(2)).squeeze(2)
    start_idx2 = torch.where(x_idx == 0, torch.tensor(0), torch.where(x_idx == K, torch.tensor(K - 2), cand_start))
    ype = yp.unsqueeze(0).expand(N, -1, -1)
    start_y = torch.gather(ype, 2, start_idx2.unsqueeze(2)).squeeze(2)
    end_y = torch.gather(ype, 2, (start_idx2 + 1).unsqueeze(2)).squeeze(2)
    return start_y + (x - start_x) * (end_y - start_y) / (end_x - start_x)


class VPSchedule:
    """NoiseScheduleVP('discrete', alphas_cumprod=...) uni_pc.py:77-153.  The sampler hands it an
    fp32 copy of alphas_cumprod (uni_pc/sampler.py:11-12)."""

    def __init__(self, alphas_cumprod_f32):
        la = 0.5 * torch.log(alphas_cumprod_f32)
        self.total_N = len(la)
        self.T = 1.0
        self.t_array = torch.linspace(0.0, 1.0, self.total_N + 1)[1:].reshape((1, -1))
        self.log_alpha_array = la.reshape((1, -1))

    def log_alpha(self, t):
        return _interp(t.reshape((-1, 1)), self.t_array, self.log_alpha_array).reshape((-1))

    def alpha(self, t):
        return torch.exp(self.log_alpha(t))

    def std(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.log_alpha(t)))

    def lam(self, t):
        lm = self.log_alpha(t)
        return lm - 0.5 * torch.log(1.0 - torch.exp(2.0 * lm))


def _ex(v, dims):
    return v[(...,) + (None,) * (dims - 1)]


@torch.no_grad()
def unipc_sample(model, betas, x_T, S, cond, uncond, guide_scale, t_start=None, callback=None,
                 trace=None, order=3):
    """UniPCSampler.sample (sampler.py:32-89) -> UniPC.sample(method='multistep', order=3,
    skip_type='time_uniform', lower_order_final=True, initial_corrector=True) uni_pc.py:683-743."""
    ns = VPSchedule(torch.cumprod(1 - betas, dim=0).clone().detach().to(torch.float32))
    dims = x_T.dim()

    def data_pred(x, t):
        """model_wrapper.model_fn :291-307 + data_prediction_fn :378-391 (no thresholding)."""
        tc = t.expand((x.shape[0])) if t.reshape((-1,)).shape[0] == 1 else t
        t_in = (tc - 1.0 / ns.total_N) * 1000.0                         # float model time (:248)
        if guide_scale == 1.0 or uncond is None:
            noise = model(x, t_in, cond)
        else:
            n_c = model(x, t_in, cond)
            n_u = model(x, t_in, uncond)
            noise = n_u + guide_scale * (n_c - n_u)
        a, s = ns.alpha(t), ns.std(t)
        return (x - _ex(s, dims) * noise) / _ex(a, dims)

    def bh_update(x, m_list, t_list, t, k, use_corrector):
        """multistep_uni_pc_bh_update :551-677 with variant 'bh1', predict_x0."""
        t_prev_0 = t_list[-1]
        lam_prev_0, lam_t = ns.lam(t_prev_0), ns.lam(t)
        m0 = m_list[-1]
        sigma_prev_0, sigma_t = ns.std(t_prev_0), ns.std(t)
        alpha_t = torch.exp(ns.log_alpha(t))
        h = lam_t - lam_prev_0
        rks, D1s = [], []
        for i in range(1, k):
            rk = ((ns.lam(t_list[-(i + 1)]) - lam_prev_0) / h)[0]
            rks.append(rk)
            D1s.append((m_list[-(i + 1)] - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks)
        hh = -h[0]
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = hh
        R, bvec, fact = [], [], 1
        for i in range(1, k + 1):
            R.append(torch.pow(rks, i - 1))
            bvec.append(h_phi_k * fact / B_h)
            fact *= (i + 1)
            h_phi_k = h_phi_k / hh - 1 / fact
        R = torch.stack(R)
        bvec = torch.tensor(bvec)
        rhos_p = None
        if len(D1s) > 0:
            rhos_p = torch.tensor([0.5]) if k == 2 else torch.linalg.solve(R[:-1, :-1], bvec[:-1])
        if use_corrector:
            rhos_c = torch.tensor([0.5]) if k == 1 else torch.linalg.solve(R, bvec)
        x_t_ = _ex(sigma_t / sigma_prev_0, dims) * x - _ex(alpha_t * h_phi_1, dims) * m0

        def comb(rhos):
            # :618-620, :637-639: stack -> 'b k c f h w -> (b f) k c h w' -> einsum('k,bkchw->bchw') ->
            # 'f c h w -> b c f h w' (b == 1 in this pipeline; same op sequence so CPU results are bit-equal)
            D = torch.stack(D1s, dim=1)
            if D.dim() > 5:
                b_, k_, c_, f_, h_, w_ = D.shape
                D = D.permute(0, 3, 1, 2, 4, 5).reshape(b_ * f_, k_, c_, h_, w_)
                r = torch.einsum('k,bkchw->bchw', rhos, D)
                return r.permute(1, 0, 2, 3).unsqueeze(0).repeat(x.shape[0], 1, 1, 1, 1)
            return torch.einsum('k,bkchw->bchw', rhos, D)

        pred_res = comb(rhos_p) if len(D1s) > 0 else 0
        x_t = x_t_ - _ex(alpha_t * B_h, dims) * pred_res
        m_t = None
        if use_corrector:
            m_t = data_pred(x_t, t)
            corr_res = comb(rhos_c[:-1]) if len(D1s) > 0 else 0
            x_t = x_t_ - _ex(alpha_t * B_h, dims) * (corr_res + rhos_c[-1] * (m_t - m0))
        return x_t, m_t

    t_0 = 1.0 / ns.total_N
    t_T = ns.T if t_start is None else t_start
    timesteps = torch.linspace(t_T, t_0, S + 1)
    assert S >= order
    x = x_T
    vec_t = timesteps[0].expand((x.shape[0]))
    m_list, t_list = [data_pred(x, vec_t)], [vec_t]
    for init_order in range(1, order):
        vec_t = timesteps[init_order].expand(x.shape[0])
        x, m_x = bh_update(x, m_list, t_list, vec_t, init_order, True)
        if m_x is None:
            m_x = data_pred(x, vec_t)
        m_list.append(m_x)
        t_list.append(vec_t)
        if trace is not None:
            trace.append(x.clone())
        if callback is not None:
            callback()
    for step in range(order, S + 1):
        vec_t = timesteps[step].expand(x.shape[0])
        k = min(order, S + 1 - step)
        x, m_x = bh_update(x, m_list, t_list, vec_t, k, step != S)
        for i in range(order - 1):
            t_list[i] = t_list[i + 1]
            m_list[i] = m_list[i + 1]
        t_list[-1] = vec_t
        if step < S:
            if m_x is None:
                m_x = data_pred(x, vec_t)
            m_list[-1] = m_x
        if trace is not None:
            trace.append(x.clone())
        if callback is not None:
            callback()
    return x
