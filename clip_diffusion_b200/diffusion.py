"""Respaced DDIM Gaussian diffusion -- the sampler side of a guidance step (host logic + stock torch ops).

The reference takes this from the un-vendored crowsonkb/guided-diffusion (clip_diffusion/models.py:90-117,129;
sampler call at sample.py:241-275).  Restated from SURVEY.md App. A.3: linear beta schedule over 1000 steps,
``ddim{N}`` respacing, learned-range sigma (6-channel model output), ``rescale_timesteps``; ``p_mean_variance``
as called from cond_fn (sample.py:149-151), ``condition_score`` and the DDIM update.  ``cond_fn`` receives the
ORIGINAL (0..999) timestep, which is why sample.py:157-159 indexes its 1000-long schedules with it.
"""
import numpy as np
import torch


_DEVICE_TABLES = {}


def _extract(arr, t, like):
    """arr[t] broadcastable against ``like``; the float32 device copy of each schedule table is cached."""
    key = (id(arr), t.device)
    tab = _DEVICE_TABLES.get(key)
    if tab is None or tab[0] is not arr:
        tab = (arr, torch.from_numpy(arr).to(device=t.device, dtype=torch.float32))
        _DEVICE_TABLES[key] = tab
    return tab[1][t].view(-1, *([1] * (like.dim() - 1)))


class SpacedDiffusion:
    def __init__(self, steps=250, original_steps=1000, rescale_timesteps=True):
        scale = 1000 / original_steps
        betas_full = np.linspace(scale * 0.0001, scale * 0.02, original_steps, dtype=np.float64)
        acp_full = np.cumprod(1.0 - betas_full)
        # "ddimN": the stride that yields exactly N steps
        use = None
        for stride in range(1, original_steps):
            if len(range(0, original_steps, stride)) == steps:
                use = list(range(0, original_steps, stride))
                break
        if use is None:
            raise ValueError("cannot create exactly %d DDIM steps out of %d" % (steps, original_steps))
        self.timestep_map = np.array(use, dtype=np.int64)
        self.original_steps = original_steps
        self.rescale_timesteps = rescale_timesteps
        last, betas = 1.0, []
        for i in use:
            betas.append(1 - acp_full[i] / last)
            last = acp_full[i]
        self.betas = np.array(betas)
        self.num_timesteps = len(betas)
        self.log_betas = np.log(self.betas)
        self._map_f = self.timestep_map.astype(np.float64) * ((1000.0 / original_steps) if rescale_timesteps else 1.0)
        a = np.cumprod(1.0 - self.betas)
        self.alphas_cumprod = a
        self.alphas_cumprod_prev = np.append(1.0, a[:-1])
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - a)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / a)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / a - 1)
        post_var = self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - a)
        self.posterior_log_variance_clipped = np.log(np.append(post_var[1], post_var[1:]))
        self.posterior_mean_coef1 = self.betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - a)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(1.0 - self.betas) / (1.0 - a)

    # the wrapped model / cond_fn see original timesteps
    def model_timesteps(self, t):
        return _extract(self._map_f, t, t)

    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        model_kwargs = model_kwargs or {}
        out = model(x, self.model_timesteps(t), **model_kwargs)
        eps, var_values = torch.split(out, x.shape[1], dim=1)
        min_log = _extract(self.posterior_log_variance_clipped, t, x)
        max_log = _extract(self.log_betas, t, x)
        frac = (var_values + 1) / 2
        log_variance = frac * max_log + (1 - frac) * min_log
        pred_xstart = _extract(self.sqrt_recip_alphas_cumprod, t, x) * x - _extract(self.sqrt_recipm1_alphas_cumprod, t, x) * eps
        if denoised_fn is not None:
            pred_xstart = denoised_fn(pred_xstart)
        if clip_denoised:
            pred_xstart = pred_xstart.clamp(-1, 1)
        mean = _extract(self.posterior_mean_coef1, t, x) * pred_xstart + _extract(self.posterior_mean_coef2, t, x) * x
        return {"mean": mean, "variance": torch.exp(log_variance), "log_variance": log_variance, "pred_xstart": pred_xstart}

    def _eps_from_xstart(self, x, t, pred_xstart):
        return (_extract(self.sqrt_recip_alphas_cumprod, t, x) * x - pred_xstart) / _extract(self.sqrt_recipm1_alphas_cumprod, t, x)

    def condition_score(self, cond_fn, out, x, t, model_kwargs=None):
        model_kwargs = model_kwargs or {}
        alpha_bar = _extract(self.alphas_cumprod, t, x)
        eps = self._eps_from_xstart(x, t, out["pred_xstart"])
        eps = eps - (1 - alpha_bar).sqrt() * cond_fn(x, self.model_timesteps(t), **model_kwargs)
        out = dict(out)
        out["pred_xstart"] = _extract(self.sqrt_recip_alphas_cumprod, t, x) * x - _extract(self.sqrt_recipm1_alphas_cumprod, t, x) * eps
        out["mean"] = _extract(self.posterior_mean_coef1, t, x) * out["pred_xstart"] + _extract(self.posterior_mean_coef2, t, x) * x
        return out

    def finish_p_mean_variance(self, x, t, pred_xstart_raw, clip_denoised=False, denoised_fn=None):
        """The sampler-side post-processing of p_mean_variance on an already computed raw x0 prediction (used when one
        grad-enabled UNet forward serves both the sampler and cond_fn)."""
        pred_xstart = pred_xstart_raw
        if denoised_fn is not None:
            pred_xstart = denoised_fn(pred_xstart)
        if clip_denoised:
            pred_xstart = pred_xstart.clamp(-1, 1)
        mean = _extract(self.posterior_mean_coef1, t, x) * pred_xstart + _extract(self.posterior_mean_coef2, t, x) * x
        return {"mean": mean, "pred_xstart": pred_xstart}

    @torch.no_grad()
    def ddim_sample(self, model, x, t, clip_denoised=False, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0, noise=None,
                    out_orig=None):
        if out_orig is None:
            out_orig = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, model_kwargs=model_kwargs)
        out = self.condition_score(cond_fn, out_orig, x, t, model_kwargs=model_kwargs) if cond_fn is not None else out_orig
        eps = self._eps_from_xstart(x, t, out["pred_xstart"])
        alpha_bar = _extract(self.alphas_cumprod, t, x)
        alpha_bar_prev = _extract(self.alphas_cumprod_prev, t, x)
        sigma = eta * ((1 - alpha_bar_prev) / (1 - alpha_bar)).sqrt() * (1 - alpha_bar / alpha_bar_prev).sqrt()
        mean_pred = out["pred_xstart"] * alpha_bar_prev.sqrt() + (1 - alpha_bar_prev - sigma ** 2).sqrt() * eps
        if noise is None:
            noise = torch.randn_like(x) if eta > 0 else torch.zeros_like(x)
        nonzero = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        return {"sample": mean_pred + nonzero * sigma * noise, "pred_xstart": out_orig["pred_xstart"]}

    def ddim_sample_loop_progressive(self, model, shape, clip_denoised=False, denoised_fn=None, cond_fn=None, model_kwargs=None, device=None,
                                     eta=0.0, skip_timesteps=0, noise=None):
        x = noise if noise is not None else torch.randn(*shape, device=device)
        for i in range(self.num_timesteps - skip_timesteps - 1, -1, -1):
            t = torch.full((shape[0],), i, device=x.device, dtype=torch.long)
            out = self.ddim_sample(model, x, t, clip_denoised, denoised_fn, cond_fn, model_kwargs, eta)
            yield out
            x = out["sample"]
