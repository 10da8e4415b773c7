"""The per-timestep guidance function (``cond_fn``) of clip_diffusion/sample.py:134-238, B200-native.

Two entry points with the SAME semantics (x, t) -> guidance gradient:

``make_conditon_function``  the reference closure restated line for line on top of the drop-in operators
    (make_cutouts, embed_image, the losses) and torch.autograd, exactly the way the unchanged sample.py
    drives them.  It exists to show the operators ARE a drop-in (same call pattern, same autograd.grad calls,
    same global-RNG consumption) and is what the parity tests compare with the oracle.

``GuidanceStep``            the fast path (north_star (4)): the CLIP part runs WITHOUT the autograd engine --
    fused cutouts -> ViT forward -> single-pass spherical loss+gradient -> ViT dgrad -> cutout backward
    accumulating straight into d(loss)/d(x_in) -- the cutout batch is sharded across ranks (every rank draws
    the full RNG record and takes its slice), ONE all-reduce of the [3,H,W] fp32 image gradient per step, TV
    loss replicated, NaN guard and RMS clamp on device (no host sync).  The UNet forward and VJP stay stock
    PyTorch and replicated.

Both read ``Config`` at call time like the reference (sample.py:162-238).
"""
import torch

from clip_diffusion_b200 import _lib
from clip_diffusion_b200.config import Config
from clip_diffusion_b200.cutouts import cutouts_backward, cutouts_forward, make_cutouts, torch_noise_state
from clip_diffusion_b200.losses import (LPIPS_loss, aesthetic_loss, square_spherical_distance_loss, structural_dissimilarity_loss,
                                        total_variational_loss)
from clip_diffusion_b200.rng_record import draw_cutout_record
from clip_diffusion_b200.utils.functional import embed_image


def make_conditon_function(diffusion, model, clip_models, text_embeddings_and_weights, get_current_timestep, aesthetic_predictors=None,
                           config=Config, init_image_tensor=None, LPIPS_model=None):
    """sample.py:134-238 (the typo in the name is the reference's).  ``get_current_timestep()`` returns the respaced
    index the outer loop maintains (sample.py:285-288).  With ``init_image_tensor`` the init-image branch
    (sample.py:220-225) is evaluated with the caller's ``LPIPS_model`` (the un-vendored VGG LPIPS; skipped when None) and
    the torch MS-SSIM of losses.py -- stock PyTorch, not part of the kernel path."""
    aesthetic_predictors = aesthetic_predictors or {}

    @torch.enable_grad()
    def conditon_function(x, t, y=None):
        x = x.detach().requires_grad_()
        batch_size = x.shape[0]
        current_timestep = get_current_timestep()
        current_timestep_tensor = torch.ones([batch_size], device=x.device, dtype=torch.long) * current_timestep
        p_mean_var = diffusion.p_mean_variance(model, x, current_timestep_tensor, clip_denoised=False, model_kwargs={"y": y})
        factor = float(diffusion.sqrt_one_minus_alphas_cumprod[current_timestep])
        denoised_prediction = p_mean_var["pred_xstart"] * factor + x * (1 - factor)
        grad_tensor = torch.zeros_like(denoised_prediction)
        current_diffusion_step = 1000 - (int(t.item()) + 1)
        n_over = config.num_overview_cuts_schedule[current_diffusion_step]
        n_inner = config.num_inner_cuts_schedule[current_diffusion_step]
        for name, clip_model in clip_models.items():
            for _ in range(config.num_cutout_batches):
                aesthetic_score = None
                cutout_images = make_cutouts(
                    input=denoised_prediction,
                    cut_size=clip_model.visual.input_resolution,
                    num_overview_cuts=n_over,
                    num_inner_cuts=n_inner,
                    inner_cut_size_power=config.inner_cut_size_power_schedule[current_diffusion_step],
                    cut_gray_portion=config.cut_gray_portion_schedule[current_diffusion_step],
                )
                image_embeddings = embed_image(clip_model, cutout_images, clip_normalize=True)
                if config.aesthetic_scale > 0 and name in aesthetic_predictors:
                    aesthetic_score = aesthetic_loss(aesthetic_predictors[name], image_embeddings)
                distances = square_spherical_distance_loss(
                    image_embeddings.unsqueeze(1), text_embeddings_and_weights[name]["embeddings"].unsqueeze(0)
                )
                distances = distances.view([n_over + n_inner, batch_size, -1])
                distance_loss = distances.mul(text_embeddings_and_weights[name]["weights"]).sum(dim=2).mean(dim=0)
                objective = distance_loss.sum() * config.clip_guidance_scale
                if aesthetic_score is not None:
                    objective = objective - aesthetic_score * config.aesthetic_scale
                grad_tensor += torch.autograd.grad(objective, denoised_prediction)[0] / config.num_cutout_batches
        denoise_loss = total_variational_loss(denoised_prediction)
        loss_sum = denoise_loss.sum() * config.denoise_scale
        if init_image_tensor is not None:  # sample.py:220-225
            dissimilarity_loss = structural_dissimilarity_loss(denoised_prediction, init_image_tensor)
            loss_sum = loss_sum + dissimilarity_loss.sum() * getattr(config, "MS_SSIM_scale", 0)
            if LPIPS_model is not None:
                loss_sum = loss_sum + LPIPS_loss(LPIPS_model, denoised_prediction, init_image_tensor).sum() * getattr(config, "LPIPS_scale", 0)
        grad_tensor += torch.autograd.grad(loss_sum, denoised_prediction)[0]
        if not torch.isnan(grad_tensor).any():
            grad = -torch.autograd.grad(denoised_prediction, x, grad_tensor)[0]
        else:
            return torch.zeros_like(x)
        magnitude = grad.square().mean().sqrt()
        return grad * magnitude.clamp(min=-config.grad_threshold, max=config.grad_threshold) / magnitude

    return conditon_function


def make_denoised_function(dynamic_thresholding_percentile=0.995):
    """Imagen-style dynamic thresholding applied to the sampler's x0 prediction (sample.py:116-132): clamp to the
    per-sample |x| quantile (at least 1) and rescale.  Plain torch ops (a selection kernel is the
    next widening step); part of the benchmarked step because the reference passes it as ``denoised_fn`` (sample.py:254,268)."""

    def denoised_function(x_start):
        if x_start.is_cuda and x_start.dtype == torch.float32:
            # radix-select kernel (csrc/select.cu): same threshold as torch.quantile, no sort
            xs = x_start.contiguous()
            b, n = xs.shape[0], xs.numel() // xs.shape[0]
            out = torch.empty_like(xs)
            ws = torch.empty(_lib.load().cg_dynamic_threshold_workspace_bytes(b), dtype=torch.uint8, device=xs.device)
            _lib.call("cg_dynamic_threshold", _lib.ptr(xs), b, n, float(dynamic_thresholding_percentile), 1.0, _lib.ptr(out), None, _lib.ptr(ws))
            return out
        threshold = torch.quantile(x_start.reshape(x_start.shape[0], -1).abs().float(), dynamic_thresholding_percentile, dim=-1)
        threshold = threshold.clamp(min=1.0).view(-1, *((1,) * (x_start.ndim - 1))).to(x_start.dtype)
        return x_start.clamp(min=-threshold, max=threshold) / threshold

    return denoised_function


def shard_range(n, rank, world_size):
    """Contiguous slice [start, stop) of n cutouts owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class GuidanceStep:
    """Fast sharded cond_fn.  ``current_timestep`` (respaced index) is set by the sampling loop before each call,
    like the closure variable of sample.py:113,285-288."""

    def __init__(self, diffusion, model, clip_models, text_embeddings_and_weights, aesthetic_predictors=None, config=Config,
                 rank=0, world_size=1, process_group=None, range_scale=0.0, record_source=None, init_image_tensor=None, LPIPS_model=None):
        self.diffusion, self.model, self.clip_models = diffusion, model, clip_models
        self.text = text_embeddings_and_weights
        self.aesthetic_predictors = aesthetic_predictors or {}
        self.config = config
        self.rank, self.world_size, self.group = rank, world_size, process_group
        self.range_scale = range_scale  # rgb_range_loss is dead code in the reference (losses.py:31-35); opt-in here (config C3)
        # optional callable (name, batch, H, W, cs, n_over, n_inner, power, gray) -> CutoutRecord replacing the global-RNG
        # draw (parity tests feed the oracle and the kernels one explicit record)
        self.record_source = record_source
        # init-image branch (sample.py:220-225): stock-PyTorch LPIPS (caller's network) + MS-SSIM terms, replicated on every rank
        self.init_image_tensor, self.LPIPS_model = init_image_tensor, LPIPS_model
        self.current_timestep = None
        self.last_records = []  # RNG records of the latest call, per (model, cutout batch)
        self.cutouts_processed = 0
        self._scratch = None
        self._ms_ws = None

    # ---- the CLIP part: d(sum of CLIP objectives)/d(x_in), local shard only ---------------------------------
    def clip_guidance_grad(self, x_in, current_diffusion_step, grad_out):
        cfg = self.config
        H, W = int(x_in.shape[-2]), int(x_in.shape[-1])
        n_over = cfg.num_overview_cuts_schedule[current_diffusion_step]
        n_inner = cfg.num_inner_cuts_schedule[current_diffusion_step]
        power = cfg.inner_cut_size_power_schedule[current_diffusion_step]
        gray = cfg.cut_gray_portion_schedule[current_diffusion_step]
        n_total = n_over + n_inner
        nb = cfg.num_cutout_batches
        self.last_records = []
        for name, clip_model in self.clip_models.items():
            tower = clip_model.visual.tower
            txt = self.text[name]["embeddings"].reshape(-1, tower.output_dim).float().contiguous()
            wts = self.text[name]["weights"].reshape(-1).float().contiguous()
            if wts.numel() == 1 and txt.shape[0] > 1:
                wts = wts.expand(txt.shape[0]).contiguous()
            for b in range(nb):
                # every rank draws the FULL record (same seed => same record), then takes its slice
                if self.record_source is not None:
                    rec = self.record_source(name, b, H, W, tower.input_resolution, n_over, n_inner, power, gray)
                else:
                    rec = draw_cutout_record(H, W, tower.input_resolution, n_over, n_inner, power, gray, noise="device")
                    rec.noise_torch = torch_noise_state(x_in.device, rec.num_cuts, tower.input_resolution)  # the reference's CUDA randn stream
                self.last_records.append(rec)
                start, stop = shard_range(n_total, self.rank, self.world_size)
                if stop <= start:
                    continue
                local = rec if self.world_size == 1 else rec.slice(start, stop)
                patches, saved = cutouts_forward(x_in, local, fmt=_lib.CG_FMT_BF16_PATCH, patch=tower.patch, kpad=tower.kpad,
                                                 augment=True, normalize=True)
                emb = tower.forward_patches(patches)
                demb = torch.empty_like(emb)
                # objective = clip_guidance_scale * mean_n sum_p w_p d[n,p]  (sample.py:194-206), divided by num_cutout_batches (:207)
                coef = float(cfg.clip_guidance_scale) / (n_total * nb)
                _lib.call("cg_spherical_loss_fwd_bwd", _lib.ptr(emb), _lib.ptr(txt), _lib.ptr(wts), emb.shape[0], txt.shape[0],
                          tower.output_dim, coef, None, _lib.ptr(demb))
                if cfg.aesthetic_scale > 0 and name in self.aesthetic_predictors:
                    with torch.enable_grad():
                        e = emb.detach().requires_grad_()
                        score = self.aesthetic_predictors[name](torch.nn.functional.normalize(e, dim=-1)).sum() / n_total
                        (ge,) = torch.autograd.grad(-score * (float(cfg.aesthetic_scale) / nb), e)
                    demb += ge
                dpatch = tower.backward_patches(demb)
                cutouts_backward(dpatch, saved, coef=1.0, dx_in=grad_out)  # accumulates into d(loss)/d(x_in)
                self.cutouts_processed += stop - start
        return grad_out

    @torch.enable_grad()
    def cond_fn(self, x, t, y=None):
        x = x.detach().requires_grad_()
        batch_size = x.shape[0]
        if batch_size != 1:
            raise ValueError("the guidance path is defined for batch size 1 (sample.py:246-251)")
        ct = self.current_timestep
        ts = torch.full([batch_size], ct, device=x.device, dtype=torch.long)
        p_mean_var = self.diffusion.p_mean_variance(self.model, x, ts, clip_denoised=False, model_kwargs={"y": y})
        return self._guidance_from_prediction(x, p_mean_var)

    def ddim_step(self, x, i, eta=0.0, denoised_fn=None, clip_denoised=False, reuse_forward=True, noise=None):
        """One guided DDIM step (what the sampler generator does per iteration, sample.py:241-287).

        reuse_forward=True: the reference evaluates the UNet twice at the same (x, t) -- once
        without grad inside the sampler's p_mean_variance and once with grad inside cond_fn (sample.py:149-151).  Both
        give the same numbers, so ONE grad-enabled forward serves both; results are identical to the two-forward path.
        ``noise`` replaces the sampler's ``randn_like`` draw (parity tests feed both sides the same tensor)."""
        self.current_timestep = i
        t = torch.full((x.shape[0],), i, device=x.device, dtype=torch.long)
        if not reuse_forward:
            return self.diffusion.ddim_sample(self.model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=self.cond_fn,
                                              model_kwargs={}, eta=eta, noise=noise)
        with torch.enable_grad():
            xg = x.detach().requires_grad_()
            p_mean_var = self.diffusion.p_mean_variance(self.model, xg, t, clip_denoised=False, model_kwargs={"y": None})
            guidance = self._guidance_from_prediction(xg, p_mean_var)
        with torch.no_grad():
            out_orig = self.diffusion.finish_p_mean_variance(x, t, p_mean_var["pred_xstart"].detach(), clip_denoised, denoised_fn)
        return self.diffusion.ddim_sample(self.model, x, t, cond_fn=lambda *_a, **_k: guidance, model_kwargs={}, eta=eta, out_orig=out_orig, noise=noise)

    def _guidance_from_prediction(self, x, p_mean_var):
        cfg = self.config
        ct = self.current_timestep
        factor = float(self.diffusion.sqrt_one_minus_alphas_cumprod[ct])
        denoised_prediction = p_mean_var["pred_xstart"] * factor + x * (1 - factor)
        x_in = denoised_prediction.detach().float().contiguous()
        # 1000 - (int(t)+1): t is the original timestep of respaced index ct (host-side integer: no device sync)
        original_t = int(self.diffusion.timestep_map[ct] * (1000.0 / self.diffusion.original_steps)) if self.diffusion.rescale_timesteps \
            else int(self.diffusion.timestep_map[ct])
        current_diffusion_step = 1000 - (original_t + 1)
        grad_tensor = torch.zeros(3, x_in.shape[-2], x_in.shape[-1], device=x.device, dtype=torch.float32)
        self.clip_guidance_grad(x_in, current_diffusion_step, grad_tensor)
        H, W = x_in.shape[-2:]
        if self._scratch is None:
            self._scratch = torch.zeros(4, device=x.device, dtype=torch.float32)
        flag, scratch = self._scratch[:2], self._scratch[2:]
        has_init = self.init_image_tensor is not None
        # replicated image-space losses (sample.py:217-218), value + gradient in one pass, accumulated into the same buffer.  Every rank adds
        # 1/world of them BEFORE the reduce, so the one collective of the step delivers the finished d(loss)/d(x_in) (exact for
        # power-of-two world sizes) and nothing but the NaN test follows it.
        share = 1.0 / self.world_size
        fused = W % 128 == 0
        if fused:  # TV + range + NaN flag of the finished gradient in ONE launch
            _lib.call("cg_image_losses_fwd_bwd", _lib.ptr(x_in), 1, 3, H, W, float(cfg.denoise_scale) * share, float(self.range_scale) * share, 1, None,
                      _lib.ptr(grad_tensor), _lib.ptr(flag))
        else:
            _lib.call("cg_tv_loss_fwd_bwd", _lib.ptr(x_in), 1, 3, H, W, float(cfg.denoise_scale) * share, 1, None, _lib.ptr(grad_tensor))
            if self.range_scale:
                _lib.call("cg_range_loss_fwd_bwd", _lib.ptr(x_in), 1, 3, H, W, float(self.range_scale) * share, 1, None, _lib.ptr(grad_tensor))
        if has_init:  # sample.py:220-225: MS-SSIM is evaluated whenever an init image exists, even at scale 0 (config.py:52)
            init = self.init_image_tensor.reshape(3, H, W).float().contiguous()
            nbytes = _lib.load().cg_ms_ssim_workspace_bytes(3, H, W)
            if nbytes == 0:
                raise _lib.ClipGuideError(_lib.load().cg_last_error().decode("utf-8", "replace"))
            if self._ms_ws is None or self._ms_ws.numel() < nbytes:
                self._ms_ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            _lib.call("cg_ms_ssim_dissimilarity_fwd_bwd", _lib.ptr(x_in), _lib.ptr(init), 3, H, W, float(getattr(cfg, "MS_SSIM_scale", 0)) * share, 1, None,
                      _lib.ptr(grad_tensor), _lib.ptr(self._ms_ws))
            if self.LPIPS_model is not None:  # the caller's (un-vendored) VGG LPIPS network: stock PyTorch, outside the kernel path
                with torch.enable_grad():
                    xi = x_in.view(1, *x_in.shape[-3:]).detach().requires_grad_()
                    extra = LPIPS_loss(self.LPIPS_model, xi, self.init_image_tensor).sum() * getattr(cfg, "LPIPS_scale", 0)
                    (gi,) = torch.autograd.grad(extra, xi)
                grad_tensor += gi.view_as(grad_tensor) * share
        if self.world_size > 1:
            torch.distributed.all_reduce(grad_tensor, group=self.group)  # the one collective of the step
        self.last_grad_tensor = grad_tensor
        if self.world_size > 1 or has_init or not fused:
            _lib.call("cg_any_nan", _lib.ptr(grad_tensor), grad_tensor.numel(), _lib.ptr(flag))
        (grad,) = torch.autograd.grad(denoised_prediction, x, grad_tensor.view_as(denoised_prediction).to(denoised_prediction.dtype))
        grad = grad.float().contiguous()
        out = torch.empty_like(grad)
        _lib.call("cg_grad_finalize", _lib.ptr(grad), grad.numel(), -1.0, float(cfg.grad_threshold), _lib.ptr(flag), _lib.ptr(out), _lib.ptr(scratch))
        return out.to(x.dtype)

    __call__ = cond_fn
